// End-of-step macroscopics of one node, evaluated from the populations and the current forces: what GridClass::rho / ::u hold
// after a completed step (src/Grid.cpp:282-299 + src/IBMNode.cpp:97-136).  Shared by the download, scan and file kernels.
#pragma once
#include "ctx.h"
#include "d2q9.cuh"

namespace life {

struct MacroArgs {
	const double *f;
	PopShift ps;          // layout of f (zeros unless cfg.inplace)
	int shifted;          // any offset non-zero
	Layout L;
	int fxy_mode;
	double fx, fy;
	const double *fxyf, *fibm;
};

// populations of the node into p[], and its macroscopics.  Every operation below rounds the same way with or without FMA
// contraction (the products are by 0.5, i.e. exact), so all kernels that inline this return bit-identical values.
__device__ __forceinline__ void node_macro_p(const MacroArgs &a, int64_t idx, double (&p)[NV], double &rho, double &ux, double &uy) {
	double mx, my;
#pragma unroll
	for (int v = 0; v < NV; v++) p[v] = __ldg(a.f + a.ps.at(v, idx, a.L.S));
	moments(p, rho, mx, my);
	double fx = a.fx, fy = a.fy;
	if (a.fxy_mode == FXY_FIELD) { fx = a.fxyf[idx]; fy = a.fxyf[a.L.S + idx]; }
	if (a.fibm) {
		// (F_xy + F_ibm)/2 as src/IBMNode.cpp:121-122; off-support F_ibm = 0 and this equals src/Grid.cpp:297-298
		ux = (mx + 0.5 * (fx + a.fibm[idx])) / rho;
		uy = (my + 0.5 * (fy + a.fibm[a.L.S + idx])) / rho;
	} else {
		ux = (mx + 0.5 * fx) / rho;
		uy = (my + 0.5 * fy) / rho;
	}
}

__device__ __forceinline__ void node_macro(const MacroArgs &a, int64_t idx, double &rho, double &ux, double &uy) {
	double p[NV];
	node_macro_p(a, idx, p, rho, ux, uy);
}

// arguments for the context's current state (lbm_io.cu)
MacroArgs macro_args(life_ctx *ctx);

}  // namespace life
