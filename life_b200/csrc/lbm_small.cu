// Small lattices: MANY time steps in ONE launch of ONE thread-block cluster.
//
// The example cases of the reference are tiny (LidDrivenCavity 101 x 101 = 10 201 nodes, ChannelFlow 501 x 51): the whole state
// (2 x 72 B/node) sits in L2 and a time step is ~1 microsecond of arithmetic, so stepping them with one launch per kernel
// (sweep, ring, boundary: 2-4 launches per step) is bound by launch latency and by the gaps between dependent launches.  Here the
// n steps of life_step_n run inside one kernel: a single cluster of 8 CTAs x 512 threads (8 SMs) strides over the nodes, and
// the phases of a step — [outlet speed] -> sweep -> ghost ring -> boundary nodes — are separated by the cluster's hardware barrier
// (barrier.cluster, release / acquire at cluster scope: the populations live in global memory, i.e. in L2, and become visible to
// the other CTAs of the cluster across the barrier) instead of by kernel boundaries.  The per-node arithmetic is the sweep's and
// the boundary kernel's own device code (lbm_bulk.cuh, lbm_boundary.cuh; this file is compiled twice like they are): in the exact
// build results are bit-identical to stepping with life_step, in the default build identical up to which product of an a*b + c*d the
// compiler contracts into the FMA (rounding level).
//
// Measured (profiles/r02_small_lattice_timing.txt): LidDrivenCavity 9.4 us/step against 10.2 with per-step launches, fully periodic
// 40 x 36: 7.4 against 26 (the per-step path needs 5+ launches there); above ~16 k nodes the 8 SMs lose to the whole-GPU kernels.
// Used by life_step_n (api.cu) when the lattice is small enough for 8 SMs to beat the launch-bound path (<= SMALL_MAX_NODES), there
// is one rank, no force field and no stored macroscopics; everything else takes the per-step path.
#include "ctx.h"
#include "d2q9.cuh"
#include <cooperative_groups.h>

namespace life {
#ifdef LIFE_EXACT
namespace exact {
#endif

#include "lbm_bulk.cuh"
#include "lbm_boundary.cuh"

constexpr int SMALL_CLUSTER = 8;
constexpr int SMALL_THREADS = 512;

struct SmallArgs {
	BulkArgs bulk;            // fin / fout are swapped every step inside the kernel
	BcArgs bc;                // likewise fprev / f
	const StepScalars *sc;    // [n] per-step scalars (host-evaluated: ramp, Womersley cosine, uniform forces)
	int n_steps;
	int convective;           // right wall is a convective outlet
	int wrap_bottom, wrap_top, left_periodic, right_periodic;
};

template <int COLL, int MODE>
__global__ void __cluster_dims__(SMALL_CLUSTER, 1, 1) __launch_bounds__(SMALL_THREADS, 1) k_steps_small(SmallArgs a) {
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	const int64_t tid = (int64_t)blockIdx.x * SMALL_THREADS + threadIdx.x;
	const int64_t nth = (int64_t)SMALL_CLUSTER * SMALL_THREADS;
	const Layout L = a.bulk.L;
	const int64_t nodes = L.nxl * L.Ny;
	double *fin = const_cast<double *>(a.bulk.fin), *fout = a.bulk.fout;

	for (int s = 0; s < a.n_steps; s++) {
		const StepScalars sc = a.sc[s];
		BulkArgs b = a.bulk;
		b.fin = fin; b.fout = fout;
		b.fup_x = sc.fxy_prev[0]; b.fup_y = sc.fxy_prev[1];
		b.fuc_x = sc.fxy_cur[0]; b.fuc_y = sc.fxy_cur[1];
		BcArgs c = a.bc;
		c.fprev = fin; c.f = fout;
		c.ramp = sc.ramp;
		c.fcur.ux = sc.fxy_cur[0]; c.fcur.uy = sc.fxy_cur[1];
		c.fprv.ux = sc.fxy_prev[0]; c.fprv.uy = sc.fxy_prev[1];

		// convective outlet speed from the state before the step (src/Grid.cpp:39-40): one CTA, finished before the boundary phase
		if (a.convective && blockIdx.x == SMALL_CLUSTER - 1) convective_speed_block(fin, PopShift{}, nullptr, L, c.fprv, const_cast<double *>(c.delU));

		// sweep: stream + collide of every node (src/Grid.cpp:65-84)
		for (int64_t n = tid; n < nodes; n += nth) {
			const int64_t col = 1 + n / L.Ny, j = n % L.Ny;
			const int64_t idx = col * L.P + j + JOFF;
			double f[NV], o[NV];
#pragma unroll
			for (int v = 0; v < NV; v++) f[v] = fin[v * L.S + idx];
			node_update<COLL, MODE>(b, idx, f, o, ibm_span(b, col, j));
#pragma unroll
			for (int v = 0; v < NV; v++) fout[v * L.S + idx + LIFE_CX(v) * L.P + LIFE_CY(v)] = o[v];
		}
		cluster.sync();

		// ghost ring: periodic wrap in y, then in x (the modulo of src/Grid.cpp:229); the x copies read what the y wrap wrote in the
		// ghost columns, hence the second barrier — only taken by lattices that are periodic somewhere
		if (a.wrap_bottom || a.wrap_top) {
			for (int64_t col = tid; col <= L.nxl + 1; col += nth) wrap_y_column(fout, PopShift{}, L, a.wrap_bottom, a.wrap_top, 0, col);
			cluster.sync();
		}
		if (a.left_periodic || a.right_periodic) {
			for (int64_t e = tid; e < 3 * L.Ny; e += nth) {
				const int k = (int)(e / L.Ny);
				const int64_t r = JOFF + e % L.Ny;
				const int vr = k == 0 ? 1 : (k == 1 ? 5 : 7), vl = k == 0 ? 2 : (k == 1 ? 6 : 8);      // cx = +1 / cx = -1
				if (a.left_periodic) fout[vr * L.S + L.at(1, r)] = fout[vr * L.S + L.at(L.nxl + 1, r)];
				if (a.right_periodic) fout[vl * L.S + L.at(L.nxl, r)] = fout[vl * L.S + L.at(0, r)];
			}
			cluster.sync();
		}

		// boundary conditions (src/Grid.cpp:87-98)
		for (int64_t k = tid; k < c.n; k += nth) bc_node<COLL>(c, k);
		cluster.sync();

		double *t = fin; fin = fout; fout = t;
	}
}

template <int COLL, int MODE>
static int launch_small(life_ctx *ctx, const SmallArgs &a) {
	k_steps_small<COLL, MODE><<<SMALL_CLUSTER, SMALL_THREADS, 0, ctx->stream>>>(a);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

#ifdef LIFE_EXACT
}  // namespace exact
using namespace exact;
int launch_steps_small_exact(life_ctx *ctx, const StepScalars *d_sc, const StepScalars &first, int n) {
#else
int launch_steps_small(life_ctx *ctx, const StepScalars *d_sc, const StepScalars &first, int n) {
#endif
	SmallArgs a{};
	int mode;
	a.bulk = make_bulk_args(ctx, first, 1, &mode);
	a.bc = make_bc_args(ctx, first);
	a.sc = d_sc;
	a.n_steps = n;
	const life_config &c = ctx->cfg;
	a.convective = c.wall_right == LIFE_CONVECTIVE;
	a.wrap_bottom = c.wall_bottom == LIFE_FLUID;
	a.wrap_top = c.wall_top == LIFE_FLUID;
	const bool tb = a.wrap_bottom || a.wrap_top;
	a.left_periodic = c.wall_left == LIFE_FLUID || tb;       // as halo.cu: exchange_x
	a.right_periodic = c.wall_right == LIFE_FLUID || tb;
	const bool cm = c.collision == LIFE_CENTRAL_MOMENTS;
	switch (mode) {
	case M_NONE: return cm ? launch_small<COLL_CM, M_NONE>(ctx, a) : launch_small<COLL_BGK, M_NONE>(ctx, a);
	case M_UNI: return cm ? launch_small<COLL_CM, M_UNI>(ctx, a) : launch_small<COLL_BGK, M_UNI>(ctx, a);
	default: return fail(ctx, LIFE_E_STATE, "persistent small-lattice kernel: unsupported force mode");
	}
}

}  // namespace life
