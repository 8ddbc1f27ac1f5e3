// Arguments and the per-node arithmetic of the bulk sweep, shared by the sweep kernels (lbm_bulk.cu) and the persistent small-lattice
// kernel (lbm_small.cu).  Included INSIDE namespace life (and, in the -DLIFE_EXACT compilation, inside life::exact): the including
// translation unit decides which variant of the arithmetic it instantiates (see lbm_bulk.cu).
#pragma once

struct BulkArgs {
	const double *fin;
	double *fout;
	PopShift ps;            // cfg.inplace: layout of the ONE buffer (fin == fout); all zero otherwise
	Layout L;
	double omega;
	double fup_x, fup_y;    // uniform force_xy entering u_n   (previous step's value)
	double fuc_x, fuc_y;    // uniform force_xy of this step   (collision)
	const double *fibm;     // IBM force planes or nullptr
	const uint8_t *fmask;   // per (column, 64-row span) flag "force_ibm written here" (ctx.h), or nullptr: read the planes everywhere
	int64_t mask_pitch;
	// generic path only
	const double *macro;    // stored rho_n, ux_n, uy_n planes (first step after an upload) or nullptr
	double *fxyf;           // force_xy field planes or nullptr
	int wom_field;          // recompute the force_xy field from rho_n (Womersley with gravity, src/Grid.cpp:55-61)
	double Drho, gx, gy, dpdx_cos, dpdy_cos, sq, Dm;
	int64_t c_first;
	int64_t tiles;          // thread blocks per column
};

// mode bits of the specialised kernels
enum { M_NONE = 0, M_UNI = 1, M_IBM = 2, M_UNI_IBM = 3, M_GENERIC = 4 };

// start-of-step macroscopics + collision of one node held in registers
// does the 64-row span holding row j of column `col` carry any IBM force?  (warp-uniform in every kernel below: a warp's rows lie
// in one span, or — QUAD — each thread's four rows do)
__device__ __forceinline__ bool ibm_span(const BulkArgs &a, int64_t col, int64_t j) {
	return a.fmask == nullptr || a.fmask[col * a.mask_pitch + (j >> 6)] != 0;
}

template <int COLL, int MODE>
__device__ __forceinline__ void node_update(const BulkArgs &a, int64_t idx, const double (&f)[NV], double (&o)[NV], bool ibm_here = true) {
	double sum, mx, my;
	moments(f, sum, mx, my);
	double Fux = 0.0, Fuy = 0.0, Fcx = 0.0, Fcy = 0.0;   // force entering u_n / force of the collision
	double rho = sum, ux, uy;
	if (MODE == M_GENERIC) {
		double fix = 0.0, fiy = 0.0;
		if (a.fibm && ibm_here) { fix = a.fibm[idx]; fiy = a.fibm[a.L.S + idx]; }
		double fpx = a.fup_x, fpy = a.fup_y, fcx = a.fuc_x, fcy = a.fuc_y;
		if (a.fxyf) { fpx = a.fxyf[idx]; fpy = a.fxyf[a.L.S + idx]; fcx = fpx; fcy = fpy; }
		if (a.macro) {
			rho = a.macro[idx]; ux = a.macro[a.L.S + idx]; uy = a.macro[2 * a.L.S + idx];
		} else {
#ifdef LIFE_EXACT
			ux = (mx + 0.5 * (fpx + fix)) / rho;      // src/IBMNode.cpp:121-122 (== src/Grid.cpp:297-298 where force_ibm is 0)
			uy = (my + 0.5 * (fpy + fiy)) / rho;
#else
			const double inv = 1.0 / rho;
			ux = (mx + 0.5 * (fpx + fix)) * inv;
			uy = (my + 0.5 * (fpy + fiy)) * inv;
#endif
		}
		if (a.wom_field) {
			fcx = (rho * a.Drho * a.gx + a.dpdx_cos) * a.sq / a.Dm;
			fcy = (rho * a.Drho * a.gy + a.dpdy_cos) * a.sq / a.Dm;
			a.fxyf[idx] = fcx;
			a.fxyf[a.L.S + idx] = fcy;
		}
		Fcx = fcx + fix; Fcy = fcy + fiy;
	} else {
		if (MODE & M_UNI) { Fux = a.fup_x; Fuy = a.fup_y; Fcx = a.fuc_x; Fcy = a.fuc_y; }
		if (MODE & M_IBM) {
			// force_ibm is zero outside the <= 9 n support sites: the two planes are only read where the span's flag is set, so
			// a lattice with bodies still moves ~144 B per node
			double fix = 0.0, fiy = 0.0;
			if (ibm_here) { fix = a.fibm[idx]; fiy = a.fibm[a.L.S + idx]; }
			Fux += fix; Fuy += fiy; Fcx += fix; Fcy += fiy;
		}
#ifdef LIFE_EXACT
		if (MODE == M_NONE) { ux = mx / rho; uy = my / rho; }      // (sum c f + 0.5 * 0) / rho, src/Grid.cpp:297-298
		else { ux = (mx + 0.5 * Fux) / rho; uy = (my + 0.5 * Fuy) / rho; }
#else
		const double inv = 1.0 / rho;
		if (MODE == M_NONE) { ux = mx * inv; uy = my * inv; }
		else { ux = (mx + 0.5 * Fux) * inv; uy = (my + 0.5 * Fuy) * inv; }
#endif
	}
	constexpr bool HASF = MODE != M_NONE;
#ifdef LIFE_EXACT
	// both operators in the reference's own evaluation order, bit for bit (d2q9.cuh)
	if (COLL == COLL_CM) collide_cm_ref(f, rho, ux, uy, Fcx, Fcy, a.omega, o);
	else collide_bgk_ref<HASF>(f, rho, ux, uy, Fcx, Fcy, a.omega, o);
#else
	if (COLL == COLL_CM) collide_cm<HASF>(f, sum, mx, my, rho, ux, uy, Fcx, Fcy, a.omega, o);
	else collide_bgk<HASF>(f, rho, ux, uy, Fcx, Fcy, a.omega, o);
#endif
}

// BulkArgs of the context's current state for a sweep from fA into fB starting at local column c_first; *mode = the M_* specialisation
static inline BulkArgs make_bulk_args(life_ctx *ctx, const StepScalars &sc, int64_t c_first, int *mode_out) {
	BulkArgs a{};
	a.fin = ctx->fA;
	a.fout = ctx->inplace ? ctx->fA : ctx->fB;
	a.ps = ctx->shift;
	a.L = ctx->L;
	a.omega = ctx->cfg.omega;
	a.fup_x = sc.fxy_prev[0]; a.fup_y = sc.fxy_prev[1];
	a.fuc_x = sc.fxy_cur[0]; a.fuc_y = sc.fxy_cur[1];
	a.fibm = ctx->fibm_any ? ctx->fibm : nullptr;
	a.fmask = (a.fibm && !ctx->fibm_full_dirty) ? ctx->fibm_mask : nullptr;     // uploaded force_ibm: sites unknown, read everywhere
	if (ctx->cfg.tune == 20) a.fmask = nullptr;     // measurement only: the round-1 behaviour (both force planes read at every node)
	a.mask_pitch = ctx->mask_pitch;
	a.c_first = c_first;
	const bool generic = ctx->stored_macro_valid || ctx->fxy_mode == FXY_FIELD;
	int mode;
	if (generic) {
		mode = M_GENERIC;
		a.macro = ctx->stored_macro_valid ? ctx->macro : nullptr;
		a.fxyf = ctx->fxy_mode == FXY_FIELD ? ctx->fxyf : nullptr;
		a.wom_field = ctx->wom_field ? 1 : 0;
		a.Drho = ctx->cfg.Drho; a.gx = ctx->cfg.gravity_x; a.gy = ctx->cfg.gravity_y;
		a.dpdx_cos = ctx->cfg.dpdx * sc.wom_cos; a.dpdy_cos = ctx->cfg.dpdy * sc.wom_cos;
		a.sq = (ctx->cfg.Dx * ctx->cfg.Dt) * (ctx->cfg.Dx * ctx->cfg.Dt);
		a.Dm = ctx->cfg.Dm;
	} else {
		const bool uni = ctx->fxy_mode == FXY_UNIFORM;
		mode = (uni ? M_UNI : 0) | (a.fibm ? M_IBM : 0);
	}
	*mode_out = mode;
	return a;
}

