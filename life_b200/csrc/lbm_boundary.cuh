// Device code of everything in the LBM step that touches O(Nx + Ny) nodes — macroscopics of single nodes, the convective outlet speed,
// the boundary conditions — shared by the per-step kernels (lbm_boundary.cu) and the persistent small-lattice kernel (lbm_small.cu).
// Included INSIDE namespace life (and, in the -DLIFE_EXACT compilation, inside life::exact).
#pragma once

// ---------------------------------------------------------------------------------------------------------------------
// shared device helpers: macroscopics of an arbitrary node recomputed from its populations
// ---------------------------------------------------------------------------------------------------------------------
struct ForceView {
	int mode;                 // FXY_*
	double ux, uy;            // uniform force_xy
	const double *field;      // force_xy planes (FXY_FIELD)
	const double *fibm;       // force_ibm planes or nullptr
	int64_t S;
	__device__ __forceinline__ void xy(int64_t idx, double &fx, double &fy) const {
		if (mode == FXY_FIELD) { fx = field[idx]; fy = field[S + idx]; }
		else { fx = ux; fy = uy; }
	}
	__device__ __forceinline__ void ibm(int64_t idx, double &fx, double &fy) const {
		if (fibm) { fx = fibm[idx]; fy = fibm[S + idx]; } else { fx = 0.0; fy = 0.0; }
	}
};

__device__ __forceinline__ void load9(const double *f, const PopShift &ps, int64_t S, int64_t idx, double (&o)[NV]) {
#pragma unroll
	for (int v = 0; v < NV; v++) o[v] = f[ps.at(v, idx, S)];
}

// rho and u of GridClass::macroscopic (src/Grid.cpp:282-299): u = (sum c f + F_xy/2)/rho — no IBM force (mid-step value)
__device__ __forceinline__ void macro_mid(const double *f, const PopShift &ps, const Layout &L, const ForceView &fv, int64_t idx, double &rho,
                                          double &ux, double &uy) {
	double p[NV], mx, my, fx, fy;
	load9(f, ps, L.S, idx, p);
	moments(p, rho, mx, my);
	fv.xy(idx, fx, fy);
	ux = (mx + 0.5 * fx) / rho;
	uy = (my + 0.5 * fy) / rho;
}

// end-of-step value: u = (sum c f + (F_xy + F_ibm)/2)/rho (src/IBMNode.cpp:121-122; equals macroscopic() off-support)
__device__ __forceinline__ void macro_end(const double *f, const PopShift &ps, const Layout &L, const ForceView &fv, int64_t idx, double &rho,
                                          double &ux, double &uy) {
	double p[NV], mx, my, fx, fy, gx, gy;
	load9(f, ps, L.S, idx, p);
	moments(p, rho, mx, my);
	fv.xy(idx, fx, fy);
	fv.ibm(idx, gx, gy);
	ux = (mx + 0.5 * (fx + gx)) / rho;
	uy = (my + 0.5 * (fy + gy)) / rho;
}

// ---------------------------------------------------------------------------------------------------------------------
// convective outlet speed (src/Grid.cpp:477-495), on the rank that owns the last three columns.
// One block; u of the previous step's end at columns Nx-1, Nx-2, Nx-3 is recomputed from the state buffer (or read from the
// uploaded macroscopics on the first step).  The column mean is a fixed-shape tree sum (deterministic).
// ---------------------------------------------------------------------------------------------------------------------
// block-level: call with ALL threads of one CTA (blockDim.x a power of two <= 1024)
__device__ __forceinline__ void convective_speed_block(const double *f, const PopShift &ps, const double *stored, const Layout &L, const ForceView &fv, double *delU) {
	const int64_t c1 = L.nxl, c2 = L.nxl - 1, c3 = L.nxl - 2;   // local columns of i = Nx-1, Nx-2, Nx-3
#ifdef LIFE_EXACT
	// the reference's serial loop (src/Grid.cpp:483-487): every thread fetches its u_x, one thread adds them up in j order
	__shared__ double s_uout;
	for (int64_t j = threadIdx.x; j < L.Ny; j += blockDim.x) {
		const int64_t idx = L.at(c1, j + JOFF);
		double rho, ux, uy;
		if (stored) ux = stored[L.S + idx];
		else macro_end(f, ps, L, fv, idx, rho, ux, uy);
		delU[2 * j] = ux;                                   // staging: overwritten with delU below, after the sum has been read
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		double acc = 0.0;
		for (int64_t j = 0; j < L.Ny; j++) acc += delU[2 * j];
		s_uout = acc / static_cast<double>(L.Ny);
	}
	__syncthreads();
	const double uOut = s_uout;
#else
	__shared__ double red[1024];
	double part = 0.0;
	for (int64_t j = threadIdx.x; j < L.Ny; j += blockDim.x) {
		const int64_t idx = L.at(c1, j + JOFF);
		double rho, ux, uy;
		if (stored) ux = stored[L.S + idx];
		else macro_end(f, ps, L, fv, idx, rho, ux, uy);
		part += ux;
	}
	red[threadIdx.x] = part;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
		__syncthreads();
	}
	const double uOut = red[0] / (double)L.Ny;
#endif
	for (int64_t j = threadIdx.x; j < L.Ny; j += blockDim.x) {
		double r, a[2], b[2], c[2];
		const int64_t i1 = L.at(c1, j + JOFF), i2 = L.at(c2, j + JOFF), i3 = L.at(c3, j + JOFF);
		if (stored) {
			a[0] = stored[L.S + i1]; a[1] = stored[2 * L.S + i1];
			b[0] = stored[L.S + i2]; b[1] = stored[2 * L.S + i2];
			c[0] = stored[L.S + i3]; c[1] = stored[2 * L.S + i3];
		} else {
			macro_end(f, ps, L, fv, i1, r, a[0], a[1]);
			macro_end(f, ps, L, fv, i2, r, b[0], b[1]);
			macro_end(f, ps, L, fv, i3, r, c[0], c[1]);
		}
		delU[2 * j] = (-uOut / 2.0) * (3.0 * a[0] - 4.0 * b[0] + c[0]);
		delU[2 * j + 1] = (-uOut / 2.0) * (3.0 * a[1] - 4.0 * b[1] + c[1]);
	}
}

static inline ForceView force_view(life_ctx *ctx, const double uni[2]) {
	ForceView fv{};
	fv.mode = ctx->fxy_mode;
	fv.ux = uni[0]; fv.uy = uni[1];
	fv.field = ctx->fxyf;
	fv.fibm = ctx->fibm_any ? ctx->fibm : nullptr;
	fv.S = ctx->L.S;
	return fv;
}

// ---------------------------------------------------------------------------------------------------------------------
// boundary conditions of one BCVec entry
// ---------------------------------------------------------------------------------------------------------------------
struct BcArgs {
	const BcNode *bc;
	int64_t n;
	const double *fprev;      // state before this step (f_n): convective outlet, stale u_n component (null when `captured` is set)
	double *f;                // freshly streamed populations (f)
	PopShift ps;              // layout of f (and of fprev in the two-buffer mode, where it is all zero)
	const double *captured;   // cfg.inplace: per BCVec entry, what is needed of the state before the sweep (k_bc_capture), else null
	const double *stored;     // uploaded u_n/rho_n if this is the first step, else nullptr
	Layout L;
	ForceView fcur;           // forces of this step (new macroscopics of neighbours)
	ForceView fprv;           // forces in effect at the end of the previous step
	const double *u_in, *rho_in, *delU;
	double ramp;
};

template <int COLL>
__device__ __forceinline__ void bc_node(const BcArgs &a, const int64_t b) {
	const BcNode bn = a.bc[b];
	const Layout &L = a.L;
	const int64_t idx = L.node(bn.il, bn.j);
	const int nx = bn.nx, ny = bn.ny, nd = bn.nd, type = bn.type;

	if (type == LIFE_CONVECTIVE) {   // convectiveBC, src/Grid.cpp:468-474: only the three incoming populations
		const double dux = a.delU[2 * bn.j], duy = a.delU[2 * bn.j + 1];
		const double p2 = a.captured ? a.captured[3 * b] : a.fprev[2 * L.S + idx];
		const double p6 = a.captured ? a.captured[3 * b + 1] : a.fprev[6 * L.S + idx];
		const double p8 = a.captured ? a.captured[3 * b + 2] : a.fprev[8 * L.S + idx];
		a.f[a.ps.at(2, idx, L.S)] = p2 + 3.0 * W1 * (dux * -1.0 + duy * 0.0);
		a.f[a.ps.at(6, idx, L.S)] = p6 + 3.0 * W2 * (dux * -1.0 + duy * -1.0);
		a.f[a.ps.at(8, idx, L.S)] = p8 + 3.0 * W2 * (dux * -1.0 + duy * 1.0);
		return;
	}

	double f[NV];
	load9(a.f, a.ps, L.S, idx, f);

	// u_n / rho_n start as the end-of-previous-step values of this node (they are only read back in one case: the normal
	// velocity of a pressure corner, which applyBCs leaves untouched)
	double un[2] = {0.0, 0.0}, rhon = 1.0;
	const bool corner = (nx != 0 && ny != 0);
	if (type == LIFE_PRESSURE && corner) {
		if (a.stored) { un[0] = a.stored[L.S + idx]; un[1] = a.stored[2 * L.S + idx]; }
		else if (a.captured) { un[0] = a.captured[3 * b]; un[1] = a.captured[3 * b + 1]; }
		else { double r; macro_end(a.fprev, a.ps, L, a.fprv, idx, r, un[0], un[1]); }
	}

	// interior neighbours along the normal (inc/Utils.h:151-161, :200-210)
	const int64_t i1 = idx + nx * L.P + ny, i2 = idx + 2 * (nx * L.P + ny);

	// applyBCs, src/Grid.cpp:312-375
	if (type == LIFE_WALL) {
		un[0] = 0.0; un[1] = 0.0;
	} else if (type == LIFE_VELOCITY) {
		un[0] = a.u_in[2 * bn.j] * a.ramp;
		un[1] = a.u_in[2 * bn.j + 1] * a.ramp;
	} else {   // free slip / pressure: tangential velocity by 2nd-order zero gradient of the NEW u
		double r1, u1[2], r2, u2[2];
		macro_mid(a.f, a.ps, L, a.fcur, i1, r1, u1[0], u1[1]);
		macro_mid(a.f, a.ps, L, a.fcur, i2, r2, u2[0], u2[1]);
		const int dt = 1 - nd;
		if (type == LIFE_FREESLIP) un[nd] = 0.0;
		else rhon = a.rho_in[bn.j];
		un[dt] = (4.0 / 3.0) * u1[dt] - (1.0 / 3.0) * u2[dt];
	}

	// regularisedBC, src/Grid.cpp:387-465
	const int nn = nd == 0 ? nx : ny;
	if (corner) {
		if (type != LIFE_PRESSURE) {   // density extrapolated along the diagonal from the NEW rho (src/Grid.cpp:394)
			double r1, r2, t0, t1;
			macro_mid(a.f, a.ps, L, a.fcur, i1, r1, t0, t1);
			macro_mid(a.f, a.ps, L, a.fcur, i2, r2, t0, t1);
			rhon = 2.0 * r1 - r2;
		}
	} else {
		double fplus = 0.0, fzero = 0.0;
#pragma unroll
		for (int v = 0; v < NV; v++) {
			const int cn = nd == 0 ? kCx[v] : kCy[v];
			if (cn == -nn) fplus += f[v];
			else if (cn == 0) fzero += f[v];
		}
		if (type != LIFE_PRESSURE) rhon = (2.0 * fplus + fzero) / (1.0 - nn * un[nd]);
		else un[nd] = nn * (1.0 - (2.0 * fplus + fzero) / rhon);
	}

	double feq[NV];
#pragma unroll
	for (int v = 0; v < NV; v++) feq[v] = equilibrium<COLL>(rhon, un[0], un[1], v);

	double Sxx = 0.0, Syy = 0.0, Sxy = 0.0;
#pragma unroll
	for (int v = 0; v < NV; v++) {
		const int cx = kCx[v], cy = kCy[v];
		double fv = f[v];
		if (corner) {
			if (cx == nx || cy == ny) {
				if (nx * cx + ny * cy == 0) fv = feq[v];                        // buried link
				else fv = feq[v] + (f[LIFE_OPP(v)] - feq[LIFE_OPP(v)]);
			}
		} else {
			const int cn = nd == 0 ? cx : cy;
			if (cn == nn) fv = feq[v] + (f[LIFE_OPP(v)] - feq[LIFE_OPP(v)]);
		}
		const double fneq = fv - feq[v];
		Sxx += cx * cx * fneq;
		Syy += cy * cy * fneq;
		Sxy += cx * cy * fneq;
	}
#pragma unroll
	for (int v = 0; v < NV; v++) {
		const int cx = kCx[v], cy = kCy[v];
		const double w = v == 0 ? W0 : (v < 5 ? W1 : W2);
		a.f[a.ps.at(v, idx, L.S)] = feq[v] + (w / (2.0 * CS4)) * (((cx * cx - CS2) * Sxx) + ((cy * cy - CS2) * Syy) + (2.0 * cx * cy * Sxy));
	}
}

static inline BcArgs make_bc_args(life_ctx *ctx, const StepScalars &sc) {
	BcArgs a{};
	a.bc = ctx->bc; a.n = ctx->n_bc;
	a.fprev = ctx->fA; a.f = ctx->fB;
	if (ctx->inplace) {      // one buffer, already advanced to the post-stream layout; the pre-sweep values come from k_bc_capture
		a.fprev = nullptr; a.f = ctx->fA;
		a.ps = ctx->shift;
		a.captured = ctx->bc_prev;
	} else if (ctx->wom_field && ctx->bc_prev) {
		a.captured = ctx->bc_prev;      // force_xy field rewritten by the sweep: pre-sweep u_n of pressure corners saved by k_bc_capture
	}
	a.stored = ctx->stored_macro_valid ? ctx->macro : nullptr;
	a.L = ctx->L;
	a.fcur = force_view(ctx, sc.fxy_cur);
	a.fprv = force_view(ctx, sc.fxy_prev);
	a.u_in = ctx->u_in; a.rho_in = ctx->rho_in; a.delU = ctx->delU;
	a.ramp = sc.ramp;
	return a;
}

// y wrap-around of one column c of the ghost ring (see lbm_boundary.cu: k_wrap_y)
__device__ __forceinline__ void wrap_y_column(double *f, const PopShift &ps, const Layout &L, int wrap_to_bottom, int wrap_to_top, int after_exchange, int64_t c) {
	if (after_exchange && (c == 0 || c == L.nxl + 1)) return;
	const bool skip_plus = after_exchange && c == 1;        // cx = +1 planes: 5 (cy=+1), 7 (cy=-1)
	const bool skip_minus = after_exchange && c == L.nxl;   // cx = -1 planes: 8 (cy=+1), 6 (cy=-1)
	const int64_t base = c * L.P;
	if (wrap_to_bottom) {   // cy = +1 populations: v = 3, 5, 8
		f[ps.at(3, base + JOFF, L.S)] = f[ps.at(3, base + JOFF + L.Ny, L.S)];
		if (!skip_plus) f[ps.at(5, base + JOFF, L.S)] = f[ps.at(5, base + JOFF + L.Ny, L.S)];
		if (!skip_minus) f[ps.at(8, base + JOFF, L.S)] = f[ps.at(8, base + JOFF + L.Ny, L.S)];
	}
	if (wrap_to_top) {      // cy = -1 populations: v = 4, 6, 7
		f[ps.at(4, base + JOFF + L.Ny - 1, L.S)] = f[ps.at(4, base + JOFF - 1, L.S)];
		if (!skip_minus) f[ps.at(6, base + JOFF + L.Ny - 1, L.S)] = f[ps.at(6, base + JOFF - 1, L.S)];
		if (!skip_plus) f[ps.at(7, base + JOFF + L.Ny - 1, L.S)] = f[ps.at(7, base + JOFF - 1, L.S)];
	}
}
