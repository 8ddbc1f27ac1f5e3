// Everything of the LBM step that touches only O(Nx + Ny) nodes:
//   * site types, BCVec and normals                       (src/Grid.cpp:925-951, :498-545) — built on the host once
//   * the y wrap-around of the ghost ring                 (periodic modulo of src/Grid.cpp:229, applied to the ring)
//   * convective outlet speed                             (src/Grid.cpp:477-495)
//   * boundary conditions + their macroscopic update      (src/Grid.cpp:87-98 → applyBCs :302-384, regularisedBC :387-465,
//                                                          convectiveBC :468-474, Utils::extrapolate / zeroGradient
//                                                          inc/Utils.h:137-217)
// The boundary kernel runs after the bulk sweep (and the halo exchange) on the freshly streamed buffer; it needs the
// *new* rho/u of up to two interior neighbours, which it recomputes from their nine populations.
//
// Compiled twice (life_b200/build.py): as is, and with -DLIFE_EXACT -fmad=false, which puts the convective-speed and boundary
// kernels into namespace life::exact behind launch_convective_speed_exact / launch_boundary_exact (cfg.exact).  The expressions
// below are written in the reference's association order, so without FMA contraction every boundary population, the outlet
// speed (then a serial sum in j order, src/Grid.cpp:483-484) and delU are the reference's doubles bit for bit.
#include "ctx.h"
#include "d2q9.cuh"
#include <cmath>

namespace life {

#ifndef LIFE_EXACT
#include "lbm_boundary.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// host: type matrix of the slab, BCVec order, normals
// ---------------------------------------------------------------------------------------------------------------------
static inline int site_type(const life_config &c, int64_t i, int64_t j) {
	int t = LIFE_FLUID;
	if (i == 0) t = c.wall_left;
	else if (i == c.Nx - 1) t = c.wall_right;
	if (j == 0) t = c.wall_bottom;          // bottom/top override left/right at the corners (src/Grid.cpp:940-945)
	else if (j == c.Ny - 1) t = c.wall_top;
	return t;
}

int build_boundary(life_ctx *ctx) {
	const life_config &c = ctx->cfg;
	const int64_t Ny = c.Ny, nxl = ctx->L.nxl;
	if (c.wall_left == LIFE_CONVECTIVE || c.wall_bottom == LIFE_CONVECTIVE || c.wall_top == LIFE_CONVECTIVE)
		return fail(ctx, LIFE_E_ARG, "Currently convective BC only supported for right boundary");   // src/Grid.cpp:919-920
	ctx->h_type.assign((size_t)(nxl * Ny), LIFE_FLUID);
	ctx->h_bc.clear();
	for (int64_t il = 0; il < nxl; il++) {
		const int64_t i = ctx->i_begin + il;
		const bool xedge = (i == 0 || i == c.Nx - 1);
		for (int64_t j = 0; j < Ny; j++) {
			if (!xedge && j != 0 && j != Ny - 1) continue;
			const int t = site_type(c, i, j);
			ctx->h_type[(size_t)(il * Ny + j)] = t;
			if (t == LIFE_FLUID) continue;
			BcNode b{};
			b.il = (int32_t)il; b.j = (int32_t)j; b.type = (int8_t)t;
			// getNormalVector, src/Grid.cpp:498-545
			int nx = 0, ny = 0, nd = 0;
			if (i == 0) { nx = 1; nd = 0; } else if (i == c.Nx - 1) { nx = -1; nd = 0; }
			if (j == 0) { ny = 1; nd = 1; } else if (j == Ny - 1) { ny = -1; nd = 1; }
			if (nx != 0 && ny != 0) {
				const int tx = site_type(c, i + nx, j), ty = site_type(c, i, j + ny);
				if (tx == LIFE_FLUID && ty == LIFE_FLUID)
					return fail(ctx, LIFE_E_ARG, "Corner node is surrounded by fluid lattice sites");   // src/Grid.cpp:527-528
				else if (tx == LIFE_FLUID) { nd = 0; ny = 0; }
				else if (ty == LIFE_FLUID) { nd = 1; nx = 0; }
			}
			b.nx = (int8_t)nx; b.ny = (int8_t)ny; b.nd = (int8_t)nd;
			ctx->h_bc.push_back(b);
		}
	}
	ctx->n_bc = (int64_t)ctx->h_bc.size();
	if (ctx->n_bc > 0) {
		LIFE_CUDA(ctx, cudaMalloc(&ctx->bc, sizeof(BcNode) * ctx->n_bc));
		LIFE_CUDA(ctx, cudaMemcpy(ctx->bc, ctx->h_bc.data(), sizeof(BcNode) * ctx->n_bc, cudaMemcpyHostToDevice));
	}
	return LIFE_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// y wrap-around: a population pushed out through the top lands in ghost row Ny; the reference's modulo puts it in row 0
// (and vice versa).  Only consumed where the receiving row is periodic (type eFluid), so only applied then.
// Covers the ghost columns too, so corner populations end up where the x exchange expects them.
// ---------------------------------------------------------------------------------------------------------------------
// `after_exchange` = second pass of a multi-rank step (after the interior sweep, concurrent with the halo receive): it must
// leave alone what the x exchange delivers — the cx = +1 planes of the first column, the cx = -1 planes of the last column —
// and the ghost columns, which were wrapped by the first pass before they were sent.
__global__ void k_wrap_y(double *f, PopShift ps, Layout L, int wrap_to_bottom, int wrap_to_top, int after_exchange) {
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c > L.nxl + 1) return;
	wrap_y_column(f, ps, L, wrap_to_bottom, wrap_to_top, after_exchange, c);
}

int launch_wrap_y(life_ctx *ctx, cudaStream_t st, bool after_exchange) {
	const int wb = ctx->cfg.wall_bottom == LIFE_FLUID, wt = ctx->cfg.wall_top == LIFE_FLUID;
	if (!wb && !wt) return LIFE_OK;
	const int64_t n = ctx->L.nxl + 2;
	k_wrap_y<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->inplace ? ctx->fA : ctx->fB, ctx->shift, ctx->L, wb, wt, after_exchange ? 1 : 0);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

#else    // LIFE_EXACT
namespace exact {
#include "lbm_boundary.cuh"
#endif   // LIFE_EXACT

__global__ void __launch_bounds__(1024) k_convective_speed(const double *f, PopShift ps, const double *stored, Layout L, ForceView fv, double *delU) {
	convective_speed_block(f, ps, stored, L, fv, delU);
}

// cfg.inplace: the sweep overwrites the state it reads, so what the boundary kernel needs of the state BEFORE the sweep is saved first,
// 3 doubles per BCVec entry: f_n[2], f_n[6], f_n[8] of a convective outlet node (src/Grid.cpp:468-474), u_n of a pressure corner (the
// normal component applyBCs leaves untouched, src/Grid.cpp:356-369).
__global__ void __launch_bounds__(128) k_bc_capture(const BcNode *bc, int64_t n, const double *f, PopShift ps, Layout L, ForceView fprv, double *out) {
	const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n) return;
	const BcNode bn = bc[b];
	const int64_t idx = L.node(bn.il, bn.j);
	if (bn.type == LIFE_CONVECTIVE) {
		out[3 * b] = f[ps.at(2, idx, L.S)];
		out[3 * b + 1] = f[ps.at(6, idx, L.S)];
		out[3 * b + 2] = f[ps.at(8, idx, L.S)];
	} else if (bn.type == LIFE_PRESSURE && bn.nx != 0 && bn.ny != 0) {
		double r, ux, uy;
		macro_end(f, ps, L, fprv, idx, r, ux, uy);
		out[3 * b] = ux;
		out[3 * b + 1] = uy;
	}
}

template <int COLL>
__global__ void __launch_bounds__(128) k_boundary(const BcArgs a) {
	const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= a.n) return;
	bc_node<COLL>(a, b);
}

#ifdef LIFE_EXACT
}  // namespace exact
using namespace exact;
int launch_convective_speed_exact(life_ctx *ctx, const StepScalars &sc) {
#else
int launch_convective_speed(life_ctx *ctx, const StepScalars &sc) {
#endif
	if (ctx->cfg.wall_right != LIFE_CONVECTIVE || ctx->i_end != ctx->cfg.Nx) return LIFE_OK;
	if (ctx->L.nxl < 3) return fail(ctx, LIFE_E_ARG, "convective outlet needs the last three columns on one rank");
	ForceView fv = force_view(ctx, sc.fxy_prev);
	k_convective_speed<<<1, 1024, 0, ctx->stream>>>(ctx->fA, ctx->shift, ctx->stored_macro_valid ? ctx->macro : nullptr, ctx->L, fv,
	                                                ctx->delU);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

#ifdef LIFE_EXACT
int launch_bc_capture_exact(life_ctx *ctx, const StepScalars &sc) {
#else
int launch_bc_capture(life_ctx *ctx, const StepScalars &sc) {
#endif
	// also with two buffers when force_xy is a field the sweep rewrites (Womersley with gravity): the retained normal velocity of a
	// pressure corner belongs to the PREVIOUS step's force (src/Grid.cpp:356-369), which is gone after the sweep
	if (!(ctx->inplace || ctx->wom_field) || ctx->n_bc == 0) return LIFE_OK;
	const bool needed = ctx->cfg.wall_right == LIFE_CONVECTIVE || ctx->cfg.wall_left == LIFE_PRESSURE || ctx->cfg.wall_right == LIFE_PRESSURE ||
	                    ctx->cfg.wall_bottom == LIFE_PRESSURE || ctx->cfg.wall_top == LIFE_PRESSURE;
	if (!needed) return LIFE_OK;
	if (!ctx->bc_prev) LIFE_CUDA(ctx, cudaMalloc(&ctx->bc_prev, sizeof(double) * 3 * (size_t)ctx->n_bc));
	k_bc_capture<<<(unsigned)((ctx->n_bc + 127) / 128), 128, 0, ctx->stream>>>(ctx->bc, ctx->n_bc, ctx->fA, ctx->shift, ctx->L,
	                                                                          force_view(ctx, sc.fxy_prev), ctx->bc_prev);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

#ifdef LIFE_EXACT
int launch_boundary_exact(life_ctx *ctx, const StepScalars &sc) {
#else
int launch_boundary(life_ctx *ctx, const StepScalars &sc) {
#endif
	if (ctx->n_bc == 0) return LIFE_OK;
	const BcArgs a = make_bc_args(ctx, sc);
	const unsigned blocks = (unsigned)((a.n + 127) / 128);
	if (ctx->cfg.collision == LIFE_CENTRAL_MOMENTS) k_boundary<COLL_CM><<<blocks, 128, 0, ctx->stream>>>(a);
	else k_boundary<COLL_BGK><<<blocks, 128, 0, ctx->stream>>>(a);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

}  // namespace life
