// Bulk sweep of the D2Q9 step: stream + collide of every node of a column range, fused in one pass over HBM.
//
// Replaces, for all nodes at once, the reference's hot loops 1 and 2 (src/Grid.cpp:65-84): streamCollide (:103-246) with
// equilibrium (:249-264) and latticeForce (:267-279), and the macroscopic pass (:282-299) — the latter is not stored at
// all: the start-of-step rho_n/u_n a node needs are recomputed from its own nine populations and the forces
// (after every completed step u == (sum c f + (F_xy + F_ibm)/2) / sum f at every node).
//
// State convention = the reference's: the buffer holds post-stream, pre-collision populations at their own node.
// Each thread reads the nine populations of its node(s) from `fin` (coalesced, 16-byte vector loads), collides, and
// pushes population v to (c + cx, r + cy) of `fout`.  The ghost ring of the layout (ctx.h) absorbs pushes that leave
// the slab or wrap around, so there is no modulo and no branch in here (the reference wraps with two `%` per
// population, src/Grid.cpp:229).
//
// Algorithmic traffic: 9 loads + 9 stores of 8 B = 144 B per lattice update (+16 B when an IBM force field is read).
//
// Variants (cfg.kernel, identical results):
//   DIRECT : one node per thread, scalar shifted stores.
//   SHUFFLE: two nodes per thread; the six populations that move in y are re-aligned across the warp with shuffles so
//            that all but two stores per warp and plane are 16-byte aligned vector stores.
//
// This file is compiled twice (life_b200/build.py): as is — the fast sweep, factored collisions, FMA contraction allowed — and with
// -DLIFE_EXACT -fmad=false, which puts the same kernels into namespace life::exact with the collision and the start-of-step
// macroscopics written in the reference's operation order (cfg.exact; d2q9.cuh: collide_bgk_ref) behind launch_bulk_exact.
#include "ctx.h"
#include "d2q9.cuh"

namespace life {
#ifdef LIFE_EXACT
namespace exact {
#endif

#include "lbm_bulk.cuh"

// ---- DIRECT: one node per thread ------------------------------------------------------------------------------------------
template <int COLL, int MODE>
__global__ void __launch_bounds__(256) k_bulk_direct(const BulkArgs a) {
	const int64_t col = a.c_first + blockIdx.x / a.tiles;
	const int64_t j = (int64_t)(blockIdx.x % a.tiles) * blockDim.x + threadIdx.x;
	if (j >= a.L.Ny) return;
	const int64_t idx = col * a.L.P + j + JOFF;
	double f[NV], o[NV];
#pragma unroll
	for (int v = 0; v < NV; v++) f[v] = __ldg(a.fin + v * a.L.S + idx);
	node_update<COLL, MODE>(a, idx, f, o, ibm_span(a, col, j));
#pragma unroll
	for (int v = 0; v < NV; v++) a.fout[v * a.L.S + idx + LIFE_CX(v) * a.L.P + LIFE_CY(v)] = o[v];
}

// ---- SHIFT: the in-place sweep (cfg.inplace), ONE population buffer -----------------------------------------------------------------------
// Population v of logical element n sits at plane element (n - off_v) mod S (ctx.h: PopShift).  A node reads its nine populations,
// collides, and writes every result back to the slot it was read from: with off_v growing by shift_v = cx*P + cy after the sweep,
// that slot IS element n + shift_v of the next step's layout — the push to the neighbour happens by renaming, not by moving data.
// No second buffer, no ordering constraint between nodes (each touches only its own nine slots), the same kernel every step, and
// the same 144 B/node of traffic.  The offsets change parity every step, so accesses are 8-byte (fully coalesced along y; the DIRECT
// variant shows what that costs against 16-byte accesses: ~1 %).
template <int COLL, int MODE>
__global__ void __launch_bounds__(256) k_bulk_shift(const BulkArgs a) {
	const int64_t col = a.c_first + blockIdx.x / a.tiles;
	const int64_t j = (int64_t)(blockIdx.x % a.tiles) * blockDim.x + threadIdx.x;
	if (j >= a.L.Ny) return;
	const int64_t idx = col * a.L.P + j + JOFF;
	int64_t at[NV];
	double f[NV], o[NV];
#pragma unroll
	for (int v = 0; v < NV; v++) {
		at[v] = a.ps.at(v, idx, a.L.S);
		f[v] = a.fout[at[v]];
	}
	node_update<COLL, MODE>(a, idx, f, o, ibm_span(a, col, j));
#pragma unroll
	for (int v = 0; v < NV; v++) a.fout[at[v]] = o[v];
}

// ---- SHUFFLE: two nodes per thread, warp-shuffle realignment of the y-moving populations -------------------------------------
// A warp owns 64 consecutive rows R..R+63 (R even) of one column.  Thread `lane` holds rows R+2*lane and R+2*lane+1.
// Population with cy = +1: row q goes to row q+1.  The aligned pair (R+2l, R+2l+1) of the destination therefore consists of
// the upper element of lane l-1 and the lower element of lane l: one shuffle-up, then lanes 1..31 issue one aligned 16-byte
// store each; lane 0 stores row R+1 alone and lane 31 stores row R+64 alone.  cy = -1 is the mirror image (shuffle-down).
// HINT selects the cache policy of the population traffic: 0 = read-only path loads + default stores, 1 = streaming
// (evict-first) loads and stores.  BLOCK is the CTA size.  Both only matter for tuning (cfg.tune); results are identical.
template <int HINT>
__device__ __forceinline__ double2 ld_pair(const double *p) {
	if (HINT == 1) return __ldcs(reinterpret_cast<const double2 *>(p));
	return __ldg(reinterpret_cast<const double2 *>(p));
}
template <int HINT>
__device__ __forceinline__ void st_pair(double *p, double x, double y) {
	if (HINT == 1) __stcs(reinterpret_cast<double2 *>(p), make_double2(x, y));
	else *reinterpret_cast<double2 *>(p) = make_double2(x, y);
}
template <int HINT>
__device__ __forceinline__ void st_one(double *p, double x) {
	if (HINT == 1) __stcs(p, x);
	else *p = x;
}

// MINB = 0 leaves the register budget to the compiler (72 registers, 3 CTAs of 256 threads per SM); note that (256, 1) is NOT the
// same thing: it lets ptxas spend 84 registers, which rounds to 2 CTAs per SM and costs 6 % (profiles/r02_bulk_tune_sweep_16384.txt)
template <int COLL, int MODE, int BLOCK = 256, int HINT = 0, int MINB = 0>
__global__ void __launch_bounds__(BLOCK, MINB) k_bulk_shuffle(const BulkArgs a) {
	const int64_t col = a.c_first + blockIdx.x / a.tiles;
	const int64_t j = ((int64_t)(blockIdx.x % a.tiles) * blockDim.x + threadIdx.x) * 2;
	const int lane = threadIdx.x & 31;
	const bool v0ok = j < a.L.Ny, v1ok = j + 1 < a.L.Ny;
	// whole warps beyond the column exit together (Ny tail): shuffles below need converged warps only among survivors
	const int64_t jw = j - 2 * lane;
	if (jw >= a.L.Ny) return;
	const int64_t idx = col * a.L.P + j + JOFF;   // even → 16-byte aligned
	double f0[NV], f1[NV], o0[NV], o1[NV];
#pragma unroll
	for (int v = 0; v < NV; v++) {
		// rows beyond Ny still lie inside the padded pitch (ghost row + padding), so the vector load is always in bounds
		const double2 t = ld_pair<HINT>(a.fin + v * a.L.S + idx);
		f0[v] = t.x; f1[v] = t.y;
	}
	const bool ih = ibm_span(a, col, j);
	if (v0ok) node_update<COLL, MODE>(a, idx, f0, o0, ih);
	if (v1ok) node_update<COLL, MODE>(a, idx + 1, f1, o1, ih);
#pragma unroll
	for (int v = 0; v < NV; v++) {
		double *dst = a.fout + v * a.L.S + idx + LIFE_CX(v) * a.L.P;
		if (LIFE_CY(v) == 0) {
			if (v1ok) st_pair<HINT>(dst, o0[v], o1[v]);
			else if (v0ok) st_one<HINT>(dst, o0[v]);
		} else if (LIFE_CY(v) == 1) {
			// destination rows j+1, j+2.  aligned pair (j, j+1) = { upper of lane-1 , own lower }
			const double up = __shfl_up_sync(0xffffffffu, o1[v], 1);
			if (lane == 0) { if (v0ok) st_one<HINT>(dst + 1, o0[v]); }
			else if (v0ok) st_pair<HINT>(dst, up, o0[v]);
			else if (j - 1 < a.L.Ny) st_one<HINT>(dst, up);            // first thread past the end still owns row j (= upper of lane-1)
			if (lane == 31 && v1ok) st_one<HINT>(dst + 2, o1[v]);
		} else {
			// destination rows j-1, j.  aligned pair (j, j+1) = { own upper , lower of lane+1 }
			const double dn = __shfl_down_sync(0xffffffffu, o0[v], 1);
			if (lane == 0 && v0ok) st_one<HINT>(dst - 1, o0[v]);
			if (lane == 31) { if (v1ok) st_one<HINT>(dst, o1[v]); }
			else if (j + 2 < a.L.Ny) st_pair<HINT>(dst, o1[v], dn);
			else if (v1ok) st_one<HINT>(dst, o1[v]);
		}
	}
}


// ---- QUAD: four nodes per thread, 32-byte accesses (sm_100: LDG.E.256 / STG.E.256) ------------------------------------------------
// Same idea as SHUFFLE with twice the span: a warp owns 128 consecutive rows (R a multiple of 4) of one column, thread `lane`
// holds rows R+4*lane .. R+4*lane+3: nine 32-byte loads, four collisions in registers, and per plane ONE aligned 32-byte store —
// directly for cy = 0; for cy = +1 the aligned group {R+4l..R+4l+3} of the destination is {row 3 of lane l-1, rows 0, 1, 2 of
// lane l} (one shuffle-up), for cy = -1 it is {rows 1, 2, 3 of lane l, row 0 of lane l+1} (one shuffle-down); only the two end
// lanes of a warp store partial groups.  Needs Ny % 128 == 0 so that every lane of every warp owns four real rows (the launcher
// falls back to SHUFFLE otherwise).  Selected with cfg.kernel = LIFE_KERNEL_QUAD; results identical to the other variants.
__device__ __forceinline__ void ld_quad(const double *p, double (&x)[4]) {
	asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void st_quad(double *p, double a, double b, double c, double d) {
	asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int COLL, int MODE, int BLOCK = 128, int MINB = 3>
__global__ void __launch_bounds__(BLOCK, MINB) k_bulk_quad(const BulkArgs a) {
	const int64_t col = a.c_first + blockIdx.x / a.tiles;
	const int64_t j = ((int64_t)(blockIdx.x % a.tiles) * BLOCK + threadIdx.x) * 4;
	const int lane = threadIdx.x & 31;
	if (j >= a.L.Ny) return;                      // whole warps (Ny % 128 == 0)
	const int64_t idx = col * a.L.P + j + JOFF;   // multiple of 4 doubles -> 32-byte aligned
	double f[4][NV], o[4][NV];
#pragma unroll
	for (int v = 0; v < NV; v++) {
		double t[4];
		ld_quad(a.fin + v * a.L.S + idx, t);
#pragma unroll
		for (int k = 0; k < 4; k++) f[k][v] = t[k];
	}
	const bool ih = ibm_span(a, col, j);
#pragma unroll
	for (int k = 0; k < 4; k++) node_update<COLL, MODE>(a, idx + k, f[k], o[k], ih);
#pragma unroll
	for (int v = 0; v < NV; v++) {
		double *dst = a.fout + v * a.L.S + idx + LIFE_CX(v) * a.L.P;
		if (LIFE_CY(v) == 0) {
			st_quad(dst, o[0][v], o[1][v], o[2][v], o[3][v]);
		} else if (LIFE_CY(v) == 1) {
			// destination rows j+1 .. j+4; aligned group (j .. j+3) = { row 3 of lane-1, own rows 0, 1, 2 }
			const double up = __shfl_up_sync(0xffffffffu, o[3][v], 1);
			if (lane == 0) { dst[1] = o[0][v]; *reinterpret_cast<double2 *>(dst + 2) = make_double2(o[1][v], o[2][v]); }
			else st_quad(dst, up, o[0][v], o[1][v], o[2][v]);
			if (lane == 31) dst[4] = o[3][v];
		} else {
			// destination rows j-1 .. j+2; aligned group (j .. j+3) = { own rows 1, 2, 3, row 0 of lane+1 }
			const double dn = __shfl_down_sync(0xffffffffu, o[0][v], 1);
			if (lane == 0) dst[-1] = o[0][v];
			if (lane == 31) { *reinterpret_cast<double2 *>(dst) = make_double2(o[1][v], o[2][v]); dst[2] = o[3][v]; }
			else st_quad(dst, o[1][v], o[2][v], o[3][v], dn);
		}
	}
}

#ifndef LIFE_EXACT
// ---- TMA: persistent CTAs, populations prefetched through shared memory by bulk asynchronous copies -----------------------------
// 2 CTAs per SM stay resident for the whole sweep and walk the (column, 512-row tile) list with stride gridDim.x.  One elected
// thread arms an mbarrier with the tile's byte count and issues nine cp.async.bulk copies (one contiguous 4 KB run per plane:
// SoA + y-fastest layout means a tile of a plane IS a contiguous run, so the 1-D bulk form needs no tensor map); the copies of
// the next TMA_STAGES-1 tiles are in flight while the current one is collided.  Stores leave from registers exactly as in the
// SHUFFLE variant.  SASS: UBLKCP.S.G + SYNCS (mbarrier), no LDG on the population path.
constexpr int TMA_TILE = 512;      // rows per tile = 2 per thread
constexpr int TMA_STAGES = 3;
constexpr int TMA_SMEM_BYTES = TMA_STAGES * NV * TMA_TILE * 8 + TMA_STAGES * 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "LIFE_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra LIFE_DONE;\n"
	    "bra LIFE_WAIT;\n"
	    "LIFE_DONE:\n"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

template <int COLL, int MODE>
__global__ void __launch_bounds__(256, 2) k_bulk_tma(const BulkArgs a, const int64_t n_tiles) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *buf = reinterpret_cast<double *>(smem_raw);                                   // [stage][plane][row]
	uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + TMA_STAGES * NV * TMA_TILE * 8);
	const int tid = threadIdx.x, lane = tid & 31;

	if (tid == 0) {
		for (int s = 0; s < TMA_STAGES; s++) mbar_init(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// elected thread: arm the stage's barrier and start the nine plane copies of `tile`
	auto issue = [&](int64_t tile, int s) {
		const int64_t col = a.c_first + tile / a.tiles;
		const int64_t j0 = (tile % a.tiles) * TMA_TILE;
		int64_t rows = a.L.P - JOFF - j0;             // never run past the end of this column's pitch
		if (rows > TMA_TILE) rows = TMA_TILE;
		const unsigned bytes = (unsigned)(rows * 8);   // P, JOFF, j0 are multiples of 16 rows: 16-byte granular
		const int64_t idx0 = col * a.L.P + j0 + JOFF;
		mbar_expect_tx(&full[s], bytes * NV);
#pragma unroll
		for (int v = 0; v < NV; v++) bulk_g2s(buf + (s * NV + v) * TMA_TILE, a.fin + v * a.L.S + idx0, bytes, &full[s]);
	};

	if (tid == 0)
		for (int s = 0; s < TMA_STAGES; s++) {
			const int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
			if (tile < n_tiles) issue(tile, s);
		}

	int s = 0;
	unsigned phase = 0;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t col = a.c_first + tile / a.tiles;
		const int64_t j = (tile % a.tiles) * TMA_TILE + 2 * tid;
		mbar_wait(&full[s], phase);
		double f0[NV], f1[NV], o0[NV], o1[NV];
#pragma unroll
		for (int v = 0; v < NV; v++) {
			const double2 t = *reinterpret_cast<const double2 *>(buf + (s * NV + v) * TMA_TILE + 2 * tid);
			f0[v] = t.x; f1[v] = t.y;
		}
		__syncthreads();                               // every thread has drained this stage: it can be refilled
		if (tid == 0) {
			const int64_t next = tile + (int64_t)TMA_STAGES * gridDim.x;
			if (next < n_tiles) issue(next, s);
		}
		if (++s == TMA_STAGES) { s = 0; phase ^= 1; }

		const bool v0ok = j < a.L.Ny, v1ok = j + 1 < a.L.Ny;
		const int64_t jw = j - 2 * lane;
		if (jw >= a.L.Ny) continue;                    // whole warp beyond the column (warp-uniform)
		const int64_t idx = col * a.L.P + j + JOFF;
		const bool ih = ibm_span(a, col, j);
		if (v0ok) node_update<COLL, MODE>(a, idx, f0, o0, ih);
		if (v1ok) node_update<COLL, MODE>(a, idx + 1, f1, o1, ih);
#pragma unroll
		for (int v = 0; v < NV; v++) {
			double *dst = a.fout + v * a.L.S + idx + LIFE_CX(v) * a.L.P;
			if (LIFE_CY(v) == 0) {
				if (v1ok) st_pair<0>(dst, o0[v], o1[v]);
				else if (v0ok) st_one<0>(dst, o0[v]);
			} else if (LIFE_CY(v) == 1) {
				const double up = __shfl_up_sync(0xffffffffu, o1[v], 1);
				if (lane == 0) { if (v0ok) st_one<0>(dst + 1, o0[v]); }
				else if (v0ok) st_pair<0>(dst, up, o0[v]);
				else if (j - 1 < a.L.Ny) st_one<0>(dst, up);
				if (lane == 31 && v1ok) st_one<0>(dst + 2, o1[v]);
			} else {
				const double dn = __shfl_down_sync(0xffffffffu, o0[v], 1);
				if (lane == 0 && v0ok) st_one<0>(dst - 1, o0[v]);
				if (lane == 31) { if (v1ok) st_one<0>(dst, o1[v]); }
				else if (j + 2 < a.L.Ny) st_pair<0>(dst, o1[v], dn);
				else if (v1ok) st_one<0>(dst, o1[v]);
			}
		}
	}
}

template <int COLL, int MODE>
static int launch_tma(life_ctx *ctx, BulkArgs a, int64_t c_count, cudaStream_t st) {
	// per launch: the attribute is per device, and one process may hold contexts on several GPUs (cheap: no synchronisation)
	LIFE_CUDA(ctx, cudaFuncSetAttribute(k_bulk_tma<COLL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
	a.tiles = (a.L.Ny + TMA_TILE - 1) / TMA_TILE;
	const int64_t n_tiles = a.tiles * c_count;
	if (n_tiles <= 0) return LIFE_OK;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	const int64_t grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;      // persistent: two resident CTAs per SM
	k_bulk_tma<COLL, MODE><<<(unsigned)grid, 256, TMA_SMEM_BYTES, st>>>(a, n_tiles);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

#endif   // !LIFE_EXACT

template <int COLL, int MODE>
static int launch_one(life_ctx *ctx, const BulkArgs &a0, int64_t c_count, cudaStream_t st) {
#ifndef LIFE_EXACT
	if (ctx->cfg.kernel == LIFE_KERNEL_TMA && !ctx->inplace) return launch_tma<COLL, MODE>(ctx, a0, c_count, st);
#endif
	BulkArgs a = a0;
	if (ctx->inplace) {
		a.tiles = (a.L.Ny + 255) / 256;
		const int64_t blocks = a.tiles * c_count;
		if (blocks <= 0) return LIFE_OK;
		if (blocks > 0x7fffffffLL) return fail(ctx, LIFE_E_ARG, "bulk sweep: grid too large");
		k_bulk_shift<COLL, MODE><<<(unsigned)blocks, 256, 0, st>>>(a);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
		return LIFE_OK;
	}
	if (ctx->cfg.kernel == LIFE_KERNEL_QUAD && a.L.Ny % 128 == 0) {
		const int block = ctx->cfg.tune == 1 ? 128 : 256;
		a.tiles = (a.L.Ny + 4 * block - 1) / (4 * block);
		const int64_t blocks = a.tiles * c_count;
		if (blocks <= 0) return LIFE_OK;
		if (blocks > 0x7fffffffLL) return fail(ctx, LIFE_E_ARG, "bulk sweep: grid too large");
		if (block == 256) k_bulk_quad<COLL, MODE, 256, 2><<<(unsigned)blocks, 256, 0, st>>>(a);
		else k_bulk_quad<COLL, MODE, 128, 3><<<(unsigned)blocks, 128, 0, st>>>(a);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
		return LIFE_OK;
	}
	const bool staged = ctx->cfg.kernel != LIFE_KERNEL_DIRECT;   // AUTO → SHUFFLE
	// cfg.tune (measurement only, force-free shuffle kernel): tens digit = cache hint, units digit = CTA size 1:128 2:256 3:512
#ifdef LIFE_EXACT
	const int tune = 0;
#else
	const int tune = (staged && MODE == M_NONE && ctx->cfg.tune < 20) ? ctx->cfg.tune : 0;
#endif
	const int threads = (tune % 10 == 1) ? 128 : ((tune % 10 == 3) ? 512 : 256);      // 2, 4, 5: 256
	const int64_t rows_per_block = staged ? 2 * threads : threads;
	a.tiles = (a.L.Ny + rows_per_block - 1) / rows_per_block;
	const int64_t blocks = a.tiles * c_count;
	if (blocks <= 0) return LIFE_OK;
	if (blocks > 0x7fffffffLL) return fail(ctx, LIFE_E_ARG, "bulk sweep: grid too large");
	if (!staged) k_bulk_direct<COLL, MODE><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (MODE != M_NONE || tune == 0 || tune == 2) k_bulk_shuffle<COLL, MODE><<<(unsigned)blocks, threads, 0, st>>>(a);
#ifndef LIFE_EXACT
	else if (tune == 1) k_bulk_shuffle<COLL, M_NONE, 128, 0><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 3) k_bulk_shuffle<COLL, M_NONE, 512, 0><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 11) k_bulk_shuffle<COLL, M_NONE, 128, 1><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 12) k_bulk_shuffle<COLL, M_NONE, 256, 1><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 13) k_bulk_shuffle<COLL, M_NONE, 512, 1><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 4) k_bulk_shuffle<COLL, M_NONE, 256, 0, 4><<<(unsigned)blocks, threads, 0, st>>>(a);    // <= 64 registers: 4 CTAs / SM
	else if (tune == 5) k_bulk_shuffle<COLL, M_NONE, 256, 0, 5><<<(unsigned)blocks, threads, 0, st>>>(a);
	else if (tune == 14) k_bulk_shuffle<COLL, M_NONE, 256, 1, 4><<<(unsigned)blocks, threads, 0, st>>>(a);
#endif
	else return fail(ctx, LIFE_E_ARG, "bulk sweep: unknown cfg.tune");
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

template <int COLL>
static int launch_coll(life_ctx *ctx, const BulkArgs &a, int mode, int64_t c_count, cudaStream_t st) {
	switch (mode) {
	case M_NONE: return launch_one<COLL, M_NONE>(ctx, a, c_count, st);
	case M_UNI: return launch_one<COLL, M_UNI>(ctx, a, c_count, st);
	case M_IBM: return launch_one<COLL, M_IBM>(ctx, a, c_count, st);
	case M_UNI_IBM: return launch_one<COLL, M_UNI_IBM>(ctx, a, c_count, st);
	default: return launch_one<COLL, M_GENERIC>(ctx, a, c_count, st);
	}
}

// Sweep local columns [c_first, c_first + c_count) (c = i_local + 1) from ctx->fA into ctx->fB.
#ifdef LIFE_EXACT
}  // namespace exact
using namespace exact;
int launch_bulk_exact(life_ctx *ctx,
#else
int launch_bulk(life_ctx *ctx,
#endif
                const StepScalars &sc, int64_t c_first, int64_t c_count, cudaStream_t st) {
	int mode;
	BulkArgs a = make_bulk_args(ctx, sc, c_first, &mode);
	if (ctx->cfg.collision == LIFE_CENTRAL_MOMENTS) return launch_coll<COLL_CM>(ctx, a, mode, c_count, st);
	return launch_coll<COLL_BGK>(ctx, a, mode, c_count, st);
}

}  // namespace life
