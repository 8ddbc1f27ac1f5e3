// Immersed-boundary marker kernels: support search, interpolation + direct-forcing force, force spreading.
//
//   k_find_support  IBMNodeClass::findSupport (src/IBMNode.cpp:139-179) with Utils::diracDelta (inc/Utils.h:220-232)
//   k_interp        IBMNodeClass::interpolate (:26-48) + forceCalc (:51-58), driven as ObjectsClass::ibmKernelInterp
//                   (src/Objects.cpp:102-117)
//   k_spread_*      IBMNodeClass::spread (:61-94), driven as ObjectsClass::ibmKernelSpread (src/Objects.cpp:120-149)
//   updateMacroscopic (:97-136) has no kernel: rho and u are never stored, every consumer evaluates them from f and the
//   current forces, which is exactly what that routine writes at the support sites.
//
// One warp per marker.  A marker has at most 9 support sites (3-point delta, suppSize = 9, inc/defs.h:39): lanes 0..8 each own
// one site for the gather and the scatter; the 25 candidate sites of the search map to lanes 0..24 and are compacted with a
// ballot, which preserves the reference's i-outer / j-inner order.
//
// Arithmetic that feeds values the host compares bit for bit (delta weights) or sums in a fixed order (interpolation, ordered
// spread) uses explicit __dmul_rn/__dadd_rn so the compiler cannot contract it into FMAs: given the same inputs these kernels
// produce the same doubles as the reference's g++ build.
//
// Spread variants:
//   atomic  (cfg.ordered == 0, the reference's `omp atomic` path): 2 atomicAdd(double) per support site.
//   ordered (cfg.ordered == 1, the reference's `omp ordered` path): bit-repeatable, no floating-point atomics.  Every (marker,
//           site) entry links itself into the list of its lattice site with one integer atomicExch; the entry left at the head
//           owns the site, orders the site's few contributions by marker index and writes their sum — the same value, in the
//           same order, as the reference's marker-ordered loop.
#include "ctx.h"
#include "d2q9.cuh"

namespace life {

constexpr int SUPP = 9;
constexpr int MAX_CONTRIB = 32;

// ---- buffers ---------------------------------------------------------------------------------------------------------------------
static int ensure_markers(life_ctx *ctx, int64_t n) {
	MarkerBuffers &m = ctx->mk;
	if (n <= m.cap) return LIFE_OK;
	int64_t cap = m.cap ? m.cap : 256;
	while (cap < n) cap *= 2;
	// contents need not survive a growth: the caller is about to overwrite everything
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(m.in); cudaFree(m.force); cudaFree(m.irho); cudaFree(m.imom);
	cudaFree(m.scount); cudaFree(m.sidx); cudaFree(m.sjdx); cudaFree(m.sdirac); cudaFree(m.next);
	if (m.h_stage) cudaFreeHost(m.h_stage);
	int32_t *keep_err = m.err;
	cudaEvent_t keep_ev = m.ev_stage;
	m = MarkerBuffers{};
	m.err = keep_err;
	m.ev_stage = keep_ev;
	// pos | vel | ds | eps live in ONE device array laid out like the pinned staging buffer, so a marker update is one copy
	LIFE_CUDA(ctx, cudaMalloc(&m.in, sizeof(double) * 6 * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.force, sizeof(double) * 2 * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.irho, sizeof(double) * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.imom, sizeof(double) * 2 * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.scount, sizeof(int32_t) * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.sidx, sizeof(int32_t) * SUPP * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.sjdx, sizeof(int32_t) * SUPP * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.sdirac, sizeof(double) * SUPP * cap));
	LIFE_CUDA(ctx, cudaMalloc(&m.next, sizeof(int32_t) * SUPP * cap));
	LIFE_CUDA(ctx, cudaMemsetAsync(m.force, 0, sizeof(double) * 2 * cap, ctx->stream));
	LIFE_CUDA(ctx, cudaMemsetAsync(m.scount, 0, sizeof(int32_t) * cap, ctx->stream));
	LIFE_CUDA(ctx, cudaMallocHost(&m.h_stage, sizeof(double) * 8 * cap));   // 6*cap upload staging + 2*cap result staging
	m.h_cap = cap;
	m.cap = cap;
	if (!ctx->mk.err) {
		LIFE_CUDA(ctx, cudaMalloc(&ctx->mk.err, sizeof(int32_t)));
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->mk.err, 0, sizeof(int32_t), ctx->stream));
	}
	if (!ctx->mk.ev_stage) LIFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->mk.ev_stage, cudaEventDisableTiming));
	return LIFE_OK;
}

void ibm_free(life_ctx *ctx) {
	MarkerBuffers &m = ctx->mk;
	cudaFree(m.in); cudaFree(m.force); cudaFree(m.irho); cudaFree(m.imom);
	cudaFree(m.scount); cudaFree(m.sidx); cudaFree(m.sjdx); cudaFree(m.sdirac); cudaFree(m.next); cudaFree(m.err);
	if (m.h_stage) cudaFreeHost(m.h_stage);
	if (m.ev_stage) cudaEventDestroy(m.ev_stage);
	m = MarkerBuffers{};
}

// ---- support search ------------------------------------------------------------------------------------------------------------
// Utils::diracDelta, inc/Utils.h:220-232, with every operation individually rounded (no FMA contraction)
__device__ __forceinline__ double dirac_delta(double dist) {
	const double a = fabs(dist);
	if (a > 1.5) return 0.0;
	else if (a > 0.5) {
		const double q = __dsub_rn(1.0, a);
		const double rad = __dadd_rn(__dmul_rn(-3.0, __dmul_rn(q, q)), 1.0);
		return __ddiv_rn(__dsub_rn(__dsub_rn(5.0, __dmul_rn(3.0, a)), sqrt(rad)), 6.0);
	} else {
		const double rad = __dsub_rn(1.0, __dmul_rn(3.0, __dmul_rn(a, a)));
		return __ddiv_rn(__dadd_rn(1.0, sqrt(rad)), 3.0);
	}
}

__global__ void __launch_bounds__(128) k_find_support(int64_t n, const double *__restrict__ pos, double Dx, int64_t Nx,
                                                      int64_t Ny, int32_t *scount, int32_t *sidx, int32_t *sjdx,
                                                      double *sdirac, int32_t *err) {
	const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (m >= n) return;
	const double px = __ddiv_rn(pos[2 * m], Dx), py = __ddiv_rn(pos[2 * m + 1], Dx);
	const int inear = (int)round(px), jnear = (int)round(py);
	// candidate (i, j) of this lane: i outer, j inner, as the reference's double loop (src/IBMNode.cpp:156-178)
	const int i = inear - 2 + lane / 5, j = jnear - 2 + lane % 5;
	const double distX = fabs(__dsub_rn(px, (double)i)), distY = fabs(__dsub_rn(py, (double)j));
	const bool ok = lane < 25 && distX < 1.5 && distY < 1.5 && i >= 0 && i <= Nx - 1 && j >= 0 && j <= Ny - 1;
	const unsigned mask = __ballot_sync(0xffffffffu, ok);
	const int slot = __popc(mask & ((1u << lane) - 1u));
	const int cnt = __popc(mask);
	if (ok && slot < SUPP) {
		sidx[m * SUPP + slot] = i;
		sjdx[m * SUPP + slot] = j;
		sdirac[m * SUPP + slot] = __dmul_rn(dirac_delta(distX), dirac_delta(distY));
	}
	if (lane < SUPP && lane >= cnt) {   // the reference clears unused entries (src/IBMNode.cpp:142)
		sidx[m * SUPP + lane] = 0;
		sjdx[m * SUPP + lane] = 0;
		sdirac[m * SUPP + lane] = 0.0;
	}
	if (lane == 0) {
		scount[m] = cnt > SUPP ? SUPP : cnt;
		if (cnt > SUPP) atomicOr(err, 1);   // "Support buffer size is not big enough" (src/IBMNode.cpp:171-172)
	}
}

// ---- clearing force_ibm at the sites the last spread touched -------------------------------------------------------------------
__global__ void k_zero_sites(int64_t n, const int32_t *__restrict__ scount, const int32_t *__restrict__ sidx,
                             const int32_t *__restrict__ sjdx, Layout L, int64_t i_begin, double *fibm, uint8_t *mask, int64_t mask_pitch) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n * SUPP) return;
	const int64_t m = e / SUPP;
	const int s = (int)(e - m * SUPP);
	if (s >= scount[m]) return;
	const int64_t il = sidx[e] - i_begin;
	if (il < 0 || il >= L.nxl) return;
	const int64_t idx = L.node(il, sjdx[e]);
	fibm[idx] = 0.0;
	fibm[L.S + idx] = 0.0;
	mask[(il + 1) * mask_pitch + (sjdx[e] >> 6)] = 0;     // every site of the span is being cleared by this launch
}

int ibm_clear_force(life_ctx *ctx) {
	if (!ctx->fibm) return LIFE_OK;
	if (ctx->fibm_full_dirty) {
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm, 0, sizeof(double) * 2 * ctx->L.S, ctx->stream));
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm_mask, 0, (size_t)(ctx->mask_pitch * (ctx->L.nxl + 2)), ctx->stream));
	} else if (ctx->fibm_sites_dirty && ctx->mk.n > 0) {
		const int64_t n = ctx->mk.n * SUPP;
		k_zero_sites<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->mk.n, ctx->mk.scount, ctx->mk.sidx, ctx->mk.sjdx,
		                                                                  ctx->L, ctx->i_begin, ctx->fibm, ctx->fibm_mask, ctx->mask_pitch);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
	}
	ctx->fibm_full_dirty = false;
	ctx->fibm_sites_dirty = false;
	return LIFE_OK;
}

// the markers are about to move: what happens to the force_ibm spread with the old supports
static int ibm_before_move(life_ctx *ctx) {
	int rc;
	// force_ibm is non-zero at the OLD supports: clear it before they are replaced — unless no step has used it yet (markers
	// set between a spread, an upload or life_read_restart and the next life_step: the reference keeps force_ibm until
	// ibmKernelInterp zeroes it, src/Objects.cpp:105); then it stays and is cleared wholesale later
	if (ctx->fibm_consumed) {
		if ((rc = ibm_clear_force(ctx))) return rc;
	} else if (ctx->fibm_sites_dirty) {
		ctx->fibm_full_dirty = true;
	}
	return LIFE_OK;
}

static int launch_find_support(life_ctx *ctx) {
	MarkerBuffers &m = ctx->mk;
	const int64_t threads = m.n * 32;
	k_find_support<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(m.n, m.pos, ctx->cfg.Dx, ctx->cfg.Nx, ctx->cfg.Ny,
	                                                                           m.scount, m.sidx, m.sjdx, m.sdirac, m.err);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

// supports of the positions the DEVICE holds (the structural solver moved the markers there, fem.cu): findSupport of every marker
// without the positions crossing PCIe
int ibm_refresh_supports(life_ctx *ctx) {
	int rc;
	if (ctx->mk.n == 0) return LIFE_OK;
	if ((rc = ibm_before_move(ctx))) return rc;
	return launch_find_support(ctx);
}

int ibm_set_markers(life_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *ds, const double *eps) {
	int rc;
	if ((rc = ibm_before_move(ctx))) return rc;
	if ((rc = ensure_markers(ctx, n))) return rc;
	MarkerBuffers &m = ctx->mk;
	m.n = n;
	m.pos = m.in; m.vel = m.in + 2 * n; m.ds = m.in + 4 * n; m.eps = m.in + 5 * n;
	if (n == 0) return LIFE_OK;
	// the staging buffer may still be feeding the previous update's copy
	if (m.stage_busy) LIFE_CUDA(ctx, cudaEventSynchronize(m.ev_stage));
	double *h = m.h_stage;
	memcpy(h, pos, sizeof(double) * 2 * n);
	memcpy(h + 2 * n, vel, sizeof(double) * 2 * n);
	memcpy(h + 4 * n, ds, sizeof(double) * n);
	memcpy(h + 5 * n, eps, sizeof(double) * n);
	LIFE_CUDA(ctx, cudaMemcpyAsync(m.in, h, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, ctx->stream));
	LIFE_CUDA(ctx, cudaEventRecord(m.ev_stage, ctx->stream));
	m.stage_busy = true;
	// asynchronous: a support overflow (src/IBMNode.cpp:171-172) is reported by the next synchronising marker call
	return launch_find_support(ctx);
}

// reads back the device error flag (after a synchronisation of ctx->stream has been enqueued by the caller)
static int check_marker_errors(life_ctx *ctx, const int32_t *host_flag) {
	if (*host_flag & 1) {
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->mk.err, 0, sizeof(int32_t), ctx->stream));
		return fail(ctx, LIFE_E_SUPPORT, "Support buffer size is not big enough for number of support points");
	}
	if (*host_flag & 2) {
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->mk.err, 0, sizeof(int32_t), ctx->stream));
		return fail(ctx, LIFE_E_SUPPORT, "ordered spread: more than 32 markers contribute to one lattice site");
	}
	return LIFE_OK;
}

int ibm_check(life_ctx *ctx) {
	if (!ctx->mk.err) return LIFE_OK;
	int32_t *herr = reinterpret_cast<int32_t *>(ctx->h_pin);
	LIFE_CUDA(ctx, cudaMemcpyAsync(herr, ctx->mk.err, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return check_marker_errors(ctx, herr);
}

// ---- interpolation + force ------------------------------------------------------------------------------------------------------
struct InterpArgs {
	int64_t n;
	const double *f;
	PopShift ps;
	Layout L;
	int64_t i_begin;
	int fxy_mode;
	double fx, fy;
	const double *fxyf;
	const int32_t *scount, *sidx, *sjdx;
	const double *sdirac, *vel;
	double velScale;
	double *irho, *imom, *force;
	int partial;     // nranks > 1: write this rank's partial sums only, the force is finished after the all-reduce
};

__global__ void __launch_bounds__(128) k_interp(const InterpArgs a) {
	const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (m >= a.n) return;
	const int cnt = a.scount[m];
	double pr = 0.0, px = 0.0, py = 0.0;
	if (lane < cnt) {
		const int64_t il = a.sidx[m * SUPP + lane] - a.i_begin;
		if (il >= 0 && il < a.L.nxl) {
			const int64_t idx = a.L.node(il, a.sjdx[m * SUPP + lane]);
			// rho, u as GridClass::macroscopic left them (src/Grid.cpp:282-299): no IBM force at this point of the step
			double p[NV], rho, mx, my;
#pragma unroll
			for (int v = 0; v < NV; v++) p[v] = a.f[a.ps.at(v, idx, a.L.S)];
			moments(p, rho, mx, my);
			double fx = a.fx, fy = a.fy;
			if (a.fxy_mode == FXY_FIELD) { fx = a.fxyf[idx]; fy = a.fxyf[a.L.S + idx]; }
			const double ux = __ddiv_rn(__dadd_rn(mx, __dmul_rn(0.5, fx)), rho);
			const double uy = __ddiv_rn(__dadd_rn(my, __dmul_rn(0.5, fy)), rho);
			const double d = a.sdirac[m * SUPP + lane];
			pr = __dmul_rn(rho, d);                       // rho * diracVal            (src/IBMNode.cpp:42)
			px = __dmul_rn(__dmul_rn(rho, ux), d);        // rho * u * diracVal        (src/IBMNode.cpp:46)
			py = __dmul_rn(__dmul_rn(rho, uy), d);
		}
	}
	// sum over the support in the reference's order s = 0, 1, ... (lane 0 collects)
	double sr = 0.0, sx = 0.0, sy = 0.0;
#pragma unroll
	for (int s = 0; s < SUPP; s++) {
		const double r = __shfl_sync(0xffffffffu, pr, s), x = __shfl_sync(0xffffffffu, px, s), y = __shfl_sync(0xffffffffu, py, s);
		if (s < cnt) { sr = __dadd_rn(sr, r); sx = __dadd_rn(sx, x); sy = __dadd_rn(sy, y); }
	}
	if (lane == 0) {
		a.irho[m] = sr;
		a.imom[2 * m] = sx;
		a.imom[2 * m + 1] = sy;
		if (!a.partial) {
			// forceCalc (src/IBMNode.cpp:51-58): force = 2 (velScale * interpRho * vel - interpMom)
			const double sc = __dmul_rn(a.velScale, sr);
			a.force[2 * m] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(sc, a.vel[2 * m]), sx));
			a.force[2 * m + 1] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(sc, a.vel[2 * m + 1]), sy));
		}
	}
}

__global__ void k_force_calc(int64_t n, const double *irho, const double *imom, const double *vel, double velScale, double *force) {
	const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	const double sc = __dmul_rn(velScale, irho[m]);
	force[2 * m] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(sc, vel[2 * m]), imom[2 * m]));
	force[2 * m + 1] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(sc, vel[2 * m + 1]), imom[2 * m + 1]));
}

// `no_sync`: leave everything enqueued (the structural solver consumes the forces on the device; errors surface at its own sync)
int ibm_interp(life_ctx *ctx, double *force_out, bool no_sync) {
	int rc;
	if ((rc = ibm_clear_force(ctx))) return rc;     // fill(force_ibm, 0), src/Objects.cpp:105
	MarkerBuffers &m = ctx->mk;
	if (m.n == 0) return LIFE_OK;
	InterpArgs a{};
	a.n = m.n;
	a.f = ctx->fA;
	a.ps = ctx->shift;
	a.L = ctx->L;
	a.i_begin = ctx->i_begin;
	a.fxy_mode = ctx->fxy_mode;
	a.fx = ctx->fxy_uniform[0]; a.fy = ctx->fxy_uniform[1];
	a.fxyf = ctx->fxyf;
	a.scount = m.scount; a.sidx = m.sidx; a.sjdx = m.sjdx; a.sdirac = m.sdirac; a.vel = m.vel;
	a.velScale = ctx->cfg.Dt / ctx->cfg.Dx;
	a.irho = m.irho; a.imom = m.imom; a.force = m.force;
	a.partial = ctx->comm != nullptr;
	const int64_t threads = m.n * 32;
	k_interp<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(a);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	if (ctx->comm) {
		// a marker whose support straddles a slab face is gathered partly on each side: add the partial sums
		LIFE_NCCL(ctx, ncclGroupStart());
		LIFE_NCCL(ctx, ncclAllReduce(m.irho, m.irho, (size_t)m.n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
		LIFE_NCCL(ctx, ncclAllReduce(m.imom, m.imom, (size_t)(2 * m.n), ncclDouble, ncclSum, ctx->comm, ctx->stream));
		LIFE_NCCL(ctx, ncclGroupEnd());
		k_force_calc<<<(unsigned)((m.n + 127) / 128), 128, 0, ctx->stream>>>(m.n, m.irho, m.imom, m.vel, a.velScale, m.force);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
	}
	if (no_sync) return LIFE_OK;
	int32_t *herr = reinterpret_cast<int32_t *>(ctx->h_pin);
	LIFE_CUDA(ctx, cudaMemcpyAsync(herr, m.err, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	if (force_out) {
		// results come back through the second half of the pinned staging buffer (the first half may feed an upload)
		double *hf = m.h_stage + 6 * m.h_cap;
		LIFE_CUDA(ctx, cudaMemcpyAsync(hf, m.force, sizeof(double) * 2 * m.n, cudaMemcpyDeviceToHost, ctx->stream));
		LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		memcpy(force_out, hf, sizeof(double) * 2 * m.n);
	} else {
		LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	m.stage_busy = false;
	return check_marker_errors(ctx, herr);
}

// ---- spread ------------------------------------------------------------------------------------------------------------------------
struct SpreadArgs {
	int64_t n;
	const int32_t *scount, *sidx, *sjdx;
	const double *sdirac, *force, *eps, *ds, *pos;
	Layout L;
	int64_t i_begin;
	double Dx;
	double *fibm;
	uint8_t *mask;        // per (column, 64-row span) flag: force_ibm written there (ctx.h)
	int64_t mask_pitch;
	int32_t *head, *next, *err;
};

// force * epsilon * ds * 1.0 * diracVal, left to right (src/IBMNode.cpp:73-74)
__device__ __forceinline__ double spread_term(double force, double eps, double ds, double dirac) {
	return __dmul_rn(__dmul_rn(__dmul_rn(force, eps), ds), dirac);
}

__global__ void __launch_bounds__(128) k_spread_atomic(const SpreadArgs a) {
	const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (m >= a.n || lane >= a.scount[m]) return;
	const int64_t il = a.sidx[m * SUPP + lane] - a.i_begin;
	if (il < 0 || il >= a.L.nxl) return;
	const int64_t idx = a.L.node(il, a.sjdx[m * SUPP + lane]);
	const double d = a.sdirac[m * SUPP + lane];
	a.mask[(il + 1) * a.mask_pitch + (a.sjdx[m * SUPP + lane] >> 6)] = 1;
	atomicAdd(a.fibm + idx, spread_term(a.force[2 * m], a.eps[m], a.ds[m], d));
	atomicAdd(a.fibm + a.L.S + idx, spread_term(a.force[2 * m + 1], a.eps[m], a.ds[m], d));
}

// Ordered spread, two launches.
//   k_site_lists : every (marker, support site) entry e = m*9 + s pushes itself onto the list of its lattice site with one
//                  atomicExch on the site's head (an int32 per node, -1 = empty): link[e] = previous head.
//   k_spread_ordered : the entry that ended up as the head of its site's list owns the site: it walks the list (its length is the
//                  number of markers touching the site, typically 3-6), orders the contributions by marker index — the order of the
//                  reference's `omp ordered` loop (src/Objects.cpp:130-139; one contribution per marker and site) — adds them up from
//                  0.0 exactly as the reference's `+=` does, writes force_ibm and the span mask, and resets the head to -1.
// No floating-point atomics, no sort over all entries, no search: the only dependent chain is the list walk.
__device__ __forceinline__ void site_list_push(const SpreadArgs &a, const int64_t e) {
	const int64_t m = e / SUPP;
	if ((int)(e - m * SUPP) >= a.scount[m]) return;
	const int64_t il = (int64_t)a.sidx[e] - a.i_begin;
	if (il < 0 || il >= a.L.nxl) return;
	a.next[e] = atomicExch(a.head + a.L.node(il, a.sjdx[e]), (int32_t)e);
}

__device__ __forceinline__ void site_owner_sum(const SpreadArgs &a, const int64_t e) {
	const int64_t m = e / SUPP;
	if ((int)(e - m * SUPP) >= a.scount[m]) return;
	const int j = a.sjdx[e];
	const int64_t il = (int64_t)a.sidx[e] - a.i_begin;
	if (il < 0 || il >= a.L.nxl) return;
	const int64_t idx = a.L.node(il, j);
	if (a.head[idx] != (int32_t)e) return;          // another entry owns this site
	int32_t who[MAX_CONTRIB];
	int32_t ent[MAX_CONTRIB];
	int cnt = 0;
	for (int32_t x = (int32_t)e; x >= 0; x = a.next[x]) {
		if (cnt == MAX_CONTRIB) { atomicOr(a.err, 2); break; }
		// insertion by marker index (lists are short and arrive nearly sorted: entries were pushed in roughly ascending order)
		const int32_t mk = x / SUPP;
		int y = cnt - 1;
		while (y >= 0 && who[y] > mk) { who[y + 1] = who[y]; ent[y + 1] = ent[y]; y--; }
		who[y + 1] = mk; ent[y + 1] = x;
		cnt++;
	}
	double sx = 0.0, sy = 0.0;
	for (int x = 0; x < cnt; x++) {
		const int64_t m2 = who[x];
		const double d = a.sdirac[ent[x]];
		sx = __dadd_rn(sx, spread_term(a.force[2 * m2], a.eps[m2], a.ds[m2], d));
		sy = __dadd_rn(sy, spread_term(a.force[2 * m2 + 1], a.eps[m2], a.ds[m2], d));
	}
	a.fibm[idx] = sx;
	a.fibm[a.L.S + idx] = sy;
	a.mask[(il + 1) * a.mask_pitch + (j >> 6)] = 1;
	a.head[idx] = -1;                               // empty again for the next spread
}

__global__ void __launch_bounds__(256) k_site_lists(const SpreadArgs a) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.n * SUPP) site_list_push(a, e);
}

__global__ void __launch_bounds__(256) k_spread_ordered(const SpreadArgs a) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.n * SUPP) site_owner_sum(a, e);
}

int ibm_spread(life_ctx *ctx) {
	int rc;
	if ((rc = ensure_fibm(ctx))) return rc;
	if ((rc = ibm_clear_force(ctx))) return rc;     // fill(force_ibm, 0), src/Objects.cpp:123
	MarkerBuffers &m = ctx->mk;
	if (m.n == 0) return LIFE_OK;
	SpreadArgs a{};
	a.n = m.n;
	a.scount = m.scount; a.sidx = m.sidx; a.sjdx = m.sjdx; a.sdirac = m.sdirac;
	a.force = m.force; a.eps = m.eps; a.ds = m.ds; a.pos = m.pos;
	a.L = ctx->L;
	a.i_begin = ctx->i_begin;
	a.Dx = ctx->cfg.Dx;
	a.fibm = ctx->fibm;
	a.mask = ctx->fibm_mask; a.mask_pitch = ctx->mask_pitch;
	a.next = m.next; a.err = m.err;
	if (ctx->cfg.ordered) {
		if (!ctx->cell_head) {
			LIFE_CUDA(ctx, cudaMalloc(&ctx->cell_head, sizeof(int32_t) * ctx->L.S));
			LIFE_CUDA(ctx, cudaMemsetAsync(ctx->cell_head, 0xff, sizeof(int32_t) * ctx->L.S, ctx->stream));
		}
		a.head = ctx->cell_head;
		// (both phases in one launch of one 1024-thread CTA was tried for example-scale marker counts: 24.9 us against 11.4 us for
		// these two whole-GPU launches, profiles/r02_ncu_summary.md)
		const unsigned eb = (unsigned)((m.n * SUPP + 255) / 256);
		k_site_lists<<<eb, 256, 0, ctx->stream>>>(a);
		k_spread_ordered<<<eb, 256, 0, ctx->stream>>>(a);
		ctx->launches += 2;
	} else {
		const int64_t threads = m.n * 32;
		k_spread_atomic<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(a);
		ctx->launches++;
	}
	LIFE_CUDA(ctx, cudaGetLastError());
	ctx->fibm_any = true;
	ctx->fibm_sites_dirty = true;
	ctx->fibm_consumed = false;
	return LIFE_OK;
}

}  // namespace life
