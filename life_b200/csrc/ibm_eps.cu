// Implicit-IBM force multiplier epsilon on the device (SURVEY.md §8f row 1 — opt-in; the north star keeps LAPACK on the host by default).
//
//   k_eps_assemble   ObjectsClass::computeEpsilon, matrix part (src/Objects.cpp:262-301): for every body, dim x dim matrix
//                    A_ij = ( sum_s delta_i(s) * delta(x_j/Dx - site_s) ) * ds_j  over the <= 9 support sites s of marker i.
//                    One thread per (i, j); operation order and rounding of the reference (explicit __dmul_rn/__dadd_rn), so A
//                    is the reference's matrix bit for bit.
//   k_eps_solve      Utils::solveLAPACK (src/Utils.cpp:288-311): dgetrf + dgetrs('T') of the row-major A handed to Fortran, i.e.
//                    LU with partial pivoting of M = A^T followed by the solve of M^T eps = 1.  One CTA per body, matrix in
//                    global memory (L2-resident: 310^2 doubles = 0.77 MB for the largest example), unblocked right-looking
//                    elimination.  Same pivoting rule as idamax (first maximum); sums are ordered differently from a blocked
//                    LAPACK, so eps agrees to rounding x cond(A), not bit for bit.
//
// On the reference's CPU path the O(dim^2 * 9) assembly with two delta evaluations per term is serial per body; with
// UNI_EPSILON (one body holding every marker: TurekHron 132, PELskin 310) it is the largest host cost once the LBM step is on
// the GPU.
#include "ctx.h"
#include <cooperative_groups.h>
#include <cstring>

namespace life {

constexpr int ESUPP = 9;
constexpr int SOLVE_THREADS = 512;

// Utils::diracDelta (inc/Utils.h:220-232), individually rounded — same function as in ibm.cu
__device__ __forceinline__ double eps_dirac(double dist) {
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

struct EpsArgs {
	const int64_t *first;      // [nb+1] offsets into members
	const int64_t *members;    // marker indices, body after body
	const int64_t *mat_off;    // [nb] offset of each body's matrix in A
	const double *pos, *ds;
	const int32_t *scount, *sidx, *sjdx;
	const double *sdirac;
	double Dx;
	double *A;                 // concatenated dim_b x dim_b matrices, A_ij at [i*dim + j]
	double *x;                 // [total members] right-hand side / solution scratch
	int32_t *piv;              // [total members]
	double *eps;               // marker array to update
	int32_t *info;             // != 0: a zero pivot was met (LAPACK's info > 0)
	int smem_dim;              // k_eps_solve: systems up to this dimension are factorised in shared memory (0: none)
};

__global__ void __launch_bounds__(256) k_eps_assemble(const EpsArgs a) {
	const int b = blockIdx.y;
	const int64_t m0 = a.first[b], dim = a.first[b + 1] - m0;
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= dim * dim) return;
	const int64_t i = e / dim, j = e - i * dim;
	const int64_t mi = a.members[m0 + i], mj = a.members[m0 + j];
	const double pxj = __ddiv_rn(a.pos[2 * mj], a.Dx), pyj = __ddiv_rn(a.pos[2 * mj + 1], a.Dx);
	const int cnt = a.scount[mi];
	double acc = 0.0;
	for (int s = 0; s < cnt; s++) {
		const double di = a.sdirac[mi * ESUPP + s];
		const double distX = fabs(__dsub_rn(pxj, (double)a.sidx[mi * ESUPP + s]));
		const double distY = fabs(__dsub_rn(pyj, (double)a.sjdx[mi * ESUPP + s]));
		const double dj = __dmul_rn(eps_dirac(distX), eps_dirac(distY));
		acc = __dadd_rn(acc, __dmul_rn(di, dj));
	}
	a.A[a.mat_off[b] + e] = __dmul_rn(acc, a.ds[mj]);     // A[i*dim+j] *= 1.0 * 1.0 * ds_j
}

// block-wide sum, result broadcast to every thread
__device__ __forceinline__ double block_sum(double v, double *sh) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	__syncthreads();               // sh may still be read from the previous call
	if (l == 0) sh[w] = v;
	__syncthreads();
	double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
	if (w == 0) {
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (l == 0) sh[32] = t;
	}
	__syncthreads();
	return sh[32];
}

__global__ void __launch_bounds__(SOLVE_THREADS) k_eps_solve(const EpsArgs a) {
	__shared__ double sh[33];
	__shared__ double s_best[SOLVE_THREADS / 32];
	__shared__ int s_idx[SOLVE_THREADS / 32];
	__shared__ int s_p;
	const int b = blockIdx.x;
	const int64_t m0 = a.first[b];
	const int dim = (int)(a.first[b + 1] - m0);
	if (dim <= 0) return;
	// memory a.A[c*dim + r] is element (r, c) of M = A^T in column-major order — exactly what dgetrf_ is handed (src/Utils.cpp:303)
	double *M = a.A + a.mat_off[b];
	double *x = a.x + m0;
	int32_t *piv = a.piv + m0;
	const int tid = threadIdx.x, T = blockDim.x;
	// small systems (the launch provides dim^2 doubles of dynamic shared memory when every system fits): factorise a shared-memory
	// copy — every elimination step is a chain of barriers, and each used to wait for an L2 round trip
	extern __shared__ double eps_sm[];
	if (a.smem_dim >= dim) {
		for (int e = tid; e < dim * dim; e += T) eps_sm[e] = M[e];
		M = eps_sm;
		__syncthreads();
	}

	for (int k = 0; k < dim; k++) {
		// pivot: first row r >= k with the largest |M(r, k)|  (idamax)
		double best = -1.0;
		int bi = k;
		for (int r = k + tid; r < dim; r += T) {
			const double v = fabs(M[(int64_t)k * dim + r]);
			if (v > best) { best = v; bi = r; }
		}
		for (int o = 16; o > 0; o >>= 1) {
			const double ob = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
		}
		if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_idx[tid >> 5] = bi; }
		__syncthreads();
		if (tid == 0) {
			double bb = s_best[0];
			int ii = s_idx[0];
			for (int w = 1; w < (T >> 5); w++)
				if (s_best[w] > bb || (s_best[w] == bb && s_idx[w] < ii)) { bb = s_best[w]; ii = s_idx[w]; }
			s_p = ii;
			piv[k] = ii;
			if (bb == 0.0) atomicExch(a.info, k + 1);
		}
		__syncthreads();
		const int p = s_p;
		if (p != k)
			for (int c = tid; c < dim; c += T) {      // interchange rows k and p of M over all columns (dlaswp)
				const double t = M[(int64_t)c * dim + k];
				M[(int64_t)c * dim + k] = M[(int64_t)c * dim + p];
				M[(int64_t)c * dim + p] = t;
			}
		__syncthreads();
		const double pv = M[(int64_t)k * dim + k];
		if (pv != 0.0) {
			const double rp = 1.0 / pv;                // dgetf2 scales by the reciprocal
			for (int r = k + 1 + tid; r < dim; r += T) M[(int64_t)k * dim + r] *= rp;
		}
		__syncthreads();
		// trailing update M(r, c) -= M(r, k) * M(k, c), r, c > k; consecutive threads walk down a column (contiguous)
		// (warps over the columns, lanes down a column: no integer division per element)
		for (int c = k + 1 + (tid >> 5); c < dim; c += (T >> 5)) {
			const double mkc = M[(int64_t)c * dim + k];
			for (int r = k + 1 + (tid & 31); r < dim; r += 32) M[(int64_t)c * dim + r] -= M[(int64_t)k * dim + r] * mkc;
		}
		__syncthreads();
	}

	// dgetrs('T'): M^T x = b with P M = L U  ->  U^T y = b, L^T z = y, x = P^T z.  b = 1 (src/Objects.cpp:304)
	for (int r = tid; r < dim; r += T) x[r] = 1.0;
	__syncthreads();
	for (int r = 0; r < dim; r++) {                    // U^T is lower triangular: (U^T)(r, c) = M(c, r) = Mmem[r*dim + c], c <= r
		double part = 0.0;
		for (int c = tid; c < r; c += T) part += M[(int64_t)r * dim + c] * x[c];
		const double s = block_sum(part, sh);
		if (tid == 0) x[r] = (x[r] - s) / M[(int64_t)r * dim + r];
		__syncthreads();
	}
	for (int r = dim - 2; r >= 0; r--) {               // L^T is unit upper triangular: (L^T)(r, c) = M(c, r), c > r
		double part = 0.0;
		for (int c = r + 1 + tid; c < dim; c += T) part += M[(int64_t)r * dim + c] * x[c];
		const double s = block_sum(part, sh);
		if (tid == 0) x[r] -= s;
		__syncthreads();
	}
	if (tid == 0)
		for (int k = dim - 1; k >= 0; k--) {           // row interchanges in reverse order
			const int p = piv[k];
			if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; }
		}
	__syncthreads();
	for (int r = tid; r < dim; r += T) a.eps[a.members[m0 + r]] = x[r];
}


// ---- the same solve for LARGE systems (UNI_EPSILON: TurekHron 132, PELskin 310 markers) on one thread-block CLUSTER ---------------------
// The single-CTA kernel above walks an L2-resident matrix and pays a global-memory round trip per elimination step (8 ms at dim =
// 310; LAPACK on the host: 1.2 ms).  Here the matrix lives in the DISTRIBUTED SHARED MEMORY of a cluster of 8 CTAs: CTA q holds
// the columns c = q, q+8, ... of M = A^T (cyclic, so the shrinking trailing matrix stays balanced), 96 KB each at dim = 310.  Per
// elimination step the owner of column k finds the pivot (idamax rule), scales the column and writes the multipliers and the pivot
// row into a broadcast buffer of EVERY CTA with remote shared-memory stores; one cluster barrier later all CTAs swap rows k / p and
// update their own columns out of their own shared memory.  The buffers alternate, so a step costs one cluster barrier and no
// global-memory access at all.  The two triangular solves of dgetrs('T') are done by CTA 0, which reads the columns it needs from its
// peers' shared memory.  Same pivots as the kernel above and as LAPACK; sums are ordered differently from a blocked LAPACK, so epsilon
// agrees to rounding x cond(A).
constexpr int EC_CLUSTER = 8;
constexpr int EC_THREADS = 512;

__global__ void __cluster_dims__(EC_CLUSTER, 1, 1) __launch_bounds__(EC_THREADS, 1) k_eps_solve_cluster(const EpsArgs a) {
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	extern __shared__ double ec_sm[];
	__shared__ double sh[33];
	__shared__ double s_best[EC_THREADS / 32];
	__shared__ int s_idx[EC_THREADS / 32];
	__shared__ int s_piv[512];
	const int q = (int)cluster.block_rank();
	const int b = blockIdx.x / EC_CLUSTER;
	const int64_t m0 = a.first[b];
	const int dim = (int)(a.first[b + 1] - m0);
	const int tid = threadIdx.x, T = EC_THREADS;
	const int ncl_max = (dim + EC_CLUSTER - 1) / EC_CLUSTER;
	double *col = ec_sm;                                   // local column lc (global column q + 8 lc) at col + lc*dim
	double *bc0 = col + (size_t)ncl_max * dim;             // broadcast buffers: [dim] multipliers, [dim] = pivot row, [dim+1] = zero-pivot flag
	double *bc1 = bc0 + dim + 2;
	double *y = bc1 + dim + 2, *z = y + dim;
	const int ncl = dim > q ? (dim - q + EC_CLUSTER - 1) / EC_CLUSTER : 0;
	const double *M = a.A + a.mat_off[b];                  // a.A[c*dim + r] = M(r, c), M = A^T column-major (see k_eps_solve)
	for (int64_t e = tid; e < (int64_t)ncl * dim; e += T) {
		const int lc = (int)(e / dim), r = (int)(e - (int64_t)lc * dim);
		col[e] = M[(int64_t)(q + EC_CLUSTER * lc) * dim + r];
	}
	cluster.sync();

	for (int k = 0; k < dim; k++) {
		double *buf = (k & 1) ? bc1 : bc0;
		if (q == k % EC_CLUSTER) {
			double *ck = col + (size_t)(k / EC_CLUSTER) * dim;
			double best = -1.0;
			int bi = k;
			for (int r = k + tid; r < dim; r += T) {
				const double v = fabs(ck[r]);
				if (v > best) { best = v; bi = r; }
			}
			for (int o = 16; o > 0; o >>= 1) {
				const double ob = __shfl_xor_sync(0xffffffffu, best, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
			}
			if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_idx[tid >> 5] = bi; }
			__syncthreads();
			if (tid == 0) {
				double bb = s_best[0];
				int ii = s_idx[0];
				for (int w = 1; w < (T >> 5); w++)
					if (s_best[w] > bb || (s_best[w] == bb && s_idx[w] < ii)) { bb = s_best[w]; ii = s_idx[w]; }
				const double t = ck[k]; ck[k] = ck[ii]; ck[ii] = t;      // row interchange inside column k
				s_idx[0] = ii;
				s_best[0] = ck[k];
				if (bb == 0.0) atomicExch(a.info, k + 1);
			}
			__syncthreads();
			const int p = s_idx[0];
			const double pv = s_best[0];
			const double rp = pv != 0.0 ? 1.0 / pv : 1.0;              // dgetf2 scales by the reciprocal
			for (int r = k + 1 + tid; r < dim; r += T) ck[r] *= rp;
			__syncthreads();
			// multipliers and pivot row into every CTA's buffer (remote shared-memory stores)
			for (int d = 0; d < EC_CLUSTER; d++) {
				double *dst = cluster.map_shared_rank(buf, d);
				for (int r = k + 1 + tid; r < dim; r += T) dst[r] = ck[r];
				if (tid == 0) dst[dim] = (double)p;
			}
		}
		cluster.sync();
		const int p = (int)buf[dim];
		if (tid == 0) s_piv[k & 511] = p;
		// dlaswp over the local columns (column k itself was interchanged by its owner), then the rank-1 update of the columns c > k
		if (p != k)
			for (int lc = tid; lc < ncl; lc += T) {
				if (q + EC_CLUSTER * lc == k) continue;
				double *c = col + (size_t)lc * dim;
				const double t = c[k]; c[k] = c[p]; c[p] = t;
			}
		__syncthreads();
		const int lc0 = k >= q ? (k - q) / EC_CLUSTER + 1 : 0;       // first local column with global index > k
		for (int lc = lc0 + (tid >> 5); lc < ncl; lc += (T >> 5)) {      // warps over the local columns, lanes down a column
			double *c = col + (size_t)lc * dim;
			const double ck = c[k];
			for (int r = k + 1 + (tid & 31); r < dim; r += 32) c[r] -= buf[r] * ck;
		}
		__syncthreads();
	}
	cluster.sync();      // every column is final

	// dgetrs('T') by CTA 0: U^T y = 1, L^T z = y, x = P^T z; column r of M sits in CTA r % 8 at local column r / 8
	if (q == 0) {
		for (int r = 0; r < dim; r++) {
			const double *cr = cluster.map_shared_rank(col, r % EC_CLUSTER) + (size_t)(r / EC_CLUSTER) * dim;
			double part = 0.0;
			for (int c = tid; c < r; c += T) part += cr[c] * y[c];
			const double s = block_sum(part, sh);
			if (tid == 0) y[r] = (1.0 - s) / cr[r];
			__syncthreads();
		}
		for (int r = tid; r < dim; r += T) z[r] = y[r];
		__syncthreads();
		for (int r = dim - 2; r >= 0; r--) {
			const double *cr = cluster.map_shared_rank(col, r % EC_CLUSTER) + (size_t)(r / EC_CLUSTER) * dim;
			double part = 0.0;
			for (int c = r + 1 + tid; c < dim; c += T) part += cr[c] * z[c];
			const double s = block_sum(part, sh);
			if (tid == 0) z[r] -= s;
			__syncthreads();
		}
		if (tid == 0)
			for (int k = dim - 1; k >= 0; k--) {
				const int p = s_piv[k];
				if (p != k) { const double t = z[k]; z[k] = z[p]; z[p] = t; }
			}
		__syncthreads();
		for (int r = tid; r < dim; r += T) a.eps[a.members[m0 + r]] = z[r];
	}
	cluster.sync();      // nobody leaves while CTA 0 still reads its shared memory
}

// uploads the group description, assembles every group's matrix; *out describes the device buffers
static int eps_prepare(life_ctx *ctx, int64_t nb, const int64_t *first, const int64_t *members, const char *who, EpsArgs *out,
                       int64_t *a_elems_out) {
	MarkerBuffers &mk = ctx->mk;
	if (nb < 0 || (nb > 0 && (!first || !members))) return fail(ctx, LIFE_E_ARG, std::string(who) + ": null array");
	if (nb > 65535) return fail(ctx, LIFE_E_ARG, std::string(who) + ": more than 65535 bodies");
	const int64_t total = nb > 0 ? first[nb] : 0;
	std::vector<int64_t> mat_off((size_t)nb);
	int64_t a_elems = 0, max_dim = 0;
	for (int64_t b = 0; b < nb; b++) {
		const int64_t dim = first[b + 1] - first[b];
		if (dim < 0) return fail(ctx, LIFE_E_ARG, std::string(who) + ": offsets must be non-decreasing");
		mat_off[(size_t)b] = a_elems;
		a_elems += dim * dim;
		if (dim > max_dim) max_dim = dim;
	}
	for (int64_t k = 0; k < total; k++)
		if (members[k] < 0 || members[k] >= mk.n) return fail(ctx, LIFE_E_ARG, std::string(who) + ": marker index out of range");
	*a_elems_out = a_elems;
	if (total == 0) return LIFE_OK;
	// device scratch: [offsets | members | mat_off] as int64, then x, A, piv, info
	const size_t n_i64 = (size_t)(nb + 1 + total + nb);
	const size_t bytes = sizeof(int64_t) * n_i64 + sizeof(double) * (size_t)(total + a_elems) + sizeof(int32_t) * (size_t)(total + 4) + 64;
	if (bytes > ctx->eps_bytes) {
		cudaFree(ctx->eps_buf);
		ctx->eps_buf = nullptr;
		ctx->eps_bytes = 0;
		LIFE_CUDA(ctx, cudaMalloc(&ctx->eps_buf, bytes));
		ctx->eps_bytes = bytes;
	}
	std::vector<int64_t> h(n_i64);
	memcpy(h.data(), first, sizeof(int64_t) * (size_t)(nb + 1));
	memcpy(h.data() + nb + 1, members, sizeof(int64_t) * (size_t)total);
	memcpy(h.data() + nb + 1 + total, mat_off.data(), sizeof(int64_t) * (size_t)nb);
	char *base = reinterpret_cast<char *>(ctx->eps_buf);
	LIFE_CUDA(ctx, cudaMemcpyAsync(base, h.data(), sizeof(int64_t) * n_i64, cudaMemcpyHostToDevice, ctx->stream));
	EpsArgs a{};
	a.first = reinterpret_cast<int64_t *>(base);
	a.members = a.first + nb + 1;
	a.mat_off = a.members + total;
	a.x = reinterpret_cast<double *>(base + sizeof(int64_t) * n_i64);
	a.A = a.x + total;
	a.piv = reinterpret_cast<int32_t *>(a.A + a_elems);
	a.info = a.piv + total;
	a.pos = mk.pos; a.ds = mk.ds;
	a.scount = mk.scount; a.sidx = mk.sidx; a.sjdx = mk.sjdx; a.sdirac = mk.sdirac;
	a.Dx = ctx->cfg.Dx;
	a.eps = mk.eps;
	LIFE_CUDA(ctx, cudaMemsetAsync(a.info, 0, sizeof(int32_t), ctx->stream));
	const dim3 grid((unsigned)((max_dim * max_dim + 255) / 256), (unsigned)nb);
	k_eps_assemble<<<grid, 256, 0, ctx->stream>>>(a);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `h` must outlive the copy
	*out = a;
	return LIFE_OK;
}

int ibm_assemble_epsilon(life_ctx *ctx, int64_t nb, const int64_t *first, const int64_t *members, double *A_out) {
	if (!A_out) return fail(ctx, LIFE_E_ARG, "life_ibm_assemble_epsilon: null output");
	EpsArgs a{};
	int64_t a_elems = 0;
	int rc = eps_prepare(ctx, nb, first, members, "life_ibm_assemble_epsilon", &a, &a_elems);
	if (rc || a_elems == 0) return rc;
	LIFE_CUDA(ctx, cudaMemcpyAsync(A_out, a.A, sizeof(double) * (size_t)a_elems, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int ibm_compute_epsilon(life_ctx *ctx, int64_t nb, const int64_t *first, const int64_t *members, double *eps_out) {
	MarkerBuffers &mk = ctx->mk;
	EpsArgs a{};
	int64_t a_elems = 0;
	int rc = eps_prepare(ctx, nb, first, members, "life_ibm_compute_epsilon", &a, &a_elems);
	if (rc) return rc;
	if (a_elems > 0) {
		// systems of 65 .. 512 markers: the cluster kernel (matrix in distributed shared memory); smaller ones — usually many of them,
		// Honami: 128 x 31 — one CTA each
		int64_t max_dim = 0;
		for (int64_t b = 0; b < nb; b++) max_dim = first[b + 1] - first[b] > max_dim ? first[b + 1] - first[b] : max_dim;
		const size_t smem = sizeof(double) * (((size_t)((max_dim + EC_CLUSTER - 1) / EC_CLUSTER)) * (size_t)max_dim + 2 * (size_t)(max_dim + 2) + 2 * (size_t)max_dim);
		if (max_dim > 64 && max_dim <= 512 && smem <= 200 * 1024 && ctx->cfg.tune != 40) {
			LIFE_CUDA(ctx, cudaFuncSetAttribute(k_eps_solve_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			k_eps_solve_cluster<<<(unsigned)(nb * EC_CLUSTER), EC_THREADS, smem, ctx->stream>>>(a);
		} else {
			a.smem_dim = max_dim <= 72 ? (int)max_dim : 0;      // 72^2 doubles = 41 KB: below the default dynamic shared-memory limit
			k_eps_solve<<<(unsigned)nb, SOLVE_THREADS, sizeof(double) * (size_t)a.smem_dim * a.smem_dim, ctx->stream>>>(a);
		}
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
	}
	if (eps_out && mk.n > 0) {
		double *he = mk.h_stage + 6 * mk.h_cap;    // result half of the pinned staging buffer
		LIFE_CUDA(ctx, cudaMemcpyAsync(he, mk.eps, sizeof(double) * mk.n, cudaMemcpyDeviceToHost, ctx->stream));
		LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		memcpy(eps_out, he, sizeof(double) * mk.n);
	} else {
		LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	return LIFE_OK;
}

}  // namespace life
