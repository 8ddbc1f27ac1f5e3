// Host <-> device movement in the reference's layout, and the on-demand macroscopic / max-speed kernels.
//   upload_field / download_field : AoS host arrays (f[id*9+v], u[id*2+d], rho[id]; id = i*Ny + j, src/Grid.cpp:70)
//                                   <-> padded SoA planes (ctx.h), staged through a device scratch buffer in column chunks
//   launch_macro                  : rho, u as GridClass holds them at the end of a step (src/Grid.cpp:282-299 +
//                                   src/IBMNode.cpp:97-136), evaluated from f and the current forces
//   launch_max_speed              : the scan of GridClass::writeInfo (src/Grid.cpp:562-588)
#include "ctx.h"
#include "d2q9.cuh"
#include "macro.cuh"

namespace life {

int ensure_scratch(life_ctx *ctx, size_t bytes) {
	if (ctx->scratch_bytes >= bytes) return LIFE_OK;
	if (ctx->scratch) cudaFree(ctx->scratch);
	ctx->scratch = nullptr;
	ctx->scratch_bytes = 0;
	LIFE_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
	ctx->scratch_bytes = bytes;
	return LIFE_OK;
}

int ensure_macro(life_ctx *ctx) {
	if (ctx->macro) return LIFE_OK;
	LIFE_CUDA(ctx, cudaMalloc(&ctx->macro, sizeof(double) * 3 * ctx->L.S));
	LIFE_CUDA(ctx, cudaMemsetAsync(ctx->macro, 0, sizeof(double) * 3 * ctx->L.S, ctx->stream));
	return LIFE_OK;
}

int ensure_fibm(life_ctx *ctx) {
	if (ctx->fibm) return LIFE_OK;
	LIFE_CUDA(ctx, cudaMalloc(&ctx->fibm, sizeof(double) * 2 * ctx->L.S));
	LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm, 0, sizeof(double) * 2 * ctx->L.S, ctx->stream));
	ctx->mask_pitch = (ctx->L.P + 63) / 64;
	LIFE_CUDA(ctx, cudaMalloc(&ctx->fibm_mask, (size_t)(ctx->mask_pitch * (ctx->L.nxl + 2))));
	LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm_mask, 0, (size_t)(ctx->mask_pitch * (ctx->L.nxl + 2)), ctx->stream));
	return LIFE_OK;
}

// AoS chunk (columns il0 .. il0+ncols-1) -> planes.  One thread per (node, component); consecutive threads read
// consecutive doubles of the AoS chunk.
__global__ void k_unpack(const double *__restrict__ src, double *__restrict__ planes, Layout L, int ncomp, int64_t il0,
                         int64_t n_elems) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_elems) return;
	const int64_t node = e / ncomp;
	const int k = (int)(e - node * ncomp);
	const int64_t il = il0 + node / L.Ny, j = node % L.Ny;
	planes[k * L.S + L.node(il, j)] = src[e];
}

__global__ void k_pack(double *__restrict__ dst, const double *__restrict__ planes, Layout L, int ncomp, int64_t il0,
                       int64_t n_elems, PopShift ps) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_elems) return;
	const int64_t node = e / ncomp;
	const int k = (int)(e - node * ncomp);
	const int64_t il = il0 + node / L.Ny, j = node % L.Ny;
	dst[e] = planes[ps.at(k, L.node(il, j), L.S)];     // ps: zeros for everything but in-place populations
}

static int64_t chunk_columns(life_ctx *ctx, int ncomp) {
	const int64_t per_col = ctx->L.Ny * ncomp;
	const int64_t target = (int64_t)(128ll << 20) / (int64_t)sizeof(double);   // 128 MiB per staging half
	int64_t cols = target / per_col;
	if (cols < 1) cols = 1;
	if (cols > ctx->L.nxl) cols = ctx->L.nxl;
	return cols;
}

static int ensure_copy_stream(life_ctx *ctx) {
	if (ctx->copy_stream) return LIFE_OK;
	LIFE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	for (int k = 0; k < 2; k++) {
		LIFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[k], cudaEventDisableTiming));
		LIFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_kernel[k], cudaEventDisableTiming));
	}
	return LIFE_OK;
}

// host chunk `h` = columns [il0, il0 + ncols) of a reference-layout array with `ncomp` doubles per node.
// Double-buffered: chunk k travels over PCIe (copy stream) while chunk k-1 is unpacked into the planes (compute stream).
int upload_field(life_ctx *ctx, const double *h, double *planes, int ncomp, int64_t il0, int64_t ncols) {
	const Layout &L = ctx->L;
	const int64_t cols = chunk_columns(ctx, ncomp);
	const size_t half = sizeof(double) * (size_t)(cols * L.Ny * ncomp);
	int rc = ensure_scratch(ctx, 2 * half);
	if (rc) return rc;
	if ((rc = ensure_copy_stream(ctx))) return rc;
	// the staging halves may still be read by kernels enqueued earlier on the compute stream
	LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_kernel[0], ctx->stream));
	LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_kernel[1], ctx->stream));
	int k = 0;
	for (int64_t c0 = 0; c0 < ncols; c0 += cols, k ^= 1) {
		const int64_t nc = (c0 + cols <= ncols) ? cols : ncols - c0;
		const int64_t n_elems = nc * L.Ny * ncomp;
		double *stage = reinterpret_cast<double *>(reinterpret_cast<char *>(ctx->scratch) + (k ? half : 0));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_kernel[k], 0));      // the kernel that last read this half is done
		LIFE_CUDA(ctx, cudaMemcpyAsync(stage, h + c0 * L.Ny * ncomp, sizeof(double) * n_elems, cudaMemcpyHostToDevice, ctx->copy_stream));
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_copy[k], ctx->copy_stream));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[k], 0));
		k_unpack<<<(unsigned)((n_elems + 255) / 256), 256, 0, ctx->stream>>>(stage, planes, L, ncomp, il0 + c0, n_elems);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_kernel[k], ctx->stream));
	}
	return LIFE_OK;
}

// Double-buffered the other way round: chunk k is packed (compute stream) while chunk k-1 travels to the host (copy stream).
int download_field(life_ctx *ctx, double *h, const double *planes, int ncomp, int64_t il0, int64_t ncols, const PopShift *ps) {
	const Layout &L = ctx->L;
	const PopShift shift = ps ? *ps : PopShift{};
	const int64_t cols = chunk_columns(ctx, ncomp);
	const size_t half = sizeof(double) * (size_t)(cols * L.Ny * ncomp);
	int rc = ensure_scratch(ctx, 2 * half);
	if (rc) return rc;
	if ((rc = ensure_copy_stream(ctx))) return rc;
	LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_copy[0], ctx->copy_stream));
	LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_copy[1], ctx->copy_stream));
	int k = 0;
	for (int64_t c0 = 0; c0 < ncols; c0 += cols, k ^= 1) {
		const int64_t nc = (c0 + cols <= ncols) ? cols : ncols - c0;
		const int64_t n_elems = nc * L.Ny * ncomp;
		double *stage = reinterpret_cast<double *>(reinterpret_cast<char *>(ctx->scratch) + (k ? half : 0));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[k], 0));             // the copy that last read this half is done
		k_pack<<<(unsigned)((n_elems + 255) / 256), 256, 0, ctx->stream>>>(stage, planes, L, ncomp, il0 + c0, n_elems, shift);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_kernel[k], ctx->stream));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_kernel[k], 0));
		LIFE_CUDA(ctx, cudaMemcpyAsync(h + c0 * L.Ny * ncomp, stage, sizeof(double) * n_elems, cudaMemcpyDeviceToHost, ctx->copy_stream));
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_copy[k], ctx->copy_stream));
	}
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

// planes[k] := vals[k] on columns [il0, il0 + ncols)
__global__ void k_fill(double *__restrict__ planes, Layout L, int ncomp, int64_t il0, int64_t n_nodes, double v0, double v1) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_nodes) return;
	const int64_t idx = L.node(il0 + e / L.Ny, e % L.Ny);
	planes[idx] = v0;
	if (ncomp > 1) planes[L.S + idx] = v1;
}

int fill_field(life_ctx *ctx, double *planes, int ncomp, int64_t il0, int64_t ncols, double v0, double v1) {
	const int64_t n = ncols * ctx->L.Ny;
	if (n <= 0) return LIFE_OK;
	k_fill<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(planes, ctx->L, ncomp, il0, n, v0, v1);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

// ---- end-of-step macroscopics of every node ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_macro(const MacroArgs a, double *out, int64_t c_first) {
	const int64_t tiles = (a.L.Ny + blockDim.x - 1) / blockDim.x;
	const int64_t col = c_first + blockIdx.x / tiles;
	const int64_t j = (int64_t)(blockIdx.x % tiles) * blockDim.x + threadIdx.x;
	if (j >= a.L.Ny) return;
	const int64_t idx = col * a.L.P + j + JOFF;
	double rho, ux, uy;
	node_macro(a, idx, rho, ux, uy);
	out[idx] = rho;
	out[a.L.S + idx] = ux;
	out[2 * a.L.S + idx] = uy;
}

MacroArgs macro_args(life_ctx *ctx) {
	MacroArgs a{};
	a.f = ctx->fA;
	a.ps = ctx->shift;
	a.shifted = 0;
	for (int v = 0; v < 9; v++) a.shifted |= ctx->shift.off[v] != 0;
	a.L = ctx->L;
	a.fxy_mode = ctx->fxy_mode;
	a.fx = ctx->fxy_uniform[0]; a.fy = ctx->fxy_uniform[1];
	a.fxyf = ctx->fxyf;
	a.fibm = ctx->fibm_any ? ctx->fibm : nullptr;
	return a;
}

int launch_macro(life_ctx *ctx, double *out_planes, int64_t il0, int64_t ncols) {
	MacroArgs a = macro_args(ctx);
	const int64_t tiles = (a.L.Ny + 255) / 256;
	const int64_t blocks = tiles * ncols;
	if (blocks <= 0) return LIFE_OK;
	k_macro<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, out_planes, il0 + 1);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

// ---- max |u| and NaN scan (src/Grid.cpp:562-588) -------------------------------------------------------------------------------
// red[0] = bits of the maximum speed (non-negative doubles order like unsigned integers), red[1] = smallest global node id
// holding a NaN speed (i-major order = the order the reference's double loop meets them), ~0 if none.
// Persistent blocks stride over (column, 512-row tile) pairs; each thread handles two adjacent rows with 16-byte loads, like the
// bulk sweep, so the scan runs at the HBM rate of its 72 B per node.
// PATH 0: uploaded macroscopics; 1: forces / shifted layout (generic node_macro); 2: force-free plain layout, 16-byte loads.
// Separate instantiations: the generic path's registers must not cost the common force-free scan its occupancy.
template <int PATH>
__global__ void __launch_bounds__(256) k_max_speed(const MacroArgs a, const double *stored, int64_t i_begin,
                                                   unsigned long long *red) {
	__shared__ unsigned long long smax[256], snan[256];
	const int64_t tiles = (a.L.Ny + 511) / 512;
	const int64_t n_tiles = tiles * a.L.nxl;
	unsigned long long vmax = 0ull, nanid = ~0ull;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t il = tile / tiles;
		const int64_t j = ((tile - il * tiles) * 256 + threadIdx.x) * 2;
		if (j >= a.L.Ny) continue;
		const int64_t idx = a.L.node(il, j);      // even row offset: 16-byte aligned
		const bool two = j + 1 < a.L.Ny;
		double ux[2], uy[2];
		if (PATH == 0) {
			ux[0] = stored[a.L.S + idx]; uy[0] = stored[2 * a.L.S + idx];
			ux[1] = two ? stored[a.L.S + idx + 1] : 0.0; uy[1] = two ? stored[2 * a.L.S + idx + 1] : 0.0;
		} else if (PATH == 1) {
			double rho;
			node_macro(a, idx, rho, ux[0], uy[0]);
			if (two) node_macro(a, idx + 1, rho, ux[1], uy[1]);
			else { ux[1] = 0.0; uy[1] = 0.0; }
		} else {
			double p0[NV], p1[NV];
#pragma unroll
			for (int v = 0; v < NV; v++) {
				// the row after the last one is the ghost row: inside the padded pitch, never out of bounds
				const double2 t = __ldg(reinterpret_cast<const double2 *>(a.f + v * a.L.S + idx));
				p0[v] = t.x; p1[v] = t.y;
			}
			double rho, mx, my;
			moments(p0, rho, mx, my);
			ux[0] = (mx + 0.5 * a.fx) / rho; uy[0] = (my + 0.5 * a.fy) / rho;
			moments(p1, rho, mx, my);
			ux[1] = two ? (mx + 0.5 * a.fx) / rho : 0.0; uy[1] = two ? (my + 0.5 * a.fy) / rho : 0.0;
		}
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const double vel = sqrt(ux[k] * ux[k] + uy[k] * uy[k]);
			if (vel != vel) {
				const unsigned long long gid = (unsigned long long)((i_begin + il) * a.L.Ny + j + k);
				if (gid < nanid) nanid = gid;
			} else {
				const unsigned long long b = (unsigned long long)__double_as_longlong(vel);
				if (b > vmax) vmax = b;
			}
		}
	}
	smax[threadIdx.x] = vmax; snan[threadIdx.x] = nanid;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if ((int)threadIdx.x < s) {
			if (smax[threadIdx.x + s] > smax[threadIdx.x]) smax[threadIdx.x] = smax[threadIdx.x + s];
			if (snan[threadIdx.x + s] < snan[threadIdx.x]) snan[threadIdx.x] = snan[threadIdx.x + s];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		atomicMax(&red[0], smax[0]);
		atomicMin(&red[1], snan[0]);
	}
}

int launch_max_speed(life_ctx *ctx, double *vmax, int32_t *has_nan, int64_t *nan_id) {
	if (!ctx->d_red) LIFE_CUDA(ctx, cudaMalloc(&ctx->d_red, 64));
	unsigned long long init[2] = {0ull, ~0ull};
	LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_red, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
	MacroArgs a = macro_args(ctx);
	int64_t blocks = ((a.L.Ny + 511) / 512) * a.L.nxl;
	const int path = ctx->stored_macro_valid ? 0 : ((a.fxy_mode == FXY_FIELD || a.fibm || a.shifted) ? 1 : 2);
	// persistent: exactly one wave of resident blocks (a fixed 8 per SM used to run 1.6 waves at the kernel's real occupancy)
	// (queried per call: occupancy is a property of the device the context lives on, and one process may hold several)
	int per_sm[3] = {0, 0, 0};
	{
		cudaError_t e = path == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], k_max_speed<0>, 256, 0)
		              : path == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], k_max_speed<1>, 256, 0)
		                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[2], k_max_speed<2>, 256, 0);
		if (e != cudaSuccess || per_sm[path] < 1) per_sm[path] = 4;
	}
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	if (blocks > (int64_t)sms * per_sm[path]) blocks = (int64_t)sms * per_sm[path];
	unsigned long long *red = reinterpret_cast<unsigned long long *>(ctx->d_red);
	const double *stored = ctx->stored_macro_valid ? ctx->macro : nullptr;
	if (path == 0) k_max_speed<0><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, stored, ctx->i_begin, red);
	else if (path == 1) k_max_speed<1><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, stored, ctx->i_begin, red);
	else k_max_speed<2><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, stored, ctx->i_begin, red);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	if (ctx->comm) {
		LIFE_NCCL(ctx, ncclGroupStart());
		LIFE_NCCL(ctx, ncclAllReduce(ctx->d_red, ctx->d_red, 1, ncclDouble, ncclMax, ctx->comm, ctx->stream));
		LIFE_NCCL(ctx, ncclAllReduce(ctx->d_red + 1, ctx->d_red + 1, 1, ncclUint64, ncclMin, ctx->comm, ctx->stream));
		LIFE_NCCL(ctx, ncclGroupEnd());
	}
	unsigned long long out[2];
	LIFE_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_red, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	double v;
	memcpy(&v, &out[0], sizeof(double));
	*vmax = v;
	*has_nan = out[1] != ~0ull;
	*nan_id = *has_nan ? (int64_t)out[1] : -1;
	return LIFE_OK;
}

}  // namespace life
