// Memory-system ceilings of the device the library runs on, measured with the library's own streaming kernels (life_membw).
//
// The bulk sweep of the D2Q9 step is an HBM-bound 1:1 read/write stream (144 B per node).  What fraction of the DRAM peak such a
// stream can reach at all is a property of the memory system (read/write turnaround, refresh, channel hashing), not of the
// kernel; these kernels measure it so that the sweep's achieved bandwidth can be put next to the best ANY kernel of the same
// read/write mix reaches on the same box (profiles/):
//   mode 0  read only     16-byte LDG, 8 independent loads in flight per thread, result folded into one store per CTA
//   mode 1  write only    16-byte STG
//   mode 2  copy          16-byte LDG + STG, 4 in flight per thread, one contiguous stream in, one out
//   mode 3  copy          32-byte LDG.256 + STG.256 (sm_100)
//   mode 4  copy          TMA: cp.async.bulk global->shared (mbarrier) then shared->global (bulk group), 16 KiB chunks, 3 stages,
//                         persistent CTAs — no LSU instruction touches global memory
//   mode 5  copy          the sweep's own access shape without its arithmetic: 9 planes in, 9 planes out, 16-byte accesses, one CTA
//                         per 4 KB run of every plane (the launch shape of k_bulk_shuffle)
//   mode 6  copy          the same with 32-byte accesses, one CTA per 8 KB run (the launch shape of k_bulk_quad)
// Nothing here is on the product path.
#include "ctx.h"
#include <cstdio>

namespace life {

__global__ void __launch_bounds__(256) k_bw_read(const double2 *__restrict__ in, double *out, int64_t n16) {
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	double acc = 0.0;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 7 * stride < n16; i += 8 * stride) {
		double2 v[8];
#pragma unroll
		for (int k = 0; k < 8; k++) v[k] = __ldg(in + i + k * stride);
#pragma unroll
		for (int k = 0; k < 8; k++) acc += v[k].x + v[k].y;
	}
	for (; i < n16; i += stride) { const double2 v = __ldg(in + i); acc += v.x + v.y; }
	if (acc == 1.2345e300) out[0] = acc;     // never true: keeps the loads alive
}

__global__ void __launch_bounds__(256) k_bw_write(double2 *out, int64_t n16) {
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = make_double2(1.0, 2.0);
}

__global__ void __launch_bounds__(256) k_bw_copy16(const double2 *__restrict__ in, double2 *out, int64_t n16) {
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 3 * stride < n16; i += 4 * stride) {
		double2 v[4];
#pragma unroll
		for (int k = 0; k < 4; k++) v[k] = __ldg(in + i + k * stride);
#pragma unroll
		for (int k = 0; k < 4; k++) out[i + k * stride] = v[k];
	}
	for (; i < n16; i += stride) out[i] = __ldg(in + i);
}

__global__ void __launch_bounds__(256) k_bw_copy32(const double *__restrict__ in, double *out, int64_t n32) {
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + stride < n32; i += 2 * stride) {
		double a[4], b[4];
		asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(in + 4 * i));
		asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b[0]), "=d"(b[1]), "=d"(b[2]), "=d"(b[3]) : "l"(in + 4 * (i + stride)));
		asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * i), "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]) : "memory");
		asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * (i + stride)), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]) : "memory");
	}
	for (; i < n32; i += stride) {
		double a[4];
		asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(in + 4 * i));
		asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * i), "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]) : "memory");
	}
}

// 9 planes in, 9 planes out, one CTA per 512-double run of a plane position: the sweep's address pattern
__global__ void __launch_bounds__(256) k_bw_planes(const double *__restrict__ in, double *out, int64_t plane, int64_t runs) {
	for (int64_t r = blockIdx.x; r < runs; r += gridDim.x) {
		const int64_t idx = r * 512 + 2 * threadIdx.x;
		double2 v[9];
#pragma unroll
		for (int p = 0; p < 9; p++) v[p] = __ldg(reinterpret_cast<const double2 *>(in + p * plane + idx));
#pragma unroll
		for (int p = 0; p < 9; p++) *reinterpret_cast<double2 *>(out + p * plane + idx) = v[p];
	}
}

// the same with 32-byte accesses: one CTA per 1024-double run
__global__ void __launch_bounds__(256) k_bw_planes32(const double *__restrict__ in, double *out, int64_t plane, int64_t runs) {
	for (int64_t r = blockIdx.x; r < runs; r += gridDim.x) {
		const int64_t idx = r * 1024 + 4 * threadIdx.x;
		double v[9][4];
#pragma unroll
		for (int p = 0; p < 9; p++)
			asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[p][0]), "=d"(v[p][1]), "=d"(v[p][2]), "=d"(v[p][3]) : "l"(in + p * plane + idx));
#pragma unroll
		for (int p = 0; p < 9; p++)
			asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + p * plane + idx), "d"(v[p][0]), "d"(v[p][1]), "d"(v[p][2]), "d"(v[p][3]) : "memory");
	}
}

// ---- TMA copy ---------------------------------------------------------------------------------------------------------------
constexpr int BW_CHUNK = 16384;   // bytes per bulk copy
constexpr int BW_STAGES = 3;

__device__ __forceinline__ uint32_t bw_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k_bw_tma(const char *__restrict__ in, char *out, int64_t chunks) {
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + BW_STAGES * BW_CHUNK);
	if (threadIdx.x != 0) return;      // one elected thread drives the whole pipeline: the copies themselves are asynchronous
	for (int s = 0; s < BW_STAGES; s++)
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bw_smem(&full[s])), "r"(1) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	auto load = [&](int64_t c, int s) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bw_smem(&full[s])), "r"(BW_CHUNK) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bw_smem(smem + s * BW_CHUNK)),
		             "l"(in + c * BW_CHUNK), "r"(BW_CHUNK), "r"(bw_smem(&full[s]))
		             : "memory");
	};
	int64_t c = blockIdx.x;
	for (int s = 0; s < BW_STAGES; s++)
		if (c + (int64_t)s * gridDim.x < chunks) load(c + (int64_t)s * gridDim.x, s);
	int s = 0;
	unsigned phase = 0;
	for (; c < chunks; c += gridDim.x) {
		asm volatile(
		    "{\n.reg .pred P1;\nBW_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra BW_DONE;\nbra BW_WAIT;\nBW_DONE:\n}" ::"r"(bw_smem(&full[s])),
		    "r"(phase)
		    : "memory");
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * BW_CHUNK), "r"(bw_smem(smem + s * BW_CHUNK)), "r"(BW_CHUNK)
		             : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		// the stage can be refilled once the store has READ it
		asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
		const int64_t next = c + (int64_t)BW_STAGES * gridDim.x;
		if (next < chunks) load(next, s);
		if (++s == BW_STAGES) { s = 0; phase ^= 1; }
	}
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace life

using namespace life;

extern "C" int life_membw(int32_t device, int32_t mode, int64_t bytes, int32_t iters, double *gbytes_per_s) {
	if (!gbytes_per_s || bytes < (1 << 20) || iters < 1) return LIFE_E_ARG;
	if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return LIFE_E_CUDA;
	int sms = 148;
	int dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	bytes = bytes / (9 * 16384) * (9 * 16384);     // whole chunks, whole planes
	char *a = nullptr, *b = nullptr;
	if (cudaMalloc(&a, bytes) != cudaSuccess) return LIFE_E_NOMEM;
	if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return LIFE_E_NOMEM; }
	cudaMemset(a, 0, bytes);
	cudaMemset(b, 0, bytes);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int64_t n16 = bytes / 16;
	const unsigned grid = (unsigned)(sms * 8);
	double moved = 2.0 * (double)bytes;
	if (mode == 4) cudaFuncSetAttribute(k_bw_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_STAGES * BW_CHUNK + 64);
	float best = 1e30f;
	for (int it = 0; it < iters + 2; it++) {
		cudaEventRecord(e0);
		switch (mode) {
		case 0: k_bw_read<<<grid, 256>>>(reinterpret_cast<const double2 *>(a), reinterpret_cast<double *>(b), n16); moved = (double)bytes; break;
		case 1: k_bw_write<<<grid, 256>>>(reinterpret_cast<double2 *>(b), n16); moved = (double)bytes; break;
		case 2: k_bw_copy16<<<grid, 256>>>(reinterpret_cast<const double2 *>(a), reinterpret_cast<double2 *>(b), n16); break;
		case 3: k_bw_copy32<<<grid, 256>>>(reinterpret_cast<const double *>(a), reinterpret_cast<double *>(b), bytes / 32); break;
		case 4: k_bw_tma<<<(unsigned)(sms * 4), 32, BW_STAGES * BW_CHUNK + 64>>>(a, b, bytes / BW_CHUNK); break;
		case 5: k_bw_planes<<<(unsigned)(bytes / 8 / 9 / 512), 256>>>(reinterpret_cast<const double *>(a), reinterpret_cast<double *>(b), bytes / 8 / 9, bytes / 8 / 9 / 512); break;
		case 6: k_bw_planes32<<<(unsigned)(bytes / 8 / 9 / 1024), 256>>>(reinterpret_cast<const double *>(a), reinterpret_cast<double *>(b), bytes / 8 / 9, bytes / 8 / 9 / 1024); break;
		default: cudaFree(a); cudaFree(b); return LIFE_E_ARG;
		}
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(a); cudaFree(b); return LIFE_E_CUDA; }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (it >= 2 && ms < best) best = ms;
	}
	cudaError_t e = cudaGetLastError();
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(a);
	cudaFree(b);
	if (e != cudaSuccess) return LIFE_E_CUDA;
	*gbytes_per_s = moved / (best * 1e-3) / 1e9;
	return LIFE_OK;
}
