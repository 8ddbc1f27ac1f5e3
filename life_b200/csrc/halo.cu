// x-direction completion of the push: move what the bulk sweep wrote into the ghost columns to where it belongs.
//
// The bulk sweep pushes the three cx = +1 populations (v = 1, 5, 7) of the last owned column into ghost column nxl+1 and the
// three cx = -1 populations (v = 2, 6, 8) of the first owned column into ghost column 0.  Their true destination is the
// first / last column of the neighbouring slab — or, at the ends of the lattice, the opposite end (the reference's periodic
// modulo, src/Grid.cpp:229), which is only consumed when the receiving column is periodic (type eFluid).
//
// Per face and step that is 3*Ny doubles (24*Ny bytes), contiguous per plane, so it is sent straight out of the ghost column
// and received straight into the destination column: no pack/unpack kernels.
//   nranks == 1 : three device-to-device copies per periodic face.
//   nranks  > 1 : one NCCL group of paired ncclSend/ncclRecv on the communication stream; life_step launches the two edge
//                 columns first so this overlaps the interior sweep.
#include "ctx.h"

namespace life {

static const int kRight[3] = {1, 5, 7};   // cx = +1
static const int kLeft[3] = {2, 6, 8};    // cx = -1

// ---- cfg.inplace: the planes are circular (ctx.h: PopShift), so a column of a plane is addressed element by element -----------------
// single rank: ghost column -> opposite real column, both faces in one launch
__global__ void k_ring_x(double *f, PopShift ps, Layout L, int left_periodic, int right_periodic) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= 3 * L.Ny) return;
	const int k = (int)(e / L.Ny);
	const int64_t r = JOFF + e % L.Ny;
	const int vr = k == 0 ? 1 : (k == 1 ? 5 : 7), vl = k == 0 ? 2 : (k == 1 ? 6 : 8);
	if (left_periodic) f[ps.at(vr, L.at(1, r), L.S)] = f[ps.at(vr, L.at(L.nxl + 1, r), L.S)];
	if (right_periodic) f[ps.at(vl, L.at(L.nxl, r), L.S)] = f[ps.at(vl, L.at(0, r), L.S)];
}
// several ranks: the two ghost columns are gathered into contiguous send buffers ([0] -> right neighbour: cx = +1 planes of ghost
// column nxl+1; [1] -> left neighbour: cx = -1 planes of ghost column 0), and what arrives is scattered into the edge columns
__global__ void k_halo_pack(const double *f, PopShift ps, Layout L, double *send) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= 3 * L.Ny) return;
	const int k = (int)(e / L.Ny);
	const int64_t r = JOFF + e % L.Ny;
	const int vr = k == 0 ? 1 : (k == 1 ? 5 : 7), vl = k == 0 ? 2 : (k == 1 ? 6 : 8);
	send[e] = f[ps.at(vr, L.at(L.nxl + 1, r), L.S)];
	send[3 * L.Ny + e] = f[ps.at(vl, L.at(0, r), L.S)];
}
__global__ void k_halo_unpack(double *f, PopShift ps, Layout L, const double *recv, int from_left, int from_right) {
	const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= 3 * L.Ny) return;
	const int k = (int)(e / L.Ny);
	const int64_t r = JOFF + e % L.Ny;
	const int vr = k == 0 ? 1 : (k == 1 ? 5 : 7), vl = k == 0 ? 2 : (k == 1 ? 6 : 8);
	if (from_left) f[ps.at(vr, L.at(1, r), L.S)] = recv[e];                    // cx = +1 populations arrive from the left
	if (from_right) f[ps.at(vl, L.at(L.nxl, r), L.S)] = recv[3 * L.Ny + e];    // cx = -1 populations arrive from the right
}

int exchange_x(life_ctx *ctx) {
	const Layout &L = ctx->L;
	const life_config &c = ctx->cfg;
	double *f = ctx->fB;
	const size_t bytes = sizeof(double) * (size_t)L.Ny;
	// Column 0 consumes what wraps in from column Nx-1 wherever it holds a fluid site: everywhere if the left wall is
	// periodic, and at its two end nodes if the bottom/top wall is periodic (their type overrides the left wall's at the
	// corners, src/Grid.cpp:940-945).  Copying the whole column is harmless: boundary nodes rebuild those populations.
	const bool tb_periodic = c.wall_bottom == LIFE_FLUID || c.wall_top == LIFE_FLUID;
	const bool left_periodic = c.wall_left == LIFE_FLUID || tb_periodic;
	const bool right_periodic = c.wall_right == LIFE_FLUID || tb_periodic;   // same for column Nx-1

	if (ctx->inplace && c.nranks <= 1) {
		if (left_periodic || right_periodic) {
			k_ring_x<<<(unsigned)((3 * L.Ny + 255) / 256), 256, 0, ctx->stream>>>(ctx->fA, ctx->shift, L, left_periodic, right_periodic);
			ctx->launches++;
			LIFE_CUDA(ctx, cudaGetLastError());
		}
		return LIFE_OK;
	}
	if (c.nranks <= 1) {
		for (int k = 0; k < 3; k++) {
			if (left_periodic) {
				const int v = kRight[k];
				LIFE_CUDA(ctx, cudaMemcpyAsync(f + v * L.S + L.at(1, JOFF), f + v * L.S + L.at(L.nxl + 1, JOFF), bytes,
				                               cudaMemcpyDeviceToDevice, ctx->stream));
			}
			if (right_periodic) {
				const int v = kLeft[k];
				LIFE_CUDA(ctx, cudaMemcpyAsync(f + v * L.S + L.at(L.nxl, JOFF), f + v * L.S + L.at(0, JOFF), bytes,
				                               cudaMemcpyDeviceToDevice, ctx->stream));
			}
		}
		return LIFE_OK;
	}

	// neighbours; -1 = none (non-periodic end of the lattice)
	const int r = c.rank, n = c.nranks;
	const int right = (r + 1 < n) ? r + 1 : (left_periodic ? 0 : -1);          // receives my cx=+1 ghost column
	const int left = (r > 0) ? r - 1 : (right_periodic ? n - 1 : -1);          // receives my cx=-1 ghost column
	// what I receive: cx=+1 populations from my left neighbour unless I am rank 0 of a non-periodic lattice, etc.
	const int from_left = (r > 0) ? r - 1 : (left_periodic ? n - 1 : -1);
	const int from_right = (r + 1 < n) ? r + 1 : (right_periodic ? 0 : -1);

	cudaStream_t cs = ctx->comm_stream;
	if (ctx->inplace) {
		const size_t n3 = (size_t)(3 * L.Ny);
		if (!ctx->halo_buf) LIFE_CUDA(ctx, cudaMalloc(&ctx->halo_buf, sizeof(double) * 4 * n3));      // send[2][3 Ny] | recv[2][3 Ny]
		double *send = ctx->halo_buf, *recv = ctx->halo_buf + 2 * n3;
		const unsigned blocks = (unsigned)((n3 + 255) / 256);
		k_halo_pack<<<blocks, 256, 0, cs>>>(ctx->fA, ctx->shift, L, send);
		LIFE_CUDA(ctx, cudaGetLastError());
		LIFE_NCCL(ctx, ncclGroupStart());
		if (right >= 0) LIFE_NCCL(ctx, ncclSend(send, n3, ncclDouble, right, ctx->comm, cs));
		if (left >= 0) LIFE_NCCL(ctx, ncclSend(send + n3, n3, ncclDouble, left, ctx->comm, cs));
		if (from_left >= 0) LIFE_NCCL(ctx, ncclRecv(recv, n3, ncclDouble, from_left, ctx->comm, cs));
		if (from_right >= 0) LIFE_NCCL(ctx, ncclRecv(recv + n3, n3, ncclDouble, from_right, ctx->comm, cs));
		LIFE_NCCL(ctx, ncclGroupEnd());
		k_halo_unpack<<<blocks, 256, 0, cs>>>(ctx->fA, ctx->shift, L, recv, from_left >= 0, from_right >= 0);
		ctx->launches += 2;
		LIFE_CUDA(ctx, cudaGetLastError());
		return LIFE_OK;
	}
	LIFE_NCCL(ctx, ncclGroupStart());
	for (int k = 0; k < 3; k++) {
		if (right >= 0)
			LIFE_NCCL(ctx, ncclSend(f + kRight[k] * L.S + L.at(L.nxl + 1, JOFF), (size_t)L.Ny, ncclDouble, right, ctx->comm, cs));
		if (left >= 0)
			LIFE_NCCL(ctx, ncclSend(f + kLeft[k] * L.S + L.at(0, JOFF), (size_t)L.Ny, ncclDouble, left, ctx->comm, cs));
		if (from_left >= 0)
			LIFE_NCCL(ctx, ncclRecv(f + kRight[k] * L.S + L.at(1, JOFF), (size_t)L.Ny, ncclDouble, from_left, ctx->comm, cs));
		if (from_right >= 0)
			LIFE_NCCL(ctx, ncclRecv(f + kLeft[k] * L.S + L.at(L.nxl, JOFF), (size_t)L.Ny, ncclDouble, from_right, ctx->comm, cs));
	}
	LIFE_NCCL(ctx, ncclGroupEnd());
	return LIFE_OK;
}

}  // namespace life
