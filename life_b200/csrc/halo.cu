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
