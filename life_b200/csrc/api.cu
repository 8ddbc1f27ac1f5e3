// C ABI of liblife_b200 (include/life_b200.h): context life cycle, state upload/download, the step driver.
#include "ctx.h"
#include <cmath>
#include <cstring>
#include <new>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace life {

static thread_local std::string g_create_error;

int fail(life_ctx *ctx, int code, const std::string &msg) {
	if (ctx) ctx->err = msg;
	else g_create_error = msg;
	return code;
}

// per-step scalars evaluated on the host with the reference's own expressions
StepScalars step_scalars(life_ctx *ctx, int32_t t) {
	const life_config &c = ctx->cfg;
	StepScalars s{};
	// getRampCoefficient, src/Grid.cpp:548-556
	s.ramp = 1.0;
	if (c.inlet_ramp > 0.0 && c.Dt * t <= c.inlet_ramp) s.ramp = (1.0 - cos(M_PI * c.Dt * t / c.inlet_ramp)) / 2.0;
	s.fxy_prev[0] = ctx->fxy_uniform[0];
	s.fxy_prev[1] = ctx->fxy_uniform[1];
	s.fxy_cur[0] = ctx->fxy_uniform[0];
	s.fxy_cur[1] = ctx->fxy_uniform[1];
	s.wom_cos = 1.0;
	if (c.womersley > 0.0) {
		// src/Grid.cpp:59-60
		s.wom_cos = cos(2.0 * M_PI * t * c.Dt / (((c.height_p * c.height_p) * M_PI) / (2.0 * (c.womersley * c.womersley) * c.nu_p)));
		if (!ctx->wom_field) {
			// gravity == 0: the force is uniform, (rho_n*Drho*0 + dpd*cos) * SQ(Dx*Dt) / Dm
			const double sq = (c.Dx * c.Dt) * (c.Dx * c.Dt);
			s.fxy_cur[0] = (0.0 + c.dpdx * s.wom_cos) * sq / c.Dm;
			s.fxy_cur[1] = (0.0 + c.dpdy * s.wom_cos) * sq / c.Dm;
		}
	}
	return s;
}

static void free_ctx(life_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->io) { io_wait(ctx); io_free(ctx); }   // completes a pending asynchronous file write first
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	fem_free(ctx);
	ibm_free(ctx);
	if (ctx->comm) ncclCommDestroy(ctx->comm);
	cudaFree(ctx->bc_prev); cudaFree(ctx->halo_buf);
	cudaFree(ctx->fA); cudaFree(ctx->fB); cudaFree(ctx->macro); cudaFree(ctx->fibm); cudaFree(ctx->fibm_mask); cudaFree(ctx->fxyf);
	cudaFree(ctx->cell_head); cudaFree(ctx->u_in); cudaFree(ctx->rho_in); cudaFree(ctx->delU); cudaFree(ctx->bc);
	cudaFree(ctx->scratch); cudaFree(ctx->d_red); cudaFree(ctx->eps_buf); cudaFree(ctx->d_steps);
	if (ctx->h_steps) cudaFreeHost(ctx->h_steps);
	if (ctx->ev_steps) cudaEventDestroy(ctx->ev_steps);
	for (int k = 0; k < 2; k++) {
		if (ctx->ev_copy[k]) cudaEventDestroy(ctx->ev_copy[k]);
		if (ctx->ev_kernel[k]) cudaEventDestroy(ctx->ev_kernel[k]);
	}
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
	for (auto &p : ctx->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
	if (ctx->ev_edge) cudaEventDestroy(ctx->ev_edge);
	if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
	if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

}  // namespace life

using namespace life;

extern "C" {

int life_abi_version(void) { return LIFE_ABI_VERSION; }

int life_nccl_unique_id(void *out128) {
	if (!out128) return fail(nullptr, LIFE_E_ARG, "life_nccl_unique_id: null buffer");
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	ncclResult_t r = ncclGetUniqueId(&id);
	if (r != ncclSuccess) return fail(nullptr, LIFE_E_NCCL, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
	memcpy(out128, &id, 128);
	return LIFE_OK;
}

const char *life_last_error(const life_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int life_create(const life_config *cfg, life_ctx **out) {
	if (!cfg || !out) return fail(nullptr, LIFE_E_ARG, "life_create: null argument");
	*out = nullptr;
	if (cfg->abi_version != LIFE_ABI_VERSION) return fail(nullptr, LIFE_E_ARG, "life_create: ABI version mismatch");
	if (cfg->Nx < 3 || cfg->Ny < 3) return fail(nullptr, LIFE_E_ARG, "life_create: lattice must be at least 3 x 3");
	if (cfg->collision != LIFE_BGK && cfg->collision != LIFE_CENTRAL_MOMENTS)
		return fail(nullptr, LIFE_E_ARG, "life_create: unknown collision operator");
	const int walls[4] = {cfg->wall_left, cfg->wall_right, cfg->wall_bottom, cfg->wall_top};
	for (int w : walls)
		if (w < LIFE_FLUID || w > LIFE_CONVECTIVE) return fail(nullptr, LIFE_E_ARG, "life_create: unknown wall type");
	if (!(cfg->Dx > 0.0) || !(cfg->Dt > 0.0) || !(cfg->Dm > 0.0))
		return fail(nullptr, LIFE_E_ARG, "life_create: Dx, Dt, Dm must be positive");
	const int nranks = cfg->nranks <= 1 ? 1 : cfg->nranks;
	const int rank = nranks == 1 ? 0 : cfg->rank;
	if (rank < 0 || rank >= nranks) return fail(nullptr, LIFE_E_ARG, "life_create: rank out of range");
	if (nranks > 1 && !cfg->nccl_id) return fail(nullptr, LIFE_E_ARG, "life_create: nranks > 1 needs nccl_id");
	if (nranks > 1 && cfg->Nx / nranks < 4) return fail(nullptr, LIFE_E_ARG, "life_create: slabs must be at least 4 columns wide");

	// device: there is no CPU fallback
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(nullptr, LIFE_E_CUDA, std::string("life_create: no CUDA device (") + cudaGetErrorString(e) + "); liblife_b200 has no CPU path");
	int dev = cfg->device;
	if (dev < 0) cudaGetDevice(&dev);
	if (dev >= ndev) return fail(nullptr, LIFE_E_ARG, "life_create: device ordinal out of range");
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, dev);
	if (e != cudaSuccess) return fail(nullptr, LIFE_E_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
	if (prop.major != 10)
		return fail(nullptr, LIFE_E_CUDA, std::string("life_create: device '") + prop.name + "' is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
	e = cudaSetDevice(dev);
	if (e != cudaSuccess) return fail(nullptr, LIFE_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));

	life_ctx *ctx = new (std::nothrow) life_ctx();
	if (!ctx) return fail(nullptr, LIFE_E_NOMEM, "life_create: out of host memory");
	ctx->cfg = *cfg;
	ctx->cfg.nranks = nranks;
	ctx->cfg.rank = rank;
	ctx->cfg.nccl_id = nullptr;
	ctx->device = dev;

	// slab of columns owned by this rank: balanced split of Nx
	life_slab_range(cfg->Nx, nranks, rank, &ctx->i_begin, &ctx->i_end);
	Layout &L = ctx->L;
	L.Ny = cfg->Ny;
	L.nxl = ctx->i_end - ctx->i_begin;
	L.P = ((JOFF + cfg->Ny + 1 + 15) / 16) * 16;
	L.S = (L.nxl + 2) * L.P;

	auto bail = [&](int code) { std::string m = ctx->err; free_ctx(ctx); g_create_error = m; return code; };
#define CK(call)                                                                      \
	do {                                                                              \
		cudaError_t e__ = (call);                                                     \
		if (e__ != cudaSuccess) {                                                     \
			ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);           \
			return bail(e__ == cudaErrorMemoryAllocation ? LIFE_E_NOMEM : LIFE_E_CUDA); \
		}                                                                             \
	} while (0)

	if (cfg->stream) ctx->stream = reinterpret_cast<cudaStream_t>(cfg->stream);
	else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
	CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&ctx->ev_edge, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));

	// + 64 doubles: the last warp of a column loads its full 64-row span even when the column ends inside it (lbm_bulk.cu), which
	// for tiny Ny (pitch < 64 rows) reaches past the ghost column of the last plane
	const size_t fbytes = sizeof(double) * (9 * (size_t)L.S + 64);
	ctx->inplace = cfg->inplace != 0;
	CK(cudaMalloc(&ctx->fA, fbytes));
	CK(cudaMemsetAsync(ctx->fA, 0, fbytes, ctx->stream));
	if (!ctx->inplace) {      // cfg.inplace: one buffer, 72 B/node
		CK(cudaMalloc(&ctx->fB, fbytes));
		CK(cudaMemsetAsync(ctx->fB, 0, fbytes, ctx->stream));
	}
	CK(cudaMalloc(&ctx->u_in, sizeof(double) * 2 * L.Ny));
	CK(cudaMalloc(&ctx->rho_in, sizeof(double) * L.Ny));
	CK(cudaMalloc(&ctx->delU, sizeof(double) * 2 * L.Ny));
	CK(cudaMemsetAsync(ctx->u_in, 0, sizeof(double) * 2 * L.Ny, ctx->stream));
	CK(cudaMemsetAsync(ctx->delU, 0, sizeof(double) * 2 * L.Ny, ctx->stream));
	CK(cudaMallocHost(&ctx->h_pin, 4096));
	ctx->h_pin_bytes = 4096;
#undef CK

	int rc = build_boundary(ctx);
	if (rc) return bail(rc);

	if (nranks > 1) {
		ncclUniqueId id;
		memcpy(&id, cfg->nccl_id, sizeof(id));
		ncclResult_t r = ncclCommInitRank(&ctx->comm, nranks, id, rank);
		if (r != ncclSuccess) {
			ctx->err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
			ctx->comm = nullptr;
			return bail(LIFE_E_NCCL);
		}
	}
	cudaError_t es = cudaStreamSynchronize(ctx->stream);
	if (es != cudaSuccess) { ctx->err = std::string("life_create: ") + cudaGetErrorString(es); return bail(LIFE_E_CUDA); }
	*out = ctx;
	return LIFE_OK;
}

int life_destroy(life_ctx *ctx) {
	free_ctx(ctx);
	return LIFE_OK;
}

int life_slab_range(int64_t Nx, int32_t nranks, int32_t rank, int64_t *i_begin, int64_t *i_end) {
	if (nranks < 1) nranks = 1;
	if (Nx < 0 || rank < 0 || rank >= nranks) return LIFE_E_ARG;
	const int64_t base = Nx / nranks, rem = Nx % nranks;
	const int64_t b = rank * base + (rank < rem ? rank : rem);
	if (i_begin) *i_begin = b;
	if (i_end) *i_end = b + base + (rank < rem ? 1 : 0);
	return LIFE_OK;
}

int life_slab(const life_ctx *ctx, int64_t *i_begin, int64_t *i_end) {
	if (!ctx) return LIFE_E_ARG;
	if (i_begin) *i_begin = ctx->i_begin;
	if (i_end) *i_end = ctx->i_end;
	return LIFE_OK;
}

int life_upload_begin(life_ctx *ctx, const double *u_in, const double *rho_in) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const Layout &L = ctx->L;
	ctx->have_state = false;
	ctx->shift = PopShift{};              // an upload delivers the plain layout
	ctx->uploading = true;
	ctx->up_macro = -1;
	ctx->up_cols = 0;
	ctx->up_fxy_seen = false;
	ctx->up_fxy_uniform = true;
	ctx->up_fxy0[0] = ctx->up_fxy0[1] = 0.0;
	ctx->up_ranges.clear();
	ctx->stored_macro_valid = false;
	ctx->fxy_mode = FXY_NONE;
	ctx->fxy_uniform[0] = ctx->fxy_uniform[1] = 0.0;
	ctx->wom_field = ctx->cfg.womersley > 0.0 && (ctx->cfg.gravity_x != 0.0 || ctx->cfg.gravity_y != 0.0);
	ctx->fibm_any = false;
	ctx->fibm_sites_dirty = false;
	ctx->fibm_full_dirty = false;
	ctx->fibm_consumed = true;
	if (ctx->fibm) {
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm, 0, sizeof(double) * 2 * L.S, ctx->stream));
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fibm_mask, 0, (size_t)(ctx->mask_pitch * (L.nxl + 2)), ctx->stream));
	}
	if (ctx->wom_field) {   // force_xy is a field the sweep recomputes every step (src/Grid.cpp:55-61)
		if (!ctx->fxyf) LIFE_CUDA(ctx, cudaMalloc(&ctx->fxyf, sizeof(double) * 2 * L.S));
		LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fxyf, 0, sizeof(double) * 2 * L.S, ctx->stream));
	}
	if (u_in) LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->u_in, u_in, sizeof(double) * 2 * L.Ny, cudaMemcpyHostToDevice, ctx->stream));
	else LIFE_CUDA(ctx, cudaMemsetAsync(ctx->u_in, 0, sizeof(double) * 2 * L.Ny, ctx->stream));
	if (rho_in) {
		LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->rho_in, rho_in, sizeof(double) * L.Ny, cudaMemcpyHostToDevice, ctx->stream));
	} else {
		std::vector<double> ones((size_t)L.Ny, 1.0);
		LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->rho_in, ones.data(), sizeof(double) * L.Ny, cudaMemcpyHostToDevice, ctx->stream));
	}
	// the host arrays belong to the caller: do not return before the copies out of them have completed
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_upload_columns(life_ctx *ctx, int64_t il0, int64_t ncols, const double *f, const double *rho, const double *u,
                        const double *force_xy, const double *force_ibm) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->uploading) return fail(ctx, LIFE_E_STATE, "life_upload_columns: call life_upload_begin first");
	const Layout &L = ctx->L;
	if (il0 < 0 || ncols < 0 || il0 + ncols > L.nxl) return fail(ctx, LIFE_E_ARG, "life_upload_columns: column range outside the slab");
	if (ncols == 0) return LIFE_OK;
	if (!f) return fail(ctx, LIFE_E_ARG, "life_upload_columns: f is required");
	if ((rho == nullptr) != (u == nullptr)) return fail(ctx, LIFE_E_ARG, "life_upload_columns: give both rho and u, or neither");
	const int has_macro = rho ? 1 : 0;
	if (ctx->up_macro >= 0 && ctx->up_macro != has_macro)
		return fail(ctx, LIFE_E_ARG, "life_upload_columns: rho/u must be given for every column range or for none");
	ctx->up_macro = has_macro;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const int64_t n = ncols * L.Ny;
	int rc;
	if ((rc = upload_field(ctx, f, ctx->fA, 9, il0, ncols))) return rc;
	if (rho) {
		if ((rc = ensure_macro(ctx))) return rc;
		if ((rc = upload_field(ctx, rho, ctx->macro, 1, il0, ncols))) return rc;
		if ((rc = upload_field(ctx, u, ctx->macro + L.S, 2, il0, ncols))) return rc;
	}

	// force_xy: none / uniform / field (src/Grid.cpp:1035-1045 makes it uniform; Womersley with gravity makes it a field).
	// It is kept as two scalars for as long as every node seen so far holds the same pair.
	double fx0 = ctx->up_fxy0[0], fy0 = ctx->up_fxy0[1];
	bool chunk_uniform = true;
	if (force_xy) {
		if (!ctx->up_fxy_seen) { fx0 = force_xy[0]; fy0 = force_xy[1]; }
		for (int64_t k = 0; k < n && chunk_uniform; k++)
			if (force_xy[2 * k] != fx0 || force_xy[2 * k + 1] != fy0) chunk_uniform = false;
	} else if (ctx->up_fxy_seen && (fx0 != 0.0 || fy0 != 0.0)) {
		chunk_uniform = false;   // this range is all zero, earlier ones were not
	}
	const bool was_uniform = ctx->up_fxy_uniform && !ctx->wom_field;
	if (was_uniform && chunk_uniform) {
		if (!ctx->up_fxy_seen) { ctx->up_fxy0[0] = force_xy ? fx0 : 0.0; ctx->up_fxy0[1] = force_xy ? fy0 : 0.0; ctx->up_fxy_seen = true; }
		ctx->up_ranges.emplace_back(il0, ncols);
	} else {
		if (!ctx->fxyf) {
			LIFE_CUDA(ctx, cudaMalloc(&ctx->fxyf, sizeof(double) * 2 * L.S));
			LIFE_CUDA(ctx, cudaMemsetAsync(ctx->fxyf, 0, sizeof(double) * 2 * L.S, ctx->stream));
		}
		if (was_uniform) {   // first non-uniform range: materialise what was held as scalars
			for (auto &r : ctx->up_ranges)
				if ((rc = fill_field(ctx, ctx->fxyf, 2, r.first, r.second, ctx->up_fxy0[0], ctx->up_fxy0[1]))) return rc;
			ctx->up_ranges.clear();
		}
		ctx->up_fxy_uniform = false;
		ctx->up_fxy_seen = true;
		if (force_xy) { if ((rc = upload_field(ctx, force_xy, ctx->fxyf, 2, il0, ncols))) return rc; }
		else if ((rc = fill_field(ctx, ctx->fxyf, 2, il0, ncols, 0.0, 0.0))) return rc;
	}

	// force_ibm of the previous step (restart)
	if (force_ibm) {
		bool any = false;
		for (int64_t k = 0; k < 2 * n && !any; k++) any = force_ibm[k] != 0.0;
		if (any) {
			if ((rc = ensure_fibm(ctx))) return rc;
			if ((rc = upload_field(ctx, force_ibm, ctx->fibm, 2, il0, ncols))) return rc;
			ctx->fibm_any = true;
			ctx->fibm_full_dirty = true;
			ctx->fibm_consumed = false;
		}
	}
	// the host arrays belong to the caller: do not return before the copies out of them have completed
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->up_cols += ncols;
	return LIFE_OK;
}

int life_upload_end(life_ctx *ctx) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->uploading) return fail(ctx, LIFE_E_STATE, "life_upload_end: no upload in progress");
	if (ctx->up_cols != ctx->L.nxl)
		return fail(ctx, LIFE_E_STATE, "life_upload_end: " + std::to_string(ctx->up_cols) + " of " + std::to_string(ctx->L.nxl) + " columns were uploaded");
	ctx->stored_macro_valid = ctx->up_macro == 1;
	if (ctx->wom_field || !ctx->up_fxy_uniform) {
		ctx->fxy_mode = FXY_FIELD;
	} else {
		ctx->fxy_uniform[0] = ctx->up_fxy0[0];
		ctx->fxy_uniform[1] = ctx->up_fxy0[1];
		const bool zero = ctx->up_fxy0[0] == 0.0 && ctx->up_fxy0[1] == 0.0;
		ctx->fxy_mode = (!zero || ctx->cfg.womersley > 0.0) ? FXY_UNIFORM : FXY_NONE;
	}
	ctx->up_ranges.clear();
	ctx->uploading = false;
	ctx->have_state = true;
	return life_sync(ctx);
}

int life_upload_state(life_ctx *ctx, const double *f, const double *rho, const double *u, const double *force_xy,
                      const double *force_ibm, const double *u_in, const double *rho_in) {
	if (!ctx) return LIFE_E_ARG;
	if (!f) return fail(ctx, LIFE_E_ARG, "life_upload_state: f is required");
	if ((rho == nullptr) != (u == nullptr)) return fail(ctx, LIFE_E_ARG, "life_upload_state: give both rho and u, or neither");
	int rc;
	if ((rc = life_upload_begin(ctx, u_in, rho_in))) return rc;
	if ((rc = life_upload_columns(ctx, 0, ctx->L.nxl, f, rho, u, force_xy, force_ibm))) return rc;
	return life_upload_end(ctx);
}

int life_download_columns(life_ctx *ctx, int64_t il0, int64_t ncols, double *f, double *rho, double *u, double *force_ibm) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_download_columns: no state uploaded");
	const Layout &L = ctx->L;
	if (il0 < 0 || ncols < 0 || il0 + ncols > L.nxl) return fail(ctx, LIFE_E_ARG, "life_download_columns: column range outside the slab");
	if (ncols == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc;
	if (f && (rc = download_field(ctx, f, ctx->fA, 9, il0, ncols, &ctx->shift))) return rc;
	if (rho || u) {
		if (!ctx->stored_macro_valid) {   // otherwise `macro` already holds exactly what the host uploaded
			if ((rc = ensure_macro(ctx))) return rc;
			if ((rc = launch_macro(ctx, ctx->macro, il0, ncols))) return rc;
		}
		if (rho && (rc = download_field(ctx, rho, ctx->macro, 1, il0, ncols))) return rc;
		if (u && (rc = download_field(ctx, u, ctx->macro + L.S, 2, il0, ncols))) return rc;
	}
	if (force_ibm) {
		if (ctx->fibm) { if ((rc = download_field(ctx, force_ibm, ctx->fibm, 2, il0, ncols))) return rc; }
		else memset(force_ibm, 0, sizeof(double) * 2 * (size_t)(ncols * L.Ny));
	}
	return life_sync(ctx);
}

int life_download_macro(life_ctx *ctx, double *rho, double *u) {
	if (!ctx) return LIFE_E_ARG;
	return life_download_columns(ctx, 0, ctx->L.nxl, nullptr, rho, u, nullptr);
}

int life_download_state(life_ctx *ctx, double *f, double *rho, double *u, double *force_ibm) {
	if (!ctx) return LIFE_E_ARG;
	return life_download_columns(ctx, 0, ctx->L.nxl, f, rho, u, force_ibm);
}

int life_max_speed(life_ctx *ctx, double *vmax, int32_t *has_nan, int64_t *nan_i, int64_t *nan_j) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_max_speed: no state uploaded");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	double v = 0.0;
	int32_t hn = 0;
	int64_t id = -1;
	int rc = launch_max_speed(ctx, &v, &hn, &id);
	if (rc) return rc;
	if (vmax) *vmax = v;
	if (has_nan) *has_nan = hn;
	if (nan_i) *nan_i = hn ? id / ctx->cfg.Ny : -1;
	if (nan_j) *nan_j = hn ? id % ctx->cfg.Ny : -1;
	return LIFE_OK;
}

int life_step(life_ctx *ctx, int32_t t) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_step: no state uploaded");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const Layout &L = ctx->L;
	const StepScalars sc = step_scalars(ctx, t);
	int rc;
	// cfg.exact selects the kernels compiled in the reference's operation order without FMA contraction (namespace life::exact)
	const bool exact = ctx->cfg.exact != 0;
	auto bulk = exact ? launch_bulk_exact : launch_bulk;
	// cfg.inplace: what the boundary kernel needs of the pre-sweep state is saved first; after the sweep the layout offsets advance
	// (that is the streaming), and ring / halo / boundary work on the advanced layout of the same buffer
	if ((ctx->inplace || ctx->wom_field) && (rc = exact ? launch_bc_capture_exact(ctx, sc) : launch_bc_capture(ctx, sc))) return rc;
	auto advance = [&]() {
		if (!ctx->inplace) return;
		static const int cx[9] = {0, 1, -1, 0, 0, 1, -1, 1, -1}, cy[9] = {0, 0, 0, 1, -1, 1, -1, -1, 1};
		for (int v = 0; v < 9; v++) {
			int64_t o = (ctx->shift.off[v] + cx[v] * L.P + cy[v]) % L.S;
			ctx->shift.off[v] = o < 0 ? o + L.S : o;
		}
	};
	if ((rc = exact ? launch_convective_speed_exact(ctx, sc) : launch_convective_speed(ctx, sc))) return rc;

	cudaEvent_t p0 = nullptr, p1 = nullptr;
	if (ctx->profiling) {
		if (ctx->prof_used == ctx->prof_events.size()) {
			cudaEvent_t a, b;
			LIFE_CUDA(ctx, cudaEventCreate(&a));
			LIFE_CUDA(ctx, cudaEventCreate(&b));
			ctx->prof_events.emplace_back(a, b);
		}
		p0 = ctx->prof_events[ctx->prof_used].first;
		p1 = ctx->prof_events[ctx->prof_used].second;
		ctx->prof_used++;
	}

	if (ctx->cfg.nranks <= 1) {
		if (p0) LIFE_CUDA(ctx, cudaEventRecord(p0, ctx->stream));
		if ((rc = bulk(ctx, sc, 1, L.nxl, ctx->stream))) return rc;
		if (p1) LIFE_CUDA(ctx, cudaEventRecord(p1, ctx->stream));
		advance();
		if ((rc = launch_wrap_y(ctx, ctx->stream, false))) return rc;
		if ((rc = exchange_x(ctx))) return rc;
	} else {
		// the two edge columns first, so their ghost columns can travel while the interior is swept
		// (cfg.inplace: the sweeps address the buffer through the layout before the step, ring / halo / boundary through the advanced
		// one; kernel arguments are taken at launch, so ctx->shift is switched back and forth around the launches)
		const PopShift before = ctx->shift;
		advance();
		const PopShift after = ctx->shift;
		ctx->shift = before;
		if ((rc = bulk(ctx, sc, 1, 1, ctx->stream))) return rc;
		if ((rc = bulk(ctx, sc, L.nxl, 1, ctx->stream))) return rc;
		ctx->shift = after;
		if ((rc = launch_wrap_y(ctx, ctx->stream, false))) return rc;   // ring corners of the ghost columns
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_edge, ctx->stream));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_edge, 0));
		if ((rc = exchange_x(ctx))) return rc;
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
		if (p0) LIFE_CUDA(ctx, cudaEventRecord(p0, ctx->stream));
		ctx->shift = before;
		if ((rc = bulk(ctx, sc, 2, L.nxl - 2, ctx->stream))) return rc;
		ctx->shift = after;
		if (p1) LIFE_CUDA(ctx, cudaEventRecord(p1, ctx->stream));
		if ((rc = launch_wrap_y(ctx, ctx->stream, true))) return rc;    // what the interior sweep pushed over the top/bottom
		LIFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
	}
	if ((rc = exact ? launch_boundary_exact(ctx, sc) : launch_boundary(ctx, sc))) return rc;

	if (!ctx->inplace) { double *tmp = ctx->fA; ctx->fA = ctx->fB; ctx->fB = tmp; }
	ctx->stored_macro_valid = false;
	ctx->fibm_consumed = true;
	ctx->fxy_uniform[0] = sc.fxy_cur[0];
	ctx->fxy_uniform[1] = sc.fxy_cur[1];
	ctx->last_t = t;
	return LIFE_OK;
}

// lattices up to this many nodes step faster inside one 8-SM cluster launch than through launch-bound per-step kernels (lbm_small.cu)
static constexpr int64_t SMALL_MAX_NODES = 16384;
static constexpr int32_t SMALL_MAX_STEPS = 1024;   // per launch

static bool small_path(const life_ctx *ctx) {
	if (ctx->cfg.tune == 30) return false;          // measurement: per-step launches only
	return ctx->cfg.nranks <= 1 && !ctx->inplace && !ctx->fibm_any && ctx->fxy_mode != life::FXY_FIELD && !ctx->stored_macro_valid && !ctx->profiling &&
	       ctx->L.nxl * ctx->L.Ny <= SMALL_MAX_NODES;
}

int life_step_n(life_ctx *ctx, int32_t t_first, int32_t n) {
	if (!ctx) return LIFE_E_ARG;
	int32_t k = 0;
	while (k < n) {
		if (n - k < 2 || !small_path(ctx)) {       // (the first step after an upload carries stored macroscopics: per-step path)
			int rc = life_step(ctx, t_first + k);
			if (rc) return rc;
			k++;
			continue;
		}
		// the remaining steps, up to SMALL_MAX_STEPS at a time, in one launch
		if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_step_n: no state uploaded");
		LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
		const int32_t m = n - k < SMALL_MAX_STEPS ? n - k : SMALL_MAX_STEPS;
		if (!ctx->d_steps) {
			LIFE_CUDA(ctx, cudaMalloc(&ctx->d_steps, sizeof(StepScalars) * SMALL_MAX_STEPS));
			LIFE_CUDA(ctx, cudaMallocHost(&ctx->h_steps, sizeof(StepScalars) * SMALL_MAX_STEPS));
			LIFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_steps, cudaEventDisableTiming));
			ctx->steps_cap = SMALL_MAX_STEPS;
		} else {
			LIFE_CUDA(ctx, cudaEventSynchronize(ctx->ev_steps));     // the previous batch's scalars have left the pinned buffer
		}
		for (int32_t q = 0; q < m; q++) {
			// the uniform force a step sees as "previous" is the one the step before it applied (Womersley: it changes every step)
			ctx->h_steps[q] = step_scalars(ctx, t_first + k + q);
			ctx->fxy_uniform[0] = ctx->h_steps[q].fxy_cur[0];
			ctx->fxy_uniform[1] = ctx->h_steps[q].fxy_cur[1];
		}
		LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_steps, ctx->h_steps, sizeof(StepScalars) * m, cudaMemcpyHostToDevice, ctx->stream));
		LIFE_CUDA(ctx, cudaEventRecord(ctx->ev_steps, ctx->stream));
		int rc = ctx->cfg.exact ? launch_steps_small_exact(ctx, ctx->d_steps, ctx->h_steps[0], m) : launch_steps_small(ctx, ctx->d_steps, ctx->h_steps[0], m);
		if (rc) return rc;
		if (m & 1) { double *tmp = ctx->fA; ctx->fA = ctx->fB; ctx->fB = tmp; }
		ctx->fibm_consumed = true;
		ctx->last_t = t_first + k + m - 1;
		k += m;
	}
	return LIFE_OK;
}

int life_sync(life_ctx *ctx) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
	if (ctx->comm) {
		ncclResult_t ar;
		LIFE_NCCL(ctx, ncclCommGetAsyncError(ctx->comm, &ar));
		if (ar != ncclSuccess) return fail(ctx, LIFE_E_NCCL, std::string("asynchronous NCCL error: ") + ncclGetErrorString(ar));
	}
	if (ctx->mk.n > 0) return ibm_check(ctx);   // marker errors latched on the device since the last check
	return LIFE_OK;
}

int64_t life_launch_count(const life_ctx *ctx) { return ctx ? ctx->launches : 0; }

int life_set_profiling(life_ctx *ctx, int32_t on) {
	if (!ctx) return LIFE_E_ARG;
	ctx->profiling = on != 0;
	ctx->prof_used = 0;
	return LIFE_OK;
}

int life_bulk_kernel_ms(life_ctx *ctx, double *avg_ms, int64_t *launches) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	double total = 0.0;
	for (size_t k = 0; k < ctx->prof_used; k++) {
		float ms = 0.f;
		LIFE_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_events[k].first, ctx->prof_events[k].second));
		total += ms;
	}
	if (avg_ms) *avg_ms = ctx->prof_used ? total / (double)ctx->prof_used : 0.0;
	if (launches) *launches = (int64_t)ctx->prof_used;
	ctx->prof_used = 0;
	return LIFE_OK;
}

int life_get_types(life_ctx *ctx, int32_t *type) {
	if (!ctx || !type) return LIFE_E_ARG;
	memcpy(type, ctx->h_type.data(), sizeof(int32_t) * ctx->h_type.size());
	return LIFE_OK;
}

int life_get_boundary(life_ctx *ctx, int64_t *n, int64_t *id, int32_t *type, int32_t *normal_x, int32_t *normal_y,
                      int32_t *normal_dir) {
	if (!ctx || !n) return LIFE_E_ARG;
	*n = ctx->n_bc;
	for (int64_t b = 0; b < ctx->n_bc; b++) {
		const BcNode &k = ctx->h_bc[(size_t)b];
		if (id) id[b] = (int64_t)k.il * ctx->L.Ny + k.j;
		if (type) type[b] = k.type;
		if (normal_x) normal_x[b] = k.nx;
		if (normal_y) normal_y[b] = k.ny;
		if (normal_dir) normal_dir[b] = k.nd;
	}
	return LIFE_OK;
}

int life_ibm_set_markers(life_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *ds, const double *epsilon) {
	if (!ctx) return LIFE_E_ARG;
	if (n < 0 || (n > 0 && (!pos || !vel || !ds || !epsilon))) return fail(ctx, LIFE_E_ARG, "life_ibm_set_markers: null array");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return ibm_set_markers(ctx, n, pos, vel, ds, epsilon);
}

int life_ibm_interp(life_ctx *ctx, double *force_out) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_ibm_interp: no state uploaded");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return ibm_interp(ctx, force_out);
}

int life_ibm_spread(life_ctx *ctx) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_ibm_spread: no state uploaded");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return ibm_spread(ctx);
}

int life_ibm_compute_epsilon(life_ctx *ctx, int64_t n_bodies, const int64_t *body_first, const int64_t *members, double *epsilon_out) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return ibm_compute_epsilon(ctx, n_bodies, body_first, members, epsilon_out);
}

int life_ibm_assemble_epsilon(life_ctx *ctx, int64_t n_bodies, const int64_t *body_first, const int64_t *members, double *A_out) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return ibm_assemble_epsilon(ctx, n_bodies, body_first, members, A_out);
}

int life_ibm_set_forces(life_ctx *ctx, const double *force) {
	if (!ctx || !force) return LIFE_E_ARG;
	if (ctx->mk.n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->mk.force, force, sizeof(double) * 2 * ctx->mk.n, cudaMemcpyHostToDevice, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_ibm_get_interp(life_ctx *ctx, double *interp_rho, double *interp_mom) {
	if (!ctx) return LIFE_E_ARG;
	if (ctx->mk.n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	if (interp_rho) LIFE_CUDA(ctx, cudaMemcpyAsync(interp_rho, ctx->mk.irho, sizeof(double) * ctx->mk.n, cudaMemcpyDeviceToHost, ctx->stream));
	if (interp_mom) LIFE_CUDA(ctx, cudaMemcpyAsync(interp_mom, ctx->mk.imom, sizeof(double) * 2 * ctx->mk.n, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_ibm_get_supports(life_ctx *ctx, int32_t *count, int32_t *idx, int32_t *jdx, double *dirac) {
	if (!ctx) return LIFE_E_ARG;
	const int64_t n = ctx->mk.n;
	if (n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	if (count) LIFE_CUDA(ctx, cudaMemcpyAsync(count, ctx->mk.scount, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (idx) LIFE_CUDA(ctx, cudaMemcpyAsync(idx, ctx->mk.sidx, sizeof(int32_t) * 9 * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (jdx) LIFE_CUDA(ctx, cudaMemcpyAsync(jdx, ctx->mk.sjdx, sizeof(int32_t) * 9 * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (dirac) LIFE_CUDA(ctx, cudaMemcpyAsync(dirac, ctx->mk.sdirac, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

}  // extern "C"
