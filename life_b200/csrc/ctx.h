// Context, device layout and internal kernel-launcher declarations of liblife_b200.
#pragma once
#include <cuda_runtime.h>
#include "nccl_dyn.h"
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/life_b200.h"

namespace life {

// ---------------------------------------------------------------------------------------------------------------------
// Device layout of one x-slab (DESIGN.md "Data layout in HBM").
//
// Every field is a set of SoA planes over the same padded 2-D geometry:
//   columns  c = 0 .. nxl+1   c = i_local + 1; c = 0 and c = nxl+1 are ghost columns
//   rows     r = 0 .. P-1     r = j + JOFF;    r = JOFF-1 and r = JOFF+Ny are ghost rows
//   element (c, r) of a plane sits at  c*P + r  doubles from the plane base; planes are S doubles apart.
// JOFF = 16 and P a multiple of 16 keep row j = 0 of every column on a 128-byte boundary.
// The ghost ring is what makes the bulk sweep branch-free: node (i,j) pushes population v to (c+cx, r+cy) with no
// modulo; wrap-around (periodic walls, src/Grid.cpp:229) and slab halos are applied afterwards to the ring only.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int JOFF = 16;

struct Layout {
	int64_t Ny;     // rows of the lattice
	int64_t nxl;    // columns owned by this rank
	int64_t P;      // column pitch in doubles
	int64_t S;      // plane stride in doubles = (nxl+2)*P
	__host__ __device__ int64_t at(int64_t c, int64_t r) const { return c * P + r; }
	__host__ __device__ int64_t node(int64_t il, int64_t j) const { return (il + 1) * P + (j + JOFF); }
};

// In-place ("shift") layout of the population planes (cfg.inplace): population v of logical element idx sits at plane element
// (idx - off[v]) mod S; off[v] grows by shift_v = cx*P + cy every step, which IS the streaming.  All zero = the plain layout, so
// every kernel that is not on the hot path of the two-buffer sweep addresses populations through at().
struct PopShift {
	int64_t off[9];
	__host__ __device__ __forceinline__ int64_t at(int v, int64_t idx, int64_t S) const {
		const int64_t i = idx - off[v];
		return v * S + (i < 0 ? i + S : i);
	}
};

// one boundary node (element of BCVec, src/Grid.cpp:947-949) with its normal precomputed (src/Grid.cpp:498-545)
struct BcNode {
	int32_t il;      // local column index
	int32_t j;
	int8_t type;     // eLatType
	int8_t nx, ny;   // normal vector
	int8_t nd;       // normal direction (0 = x, 1 = y)
	int32_t pad;
};

// how force_xy is represented
enum { FXY_NONE = 0, FXY_UNIFORM = 1, FXY_FIELD = 2 };

struct StepScalars {      // per-step host-computed scalars
	double ramp;          // getRampCoefficient (src/Grid.cpp:548-556)
	double fxy_prev[2];   // uniform force_xy in force while the state was produced (enters u_n)
	double fxy_cur[2];    // uniform force_xy of this step (collision forcing, new macroscopics)
	double wom_cos;       // cos(2 pi t Dt / T_w) of this step (Womersley, src/Grid.cpp:59-60)
};

struct MarkerBuffers {
	int64_t n = 0, cap = 0;
	double *in = nullptr;             // device: pos[2n] | vel[2n] | ds[n] | eps[n] in one array (one copy per update)
	double *pos = nullptr, *vel = nullptr, *ds = nullptr, *eps = nullptr;   // views into `in`: pos[2n] = x0 y0 x1 y1 ...
	double *force = nullptr, *irho = nullptr, *imom = nullptr;
	int32_t *scount = nullptr, *sidx = nullptr, *sjdx = nullptr;
	double *sdirac = nullptr;
	int32_t *next = nullptr;          // ordered spread: link of every (marker, site) entry in its site's list, [9 * cap]
	int32_t *err = nullptr;           // device flag: support overflow
	double *h_stage = nullptr;        // pinned host staging: 6*cap doubles up, 2*cap doubles down
	int64_t h_cap = 0;
	cudaEvent_t ev_stage = nullptr;   // completion of the last upload out of h_stage
	bool stage_busy = false;
};

struct IoState;   // lbm_file.cu
struct FemState;  // fem.cu

}  // namespace life

struct life_ctx {
	life_config cfg;
	int device = 0;
	cudaStream_t stream = nullptr;        // everything is enqueued here
	bool own_stream = false;
	cudaStream_t comm_stream = nullptr;   // halo exchange
	cudaEvent_t ev_edge = nullptr, ev_comm = nullptr;
	ncclComm_t comm = nullptr;
	int64_t i_begin = 0, i_end = 0;       // global columns owned
	life::Layout L{};

	double *fA = nullptr, *fB = nullptr;  // population buffers, 9 planes each; fA holds the current state (cfg.inplace: fB stays null)
	bool inplace = false;
	life::PopShift shift{};               // where the populations of fA sit (all zero unless cfg.inplace)
	double *halo_buf = nullptr;           // cfg.inplace, nranks > 1: contiguous send / receive staging of the two faces, 4 * 3 * Ny doubles
	double *bc_prev = nullptr;            // cfg.inplace: what the boundary kernel needs of the state BEFORE the sweep, 3 doubles per BCVec entry
	double *macro = nullptr;              // rho, ux, uy planes (3*S), lazily allocated
	double *fibm = nullptr;               // force_ibm planes (2*S), lazily allocated
	uint8_t *fibm_mask = nullptr;         // one byte per (column, 64-row span): 1 where a spread has written force_ibm.  Exact while
	                                      // fibm_full_dirty is false; lets the sweep skip the two force planes everywhere else
	int64_t mask_pitch = 0;               // bytes per column = ceil(P / 64)
	double *fxyf = nullptr;               // force_xy planes (2*S), only in FXY_FIELD mode
	int32_t *cell_head = nullptr;         // ordered spread: head of the entry list of every lattice site (S ints, -1 = empty)
	double *u_in = nullptr, *rho_in = nullptr, *delU = nullptr;   // [Ny*2], [Ny], [Ny*2]
	life::BcNode *bc = nullptr;
	int64_t n_bc = 0;
	std::vector<life::BcNode> h_bc;
	std::vector<int32_t> h_type;          // local type matrix (host copy)

	int fxy_mode = life::FXY_NONE;
	double fxy_uniform[2] = {0, 0};       // current uniform force_xy (what GridClass::force_xy holds right now)
	bool wom_field = false;               // Womersley with gravity: force_xy is a field recomputed every step
	bool have_state = false;
	// streaming upload in progress (life_upload_begin .. life_upload_end)
	bool uploading = false;
	int up_macro = -1;                    // -1 undecided, 0 rho/u derived from f, 1 rho/u uploaded
	int64_t up_cols = 0;                  // columns received so far
	bool up_fxy_seen = false, up_fxy_uniform = true;
	double up_fxy0[2] = {0, 0};
	std::vector<std::pair<int64_t, int64_t>> up_ranges;   // column ranges received while force_xy was still uniform
	bool stored_macro_valid = false;      // `macro` holds uploaded rho_n/u_n to be used by the next step
	bool fibm_any = false;                // force_ibm may be non-zero somewhere
	bool fibm_sites_dirty = false;        // non-zero only at the current support sites
	bool fibm_full_dirty = false;         // non-zero anywhere (uploaded)
	bool fibm_consumed = true;            // a life_step has used the current force_ibm (it may be cleared when the markers move)
	int32_t last_t = 0;

	life::MarkerBuffers mk;
	life::IoState *io = nullptr;          // staging, snapshot and worker of the device-fed file paths (lbm_file.cu)
	life::FemState *fem = nullptr;        // flexible bodies of the device structural solver (fem.cu)
	double *scratch = nullptr;            // device staging for upload / download: two halves, so the PCIe copy of one chunk overlaps
	size_t scratch_bytes = 0;             // the pack / unpack kernel of the other (copies on copy_stream, kernels on stream)
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_kernel[2] = {nullptr, nullptr};
	double *d_red = nullptr;              // reduction scratch (max speed etc.)
	life::StepScalars *d_steps = nullptr, *h_steps = nullptr;   // per-step scalars of a multi-step launch (lbm_small.cu): device + pinned host
	int32_t steps_cap = 0;
	cudaEvent_t ev_steps = nullptr;       // the last upload out of h_steps has completed
	void *eps_buf = nullptr;              // epsilon assembly / LU scratch (ibm_eps.cu)
	size_t eps_bytes = 0;
	void *h_pin = nullptr;                // small pinned host buffer
	size_t h_pin_bytes = 0;

	int64_t launches = 0;
	bool profiling = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
	size_t prof_used = 0;
	std::string err;
};

namespace life {

// error helpers -----------------------------------------------------------------------------------------------------------
int fail(life_ctx *ctx, int code, const std::string &msg);
#define LIFE_CUDA(ctx, call)                                                                                  \
	do {                                                                                                      \
		cudaError_t e__ = (call);                                                                             \
		if (e__ != cudaSuccess)                                                                               \
			return life::fail(ctx, LIFE_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
	} while (0)
#define LIFE_NCCL(ctx, call)                                                                                  \
	do {                                                                                                      \
		ncclResult_t r__ = (call);                                                                            \
		if (r__ != ncclSuccess)                                                                               \
			return life::fail(ctx, LIFE_E_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__));        \
	} while (0)

// launchers (each returns a LIFE_* code) ----------------------------------------------------------------------------------
// lbm_bulk.cu
int launch_bulk(life_ctx *ctx, const StepScalars &sc, int64_t c_first, int64_t c_count, cudaStream_t st);
int launch_bulk_exact(life_ctx *ctx, const StepScalars &sc, int64_t c_first, int64_t c_count, cudaStream_t st);   // cfg.exact
// lbm_small.cu: n steps of a small single-rank lattice in one launch of one thread-block cluster (d_sc: per-step scalars on the device)
int launch_steps_small(life_ctx *ctx, const StepScalars *d_sc, const StepScalars &first, int n);
int launch_steps_small_exact(life_ctx *ctx, const StepScalars *d_sc, const StepScalars &first, int n);
// lbm_boundary.cu
int build_boundary(life_ctx *ctx);
int launch_convective_speed(life_ctx *ctx, const StepScalars &sc);
int launch_wrap_y(life_ctx *ctx, cudaStream_t st, bool after_exchange);
int launch_boundary(life_ctx *ctx, const StepScalars &sc);
int launch_bc_capture(life_ctx *ctx, const StepScalars &sc);     // cfg.inplace: before the sweep
int launch_convective_speed_exact(life_ctx *ctx, const StepScalars &sc);   // cfg.exact: the same two in the reference's operation order
int launch_boundary_exact(life_ctx *ctx, const StepScalars &sc);
int launch_bc_capture_exact(life_ctx *ctx, const StepScalars &sc);
// halo.cu
int exchange_x(life_ctx *ctx);
// lbm_io.cu
int upload_field(life_ctx *ctx, const double *h, double *planes, int ncomp, int64_t il0, int64_t ncols);
int download_field(life_ctx *ctx, double *h, const double *planes, int ncomp, int64_t il0, int64_t ncols, const PopShift *ps = nullptr);
int fill_field(life_ctx *ctx, double *planes, int ncomp, int64_t il0, int64_t ncols, double v0, double v1);
int launch_macro(life_ctx *ctx, double *out_planes, int64_t il0, int64_t ncols);
int launch_max_speed(life_ctx *ctx, double *vmax, int32_t *has_nan, int64_t *nan_id);
int ensure_scratch(life_ctx *ctx, size_t bytes);
int ensure_macro(life_ctx *ctx);
int ensure_fibm(life_ctx *ctx);
// ibm.cu
int ibm_set_markers(life_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *ds, const double *eps);
int ibm_interp(life_ctx *ctx, double *force_out, bool no_sync = false);
int ibm_refresh_supports(life_ctx *ctx);
int ibm_spread(life_ctx *ctx);
int ibm_clear_force(life_ctx *ctx);
int ibm_check(life_ctx *ctx);
void ibm_free(life_ctx *ctx);
// ibm_eps.cu
int ibm_compute_epsilon(life_ctx *ctx, int64_t nb, const int64_t *first, const int64_t *members, double *eps_out);
int ibm_assemble_epsilon(life_ctx *ctx, int64_t nb, const int64_t *first, const int64_t *members, double *A_out);

// fem.cu
void fem_free(life_ctx *ctx);
// lbm_file.cu
int io_wait(life_ctx *ctx);
void io_free(life_ctx *ctx);

StepScalars step_scalars(life_ctx *ctx, int32_t t);

}  // namespace life
