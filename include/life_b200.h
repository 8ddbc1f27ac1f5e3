/*
 * life_b200.h — C ABI of the B200-native LIFE hot path (D2Q9 lattice-Boltzmann step + immersed-boundary
 * interpolate / spread), implemented in hand-written sm_100a CUDA (life_b200/csrc/).
 *
 * The reference (joconnor22/LIFE v1.0.3) has no plugin/FFI layer: it is one statically linked C++ binary in which
 * GridClass exposes its arrays to ObjectsClass / IBMNodeClass through `friend` (inc/Grid.h:32-33).  The seam this
 * header cuts is therefore the set of member-function *bodies* on the hot path; every entry point below names the
 * reference function(s) whose body it replaces (file:line relative to the reference tree).  INTEGRATION.md shows
 * the patch a LIFE maintainer applies to call them.
 *
 * Conventions
 *   - plain C, no CUDA/torch types in any signature; every function returns 0 on success or a LIFE_E_* code and
 *     latches a message readable with life_last_error().  The reference's own convention is ERROR(msg) → print +
 *     exit(99) (inc/Utils.h:72-77); the host shim (life_b200/host) maps non-zero to exactly that.
 *   - all host arrays are in the REFERENCE layout: node id = i*Ny + j (x-major, y fastest; src/Grid.cpp:70),
 *     f[id*9 + v], u[id*2 + d], rho[id], D2Q9 numbering of src/Grid.cpp:1247-1250.  The library converts to its own
 *     device layout (SoA planes, padded column pitch, ghost row/column ring).
 *   - with nranks > 1 each rank (= one process, one GPU) owns the x-slab of columns [i_begin, i_end) returned by
 *     life_slab(); "the lattice" in every array argument below then means that slab (a contiguous chunk of the
 *     global reference array because x is the slow index).
 *   - caller owns every pointer passed in; the library never keeps a host pointer after the call returns.
 *   - one host thread per context, no re-entrancy.  Work is enqueued on the context's CUDA stream;
 *     functions that return data to the host synchronise that stream, the others return after enqueue.
 *   - there is NO CPU fallback: life_create fails with LIFE_E_CUDA if no sm_100 device is usable.
 */
#ifndef LIFE_B200_H
#define LIFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIFE_ABI_VERSION 2

/* lattice-site / wall types: values of eLatType, inc/defs.h:52 */
enum {
	LIFE_FLUID = 0,      /* eFluid      (a wall of this type is periodic) */
	LIFE_WALL = 1,       /* eWall       */
	LIFE_VELOCITY = 2,   /* eVelocity   */
	LIFE_FREESLIP = 3,   /* eFreeSlip   */
	LIFE_PRESSURE = 4,   /* ePressure   */
	LIFE_CONVECTIVE = 5  /* eConvective (right wall only, src/Grid.cpp:919-922) */
};

/* collision operator: BGK (src/Grid.cpp:237-244) or central moments (#define CENTRAL_MOMENTS, src/Grid.cpp:106-233) */
enum { LIFE_BGK = 0, LIFE_CENTRAL_MOMENTS = 1 };

/* error codes */
enum {
	LIFE_OK = 0,
	LIFE_E_ARG = 1,      /* bad argument / inconsistent configuration            */
	LIFE_E_CUDA = 2,     /* CUDA runtime error, or no usable sm_100 device       */
	LIFE_E_NCCL = 3,     /* NCCL error                                           */
	LIFE_E_STATE = 4,    /* call out of order (e.g. step before upload)          */
	LIFE_E_SUPPORT = 5,  /* a marker has more than 9 support sites (src/IBMNode.cpp:171-172) */
	LIFE_E_NOMEM = 6,
	LIFE_E_IO = 7        /* file could not be opened / read / written (device-fed file paths) */
};

/* kernel selection for the bulk stream+collide sweep (all produce identical results; bench.py compares them) */
enum {
	LIFE_KERNEL_AUTO = 0,
	LIFE_KERNEL_DIRECT = 1,   /* one node per thread, shifted scalar stores                                          */
	LIFE_KERNEL_SHUFFLE = 2,  /* two nodes per thread, 16-byte loads, y-moving populations re-aligned with warp shuffles */
	LIFE_KERNEL_TMA = 3,      /* persistent CTAs, column tiles moved by bulk async copies through shared memory      */
	LIFE_KERNEL_QUAD = 4      /* four nodes per thread, 32-byte LDG.256 / STG.256 (needs Ny % 128 == 0, else SHUFFLE) */
};

/*
 * Everything the kernels read from inc/params.h and from the GridClass constructor (src/Grid.cpp:1232-1289),
 * as run-time values.  Zero-initialise, then fill.
 */
typedef struct life_config {
	int32_t abi_version;       /* LIFE_ABI_VERSION */
	int32_t collision;         /* LIFE_BGK | LIFE_CENTRAL_MOMENTS                      (params.h:26)      */
	int64_t Nx, Ny;            /* GLOBAL lattice size; 64-bit (the reference's int overflows at 15447^2) (params.h:46-47) */
	double omega;              /* relaxation frequency                                  (params.h:83-94)   */
	int32_t wall_left, wall_right, wall_bottom, wall_top; /* LIFE_* site types          (params.h:65-68)   */
	double inlet_ramp;         /* INLET_RAMP in seconds, <= 0: no ramp                  (params.h:27)      */
	double Dx, Dt, Dm, Drho;   /* lattice scalings                                      (Grid.cpp:1257-1260) */
	double womersley;          /* WOMERSLEY number, <= 0: steady force_xy               (params.h:28)      */
	double height_p, nu_p;     /* used by the Womersley period only                     (Grid.cpp:59-60)   */
	double gravity_x, gravity_y, dpdx, dpdy;  /* used by the Womersley forcing only     (Grid.cpp:59-60)   */
	int32_t ordered;           /* ORDERED: spread sums each site's contributions in marker order, bit-repeatable (params.h:30) */
	int32_t device;            /* CUDA device ordinal; -1 = the calling thread's current device               */
	void *stream;              /* cudaStream_t to enqueue on; NULL = the library creates its own               */
	int32_t rank, nranks;      /* x-slab decomposition; nranks <= 1: single GPU                                */
	const void *nccl_id;       /* 128-byte ncclUniqueId shared by all ranks (life_nccl_unique_id); NULL iff nranks <= 1 */
	int32_t kernel;            /* LIFE_KERNEL_*                                                                 */
	int32_t tune;              /* measurement only: launch-shape / cache-hint variant of the bulk sweep, 0 = default (csrc/lbm_bulk.cu) */
	int32_t exact;             /* 1: the LBM step in the REFERENCE'S OPERATION ORDER without FMA contraction — BGK collision, macroscopic,
	                              regularised / convective boundaries, the outlet mean as a serial sum in j order (src/Grid.cpp:237-299,
	                              :387-495) — so that fields, marker forces and files equal the reference's g++ build bit for bit (the
	                              reference's own regression protocol is `diff -r`, testing/run-tests.sh:100).  Both collision operators
	                              (central moments: the nine expanded polynomials of src/Grid.cpp:143-223); ~3-4x the arithmetic of the
	                              default factored collisions. */
	int32_t inplace;           /* 1: ONE population buffer (72 B/node resident instead of 144): the sweep collides in place and streaming is
	                              implicit — population v of node n lives at plane element (n - t_steps * shift_v) mod S, shift_v = cx * pitch + cy,
	                              so "pushing" it to its neighbour is an offset update, not a data movement (csrc/lbm_bulk.cu: k_bulk_shift).
	                              Same results, same 144 B/node of traffic per step; every state-returning call still delivers the
	                              reference convention.  32768^2 then needs 77 GB and leaves room for the asynchronous restart snapshot. */
	int32_t reserved[4];
} life_config;

typedef struct life_ctx life_ctx;

/* ---- life cycle ------------------------------------------------------------------------------------------------- */

int life_abi_version(void);

/* Fill `out128` with a fresh ncclUniqueId (rank 0 calls this and broadcasts the bytes by any means). */
int life_nccl_unique_id(void *out128);

/* Allocation + lattice constants + site types / boundary list: replaces the GridClass constructor's allocation
 * (src/Grid.cpp:1267-1279) and the type / BCVec construction of initialiseGrid (src/Grid.cpp:925-951). */
int life_create(const life_config *cfg, life_ctx **out);
int life_destroy(life_ctx *ctx);

/* Message of the last failure on this context (ctx == NULL: last failure of life_create on this thread). */
const char *life_last_error(const life_ctx *ctx);

/* Columns [*i_begin, *i_end) of the global lattice owned by this context. */
int life_slab(const life_ctx *ctx, int64_t *i_begin, int64_t *i_end);

/* The same partition without a context (pure arithmetic, usable before any device exists): the balanced split of Nx
 * columns over nranks slabs, the first Nx % nranks slabs one column wider.  The host uses it to cut the global
 * reference arrays (contiguous in x, src/Grid.cpp:70) into the per-rank chunks life_upload_state expects. */
int life_slab_range(int64_t Nx, int32_t nranks, int32_t rank, int64_t *i_begin, int64_t *i_end);

/* ---- state in / out --------------------------------------------------------------------------------------------- */

/*
 * Receives what initialiseGrid (src/Grid.cpp:954-1058) or readRestart (src/Grid.cpp:1072-1160) produced on the host.
 *   f          [n*9]  post-stream populations (required)
 *   rho, u     [n], [n*2]  start-of-step macroscopics (u_n / rho_n of the first step).  Both NULL: derived from f
 *                     and the forces, which is what they equal after any completed step.
 *   force_xy   [n*2]  Cartesian body force, NULL = 0      (src/Grid.cpp:1035-1045)
 *   force_ibm  [n*2]  spread IBM force of the last step, NULL = 0 (restart only)
 *   u_in       [Ny*2] inlet velocity profile, NULL = 0    (src/Grid.cpp:954-996)
 *   rho_in     [Ny]   pressure-boundary density, NULL = 1 (src/Grid.cpp:1277-1278)
 * n = (i_end - i_begin) * Ny.
 */
int life_upload_state(life_ctx *ctx, const double *f, const double *rho, const double *u, const double *force_xy,
                      const double *force_ibm, const double *u_in, const double *rho_in);

/*
 * Streaming form of life_upload_state for lattices whose host image does not fit in host RAM at once, or is read from a
 * restart file in pieces (readRestart, src/Grid.cpp:1072-1160, reads node by node): the reference's 228 B/node of host
 * arrays are 59 GB at 16384^2.  life_upload_begin takes the per-lattice inputs and resets the state;
 * life_upload_columns hands over the same arrays as life_upload_state restricted to LOCAL columns [il0, il0 + ncols) of
 * this rank's slab (each column exactly once, any order; rho/u either for every range or for none);
 * life_upload_end completes it.  life_upload_state == begin + columns(0, all) + end.
 */
int life_upload_begin(life_ctx *ctx, const double *u_in, const double *rho_in);
int life_upload_columns(life_ctx *ctx, int64_t il0, int64_t ncols, const double *f, const double *rho, const double *u,
                        const double *force_xy, const double *force_ibm);
int life_upload_end(life_ctx *ctx);

/* Streaming form of life_download_state: local columns [il0, il0 + ncols); any pointer may be NULL. */
int life_download_columns(life_ctx *ctx, int64_t il0, int64_t ncols, double *f, double *rho, double *u, double *force_ibm);

/* rho [n], u [n*2] as GridClass::rho / GridClass::u hold them at the end of a step (either may be NULL).
 * Replaces the reads of writeInfo (src/Grid.cpp:559-588) and writeVTK (src/Grid.cpp:858-889). */
int life_download_macro(life_ctx *ctx, double *rho, double *u);

/* Full state for writeRestart (src/Grid.cpp:1192-1221): f [n*9] in the reference convention (post-stream,
 * pre-collision, at its own node), rho, u, force_ibm [n*2].  Any pointer may be NULL. */
int life_download_state(life_ctx *ctx, double *f, double *rho, double *u, double *force_ibm);

/* The scan of writeInfo (src/Grid.cpp:562-588): maximum of sqrt(ux^2+uy^2), and whether any node is NaN
 * (*has_nan, with the first such node in i-major order in *nan_i, *nan_j — global indices).  All-ranks collective
 * when nranks > 1 (every rank gets the global answer). */
int life_max_speed(life_ctx *ctx, double *vmax, int32_t *has_nan, int64_t *nan_i, int64_t *nan_j);

/* ---- the lattice-Boltzmann step ----------------------------------------------------------------------------------- */

/* One pass of GridClass::lbmKernel (src/Grid.cpp:36-100) at time step t (GridClass::t, 1-based after a fresh start):
 * convective speed (:477-495), [Womersley forcing :52-62], stream + collide (:103-246), macroscopic (:282-299),
 * boundary conditions (:302-474) with the inlet ramp at Dt*t (:548-556), and the slab halo exchange.  Asynchronous. */
int life_step(life_ctx *ctx, int32_t t);

/* n consecutive life_step calls for t_first, t_first+1, ... (only valid without bodies, since the reference runs
 * objectKernel between steps). */
int life_step_n(life_ctx *ctx, int32_t t_first, int32_t n);

/* Block until everything enqueued so far has finished; surfaces asynchronous CUDA/NCCL errors. */
int life_sync(life_ctx *ctx);

/* ---- immersed boundary ----------------------------------------------------------------------------------------------- */

/*
 * Marker state the host owns (written by the FEM solver src/FEMBody.cpp:192-193, computeDs src/IBMNode.cpp:182-204
 * and computeEpsilon src/Objects.cpp:235-321): pos, vel [n*2] in physical units, ds, epsilon [n].
 * ALL markers of the simulation on every rank.  The device recomputes every marker's support exactly as
 * IBMNodeClass::findSupport does (src/IBMNode.cpp:139-179, delta of inc/Utils.h:220-232).
 * May be called every sub-iteration.  Asynchronous (the arrays are copied out before it returns): if a marker collects more
 * than 9 sites, LIFE_E_SUPPORT is returned by the next synchronising call (life_ibm_interp, life_sync).
 */
int life_ibm_set_markers(life_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *ds,
                         const double *epsilon);

/* ObjectsClass::ibmKernelInterp (src/Objects.cpp:102-117): clears force_ibm, then per marker
 * IBMNodeClass::interpolate (src/IBMNode.cpp:26-48) + forceCalc (:51-58).  force_out [n*2], lattice units.
 * Synchronises (the host FEM consumes the forces, src/FEMElement.cpp:53). */
int life_ibm_interp(life_ctx *ctx, double *force_out);

/* ObjectsClass::ibmKernelSpread (src/Objects.cpp:120-149): clears force_ibm, IBMNodeClass::spread
 * (src/IBMNode.cpp:61-94) with the forces of the last life_ibm_interp (or life_ibm_set_forces) and the
 * supports / ds / epsilon of the last life_ibm_set_markers; updateMacroscopic (:97-136) is implicit (rho, u are
 * always evaluated from f and the current forces).  Asynchronous. */
int life_ibm_spread(life_ctx *ctx);

/*
 * OPTIONAL (the default integration keeps this on the host with LAPACK):
 * ObjectsClass::computeEpsilon (src/Objects.cpp:235-321) with Utils::solveLAPACK (src/Utils.cpp:288-311) on the device.
 * For each of n_bodies groups of markers — group b = members[body_first[b] .. body_first[b+1]), indices into the arrays of the
 * last life_ibm_set_markers; one group holding every marker reproduces UNI_EPSILON — assemble
 * A_ij = ds_j * sum_s delta_i(s) delta(x_j/Dx - site_s) over the supports of marker i (bit-exact) and solve A eps = 1 by LU with
 * partial pivoting (dgetrf/dgetrs('T') semantics; eps agrees with LAPACK to rounding x cond(A)).  Updates the device's epsilon
 * of the member markers (non-members keep theirs) and returns all n marker epsilons in epsilon_out (may be NULL).  Synchronises.
 */
int life_ibm_compute_epsilon(life_ctx *ctx, int64_t n_bodies, const int64_t *body_first, const int64_t *members,
                             double *epsilon_out);

/*
 * OPTIONAL, the data-parallel half of the above only: assemble the matrices (bit-exact) and return them, body after body,
 * A_out[ sum_{b'<b} dim_b'^2 + i*dim_b + j ] = A_ij of body b — exactly the row-major array computeEpsilon hands to
 * Utils::solveLAPACK (src/Objects.cpp:304-307).  The host keeps its own LAPACK solve, so epsilon is bit-identical to the
 * reference's while the O(dim^2 * 9) delta evaluations (serial per body on the CPU) run on the GPU.  Synchronises.
 */
int life_ibm_assemble_epsilon(life_ctx *ctx, int64_t n_bodies, const int64_t *body_first, const int64_t *members, double *A_out);

/* Overwrite the marker forces kept on the device (restart: src/Objects.cpp:1226-1257 stores them). */
int life_ibm_set_forces(life_ctx *ctx, const double *force);

/* Interpolated density / momentum of the last life_ibm_interp (IBMNodeClass::interpRho / interpMom). */
int life_ibm_get_interp(life_ctx *ctx, double *interp_rho, double *interp_mom);

/* ---- device-fed files ------------------------------------------------------------------------------ */
/*
 * The reference writes its fluid files from host arrays, one 8-byte ofstream::write per value (src/Grid.cpp:857-889,
 * :1192-1221).  These entry points produce the SAME BYTES straight from the device state — an on-device pack kernel lays the
 * values out in file order (the .vti blocks are j-major, i.e. a transpose of the lattice; the restart records are 120-byte
 * AoS), double-buffered pinned staging brings them to the host and pwrite() puts them at their final offsets — so the 228 B/node
 * host mirrors are never needed and, in LIFE_IO_ASYNC mode, the time loop keeps stepping while the file is written
 * (the state is frozen in a device snapshot first; if that does not fit in HBM the call quietly runs synchronously).
 * With nranks > 1 every rank writes its own byte ranges of the one shared file.
 */
enum { LIFE_IO_SYNC = 0, LIFE_IO_ASYNC = 1 };

/* Everything of Results/VTK/Fluid.<t>.vti that is not array data (src/Grid.cpp:796-855, :891-898): the XML head up to and
 * including the '_' that opens the raw appended block, and the tail that follows the last array.  Pure host arithmetic, no
 * device needed.  head/tail may be NULL to query the lengths. */
int life_vtk_frame(int64_t Nx, int64_t Ny, double Dx, char *head, int64_t head_cap, int64_t *head_len, char *tail,
                   int64_t tail_cap, int64_t *tail_len);

/* GridClass::writeVTK (src/Grid.cpp:790-898) for the current state: Density = rho*Drho, Pressure = ref_P + (rho - rho_p/Drho)
 * * SQ(c_s) * Dm / (Dx * SQ(Dt)), Velocity = (u * (Dx/Dt), 0), evaluated in the reference's operation order, written to
 * `path` byte for byte as the reference's writer would from the same rho / u.  rho_p, ref_P: inc/params.h:51,105. */
int life_write_vtk(life_ctx *ctx, const char *path, double rho_p, double ref_P, int32_t mode);

/* GridClass::writeRestart (src/Grid.cpp:1163-1229): header (t, Nx, Ny, omega, Dx, Dt, Dm), then per node i, j, rho, u,
 * force_ibm, f[9]; written to `path`.temp and renamed onto `path` when complete, as the reference does. */
int life_write_restart(life_ctx *ctx, const char *path, int32_t t, int32_t mode);

/* Completes the pending LIFE_IO_ASYNC write, if any, and returns its status.  Collective when nranks > 1.  Implied by the
 * next life_write_*, by life_read_restart and by life_destroy. */
int life_io_wait(life_ctx *ctx);

/* Non-blocking: *busy = 1 while the worker of a LIFE_IO_ASYNC write is still packing / copying / writing (this rank's part),
 * 0 once life_io_wait would return without waiting for it (or if nothing is pending). */
int life_io_busy(life_ctx *ctx, int32_t *busy);

/* Wall seconds and bytes of the last completed write (this rank's part), and whether it ran asynchronously. */
int life_io_stats(life_ctx *ctx, double *seconds, int64_t *bytes, int32_t *was_async);

/* Size in bytes of each of the two pinned host / device staging buffers the file paths stream through (default 32 MiB;
 * tests shrink it to force many chunks). */
int life_io_set_staging(life_ctx *ctx, int64_t bytes);

/* GridClass::readRestart (src/Grid.cpp:1072-1160) straight into the device state: checks the header against the
 * configuration and every record's (i, j) against its position, with the reference's own messages on mismatch, and is
 * otherwise life_upload_state(f, rho, u, force_xy, force_ibm, u_in, rho_in) with the file's contents.  force_xy (2 doubles,
 * NULL = 0) is the uniform body force initialiseGrid set (src/Grid.cpp:1035-1045; the file does not hold it).
 * *t_out receives the time step to continue from (GridClass::tOffset).  Each rank reads its own slab. */
int life_read_restart(life_ctx *ctx, const char *path, const double *force_xy, const double *u_in, const double *rho_in,
                      int32_t *t_out);

/* ---- structural solver of the flexible bodies -------------------------------------------------------- */
/*
 * OPTIONAL (the north star keeps the FEM host-side, and so does the host program by default).  The solver core
 * (csrc/fem_core.h) is checked on the CPU against the compiled reference — serially for its arithmetic, as real threads under
 * ThreadSanitizer for its barriers —, these entry points inside live runs of the compiled reference on a B200 for all four flexible
 * examples (tests/test_gpu_fem.py), and the host program binds them with LIFE_B200_DEVICE_FEM = 1 / 2 (life_host.cpp).  Its own LU
 * agrees with LAPACK to rounding, not bit for bit.
 *
 * One CTA per filament: FEMBodyClass::dynamicFEM (src/FEMBody.cpp:26-68: corotational 2-node beam elements, Newmark-beta,
 * Newton-Raphson over a dense LU), resetValues + predictor (:341-349, :259-289) and the Aitken-relaxed update
 * (src/Objects.cpp:195-208).  Marker forces / epsilon are read from, and marker positions / velocities written to, the device
 * arrays of the immersed-boundary path (life_ibm_set_markers / life_ibm_interp), through each body's marker list.
 */
typedef struct life_fem_body {
	int32_t n_nodes;              /* FEM nodes; elements = n_nodes - 1, DOFs = 3 * n_nodes (x, y, angle per node)              */
	int32_t n_bc;                 /* leading DOFs removed by the boundary condition (3 = clamped, 2 = supported)               */
	int32_t n_markers;            /* IBM markers of this body                                                                  */
	double alpha, delta;          /* Newmark parameters                                                   (params.h:76-77)     */
	double gravity_x, gravity_y;  /*                                                                      (params.h:57-58)     */
	double ref_L;                 /* reference length of the Newton-Raphson residual                      (params.h:103)       */
	const double *pos0;           /* [2 * n_nodes] initial node positions (physical units)                                     */
	const double *angle0;         /* [n_nodes]                                                                                 */
	const double *element;        /* [(n_nodes - 1) * 5] per element: L0, A, I, E, rho      (src/FEMElement.cpp:265-278)       */
	const int32_t *marker;        /* [n_markers] index of the body's k-th marker in the arrays of life_ibm_set_markers         */
	const int32_t *marker_element;/* [n_markers] posMap: element carrying the marker ...    (src/FEMBody.cpp:292-338)          */
	const double *marker_zeta;    /* [n_markers]         ... and its local coordinate in [-1, 1]                               */
	const int32_t *map_first;     /* [n_elements + 1] forceMap, CSR over elements: entries map_first[e] .. map_first[e+1]      */
	const int32_t *map_marker;    /*   body-local marker loading the element ...                                                */
	const double *map_zeta1, *map_zeta2;  /* ... over the local range [zeta1, zeta2]                                           */
} life_fem_body;

/* Builds the device image of n_bodies flexible bodies (replaces any previous set).  Dt and Dm come from the configuration. */
int life_fem_create(life_ctx *ctx, int32_t n_bodies, const life_fem_body *bodies);

/* The eleven state vectors of one body, [11 * 3 * n_nodes]: U, Udot, Udotdot, U_n, Udot_n, Udotdot_n, U_km1, R_k, R_km1, U_nm1, U_nm2. */
int life_fem_set_state(life_ctx *ctx, int32_t body, const double *state);
int life_fem_get_state(life_ctx *ctx, int32_t body, double *state);

/* Start of time step t, sub-iteration 0: resetValues + predictor for every body; markers move accordingly.  Asynchronous. */
int life_fem_predict(life_ctx *ctx, int32_t t);
/* Later sub-iterations: U := U_km1 + relax * (U - U_km1), velocities and markers updated.  Asynchronous. */
int life_fem_relax(life_ctx *ctx, double relax);
/* dynamicFEM of every body with the marker forces of the last life_ibm_interp.  sums [3] = subRes, subNum, subDen added over the
 * bodies in body order (ObjectsClass::femKernel, src/Objects.cpp:63-98); per_body [5 * n_bodies] = subRes, subNum, subDen, resNR,
 * itNR of each (may be NULL).  Synchronises. */
int life_fem_dynamic(life_ctx *ctx, double *sums, double *per_body);

/* Marker positions / velocities as the device holds them (after life_fem_*), [2 * n] each; either may be NULL. */
int life_ibm_get_markers(life_ctx *ctx, double *pos, double *vel);

/*
 * The sub-iteration loop of ObjectsClass::objectKernel (src/Objects.cpp:33-52) with the markers RESIDENT on the device: per
 * sub-iteration nothing crosses PCIe but the three residual sums (and, for large UNI_EPSILON systems whose LU stays on the host,
 * the epsilon matrix).
 *   life_fsi_move   recomputeObjectVals (src/Objects.cpp:152-232) without its last line: predictor at time step t (sub_it == 0) or
 *                   the Aitken-relaxed update with `relax` (sub_it >= 1) of every flexible body, then IBMNodeClass::findSupport
 *                   (src/IBMNode.cpp:139-179) of every marker and computeDs (:182-204, bit-exact) of the flexible bodies' markers,
 *                   all from the positions the device holds.  Asynchronous.
 *   (epsilon)       computeEpsilon is the caller's choice: life_ibm_compute_epsilon(..., NULL) keeps it on the device;
 *                   life_ibm_assemble_epsilon + the host's LAPACK + life_ibm_set_epsilon keeps the solve bit-identical to the reference.
 *   life_fsi_force  ibmKernelInterp (src/Objects.cpp:102-117; forces stay on the device) + femKernel (:63-98; = life_fem_dynamic):
 *                   sums [3] = subRes, subNum, subDen over the bodies in body order, per_body [5 * n_bodies] (may be NULL).  Synchronises.
 * life_ibm_spread closes the step as before.  life_ibm_get_markers / life_ibm_get_marker_state / life_fem_get_state bring the body
 * state back when a host writer needs it (TotalForces.out, IBM / FEM restart files, body VTK).
 */
int life_fsi_move(life_ctx *ctx, int32_t t, int32_t sub_it, double relax);
int life_fsi_force(life_ctx *ctx, double *sums, double *per_body);

/* Overwrite the epsilon of all n markers on the device (the host solved the systems life_ibm_assemble_epsilon returned). */
int life_ibm_set_epsilon(life_ctx *ctx, const double *epsilon);

/* Marker forces [2 * n] (lattice units), ds [n], epsilon [n] as the device holds them; any pointer may be NULL. */
int life_ibm_get_marker_state(life_ctx *ctx, double *force, double *ds, double *epsilon);

/* ---- test hooks (bit-exact integer-map checks)------------------------------------------------------------------------ */

/* Supports as IBMNodeClass::supp holds them: count [n]; idx, jdx, dirac [n*9] in the reference's i-outer/j-inner order. */
int life_ibm_get_supports(life_ctx *ctx, int32_t *count, int32_t *idx, int32_t *jdx, double *dirac);

/* Boundary list: *n entries (pass arrays sized 2*(Nx_local+Ny), or NULL to query *n).  id = i_local*Ny + j in BCVec
 * order (src/Grid.cpp:925-951); type = eLatType; normal_x/normal_y/normal_dir as getNormalVector returns them
 * (src/Grid.cpp:498-545). */
int life_get_boundary(life_ctx *ctx, int64_t *n, int64_t *id, int32_t *type, int32_t *normal_x, int32_t *normal_y,
                      int32_t *normal_dir);

/* Site type of every node [n] (GridClass::type). */
int life_get_types(life_ctx *ctx, int32_t *type);

/* ---- measurement hooks ---------------------------------------------------------------------------------------------------- */

/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t life_launch_count(const life_ctx *ctx);

/* Device milliseconds of the bulk stream+collide kernel, averaged over the life_step calls since the last call of this
 * function (CUDA events recorded around that kernel on its own stream); *launches receives the count. */
int life_bulk_kernel_ms(life_ctx *ctx, double *avg_ms, int64_t *launches);

/* Enable (1) / disable (0) the per-launch event timing read by life_bulk_kernel_ms. */
int life_set_profiling(life_ctx *ctx, int32_t on);

/* What the memory system of `device` (-1: current) delivers to plain streaming kernels, best of `iters` launches over `bytes`
 * (csrc/membw.cu): mode 0 read only, 1 write only, 2 copy with 16-byte LDG/STG, 3 copy with 32-byte LDG/STG, 4 copy through TMA
 * bulk copies (global -> shared -> global), 5 / 6 copy of 9 planes into 9 planes with 16- / 32-byte accesses (the sweep's address
 * pattern and launch shape without its arithmetic).
 * GB/s of bytes read + written.  The ceiling the sweep's achieved bandwidth is put next to; not on the product path. */
int life_membw(int32_t device, int32_t mode, int64_t bytes, int32_t iters, double *gbytes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LIFE_B200_H */
