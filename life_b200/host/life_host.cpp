// life_host.cpp — the host side of the drop-in: LIFE's own C++ program with the hot path routed to liblife_b200.
//
// LIFE (joconnor22/LIFE v1.0.3) has no plugin interface; the seam is a set of member-function bodies (SURVEY.md §8b,
// include/life_b200.h).  This file DEFINES those member functions for the reference's own classes, compiled against the
// reference's own headers (inc/Grid.h, inc/Objects.h, the case's params.h), and forwards each one to the C ABI:
//
//   GridClass::lbmKernel()            (src/Grid.cpp:36-100)     -> life_step
//   ObjectsClass::ibmKernelInterp()   (src/Objects.cpp:102-117) -> life_ibm_set_markers + life_ibm_interp
//   ObjectsClass::ibmKernelSpread()   (src/Objects.cpp:120-149) -> life_ibm_spread
//   ObjectsClass::computeEpsilon()    (src/Objects.cpp:235-321), selected by LIFE_B200_DEVICE_EPSILON:
//                                     =1 (default) -> life_ibm_assemble_epsilon: the O(n^2 * 9) delta evaluations of the matrix on the
//                                     GPU (bit-exact), then the reference's own Utils::solveLAPACK on the host, as the north star
//                                     prescribes for the solve: epsilon is bit-identical to the reference's;
//                                     =0 -> the reference's own assembly + LAPACK solve, untouched;
//                                     =2 -> life_ibm_compute_epsilon (assembly and LU both on the GPU);
//                                     =3 -> per body: GPU LU for small systems (<= 64 markers, many of them: Honami), GPU
//                                     assembly + host LAPACK for large ones (UNI_EPSILON: TurekHron 132, PELskin 310)
//   ObjectsClass::recomputeObjectVals / femKernel (src/Objects.cpp:152-232, :63-98), selected with
//                                     LIFE_B200_DEVICE_FEM=1 (off by default: the north star keeps the FEM host-side, and the device
//                                     solver's own LU is within rounding of LAPACK, not bit-identical):
//                                     predictor / relaxed update / dynamicFEM of all flexible bodies on the device
//                                     (life_fem_predict / _relax / _dynamic), one CTA per filament; the host's FEM state and marker
//                                     positions are refreshed from the device after each call, so every host writer and the
//                                     support / ds / epsilon code keep working unchanged
//   GridClass::writeInfo / writeVTK / writeRestart / readRestart (src/Grid.cpp:559, :790, :1163, :1072), SURVEY.md §8f row 2:
//                                     the device-fed file paths — life_max_speed for the scan of writeInfo, life_write_vtk and
//                                     life_write_restart (asynchronous: the time loop goes on while the file is written) produce
//                                     the reference's files byte for byte from the device state, life_read_restart streams
//                                     Fluid.restart straight into it.  The 228 B/node host mirrors are then never refreshed.
//                                     LIFE_B200_HOST_IO=1 selects the round-trip instead: refresh the host mirrors
//                                     (life_download_macro / life_download_state), then run the reference's own writer / reader
//
// Everything else — main(), params.h / geometry.config, geometryReadIn, the IBMBodyClass constructors, host findSupport /
// computeDs / computeEpsilon (LAPACK), the corotational FEM + Newmark solve, the Aitken-relaxed sub-iteration loop,
// VTK / log / force / tip output and the restart files — is the UNMODIFIED reference, compiled from its sources where they lie.
//
// Two ways to put these definitions in front of the reference's:
//   (a) a maintainer deletes the three bodies above from Grid.cpp / Objects.cpp and adds this file to the makefile
//       (INTEGRATION.md shows the patch), or
//   (b) with no source change at all: build the reference sources as a shared object and link this file into the
//       executable — the dynamic linker resolves GridClass::lbmKernel etc. to the executable's definitions first
//       (ELF symbol interposition; the reference's functions are ordinary default-visibility symbols).
//       life_b200/host/Makefile does (b); the reference's own bodies stay reachable through dlsym(RTLD_NEXT, ...), which
//       is how the I/O wrappers below call the original writers.
//
// There is no CPU fallback: if liblife_b200 cannot create its context (no B200), the program exits through the reference's
// ERROR() convention (inc/Utils.h:72-77: message + exit(99)).
#include "Grid.h"
#include "Objects.h"
#include "Utils.h"
#include "FEMBody.h"
#include "life_b200.h"
#include "rank_team.h"
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace {

RankTeam team;

struct DeviceSide {
	std::vector<life_ctx *> ctxs;        // one per rank = per GPU (LIFE_B200_GPUS, default 1); x-slabs of the lattice
	std::vector<int64_t> i0, i1;         // columns [i0, i1) of each rank
	life_ctx *ctx = nullptr;             // rank 0 (what single-context calls use)
	int nranks = 1;
	bool uploaded = false;       // the device holds a state (life_upload_state or life_read_restart has run)
	bool macro_stale = false;    // host rho / u are older than the device state
	bool full_stale = false;     // host f / force_ibm are older than the device state
	std::vector<double> pos, vel, ds, eps, force;   // marker staging (SoA)
	long steps = 0, interps = 0, spreads = 0, eps_solves = 0;
	std::vector<int64_t> grp_first, grp_members;
	std::vector<double> eps_mat;
	// wall-clock accounting (seconds spent inside each replaced body; the rest of the program is the reference's host code)
	double t_step = 0, t_interp = 0, t_spread = 0, t_eps = 0, t_io = 0, t_first = 0, t_begin = 0;
	double t_info = 0, t_vtk = 0, t_restart = 0;   // parts of t_io
	double t_fem = 0;   // inside the optional device FEM bindings
	long fem_calls = 0;
	int deferred = 0, deferred_first = 0;   // body-free cases: time steps not yet handed to the library (flush_steps)
	double t_object = 0;   // inside ObjectsClass::objectKernel (whichever implementation runs it)
	long sub_its = 0;
} dev;
// the same C-ABI call on every rank's context, concurrently
#define ON_ALL_RANKS(r, ...) team.run([&](int r) { __VA_ARGS__; })

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Timed {
	double &acc, t0;
	explicit Timed(double &a) : acc(a), t0(now()) {}
	~Timed() { acc += now() - t0; }
};

[[noreturn]] void die(const char *where, int rc) {
	team.abandon = true;      // the reference's ERROR() is exit(99): possibly from a rank's thread, with the other ranks mid-call
	std::string msg;
	for (life_ctx *c : dev.ctxs) { const char *m = c ? life_last_error(c) : ""; if (m && *m) { msg = m; break; } }
	ERROR(std::string("liblife_b200: ") + where + " failed (" + std::to_string(rc) + "): " + msg);
	std::abort();
}
#define LIFE_CK(call)                      \
	do {                                   \
		int rc__ = (call);                 \
		if (rc__ != LIFE_OK) die(#call, rc__); \
	} while (0)

// marshal params.h (compile time) + the GridClass scalings into the run-time configuration of the C ABI
life_config make_config(const GridClass &g) {
	life_config c{};
	c.abi_version = LIFE_ABI_VERSION;
#ifdef CENTRAL_MOMENTS
	c.collision = LIFE_CENTRAL_MOMENTS;
#else
	c.collision = LIFE_BGK;
#endif
	c.Nx = Nx;
	c.Ny = Ny;
	c.omega = omega;
	c.wall_left = WALL_LEFT;       // eLatType values == LIFE_* values (inc/defs.h:52)
	c.wall_right = WALL_RIGHT;
	c.wall_bottom = WALL_BOTTOM;
	c.wall_top = WALL_TOP;
#ifdef INLET_RAMP
	c.inlet_ramp = INLET_RAMP;
#else
	c.inlet_ramp = -1.0;
#endif
	c.Dx = g.Dx;
	c.Dt = g.Dt;
	c.Dm = g.Dm;
	c.Drho = g.Drho;
#ifdef WOMERSLEY
	c.womersley = WOMERSLEY;
#else
	c.womersley = -1.0;
#endif
	c.height_p = height_p;
	c.nu_p = nu_p;
	c.gravity_x = gravityX;
	c.gravity_y = gravityY;
	c.dpdx = dpdx;
	c.dpdy = dpdy;
#ifdef ORDERED
	c.ordered = 1;
#else
	c.ordered = 0;
#endif
	c.device = -1;      // rank r of a multi-GPU run takes device r (ensure_context)
	c.stream = nullptr;
	c.rank = 0;
	c.nranks = 1;
	const char *k = std::getenv("LIFE_B200_KERNEL");
	c.kernel = k ? std::atoi(k) : LIFE_KERNEL_AUTO;
	// LIFE_B200_EXACT=1: the step in the reference's operation order (bitwise equal Results/, the reference's own `diff -r` protocol)
	const char *x = std::getenv("LIFE_B200_EXACT");
	c.exact = x ? std::atoi(x) : 0;
	// LIFE_B200_INPLACE=1: one population buffer (72 B/node resident), in-place sweep
	const char *ip = std::getenv("LIFE_B200_INPLACE");
	c.inplace = ip ? std::atoi(ip) : 0;
	return c;
}

void flush_steps();
void report() {
	if (!dev.ctx || team.abandon) return;
	flush_steps();
	ON_ALL_RANKS(r, life_sync(dev.ctxs[r]));
	bool io_failed = false;
	ON_ALL_RANKS(r, if (life_io_wait(dev.ctxs[r]) != LIFE_OK) io_failed = true);   // the last asynchronous file write (collective)
	if (io_failed) std::fprintf(stderr, "\n[life_b200] file output failed: %s", life_last_error(dev.ctx));
	const double wall = now() - dev.t_begin;
	std::fprintf(stderr, "\n[life_b200] wall %.3f s since the device context was created, of which context + first upload / restart read %.3f s; "
	                     "inside life_step %.3f s, interp %.3f s, spread %.3f s, epsilon %.3f s, output calls %.3f s (writeInfo %.3f, writeVTK %.3f, "
	                     "writeRestart %.3f); remaining host code %.3f s",
	             wall, dev.t_first, dev.t_step, dev.t_interp, dev.t_spread, dev.t_eps, dev.t_io, dev.t_info, dev.t_vtk, dev.t_restart,
	             wall - dev.t_first - dev.t_step - dev.t_interp - dev.t_spread - dev.t_eps - dev.t_io);
	if (dev.steps > 1)
		std::fprintf(stderr, "\n[life_b200] steady state %.1f us per time step = %.1f MLUPS (%ld x %ld lattice, %ld steps, all host work and output included)",
		             1e6 * (wall - dev.t_first) / (double)dev.steps, (double)Nx * Ny * dev.steps / (wall - dev.t_first) / 1e6,
		             (long)Nx, (long)Ny, dev.steps);
	if (dev.fem_calls) std::fprintf(stderr, "\n[life_b200] device FEM: %ld life_fem_dynamic calls, %.3f s inside the FEM bindings", dev.fem_calls, dev.t_fem);
	if (dev.sub_its) std::fprintf(stderr, "\n[life_b200] objectKernel: %.3f s, %ld sub-iterations = %.1f us per sub-iteration (interp / FEM / support / epsilon / spread, host and device)",
	                              dev.t_object, dev.sub_its, 1e6 * dev.t_object / (double)dev.sub_its);
	std::fprintf(stderr, "\n[life_b200] %ld life_step, %ld life_ibm_interp, %ld life_ibm_spread, %ld life_ibm_compute_epsilon, %lld kernel launches (rank 0), %d GPU(s)\n",
	             dev.steps, dev.interps, dev.spreads, dev.eps_solves, (long long)life_launch_count(dev.ctx), dev.nranks);
	ON_ALL_RANKS(r, life_destroy(dev.ctxs[r]));
	dev.ctxs.clear();
	dev.ctx = nullptr;
}

template <typename Fn>
Fn next_symbol(const char *mangled) {
	void *p = dlsym(RTLD_NEXT, mangled);
	if (!p) ERROR(std::string("life_host: reference symbol not found: ") + mangled);
	return reinterpret_cast<Fn>(p);
}

}  // namespace

namespace {

// LIFE_B200_HOST_IO=1: output / restart through the host mirrors and the reference's own writers instead of the device-fed paths
bool host_io() {
	static const bool on = [] { const char *e = std::getenv("LIFE_B200_HOST_IO"); return e && std::atoi(e) != 0; }();
	return on;
}

void ensure_context(const GridClass &g) {
	if (dev.ctx) return;
	const double t0 = now();
	const life_config c0 = make_config(g);
	// LIFE_B200_GPUS=N: the lattice is cut into N x-slabs, one per GPU of this box, one host thread each
	const char *ng = std::getenv("LIFE_B200_GPUS");
	dev.nranks = ng ? std::max(1, std::atoi(ng)) : 1;
	unsigned char nccl_id[128] = {0};
	if (dev.nranks > 1 && life_nccl_unique_id(nccl_id) != LIFE_OK)
		ERROR(std::string("liblife_b200: life_nccl_unique_id failed: ") + life_last_error(nullptr));
	dev.ctxs.assign((size_t)dev.nranks, nullptr);
	dev.i0.assign((size_t)dev.nranks, 0);
	dev.i1.assign((size_t)dev.nranks, 0);
	team.start(dev.nranks);
	std::vector<std::string> errs((size_t)dev.nranks);
	ON_ALL_RANKS(r, {
		life_config c = c0;
		if (dev.nranks > 1) { c.rank = r; c.nranks = dev.nranks; c.device = r; c.nccl_id = nccl_id; }
		if (life_create(&c, &dev.ctxs[(size_t)r]) != LIFE_OK) errs[(size_t)r] = life_last_error(nullptr);     // (thread-local message)
		else life_slab(dev.ctxs[(size_t)r], &dev.i0[(size_t)r], &dev.i1[(size_t)r]);
	});
	for (int r = 0; r < dev.nranks; r++)
		if (!dev.ctxs[(size_t)r]) ERROR(std::string("liblife_b200: life_create failed on rank ") + std::to_string(r) + ": " + errs[(size_t)r]);
	dev.ctx = dev.ctxs[0];
	dev.t_begin = t0;      // the wall clock of report() starts before the CUDA context is created
	std::atexit(report);
}

// The device takes over what initialiseGrid produced, at the first call that needs the state: the first step, or — fresh start,
// t = 0 — the first writeInfo / writeVTK of main() (src/main.cpp:61-64), so that even the initial files come from the device paths.
void ensure_state(GridClass &g) {
	if (dev.uploaded) return;
	const double t0 = now();
	ensure_context(g);
	// x is the slow index (src/Grid.cpp:70): a rank's slab is one contiguous chunk of every reference array
	ON_ALL_RANKS(r, {
		const size_t o = (size_t)dev.i0[(size_t)r] * Ny;
		LIFE_CK(life_upload_state(dev.ctxs[(size_t)r], g.f.data() + o * nVels, g.rho.data() + o, g.u.data() + o * dims, g.force_xy.data() + o * dims,
		                          g.force_ibm.data() + o * dims, g.u_in.data(), g.rho_in.data()));
	});
	dev.uploaded = true;
	dev.t_first += now() - t0;
}

}  // namespace

// ---- GridClass::lbmKernel -------------------------------------------------------------------------------------------------------
namespace {
// Without bodies nothing observes the lattice between output calls, so consecutive steps are handed over together: life_step_n runs
// them in ONE launch for small lattices (csrc/lbm_small.cu), instead of 2-4 launch-bound kernels per step.
void flush_steps() {
	if (dev.deferred == 0) return;
	Timed timed(dev.t_step);
	const int t0 = dev.deferred_first, n = dev.deferred;
	dev.deferred = 0;
	ON_ALL_RANKS(r, LIFE_CK(life_step_n(dev.ctxs[(size_t)r], t0, n)));
}
}  // namespace

void GridClass::lbmKernel() {
	ensure_state(*this);   // first step of a run whose output goes through the host mirrors, or of a restart read on the host
	dev.steps++;
	dev.macro_stale = dev.full_stale = true;
	if (!oPtr->hasIBM && dev.nranks == 1) {      // deferred: flushed by the next call that looks at the state (writeInfo / writeVTK / writeRestart / exit)
		if (dev.deferred == 0) dev.deferred_first = t;
		dev.deferred++;
		return;
	}
	Timed timed(dev.t_step);
	ON_ALL_RANKS(r, LIFE_CK(life_step(dev.ctxs[(size_t)r], t)));
}

namespace {
// iNode[i].pos / vel / ds / epsilon -> SoA staging -> device; the device rebuilds every marker's support from pos exactly as
// the host's findSupport did (bit-exact map)
void send_markers(std::vector<IBMNodeClass> &iNode) {
	const size_t n = iNode.size();
	dev.pos.resize(2 * n); dev.vel.resize(2 * n); dev.ds.resize(n); dev.eps.resize(n); dev.force.resize(2 * n);
	for (size_t i = 0; i < n; i++) {
		dev.pos[2 * i] = iNode[i].pos[eX]; dev.pos[2 * i + 1] = iNode[i].pos[eY];
		dev.vel[2 * i] = iNode[i].vel[eX]; dev.vel[2 * i + 1] = iNode[i].vel[eY];
		dev.ds[i] = iNode[i].ds;
		dev.eps[i] = iNode[i].epsilon;
	}
	// every rank holds all markers and gathers / spreads on the support sites it owns
	ON_ALL_RANKS(r, LIFE_CK(life_ibm_set_markers(dev.ctxs[(size_t)r], (int64_t)n, dev.pos.data(), dev.vel.data(), dev.ds.data(), dev.eps.data())));
}
}  // namespace

// ---- ObjectsClass::computeEpsilon (optional) ----------------------------------------------------------------------------------------
void ObjectsClass::computeEpsilon() {
	Timed timed(dev.t_eps);
	static const int on_device = [] { const char *e = std::getenv("LIFE_B200_DEVICE_EPSILON"); return e ? std::atoi(e) : 1; }();
	if (!on_device || !dev.uploaded) {
		// default, and always during construction (t = 0, before the first step creates the context): the reference's own
		// assembly + LAPACK solve
		using Fn = void (*)(ObjectsClass *);
		static Fn orig = next_symbol<Fn>("_ZN12ObjectsClass14computeEpsilonEv");
		orig(this);
		return;
	}
	// marker groups exactly as src/Objects.cpp:238-262 forms them: every marker in one body under UNI_EPSILON (that temporary
	// body counts as flexible), otherwise one group per flexible body (t > 0 here, so rigid bodies keep their epsilon)
	dev.grp_first.assign(1, 0);
	dev.grp_members.clear();
#ifdef UNI_EPSILON
	for (size_t i = 0; i < iNode.size(); i++) dev.grp_members.push_back((int64_t)i);
	dev.grp_first.push_back((int64_t)dev.grp_members.size());
#else
	for (size_t ib = 0; ib < iBody.size(); ib++) {
		if (iBody[ib].flex != eFlexible) continue;
		for (size_t k = 0; k < iBody[ib].node.size(); k++) dev.grp_members.push_back((int64_t)(iBody[ib].node[k] - &iNode[0]));
		dev.grp_first.push_back((int64_t)dev.grp_members.size());
	}
#endif
	send_markers(iNode);
	// split the groups by where their LU runs
	std::vector<int64_t> dfirst(1, 0), dmem, hfirst(1, 0), hmem;
	const int64_t nb = (int64_t)dev.grp_first.size() - 1;
	for (int64_t b = 0; b < nb; b++) {
		const int64_t lo = dev.grp_first[b], hi = dev.grp_first[b + 1];
		const bool lu_on_device = on_device == 2 || (on_device == 3 && hi - lo <= 64);
		std::vector<int64_t> &first = lu_on_device ? dfirst : hfirst, &mem = lu_on_device ? dmem : hmem;
		mem.insert(mem.end(), dev.grp_members.begin() + lo, dev.grp_members.begin() + hi);
		first.push_back((int64_t)mem.size());
	}
	if (dfirst.size() > 1) {
		// (the matrix only depends on the markers: every rank solves the same systems and updates its own copy of epsilon)
		ON_ALL_RANKS(r, LIFE_CK(life_ibm_compute_epsilon(dev.ctxs[(size_t)r], (int64_t)dfirst.size() - 1, dfirst.data(), dmem.data(), r == 0 ? dev.eps.data() : nullptr)));
		for (size_t k = 0; k < dmem.size(); k++) iNode[(size_t)dmem[k]].epsilon = dev.eps[(size_t)dmem[k]];
	}
	if (hfirst.size() > 1) {
		// matrices from the GPU, then exactly src/Objects.cpp:303-311: b = 1, Utils::solveLAPACK, scatter to the markers
		const int64_t nh = (int64_t)hfirst.size() - 1;
		std::vector<size_t> off((size_t)nh + 1, 0);
		for (int64_t b = 0; b < nh; b++) { const size_t d = (size_t)(hfirst[b + 1] - hfirst[b]); off[(size_t)b + 1] = off[(size_t)b] + d * d; }
		dev.eps_mat.resize(off[(size_t)nh]);
		LIFE_CK(life_ibm_assemble_epsilon(dev.ctx, nh, hfirst.data(), hmem.data(), dev.eps_mat.data()));
#pragma omp parallel for schedule(guided)
		for (int64_t b = 0; b < nh; b++) {
			const size_t d = (size_t)(hfirst[b + 1] - hfirst[b]);
			std::vector<double> A(dev.eps_mat.begin() + off[(size_t)b], dev.eps_mat.begin() + off[(size_t)b + 1]);
			std::vector<double> rhs(d, 1.0);
			const std::vector<double> sol = Utils::solveLAPACK(A, rhs);
			for (size_t i = 0; i < d; i++) iNode[(size_t)hmem[(size_t)hfirst[b] + i]].epsilon = sol[i];
		}
	}
	dev.eps_solves++;
}

// ---- ObjectsClass::recomputeObjectVals / femKernel (optional, LIFE_B200_DEVICE_FEM=1) ------------------------------------------------
namespace {

// LIFE_B200_DEVICE_FEM: 0 (default) the reference's host FEM; 1 the structural solver on the device, host state refreshed after every
// call (the reference's own support / ds / epsilon code keeps running on the host); 2 the whole sub-iteration loop RESIDENT on the
// device (ObjectsClass::objectKernel below): per sub-iteration only the residual sums come back, the host's body state is refreshed
// when a writer needs it
int device_fem_level() {
	static const int lvl = [] { const char *e = std::getenv("LIFE_B200_DEVICE_FEM"); return e ? std::atoi(e) : 0; }();
	return lvl;
}
bool device_fem() { return device_fem_level() != 0; }

struct DeviceFem {
	bool ready = false;
	bool host_stale = false;               // resident loop: the host's markers / FEM state are older than the device's
	std::vector<IBMBodyClass *> body;      // the flexible bodies, in iBody order (= the order femKernel adds their residuals)
	std::vector<double> state, pos, vel, per_body;
} dfem;

std::vector<double> *fem_vector(FEMBodyClass *s, int k) {   // the order of life_fem_set_state / life_fem_get_state
	std::vector<double> *v[11] = {&s->U, &s->Udot, &s->Udotdot, &s->U_n, &s->Udot_n, &s->Udotdot_n, &s->U_km1, &s->R_k, &s->R_km1, &s->U_nm1, &s->U_nm2};
	return v[k];
}

// describe the reference's flexible bodies to the library (once) and hand over their current state
void fem_setup(ObjectsClass &o) {
	if (dfem.ready) return;
	std::vector<life_fem_body> desc;
	std::vector<std::vector<double>> dbl;
	std::vector<std::vector<int32_t>> ints;
	for (size_t ib = 0; ib < o.iBody.size(); ib++)
		if (o.iBody[ib].flex == eFlexible) dfem.body.push_back(&o.iBody[ib]);
	dbl.reserve(dfem.body.size() * 6);
	ints.reserve(dfem.body.size() * 4);
	for (IBMBodyClass *b : dfem.body) {
		FEMBodyClass *s = b->sBody;
		const size_t n = s->node.size(), ne = s->element.size();
		std::vector<double> pos0(2 * n), angle0(n), el(5 * ne), pz(s->posMap.size()), z1, z2;
		std::vector<int32_t> mk(b->node.size()), pe(s->posMap.size()), first(ne + 1, 0), fm;
		for (size_t i = 0; i < n; i++) { pos0[2 * i] = s->node[i].pos0[eX]; pos0[2 * i + 1] = s->node[i].pos0[eY]; angle0[i] = s->node[i].angle0; }
		for (size_t e = 0; e < ne; e++) {
			const FEMElementClass &E = s->element[e];
			el[5 * e] = E.L0; el[5 * e + 1] = E.A; el[5 * e + 2] = E.I; el[5 * e + 3] = E.E; el[5 * e + 4] = E.rho;
			first[e] = (int32_t)fm.size();
			for (size_t k = 0; k < E.forceMap.size(); k++) { fm.push_back(E.forceMap[k].nodeID); z1.push_back(E.forceMap[k].zeta1); z2.push_back(E.forceMap[k].zeta2); }
		}
		first[ne] = (int32_t)fm.size();
		for (size_t i = 0; i < s->posMap.size(); i++) { pe[i] = s->posMap[i].elID; pz[i] = s->posMap[i].zeta; }
		for (size_t i = 0; i < b->node.size(); i++) mk[i] = (int32_t)(b->node[i] - &o.iNode[0]);
		life_fem_body d{};
		d.n_nodes = (int32_t)n; d.n_bc = s->bcDOFs; d.n_markers = (int32_t)b->node.size();
		d.alpha = alpha; d.delta = delta; d.gravity_x = gravityX; d.gravity_y = gravityY; d.ref_L = ref_L;
		dbl.push_back(std::move(pos0)); d.pos0 = dbl.back().data();
		dbl.push_back(std::move(angle0)); d.angle0 = dbl.back().data();
		dbl.push_back(std::move(el)); d.element = dbl.back().data();
		dbl.push_back(std::move(pz)); d.marker_zeta = dbl.back().data();
		dbl.push_back(std::move(z1)); d.map_zeta1 = dbl.back().data();
		dbl.push_back(std::move(z2)); d.map_zeta2 = dbl.back().data();
		ints.push_back(std::move(mk)); d.marker = ints.back().data();
		ints.push_back(std::move(pe)); d.marker_element = ints.back().data();
		ints.push_back(std::move(first)); d.map_first = ints.back().data();
		ints.push_back(std::move(fm)); d.map_marker = ints.back().data();
		desc.push_back(d);
	}
	send_markers(o.iNode);     // the marker arrays the solver reads and writes must exist on the device
	ON_ALL_RANKS(r, LIFE_CK(life_fem_create(dev.ctxs[(size_t)r], (int32_t)desc.size(), desc.data())));      // every rank holds every body, like the markers
	for (size_t k = 0; k < dfem.body.size(); k++) {
		FEMBodyClass *s = dfem.body[k]->sBody;
		dfem.state.resize(11 * (size_t)s->bodyDOFs);
		for (int v = 0; v < 11; v++) std::copy(fem_vector(s, v)->begin(), fem_vector(s, v)->end(), dfem.state.begin() + (size_t)v * s->bodyDOFs);
		ON_ALL_RANKS(r, LIFE_CK(life_fem_set_state(dev.ctxs[(size_t)r], (int32_t)k, dfem.state.data())));
	}
	dfem.ready = true;
}

// device -> host: FEM state vectors (and the geometry that follows from U), marker positions and velocities of the flexible bodies
// `geometry_from_U_km1`: after the predictor / relaxed update the displacement the geometry belongs to sits in U_km1 (the
// reference updates the geometry first and swaps U with U_km1 afterwards, src/FEMBody.cpp:279-288, src/Objects.cpp:199-208)
void fem_refresh_host(ObjectsClass &o, bool geometry_from_U_km1) {
	const size_t nm = o.iNode.size();
	dfem.pos.resize(2 * nm); dfem.vel.resize(2 * nm);
	LIFE_CK(life_ibm_get_markers(dev.ctx, dfem.pos.data(), dfem.vel.data()));
	for (size_t k = 0; k < dfem.body.size(); k++) {
		IBMBodyClass *b = dfem.body[k];
		FEMBodyClass *s = b->sBody;
		dfem.state.resize(11 * (size_t)s->bodyDOFs);
		LIFE_CK(life_fem_get_state(dev.ctx, (int32_t)k, dfem.state.data()));
		for (int v = 0; v < 11; v++) std::copy(dfem.state.begin() + (size_t)v * s->bodyDOFs, dfem.state.begin() + (size_t)(v + 1) * s->bodyDOFs, fem_vector(s, v)->begin());
		if (geometry_from_U_km1) s->U.swap(s->U_km1);
		s->updateFEMValues();
		if (geometry_from_U_km1) s->U.swap(s->U_km1);
		for (size_t i = 0; i < b->node.size(); i++) {
			const size_t g = (size_t)(b->node[i] - &o.iNode[0]);
			b->node[i]->pos[eX] = dfem.pos[2 * g]; b->node[i]->pos[eY] = dfem.pos[2 * g + 1];
			b->node[i]->vel[eX] = dfem.vel[2 * g]; b->node[i]->vel[eY] = dfem.vel[2 * g + 1];
		}
	}
}

}  // namespace

namespace {
// device -> host, everything the reference's writers read of the bodies (TotalForces.out: marker force, epsilon, ds; IBM.restart / body
// VTK: pos, vel, force; FEM.restart / tips: the state vectors and the geometry that follows from U)
void fem_refresh_all(ObjectsClass &o) {
	if (!dfem.host_stale) return;
	fem_refresh_host(o, false);
	const size_t n = o.iNode.size();
	std::vector<double> force(2 * n), ds(n), eps(n);
	LIFE_CK(life_ibm_get_marker_state(dev.ctx, force.data(), ds.data(), eps.data()));
	for (size_t i = 0; i < n; i++) {
		o.iNode[i].force[eX] = force[2 * i]; o.iNode[i].force[eY] = force[2 * i + 1];
		o.iNode[i].ds = ds[i];
		o.iNode[i].epsilon = eps[i];
	}
	dfem.host_stale = false;
}
}  // namespace

// ---- ObjectsClass::objectKernel (optional, LIFE_B200_DEVICE_FEM=2): the sub-iteration loop resident on the device ----------------------
void ObjectsClass::objectKernel() {
	Timed whole(dev.t_object);
	struct Count { int &it; ~Count() { dev.sub_its += it > 0 ? it : 1; } } count{subIt};
	if (device_fem_level() < 2 || !dev.uploaded || !hasFlex) {
		using Fn = void (*)(ObjectsClass *);
		static Fn orig = next_symbol<Fn>("_ZN12ObjectsClass12objectKernelEv");
		orig(this);
		return;
	}
	Timed timed(dev.t_fem);
	fem_setup(*this);
	// epsilon groups exactly as computeEpsilon forms them after t = 0 (src/Objects.cpp:238-262)
	if (dev.grp_first.size() < 2) {
		dev.grp_first.assign(1, 0);
		dev.grp_members.clear();
#ifdef UNI_EPSILON
		for (size_t i = 0; i < iNode.size(); i++) dev.grp_members.push_back((int64_t)i);
		dev.grp_first.push_back((int64_t)dev.grp_members.size());
#else
		for (size_t ib = 0; ib < iBody.size(); ib++) {
			if (iBody[ib].flex != eFlexible) continue;
			for (size_t k = 0; k < iBody[ib].node.size(); k++) dev.grp_members.push_back((int64_t)(iBody[ib].node[k] - &iNode[0]));
			dev.grp_first.push_back((int64_t)dev.grp_members.size());
		}
#endif
	}
	const int64_t nb = (int64_t)dev.grp_first.size() - 1;
	int64_t largest = 0;
	for (int64_t b = 0; b < nb; b++) largest = std::max(largest, dev.grp_first[b + 1] - dev.grp_first[b]);
	static const int lu_limit = [] { const char *e = std::getenv("LIFE_B200_DEVICE_LU_MAX"); return e ? std::atoi(e) : 96; }();       // larger systems: matrix from the device, the host's LAPACK (measured faster than the cluster LU of csrc/ibm_eps.cu at n = 132, equal at 310; LIFE_B200_DEVICE_LU_MAX=512 keeps them on the device)
	subIt = 0;
	const int MAXIT = 20;
	double sums[3] = {0, 0, 0};
	dfem.per_body.resize(5 * dfem.body.size());
	do {
		// the Aitken factor, src/Objects.cpp:178-188
		if (subIt == 1) relax = static_cast<double>(Utils::sgn(relax) * min(fabs(relax), relaxMax));
		else if (subIt > 1) relax = -relax * subNum / subDen;
		ON_ALL_RANKS(r, LIFE_CK(life_fsi_move(dev.ctxs[(size_t)r], gPtr->t, subIt, relax)));
		if (largest <= lu_limit) {
			ON_ALL_RANKS(r, LIFE_CK(life_ibm_compute_epsilon(dev.ctxs[(size_t)r], nb, dev.grp_first.data(), dev.grp_members.data(), nullptr)));
		} else {
			// large UNI_EPSILON systems: matrix from the device, the reference's own LAPACK solve on the host (src/Objects.cpp:303-311)
			std::vector<size_t> off((size_t)nb + 1, 0);
			for (int64_t b = 0; b < nb; b++) { const size_t d = (size_t)(dev.grp_first[b + 1] - dev.grp_first[b]); off[(size_t)b + 1] = off[(size_t)b] + d * d; }
			dev.eps_mat.resize(off[(size_t)nb]);
			dev.eps.resize(iNode.size());
			LIFE_CK(life_ibm_get_marker_state(dev.ctx, nullptr, nullptr, dev.eps.data()));
			LIFE_CK(life_ibm_assemble_epsilon(dev.ctx, nb, dev.grp_first.data(), dev.grp_members.data(), dev.eps_mat.data()));
			for (int64_t b = 0; b < nb; b++) {
				const size_t d = (size_t)(dev.grp_first[b + 1] - dev.grp_first[b]);
				std::vector<double> A(dev.eps_mat.begin() + off[(size_t)b], dev.eps_mat.begin() + off[(size_t)b + 1]), rhs(d, 1.0);
				const std::vector<double> sol = Utils::solveLAPACK(A, rhs);
				for (size_t i = 0; i < d; i++) dev.eps[(size_t)dev.grp_members[(size_t)dev.grp_first[b] + i]] = sol[i];
			}
			ON_ALL_RANKS(r, LIFE_CK(life_ibm_set_epsilon(dev.ctxs[(size_t)r], dev.eps.data())));
		}
		dev.eps_solves++;
		ON_ALL_RANKS(r, { double other[3]; LIFE_CK(life_fsi_force(dev.ctxs[(size_t)r], r == 0 ? sums : other, r == 0 ? dfem.per_body.data() : nullptr)); });
		dev.interps++;
		dev.fem_calls++;
		subRes = sqrt(sums[0]) / (ref_L * sqrt(static_cast<double>(simDOFs)));      // src/Objects.cpp:95-97
		subNum = sums[1];
		subDen = sums[2];
		subIt++;
	} while (subIt < MAXIT && subRes > subTol);
	for (size_t k = 0; k < dfem.body.size(); k++) {
		FEMBodyClass *s = dfem.body[k]->sBody;
		s->resNR = dfem.per_body[5 * k + 3]; s->itNR = (int)dfem.per_body[5 * k + 4];
	}
	ON_ALL_RANKS(r, LIFE_CK(life_ibm_spread(dev.ctxs[(size_t)r])));
	dev.spreads++;
	dev.macro_stale = dev.full_stale = true;
	dfem.host_stale = true;
	if (subIt == MAXIT) ERROR("Subiteration scheme hit " + to_string(subIt) + " iterations...exiting");
}

void ObjectsClass::recomputeObjectVals() {
	if (!device_fem() || !dev.uploaded) {
		using Fn = void (*)(ObjectsClass *);
		static Fn orig = next_symbol<Fn>("_ZN12ObjectsClass19recomputeObjectValsEv");
		orig(this);
		return;
	}
	{
		Timed timed(dev.t_fem);
		fem_setup(*this);
		if (subIt == 0) {
			ON_ALL_RANKS(r, LIFE_CK(life_fem_predict(dev.ctxs[(size_t)r], gPtr->t)));                    // resetValues + predictor, src/Objects.cpp:160-174
		} else {
			if (subIt == 1) relax = static_cast<double>(Utils::sgn(relax) * min(fabs(relax), relaxMax));   // src/Objects.cpp:183-188
			else relax = -relax * subNum / subDen;
			ON_ALL_RANKS(r, LIFE_CK(life_fem_relax(dev.ctxs[(size_t)r], relax)));                    // src/Objects.cpp:195-208
		}
		fem_refresh_host(*this, true);
	}
	// supports and ds of the moved markers, then epsilon: the reference's own host code (src/Objects.cpp:214-231)
#pragma omp parallel for schedule(guided)
	for (size_t i = 0; i < iNode.size(); i++) {
		if (iNode[i].iPtr->flex == eFlexible) {
			iNode[i].findSupport();
			iNode[i].computeDs();
		}
	}
	computeEpsilon();
}

void ObjectsClass::femKernel() {
	if (!device_fem() || !dev.uploaded) {
		using Fn = void (*)(ObjectsClass *);
		static Fn orig = next_symbol<Fn>("_ZN12ObjectsClass9femKernelEv");
		orig(this);
		return;
	}
	Timed timed(dev.t_fem);
	fem_setup(*this);
	// marker forces and epsilon are on the device already (the life_ibm_interp that precedes this call, src/Objects.cpp:40-47)
	double sums[3];
	dfem.per_body.resize(5 * dfem.body.size());
	ON_ALL_RANKS(r, { double other[3]; LIFE_CK(life_fem_dynamic(dev.ctxs[(size_t)r], r == 0 ? sums : other, r == 0 ? dfem.per_body.data() : nullptr)); });
	for (size_t k = 0; k < dfem.body.size(); k++) {
		FEMBodyClass *s = dfem.body[k]->sBody;
		s->subRes = dfem.per_body[5 * k]; s->subNum = dfem.per_body[5 * k + 1]; s->subDen = dfem.per_body[5 * k + 2];
		s->resNR = dfem.per_body[5 * k + 3]; s->itNR = (int)dfem.per_body[5 * k + 4];
	}
	fem_refresh_host(*this, false);
	subRes = sqrt(sums[0]) / (ref_L * sqrt(static_cast<double>(simDOFs)));      // src/Objects.cpp:95-97
	subNum = sums[1];
	subDen = sums[2];
	dev.fem_calls++;
}

// ---- ObjectsClass::ibmKernelInterp ----------------------------------------------------------------------------------------------
void ObjectsClass::ibmKernelInterp() {
	Timed timed(dev.t_interp);
	const size_t n = iNode.size();
	send_markers(iNode);
	// collective: partial sums of markers whose support straddles a slab face are added across ranks; every rank ends with all forces
	ON_ALL_RANKS(r, LIFE_CK(life_ibm_interp(dev.ctxs[(size_t)r], r == 0 ? dev.force.data() : nullptr)));
	for (size_t i = 0; i < n; i++) {
		iNode[i].force[eX] = dev.force[2 * i];      // consumed by the host FEM (src/FEMElement.cpp:53) and writeTotalForces
		iNode[i].force[eY] = dev.force[2 * i + 1];
	}
	dev.interps++;
	dev.full_stale = true;   // force_ibm was cleared
}

// ---- ObjectsClass::ibmKernelSpread ----------------------------------------------------------------------------------------------
void ObjectsClass::ibmKernelSpread() {
	// supports, ds, epsilon: those of the last ibmKernelInterp (the host recomputes them only in recomputeObjectVals, which
	// always precedes an interp); forces: those the last interp left on the device
	Timed timed(dev.t_spread);
	ON_ALL_RANKS(r, LIFE_CK(life_ibm_spread(dev.ctxs[(size_t)r])));
	dev.spreads++;
	dev.macro_stale = dev.full_stale = true;
}

// ---- output and restart: device-fed by default, through the host mirrors with LIFE_B200_HOST_IO=1 -------------------------------------
void GridClass::writeInfo() {
	flush_steps();
	if (oPtr && oPtr->hasFlex) fem_refresh_all(*oPtr);      // the resident sub-iteration loop leaves the host's body state behind

	if (!host_io()) ensure_state(*this);
	Timed timed(dev.t_io), part(dev.t_info);
	if (host_io()) {
		if (dev.uploaded && dev.macro_stale) {
			ON_ALL_RANKS(r, LIFE_CK(life_download_macro(dev.ctxs[(size_t)r], rho.data() + (size_t)dev.i0[(size_t)r] * Ny, u.data() + (size_t)dev.i0[(size_t)r] * Ny * dims)));
			dev.macro_stale = false;
		}
		using Fn = void (*)(GridClass *);
		static Fn orig = next_symbol<Fn>("_ZN9GridClass9writeInfoEv");
		orig(this);
		return;
	}
	// the scan of src/Grid.cpp:562-588 on the device, then the same report (src/Grid.cpp:590-616)
	double vmax = 0.0;
	int32_t blown = 0;
	int64_t bi = -1, bj = -1;
	ON_ALL_RANKS(r, {      // collective: every rank gets the global answer
		double v; int32_t b; int64_t x, y;
		LIFE_CK(life_max_speed(dev.ctxs[(size_t)r], &v, &b, &x, &y));
		if (r == 0) { vmax = v; blown = b; bi = x; bj = y; }
	});
	if (blown) {
#ifdef VTK
		Utils::writeVTK(*this);
		ON_ALL_RANKS(r, life_io_wait(dev.ctxs[(size_t)r]));
#endif
		ERROR("Simulation blew up (t = " + to_string(t) + ") at i = " + to_string(bi) + ", j = " + to_string(bj) + "...exiting");
	}
	loopTime = t == tOffset ? 0.0 : (omp_get_wtime() - startTime) / (t - tOffset);
	const array<int, 3> left = Utils::secs2hms(loopTime * (tOffset + nSteps - t));
	const double vphys = vmax * Dx / Dt;
	cout << endl << endl
	     << "Time step " << t << " of " << tOffset + nSteps << endl
	     << setprecision(4) << "Simulation has done " << t * Dt << " of " << (tOffset + nSteps) * Dt << " seconds" << endl
	     << "Time to finish = " << left[0] << " [h] " << left[1] << " [m] " << left[2] << " [s]" << endl
	     << setprecision(4) << "MLUPS = " << Nx * Ny / (1000000.0 * loopTime) << endl
	     << setprecision(5) << "Max Velocity = " << vmax << endl
	     << setprecision(5) << "Max Velocity (m/s) = " << vphys << endl
	     << setprecision(5) << "Max Reynolds number = " << (vphys)*ref_L / ref_nu;
}

void GridClass::writeVTK() {
	flush_steps();
	if (oPtr && oPtr->hasFlex) fem_refresh_all(*oPtr);      // the resident sub-iteration loop leaves the host's body state behind

	if (!(host_io() || bigEndian)) ensure_state(*this);
	Timed timed(dev.t_io), part(dev.t_vtk);
	if (host_io() || bigEndian) {
		if (dev.uploaded && dev.macro_stale) {
			ON_ALL_RANKS(r, LIFE_CK(life_download_macro(dev.ctxs[(size_t)r], rho.data() + (size_t)dev.i0[(size_t)r] * Ny, u.data() + (size_t)dev.i0[(size_t)r] * Ny * dims)));
			dev.macro_stale = false;
		}
		using Fn = void (*)(GridClass *);
		static Fn orig = next_symbol<Fn>("_ZN9GridClass8writeVTKEv");
		orig(this);
		return;
	}
	const string name = "Results/VTK/Fluid." + to_string(t) + ".vti";
	// every rank writes its own byte ranges of the one file (csrc/lbm_file.cu)
	ON_ALL_RANKS(r, LIFE_CK(life_write_vtk(dev.ctxs[(size_t)r], name.c_str(), rho_p, ref_P, LIFE_IO_ASYNC)));
	// placeholders for body files of an earlier run, as src/Grid.cpp:900-912
	if (!oPtr->hasIBM && boost::filesystem::exists("Results/VTK/IBM.0.vtp")) {
		string blank = "Results/VTK/IBM." + to_string(t) + ".vtp";
		oPtr->writeEmptyVTK(blank);
	}
#ifdef VTK_FEM
	if (!oPtr->hasFlex && boost::filesystem::exists("Results/VTK/FEM.0.vtp")) {
		string blank = "Results/VTK/FEM." + to_string(t) + ".vtp";
		oPtr->writeEmptyVTK(blank);
	}
#endif
}

void GridClass::writeRestart() {
	flush_steps();
	if (oPtr && oPtr->hasFlex) fem_refresh_all(*oPtr);      // the resident sub-iteration loop leaves the host's body state behind

	if (!(host_io() || bigEndian)) ensure_state(*this);
	Timed timed(dev.t_io), part(dev.t_restart);
	if (host_io() || bigEndian) {
		if (dev.uploaded && dev.full_stale) {
			ON_ALL_RANKS(r, {
				const size_t o = (size_t)dev.i0[(size_t)r] * Ny;
				LIFE_CK(life_download_state(dev.ctxs[(size_t)r], f.data() + o * nVels, rho.data() + o, u.data() + o * dims, force_ibm.data() + o * dims));
			});
			dev.full_stale = dev.macro_stale = false;
		}
		using Fn = void (*)(GridClass *);
		static Fn orig = next_symbol<Fn>("_ZN9GridClass12writeRestartEv");
		orig(this);
		return;
	}
	ON_ALL_RANKS(r, LIFE_CK(life_write_restart(dev.ctxs[(size_t)r], "Results/Restart/Fluid.restart", t, LIFE_IO_ASYNC)));
}

void GridClass::readRestart() {
	// force_xy is not in the file; initialiseGrid made it uniform (src/Grid.cpp:1035-1045), which is what the device path takes
	bool uniform = !force_xy.empty();
	for (size_t k = 2; k < force_xy.size() && uniform; k += 2) uniform = force_xy[k] == force_xy[0] && force_xy[k + 1] == force_xy[1];
	if (host_io() || bigEndian || !uniform) {
		using Fn = void (*)(GridClass *);
		static Fn orig = next_symbol<Fn>("_ZN9GridClass11readRestartEv");
		orig(this);
		return;
	}
	Timed timed(dev.t_first);
	ensure_context(*this);
	int32_t t_file = 0;
	int rc = LIFE_OK;
	ON_ALL_RANKS(r, {      // every rank reads its own slab of the file
		int32_t tf = 0;
		const int rr = life_read_restart(dev.ctxs[(size_t)r], "Results/Restart/Fluid.restart", force_xy.data(), u_in.data(), rho_in.data(), &tf);
		if (r == 0) { rc = rr; t_file = tf; } else if (rr != LIFE_OK) rc = rc == LIFE_OK ? rr : rc;
	});
	if (rc != LIFE_OK) die("life_read_restart", rc);   // the reference's own messages (src/Grid.cpp:1080, :1104, :1138)
	tOffset = t_file;
	t = tOffset;
	dev.uploaded = true;
	dev.macro_stale = dev.full_stale = true;   // the host mirrors keep their initial values and are not read again
}
