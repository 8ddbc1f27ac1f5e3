// The structural solver of the flexible bodies on the device: one CTA per filament running fem_core.h.
//
// STATUS: its logic and barrier placement are checked on the CPU (tests/test_fem_core.py: the same source run serially and as real
// threads under ThreadSanitizer against the compiled reference), the device entry points inside live runs of the compiled reference
// on a B200 (tests/test_gpu_fem.py, all four flexible examples), and the host program binds it on request (life_host.cpp,
// LIFE_B200_DEVICE_FEM = 1: solver on the device; = 2: the whole sub-iteration loop resident on the device, life_fsi_move /
// life_fsi_force below).  The default keeps the reference's host FEM: the north star leaves it host-side, and for ONE filament the
// host is faster (profiles/r02_fsi_timing_programs.txt); with many filaments (Honami, 128) the resident loop is 3x faster.
//
// Data: every body's constant description, geometry, dense M / K (dim^2 doubles each; 63 x 63 = 31 KB: L2-resident) and state
// vectors live in one device arena; `Body` views (pointers into it) sit in a device array indexed by blockIdx.x.  The marker
// forces / epsilon it reads and the marker positions / velocities it writes are the arrays of the immersed-boundary path
// (ctx->mk), addressed through each body's marker list — so between life_ibm_interp and the next support search nothing has to
// cross PCIe except the three residual sums of the Aitken loop.
#include "ctx.h"
#include "fem_core.h"
#include <cstring>

namespace life {

using life_fem::Body;
using life_fem::Lane;

struct FemState {
	int n_bodies = 0;
	std::vector<Body> h_bodies;          // host copies of the views (device pointers inside)
	std::vector<int> marker_first;       // CSR over bodies into marker ids
	int max_dof = 0;                     // largest n_dof of any body: sizes the shared-memory work space of k_fem<FEM_DYNAMIC>
	Body *d_bodies = nullptr;
	double *d_arena = nullptr;
	int *d_ints = nullptr;
	int *d_marker_first = nullptr, *d_marker_ids = nullptr;
	double *d_results = nullptr;         // per body: subRes, subNum, subDen, resNR, itNR
	int64_t n_markers_needed = 0;        // largest marker index + 1 referenced by any body
};

enum { FEM_DYNAMIC = 0, FEM_PREDICT = 1, FEM_RELAX = 2 };

constexpr int FEM_THREADS = 128;

template <int OP>
__global__ void __launch_bounds__(FEM_THREADS) k_fem(const Body *__restrict__ bodies, const int *__restrict__ mfirst, const int *__restrict__ mids,
                                                     const double *__restrict__ force, const double *__restrict__ eps, double *pos, double *vel,
                                                     int t, double relax, double *results) {
	__shared__ Body b;
	extern __shared__ double fem_sm[];
	if (threadIdx.x == 0) {
		b = bodies[blockIdx.x];
		if (OP == FEM_DYNAMIC) {
			// The Newton-Raphson work space — dense M and K (dim^2 each), the vectors of the Newmark system, the pivots — lives in shared
			// memory for the duration of the call: the LU is a chain of ~5 barriers per column, and every one of them used to wait for
			// an L2 round trip.  Nothing in it survives the call (M, K are rebuilt by every iteration; R, F, delU are per call).
			const size_t d = (size_t)b.n_dof;
			double *p = fem_sm;
			b.M = p; p += d * d;
			b.K = p; p += d * d;
			b.R = p; p += d;
			b.F = p; p += d;
			b.delU = p; p += d;
			b.work = p; p += d;
			b.piv = reinterpret_cast<int *>(p);
		}
	}
	__syncthreads();
	const Lane l{(int)threadIdx.x, (int)blockDim.x};
	const int *marker = mids + mfirst[blockIdx.x];
	if (OP == FEM_DYNAMIC) {
		life_fem::fem_dynamic(b, l, force, eps, pos, vel, marker);
		if (threadIdx.x == 0) {
			double *r = results + 5 * blockIdx.x;
			r[0] = b.scal[1]; r[1] = b.scal[2]; r[2] = b.scal[3]; r[3] = b.scal[0]; r[4] = b.scal[4];
		}
	} else if (OP == FEM_PREDICT) {
		life_fem::fem_predict(b, l, t, pos, vel, marker);
	} else {
		life_fem::fem_relax(b, l, relax, pos, vel, marker);
	}
}

// IBMNodeClass::computeDs (src/IBMNode.cpp:182-204) of the markers of every flexible body, after the solver moved them
__global__ void __launch_bounds__(FEM_THREADS) k_fem_ds(const Body *__restrict__ bodies, const int *__restrict__ mfirst, const int *__restrict__ mids,
                                               const double *__restrict__ pos, double Dx, double *ds) {
	__shared__ Body b;
	if (threadIdx.x == 0) b = bodies[blockIdx.x];
	__syncthreads();
	life_fem::compute_ds(b, Lane{(int)threadIdx.x, (int)blockDim.x}, pos, Dx, ds, mids + mfirst[blockIdx.x]);
}

void fem_free(life_ctx *ctx) {
	FemState *f = ctx->fem;
	if (!f) return;
	cudaFree(f->d_bodies); cudaFree(f->d_arena); cudaFree(f->d_ints); cudaFree(f->d_marker_first); cudaFree(f->d_marker_ids);
	cudaFree(f->d_results);
	delete f;
	ctx->fem = nullptr;
}

static int fem_ready(life_ctx *ctx, const char *who) {
	if (!ctx->fem || ctx->fem->n_bodies == 0) return fail(ctx, LIFE_E_STATE, std::string(who) + ": call life_fem_create first");
	if (ctx->mk.n < ctx->fem->n_markers_needed)
		return fail(ctx, LIFE_E_STATE, std::string(who) + ": the bodies refer to markers that life_ibm_set_markers has not supplied");
	return LIFE_OK;
}

template <int OP>
static int fem_launch(life_ctx *ctx, int t, double relax) {
	FemState *f = ctx->fem;
	size_t smem = 0;
	if (OP == FEM_DYNAMIC) {
		const size_t d = (size_t)f->max_dof;
		smem = sizeof(double) * (2 * d * d + 4 * d) + sizeof(int) * d + 16;
		if (smem > 200 * 1024) return fail(ctx, LIFE_E_ARG, "life_fem_dynamic: a body has too many degrees of freedom for the shared-memory work space");
		LIFE_CUDA(ctx, cudaFuncSetAttribute(k_fem<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	}
	k_fem<OP><<<(unsigned)f->n_bodies, FEM_THREADS, smem, ctx->stream>>>(f->d_bodies, f->d_marker_first, f->d_marker_ids, ctx->mk.force, ctx->mk.eps,
	                                                                     ctx->mk.pos, ctx->mk.vel, t, relax, f->d_results);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

}  // namespace life

using namespace life;

extern "C" {

static int fem_create_impl(life_ctx *ctx, int32_t n_bodies, const life_fem_body *desc);

int life_fem_create(life_ctx *ctx, int32_t n_bodies, const life_fem_body *desc) {
	if (!ctx) return LIFE_E_ARG;
	const int rc = fem_create_impl(ctx, n_bodies, desc);
	if (rc != LIFE_OK) {      // never leave a half-built set behind: later life_fem_* calls must fail with LIFE_E_STATE, not fault
		const std::string msg = ctx->err;
		fem_free(ctx);
		ctx->err = msg;
	}
	return rc;
}

static int fem_create_impl(life_ctx *ctx, int32_t n_bodies, const life_fem_body *desc) {
	if (n_bodies < 0 || (n_bodies > 0 && !desc)) return fail(ctx, LIFE_E_ARG, "life_fem_create: null description");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	fem_free(ctx);
	if (n_bodies == 0) return LIFE_OK;
	FemState *f = new (std::nothrow) FemState();
	if (!f) return fail(ctx, LIFE_E_NOMEM, "life_fem_create: out of host memory");
	ctx->fem = f;
	f->n_bodies = n_bodies;
	// sizes
	size_t nd = 0, ni = 0, nm = 0;
	for (int k = 0; k < n_bodies; k++) {
		const life_fem_body &d = desc[k];
		if (d.n_nodes < 2 || d.n_bc < 0 || d.n_bc >= 3 * d.n_nodes || d.n_markers < 0 || !d.pos0 || !d.angle0 || !d.element || !d.map_first ||
		    (d.n_markers > 0 && (!d.marker || !d.marker_element || !d.marker_zeta)))
			return fail(ctx, LIFE_E_ARG, "life_fem_create: inconsistent body description");
		const size_t n = (size_t)d.n_nodes, ne = n - 1, dim = 3 * n, nmap = (size_t)d.map_first[ne];
		nd += 3 * n * 2 + 5 * ne + 72 * ne + (size_t)d.n_markers + 2 * nmap + 2 * ne + 48 * ne + 2 * dim * dim + 4 * dim + 8 + 11 * dim;
		ni += (size_t)d.n_markers + ne + 1 + nmap + dim;
		nm += (size_t)d.n_markers;
	}
	std::vector<double> hd(nd, 0.0);
	std::vector<int> hi(ni, 0), mids(nm, 0);
	f->marker_first.assign((size_t)n_bodies + 1, 0);
	LIFE_CUDA(ctx, cudaMalloc(&f->d_arena, sizeof(double) * nd));
	LIFE_CUDA(ctx, cudaMalloc(&f->d_ints, sizeof(int) * ni));
	f->h_bodies.resize((size_t)n_bodies);
	size_t od = 0, oi = 0, om = 0;
	// D / I hand out the next piece of the arenas: the host pointer to fill now, the device pointer for the view
	auto D = [&](size_t n, double *&host) { host = hd.data() + od; double *dev = f->d_arena + od; od += n; return dev; };
	auto I = [&](size_t n, int *&host) { host = hi.data() + oi; int *dev = f->d_ints + oi; oi += n; return dev; };
	for (int k = 0; k < n_bodies; k++) {
		const life_fem_body &d = desc[k];
		Body &b = f->h_bodies[(size_t)k];
		const int n = d.n_nodes, ne = n - 1, dim = 3 * n, nmap = d.map_first[ne];
		b.n_nodes = n; b.n_el = ne; b.n_dof = dim; b.n_bc = d.n_bc; b.n_ibm = d.n_markers;
		if (dim > f->max_dof) f->max_dof = dim;
		b.alpha = d.alpha; b.delta = d.delta; b.Dt = ctx->cfg.Dt; b.Dm = ctx->cfg.Dm; b.gravityX = d.gravity_x; b.gravityY = d.gravity_y; b.ref_L = d.ref_L;
		double *h, *hL0, *hA, *hI, *hE, *hrho, *hM, *hK;
		int *g;
		b.pos0 = D(2 * (size_t)n, h); memcpy(h, d.pos0, sizeof(double) * 2 * n);
		b.angle0 = D((size_t)n, h); memcpy(h, d.angle0, sizeof(double) * n);
		b.L0 = D((size_t)ne, hL0); b.A = D((size_t)ne, hA); b.I = D((size_t)ne, hI); b.E = D((size_t)ne, hE); b.rho = D((size_t)ne, hrho);
		b.Mloc = D(36 * (size_t)ne, hM); b.KLloc = D(36 * (size_t)ne, hK);
		for (int e = 0; e < ne; e++) {
			hL0[e] = d.element[5 * e]; hA[e] = d.element[5 * e + 1]; hI[e] = d.element[5 * e + 2]; hE[e] = d.element[5 * e + 3]; hrho[e] = d.element[5 * e + 4];
			if (!(hL0[e] > 0.0)) return fail(ctx, LIFE_E_ARG, "life_fem_create: element length must be positive");
			// FEMElementClass::setLocalMatrices (src/FEMElement.cpp:203-253), once per element, on the host
			const double L0 = hL0[e], A = hA[e], Im = hI[e], E = hE[e], rho = hrho[e];
			double *M = hM + 36 * e, *K = hK + 36 * e;
			const double C1 = rho * A * L0 / 420.0, L2 = L0 * L0, L3 = L0 * L0 * L0;
			M[0] = C1 * 140.0; M[3] = C1 * 70.0; M[7] = C1 * 156.0; M[8] = C1 * 22.0 * L0; M[10] = C1 * 54; M[11] = C1 * (-13.0 * L0);
			M[14] = C1 * 4.0 * L2; M[16] = C1 * 13.0 * L0; M[17] = C1 * (-3.0 * L2); M[21] = C1 * 140.0; M[28] = C1 * 156.0;
			M[29] = C1 * (-22.0 * L0); M[35] = C1 * 4.0 * L2;
			K[0] = E * A / L0; K[3] = -E * A / L0; K[7] = 12.0 * E * Im / L3; K[8] = 6.0 * E * Im / L2; K[10] = -12.0 * E * Im / L3;
			K[11] = 6.0 * E * Im / L2; K[14] = 4.0 * E * Im / L0; K[16] = -6.0 * E * Im / L2; K[17] = 2.0 * E * Im / L0; K[21] = E * A / L0;
			K[28] = 12.0 * E * Im / L3; K[29] = -6.0 * E * Im / L2; K[35] = 4.0 * E * Im / L0;
			for (int i = 1; i < 6; i++)
				for (int j = 0; j < i; j++) { M[i * 6 + j] = M[j * 6 + i]; K[i * 6 + j] = K[j * 6 + i]; }
		}
		b.pm_zeta = D((size_t)d.n_markers, h); if (d.n_markers) memcpy(h, d.marker_zeta, sizeof(double) * d.n_markers);
		b.fm_z1 = D((size_t)nmap, h); if (nmap) memcpy(h, d.map_zeta1, sizeof(double) * nmap);
		b.fm_z2 = D((size_t)nmap, h); if (nmap) memcpy(h, d.map_zeta2, sizeof(double) * nmap);
		b.pm_el = I((size_t)d.n_markers, g); if (d.n_markers) memcpy(g, d.marker_element, sizeof(int) * d.n_markers);
		b.fm_first = I((size_t)ne + 1, g); memcpy(g, d.map_first, sizeof(int) * (ne + 1));
		b.fm_node = I((size_t)nmap, g); if (nmap) memcpy(g, d.map_marker, sizeof(int) * nmap);
		b.piv = I((size_t)dim, g);
		for (int i = 0; i < d.n_markers; i++) {
			if (d.marker_element[i] < 0 || d.marker_element[i] >= ne || d.marker[i] < 0) return fail(ctx, LIFE_E_ARG, "life_fem_create: marker map out of range");
			mids[om + (size_t)i] = d.marker[i];
			if ((int64_t)d.marker[i] + 1 > f->n_markers_needed) f->n_markers_needed = (int64_t)d.marker[i] + 1;
		}
		for (int i = 0; i < nmap; i++)
			if (d.map_marker[i] < 0 || d.map_marker[i] >= d.n_markers) return fail(ctx, LIFE_E_ARG, "life_fem_create: force map out of range");
		f->marker_first[(size_t)k] = (int)om;
		om += (size_t)d.n_markers;
		b.pos = D(2 * (size_t)n, h); b.angle = D((size_t)n, h);
		b.L = D((size_t)ne, h); b.elangle = D((size_t)ne, h); b.T = D(36 * (size_t)ne, h); b.Floc = D(6 * (size_t)ne, h); b.Rel = D(6 * (size_t)ne, h);
		b.M = D((size_t)dim * dim, h); b.K = D((size_t)dim * dim, h);
		b.R = D((size_t)dim, h); b.F = D((size_t)dim, h); b.delU = D((size_t)dim, h); b.work = D((size_t)dim, h);
		b.scal = D(8, h);
		double **vecs[11] = {&b.U, &b.Udot, &b.Udotdot, &b.U_n, &b.Udot_n, &b.Udotdot_n, &b.U_km1, &b.R_k, &b.R_km1, &b.U_nm1, &b.U_nm2};
		for (int v = 0; v < 11; v++) *vecs[v] = D((size_t)dim, h);
	}
	f->marker_first[(size_t)n_bodies] = (int)om;
	if (od > nd || oi > ni) return fail(ctx, LIFE_E_ARG, "life_fem_create: internal size mismatch");
	LIFE_CUDA(ctx, cudaMemcpy(f->d_arena, hd.data(), sizeof(double) * nd, cudaMemcpyHostToDevice));
	LIFE_CUDA(ctx, cudaMemcpy(f->d_ints, hi.data(), sizeof(int) * ni, cudaMemcpyHostToDevice));
	LIFE_CUDA(ctx, cudaMalloc(&f->d_bodies, sizeof(Body) * (size_t)n_bodies));
	LIFE_CUDA(ctx, cudaMemcpy(f->d_bodies, f->h_bodies.data(), sizeof(Body) * (size_t)n_bodies, cudaMemcpyHostToDevice));
	LIFE_CUDA(ctx, cudaMalloc(&f->d_marker_first, sizeof(int) * ((size_t)n_bodies + 1)));
	LIFE_CUDA(ctx, cudaMemcpy(f->d_marker_first, f->marker_first.data(), sizeof(int) * ((size_t)n_bodies + 1), cudaMemcpyHostToDevice));
	LIFE_CUDA(ctx, cudaMalloc(&f->d_marker_ids, sizeof(int) * (nm ? nm : 1)));
	if (nm) LIFE_CUDA(ctx, cudaMemcpy(f->d_marker_ids, mids.data(), sizeof(int) * nm, cudaMemcpyHostToDevice));
	LIFE_CUDA(ctx, cudaMalloc(&f->d_results, sizeof(double) * 5 * (size_t)n_bodies));
	LIFE_CUDA(ctx, cudaMemset(f->d_results, 0, sizeof(double) * 5 * (size_t)n_bodies));
	return LIFE_OK;
}

// state [11 * dim]: U, Udot, Udotdot, U_n, Udot_n, Udotdot_n, U_km1, R_k, R_km1, U_nm1, U_nm2 — contiguous in the arena in that order
int life_fem_set_state(life_ctx *ctx, int32_t body, const double *state) {
	if (!ctx || !state) return LIFE_E_ARG;
	if (!ctx->fem || body < 0 || body >= ctx->fem->n_bodies) return fail(ctx, LIFE_E_ARG, "life_fem_set_state: no such body");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const Body &b = ctx->fem->h_bodies[(size_t)body];
	LIFE_CUDA(ctx, cudaMemcpyAsync(b.U, state, sizeof(double) * 11 * (size_t)b.n_dof, cudaMemcpyHostToDevice, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_fem_get_state(life_ctx *ctx, int32_t body, double *state) {
	if (!ctx || !state) return LIFE_E_ARG;
	if (!ctx->fem || body < 0 || body >= ctx->fem->n_bodies) return fail(ctx, LIFE_E_ARG, "life_fem_get_state: no such body");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const Body &b = ctx->fem->h_bodies[(size_t)body];
	LIFE_CUDA(ctx, cudaMemcpyAsync(state, b.U, sizeof(double) * 11 * (size_t)b.n_dof, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_fem_predict(life_ctx *ctx, int32_t t) {
	if (!ctx) return LIFE_E_ARG;
	int rc = fem_ready(ctx, "life_fem_predict");
	if (rc) return rc;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return fem_launch<FEM_PREDICT>(ctx, t, 0.0);
}

int life_fem_relax(life_ctx *ctx, double relax) {
	if (!ctx) return LIFE_E_ARG;
	int rc = fem_ready(ctx, "life_fem_relax");
	if (rc) return rc;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return fem_launch<FEM_RELAX>(ctx, 0, relax);
}

int life_fem_dynamic(life_ctx *ctx, double *sums, double *per_body) {
	if (!ctx) return LIFE_E_ARG;
	int rc = fem_ready(ctx, "life_fem_dynamic");
	if (rc) return rc;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	if ((rc = fem_launch<FEM_DYNAMIC>(ctx, 0, 0.0))) return rc;
	FemState *f = ctx->fem;
	std::vector<double> r(5 * (size_t)f->n_bodies);
	LIFE_CUDA(ctx, cudaMemcpyAsync(r.data(), f->d_results, sizeof(double) * r.size(), cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (per_body) memcpy(per_body, r.data(), sizeof(double) * r.size());
	if (sums) {   // in body order, as the ORDERED femKernel adds them (src/Objects.cpp:74-92)
		sums[0] = sums[1] = sums[2] = 0.0;
		for (int k = 0; k < f->n_bodies; k++) { sums[0] += r[5 * (size_t)k]; sums[1] += r[5 * (size_t)k + 1]; sums[2] += r[5 * (size_t)k + 2]; }
	}
	return LIFE_OK;
}

// ---- the sub-iteration loop without the markers leaving the device (src/Objects.cpp:33-52) ------------------------------------------
int life_fsi_move(life_ctx *ctx, int32_t t, int32_t sub_it, double relax) {
	if (!ctx) return LIFE_E_ARG;
	int rc = fem_ready(ctx, "life_fsi_move");
	if (rc) return rc;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	// recomputeObjectVals, src/Objects.cpp:152-232: predictor (first sub-iteration) or the relaxed update, then findSupport and
	// computeDs of the moved markers.  computeEpsilon is the caller's next call (device LU or device assembly + host LAPACK).
	if (sub_it == 0) rc = fem_launch<FEM_PREDICT>(ctx, t, 0.0);
	else rc = fem_launch<FEM_RELAX>(ctx, 0, relax);
	if (rc) return rc;
	if ((rc = ibm_refresh_supports(ctx))) return rc;
	FemState *f = ctx->fem;
	k_fem_ds<<<(unsigned)f->n_bodies, FEM_THREADS, 0, ctx->stream>>>(f->d_bodies, f->d_marker_first, f->d_marker_ids, ctx->mk.pos, ctx->cfg.Dx, ctx->mk.ds);
	ctx->launches++;
	LIFE_CUDA(ctx, cudaGetLastError());
	return LIFE_OK;
}

int life_fsi_force(life_ctx *ctx, double *sums, double *per_body) {
	if (!ctx) return LIFE_E_ARG;
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "life_fsi_force: no state uploaded");
	int rc = fem_ready(ctx, "life_fsi_force");
	if (rc) return rc;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	if ((rc = ibm_interp(ctx, nullptr, true))) return rc;        // ibmKernelInterp, forces stay on the device
	if ((rc = life_fem_dynamic(ctx, sums, per_body))) return rc;  // femKernel; its read-back of the residual sums is the one synchronisation
	return ibm_check(ctx);
}

int life_ibm_set_epsilon(life_ctx *ctx, const double *epsilon) {
	if (!ctx || !epsilon) return LIFE_E_ARG;
	if (ctx->mk.n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->mk.eps, epsilon, sizeof(double) * ctx->mk.n, cudaMemcpyHostToDevice, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_ibm_get_marker_state(life_ctx *ctx, double *force, double *ds, double *epsilon) {
	if (!ctx) return LIFE_E_ARG;
	if (ctx->mk.n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	const int64_t n = ctx->mk.n;
	if (force) LIFE_CUDA(ctx, cudaMemcpyAsync(force, ctx->mk.force, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ds) LIFE_CUDA(ctx, cudaMemcpyAsync(ds, ctx->mk.ds, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (epsilon) LIFE_CUDA(ctx, cudaMemcpyAsync(epsilon, ctx->mk.eps, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

int life_ibm_get_markers(life_ctx *ctx, double *pos, double *vel) {
	if (!ctx) return LIFE_E_ARG;
	if (ctx->mk.n == 0) return LIFE_OK;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	if (pos) LIFE_CUDA(ctx, cudaMemcpyAsync(pos, ctx->mk.pos, sizeof(double) * 2 * ctx->mk.n, cudaMemcpyDeviceToHost, ctx->stream));
	if (vel) LIFE_CUDA(ctx, cudaMemcpyAsync(vel, ctx->mk.vel, sizeof(double) * 2 * ctx->mk.n, cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LIFE_OK;
}

}  // extern "C"
