// fem_core.h — the structural solver of one flexible body as cooperative-thread-array code.
//
// The same source runs as one CTA per filament on the device (csrc/fem.cu: every loop strides by the CTA size, phases are separated
// by FEM_SYNC) and, with a CTA of one thread and FEM_SYNC a no-op, serially on the host.  The serial instantiation is what
// tests/test_fem_core.py holds against the compiled reference (logic, operation by operation); the barrier placement is checked
// on the host as well, by running the CTA as real threads with FEM_SYNC a pthread barrier under ThreadSanitizer
// (tests/native/fem_core_race.cpp); the device instantiation by tests/test_gpu_fem.py on a B200.
//
// What it computes, per body and sub-iteration (reference: FEMBodyClass, src/FEMBody.cpp / FEMElementClass, src/FEMElement.cpp):
//   fem_dynamic   dynamicFEM (src/FEMBody.cpp:26-68): U := U_n; load vector from the marker forces (loadVector, src/FEMElement.cpp:27-69);
//                 Newton-Raphson (:71-87): internal forces + mass + tangent stiffness of every corotational 2-node beam element
//                 (forceVector / massMatrix / stiffMatrix, src/FEMElement.cpp:72-138), Newmark effective system (setNewmark :112-126),
//                 dense LU solve of the unconstrained DOFs, geometry update; then finishNewmark (:129-144), marker positions /
//                 velocities (updateIBMValues :147-195) and the residual sums of the Aitken loop (subResidual :244-256)
//   fem_predict   resetValues + predictor (src/FEMBody.cpp:341-349, :259-289)
//   fem_relax     the relaxed update of later sub-iterations (src/Objects.cpp:195-208)
//
// Differences from the reference that are deliberate: the LU is an in-place right-looking factorisation with partial pivoting
// written here (the reference calls LAPACK dgetrf/dgetrs: same pivots, different summation order, results agree to rounding);
// element contributions are accumulated per element first and gathered per DOF afterwards (no write conflicts between threads).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FEM_FN __device__ __forceinline__
#define FEM_SYNC() __syncthreads()
#else
#define FEM_FN static inline
#ifndef FEM_SYNC               // a host build may supply its own barrier (tests/native/fem_core_race.cpp runs the CTA as real threads)
#define FEM_SYNC() ((void)0)
#endif
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace life_fem {

struct Lane {   // this thread's place in the CTA that owns the body
	int tid, n;
};

struct Body {
	int n_nodes, n_el, n_dof, n_bc, n_ibm;
	double alpha, delta, Dt, Dm, gravityX, gravityY, ref_L;
	// constant description
	const double *pos0, *angle0;                    // nodes [2n], [n]
	const double *L0, *A, *I, *E, *rho;             // elements
	const double *Mloc, *KLloc;                     // element local mass / linear stiffness [36 each] (fem_local_matrices)
	const int *pm_el;  const double *pm_zeta;       // per marker: element, local coordinate
	const int *fm_first, *fm_node;  const double *fm_z1, *fm_z2;   // per element: markers loading it and their ranges
	// geometry derived from U
	double *pos, *angle;                            // nodes
	double *L, *elangle, *T, *Floc;                 // elements: length, angle, transformation [36], local internal forces [6]
	// work space
	double *Rel;                                    // per-element 6-vectors (load / internal force contributions)
	double *M, *K, *R, *F, *delU, *work;            // [dim*dim] x2, [dim] x4
	int *piv;                                       // [dim]
	double *scal;                                   // [8]: resNR, subRes, subNum, subDen, itNR, pivot scratch
	// state
	double *U, *Udot, *Udotdot, *U_n, *Udot_n, *Udotdot_n, *U_km1, *R_k, *R_km1, *U_nm1, *U_nm2;
};

FEM_FN double shift_angle(double a) {   // Utils::shiftAngle, inc/Utils.h:235-246
	a = fmod(a + M_PI, 2.0 * M_PI);
	if (a < 0.0) a += 2.0 * M_PI;
	return a - M_PI;
}

FEM_FN void mat6_mul(const double *A, const double *B, double *C) {
	for (int i = 0; i < 6; i++)
		for (int j = 0; j < 6; j++) {
			double s = 0.0;
			for (int k = 0; k < 6; k++) s += A[i * 6 + k] * B[k * 6 + j];
			C[i * 6 + j] = s;
		}
}
// C = A^T * B
FEM_FN void mat6_tmul(const double *A, const double *B, double *C) {
	for (int i = 0; i < 6; i++)
		for (int j = 0; j < 6; j++) {
			double s = 0.0;
			for (int k = 0; k < 6; k++) s += A[k * 6 + i] * B[k * 6 + j];
			C[i * 6 + j] = s;
		}
}
FEM_FN void mat6_tvec(const double *A, const double *x, double *y) {   // y = A^T x
	for (int i = 0; i < 6; i++) {
		double s = 0.0;
		for (int j = 0; j < 6; j++) s += A[j * 6 + i] * x[j];
		y[i] = s;
	}
}
FEM_FN void mat6_vec(const double *A, const double *x, double *y) {
	for (int i = 0; i < 6; i++) {
		double s = 0.0;
		for (int j = 0; j < 6; j++) s += A[i * 6 + j] * x[j];
		y[i] = s;
	}
}

// FEMElementClass::setLocalMatrices, src/FEMElement.cpp:203-253 (called once per element when a body is set up)
FEM_FN void fem_local_matrices(double L0, double A, double I, double E, double rho, double *M, double *K) {
	for (int k = 0; k < 36; k++) { M[k] = 0.0; K[k] = 0.0; }
	const double C1 = rho * A * L0 / 420.0, L2 = L0 * L0, L3 = L0 * L0 * L0;
	M[0] = C1 * 140.0; M[3] = C1 * 70.0;
	M[7] = C1 * 156.0; M[8] = C1 * 22.0 * L0; M[10] = C1 * 54; M[11] = C1 * (-13.0 * L0);
	M[14] = C1 * 4.0 * L2; M[16] = C1 * 13.0 * L0; M[17] = C1 * (-3.0 * L2);
	M[21] = C1 * 140.0;
	M[28] = C1 * 156.0; M[29] = C1 * (-22.0 * L0);
	M[35] = C1 * 4.0 * L2;
	K[0] = E * A / L0; K[3] = -E * A / L0;
	K[7] = 12.0 * E * I / L3; K[8] = 6.0 * E * I / L2; K[10] = -12.0 * E * I / L3; K[11] = 6.0 * E * I / L2;
	K[14] = 4.0 * E * I / L0; K[16] = -6.0 * E * I / L2; K[17] = 2.0 * E * I / L0;
	K[21] = E * A / L0;
	K[28] = 12.0 * E * I / L3; K[29] = -6.0 * E * I / L2;
	K[35] = 4.0 * E * I / L0;
	for (int i = 1; i < 6; i++)
		for (int j = 0; j < i; j++) { M[i * 6 + j] = M[j * 6 + i]; K[i * 6 + j] = K[j * 6 + i]; }
}

// FEMBodyClass::updateFEMValues, src/FEMBody.cpp:198-223 (+ setElementTransform, src/FEMElement.cpp:255-262)
FEM_FN void update_geometry(const Body &b, Lane l) {
	for (int n = l.tid; n < b.n_nodes; n += l.n) {
		b.pos[2 * n] = b.pos0[2 * n] + b.U[3 * n];
		b.pos[2 * n + 1] = b.pos0[2 * n + 1] + b.U[3 * n + 1];
		b.angle[n] = b.angle0[n] + b.U[3 * n + 2];
	}
	FEM_SYNC();
	for (int e = l.tid; e < b.n_el; e += l.n) {
		const double vx = b.pos[2 * (e + 1)] - b.pos[2 * e], vy = b.pos[2 * (e + 1) + 1] - b.pos[2 * e + 1];
		const double ang = atan2(vy, vx);
		b.elangle[e] = ang;
		b.L[e] = sqrt(vx * vx + vy * vy);
		double *T = b.T + 36 * e;
		const double c = cos(ang), s = sin(ang);
		for (int k = 0; k < 36; k++) T[k] = 0.0;
		T[0] = T[7] = T[21] = T[28] = c;
		T[1] = T[22] = s;
		T[6] = T[27] = -s;
		T[14] = T[35] = 1.0;
	}
	FEM_SYNC();
}

// gather per-element 6-vectors into a global vector: DOF d of node n takes entry 3+j of element n-1 and entry j of element n
FEM_FN void gather_elements(const Body &b, Lane l, double *out) {
	for (int d = l.tid; d < b.n_dof; d += l.n) {
		const int n = d / 3, j = d - 3 * n;
		double s = 0.0;
		if (n > 0) s += b.Rel[6 * (n - 1) + 3 + j];
		if (n < b.n_el) s += b.Rel[6 * n + j];
		out[d] = s;
	}
	FEM_SYNC();
}

// FEMBodyClass::constructRVector (src/FEMBody.cpp:226-234) over FEMElementClass::loadVector (src/FEMElement.cpp:27-69).
// force [2 * markers] in lattice units and epsilon [markers] are indexed through `marker` (the body's k-th marker in the global
// marker arrays), or directly when marker == nullptr.
FEM_FN void load_vector(const Body &b, Lane l, const double *force, const double *epsilon, const int *marker) {
	const double forceScale = b.Dm / (b.Dt * b.Dt);
	for (int e = l.tid; e < b.n_el; e += l.n) {
		const double *T = b.T + 36 * e;
		const double L = b.L[e];
		const double wx = b.rho[e] * b.A[e] * b.gravityX, wy = b.rho[e] * b.A[e] * b.gravityY;
		double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
		for (int k = b.fm_first[e]; k < b.fm_first[e + 1]; k++) {
			const int loc = b.fm_node[k], nd = marker ? marker[loc] : loc;
			const double a = b.fm_z1[k], c = b.fm_z2[k];
			const double sc = -epsilon[nd] * forceScale;
			const double gx = sc * force[2 * nd] + wx, gy = sc * force[2 * nd + 1] + wy;
			const double Fx = T[0] * gx + T[1] * gy, Fy = T[6] * gx + T[7] * gy;
			const double a2 = a * a, c2 = c * c, a3 = a2 * a, c3 = c2 * c, a4 = a2 * a2, c4 = c2 * c2;
			double R[6], RG[6];
			R[0] = Fx * 0.5 * L * (0.5 * c - 0.5 * a + 0.25 * a2 - 0.25 * c2);
			R[1] = Fy * 0.5 * L * (0.5 * c - 0.5 * a - a4 / 16.0 + c4 / 16.0 + 3.0 * a2 / 8.0 - 3.0 * c2 / 8.0);
			R[2] = Fy * 0.5 * L * (L * (-a4 + c4) / 32.0 - L * (-a3 + c3) / 24.0 - L * (-a2 + c2) / 16.0 + L * (c - a) / 8.0);
			R[3] = Fx * 0.5 * L * (-0.25 * a2 + 0.25 * c2 + 0.5 * c - 0.5 * a);
			R[4] = Fy * 0.5 * L * (0.5 * c - 0.5 * a + a4 / 16.0 - c4 / 16.0 - 3.0 * a2 / 8.0 + 3.0 * c2 / 8.0);
			R[5] = Fy * 0.5 * L * (L * (-a4 + c4) / 32.0 + L * (-a3 + c3) / 24.0 - L * (-a2 + c2) / 16.0 - L * (c - a) / 8.0);
			mat6_tvec(T, R, RG);
			for (int i = 0; i < 6; i++) acc[i] += RG[i];
		}
		for (int i = 0; i < 6; i++) b.Rel[6 * e + i] = acc[i];
	}
	FEM_SYNC();
	gather_elements(b, l, b.R);
}

// FEMBodyClass::buildGlobalMatrices (src/FEMBody.cpp:90-109): forceVector, massMatrix, stiffMatrix of every element
// (src/FEMElement.cpp:72-138).  Adjacent elements share a node, i.e. a 3x3 block of M and K: even and odd elements are assembled
// in two passes so that no two threads add to the same entry.
FEM_FN void build_matrices(const Body &b, Lane l) {
	const int dim = b.n_dof;
	for (int k = l.tid; k < dim * dim; k += l.n) { b.M[k] = 0.0; b.K[k] = 0.0; }
	FEM_SYNC();
	for (int colour = 0; colour < 2; colour++) {
		for (int e = 2 * l.tid + colour; e < b.n_el; e += 2 * l.n) {
			const double *T = b.T + 36 * e;
			const double L = b.L[e], L0 = b.L0[e], E = b.E[e], A = b.A[e], I = b.I[e];
			// forceVector
			const double u = (L * L - L0 * L0) / (L + L0);
			const double th1 = shift_angle(b.angle[e] - b.elangle[e]), th2 = shift_angle(b.angle[e + 1] - b.elangle[e]);
			const double F0 = (E * A / L0) * u;
			const double M1 = (2 * E * I / L0) * (2.0 * th1 + th2), M2 = (2 * E * I / L0) * (th1 + 2.0 * th2);
			double *F = b.Floc + 6 * e;
			F[0] = -F0; F[1] = (1.0 / L0) * (M1 + M2); F[2] = M1; F[3] = F0; F[4] = -(1.0 / L0) * (M1 + M2); F[5] = M2;
			mat6_tvec(T, F, b.Rel + 6 * e);
			// massMatrix: T^T M T ; stiffMatrix: T^T (K_L + K_NL) T
			double tmp[36], G[36], KK[36];
			mat6_tmul(T, b.Mloc + 36 * e, tmp);
			mat6_mul(tmp, T, G);
			for (int i = 0; i < 6; i++)
				for (int j = 0; j < 6; j++) b.M[(3 * e + i) * dim + 3 * e + j] += G[i * 6 + j];
			const double V0 = F[4];
			for (int k = 0; k < 36; k++) KK[k] = b.KLloc[36 * e + k];
			KK[0 * 6 + 1] += -V0 / L0; KK[0 * 6 + 4] += V0 / L0;
			KK[1 * 6 + 0] += -V0 / L0; KK[1 * 6 + 1] += F0 / L0; KK[1 * 6 + 3] += V0 / L0; KK[1 * 6 + 4] += -F0 / L0;
			KK[3 * 6 + 1] += V0 / L0; KK[3 * 6 + 4] += -V0 / L0;
			KK[4 * 6 + 0] += V0 / L0; KK[4 * 6 + 1] += -F0 / L0; KK[4 * 6 + 3] += -V0 / L0; KK[4 * 6 + 4] += F0 / L0;
			mat6_tmul(T, KK, tmp);
			mat6_mul(tmp, T, G);
			for (int i = 0; i < 6; i++)
				for (int j = 0; j < 6; j++) b.K[(3 * e + i) * dim + 3 * e + j] += G[i * 6 + j];
		}
		FEM_SYNC();
	}
	gather_elements(b, l, b.F);
}

// FEMBodyClass::setNewmark, src/FEMBody.cpp:112-126: F := R - F + M (a0 (U_n - U) + a2 Udot + a3 Udotdot), K := K + a0 M
FEM_FN void newmark_system(const Body &b, Lane l) {
	const int dim = b.n_dof;
	const double a0 = 1.0 / (b.alpha * b.Dt * b.Dt), a2 = 1.0 / (b.alpha * b.Dt), a3 = 1.0 / (2.0 * b.alpha) - 1.0;
	for (int i = l.tid; i < dim; i += l.n) b.work[i] = a0 * (b.U_n[i] - b.U[i]) + a2 * b.Udot[i] + a3 * b.Udotdot[i];
	FEM_SYNC();
	for (int i = l.tid; i < dim; i += l.n) {
		double s = 0.0;
		for (int j = 0; j < dim; j++) s += b.M[i * dim + j] * b.work[j];
		b.F[i] = b.R[i] - b.F[i] + s;
	}
	for (int k = l.tid; k < dim * dim; k += l.n) b.K[k] += a0 * b.M[k];
	FEM_SYNC();
}

// delU := K^-1 F on the DOFs n_bc .. dim-1, zero on the constrained ones (Utils::solveLAPACK(K, F, bcDOFs), src/Utils.cpp:288-311).
// In-place LU with partial pivoting (K is rebuilt by the next Newton-Raphson iteration), then the two triangular solves.
FEM_FN void solve_system(const Body &b, Lane l) {
	const int dim = b.n_dof, o = b.n_bc, n = dim - o;
	double *A = b.K, *x = b.delU;
	for (int i = l.tid; i < dim; i += l.n) x[i] = i < o ? 0.0 : b.F[i];
	FEM_SYNC();
	for (int k = 0; k < n; k++) {
#if defined(__CUDA_ARCH__)
		if (l.tid < 32) {   // pivot: largest |A[i][k]|, i >= k (first one wins, as idamax) — one warp, a shuffle reduction
			double best = -1.0;
			int p = k;
			for (int i = k + l.tid; i < n; i += 32) {
				const double v = fabs(A[(o + i) * dim + o + k]);
				if (v > best) { best = v; p = i; }
			}
			for (int w = 16; w > 0; w >>= 1) {
				const double ob = __shfl_xor_sync(0xffffffffu, best, w);
				const int op = __shfl_xor_sync(0xffffffffu, p, w);
				if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
			}
			if (l.tid == 0) b.piv[k] = p;
		}
#else
		if (l.tid == 0) {   // pivot: largest |A[i][k]|, i >= k (first one wins, as idamax)
			int p = k;
			double best = fabs(A[(o + k) * dim + o + k]);
			for (int i = k + 1; i < n; i++) {
				const double v = fabs(A[(o + i) * dim + o + k]);
				if (v > best) { best = v; p = i; }
			}
			b.piv[k] = p;
		}
#endif
		FEM_SYNC();
		const int p = b.piv[k];
		if (p != k) {
			for (int j = l.tid; j < n; j += l.n) {
				const double t = A[(o + k) * dim + o + j];
				A[(o + k) * dim + o + j] = A[(o + p) * dim + o + j];
				A[(o + p) * dim + o + j] = t;
			}
			if (l.tid == 0) { const double t = x[o + k]; x[o + k] = x[o + p]; x[o + p] = t; }
			FEM_SYNC();
		}
		// One phase for everything below the pivot row: a row belongs to one group of lanes, which forms the multiplier L[i][k] from
		// the UNSCALED entry (every lane on its own; it is not written back — L is never read again, the forward substitution is
		// folded in here), updates the row and the right-hand side.  Lanes walk along the row (contiguous), no integer division per
		// element.  Three barriers per column instead of five.
		{
			const double inv = 1.0 / A[(o + k) * dim + o + k];
			const int W = l.n >= 32 ? 32 : 1, G = l.n / W;
			const int lane = l.tid % W, grp = l.tid / W;
			for (int i = k + 1 + grp; i < n; i += G) {
				const double lik = A[(o + i) * dim + o + k] * inv;
				for (int j = k + 1 + lane; j < n; j += W) A[(o + i) * dim + o + j] -= lik * A[(o + k) * dim + o + j];
				if (lane == 0) x[o + i] -= lik * x[o + k];
			}
		}
		FEM_SYNC();
	}
	// back substitution, one barrier per column: x[k] / U[k][k] is formed by every thread (nobody writes x[k] any more), kept in
	// `work`, and the rows above are updated
	double *y = b.work;
	for (int k = n - 1; k >= 0; k--) {
		const double xk = x[o + k] / A[(o + k) * dim + o + k];
		if (l.tid == 0) y[k] = xk;
		for (int i = l.tid; i < k; i += l.n) x[o + i] -= A[(o + i) * dim + o + k] * xk;
		FEM_SYNC();
	}
	for (int i = l.tid; i < n; i += l.n) x[o + i] = y[i];
	FEM_SYNC();
}

// FEMBodyClass::finishNewmark, src/FEMBody.cpp:129-144
FEM_FN void finish_newmark(const Body &b, Lane l) {
	const double Dt = b.Dt;
	const double a6 = 1.0 / (b.alpha * Dt * Dt), a7 = -1.0 / (b.alpha * Dt), a8 = -(1.0 / (2.0 * b.alpha) - 1.0);
	const double a9 = Dt * (1.0 - b.delta), a10 = b.delta * Dt;
	for (int i = l.tid; i < b.n_dof; i += l.n) {
		const double acc = a6 * (b.U[i] - b.U_n[i]) + a7 * b.Udot_n[i] + a8 * b.Udotdot_n[i];
		b.Udotdot[i] = acc;
		b.Udot[i] = b.Udot_n[i] + a9 * b.Udotdot_n[i] + a10 * acc;
	}
	FEM_SYNC();
}

FEM_FN void shape_funs(const double *v, double zeta, double L, double *r) {   // FEMElementClass::shapeFuns, src/FEMElement.cpp:141-160
	const double h = (zeta + 1.0) / 2.0, h2 = h * h, h3 = h2 * h;
	const double N0 = 1.0 - h, N1 = 1.0 - 3.0 * h2 + 2.0 * h3, N2 = (h - 2.0 * h2 + h3) * L;
	const double N3 = h, N4 = 3.0 * h2 - 2.0 * h3, N5 = (-h2 + h3) * L;
	r[0] = v[0] * N0 + v[3] * N3;
	r[1] = v[1] * N1 + v[2] * N2 + v[4] * N4 + v[5] * N5;
}

// FEMBodyClass::updateIBMValues, src/FEMBody.cpp:147-195: marker positions / velocities (physical units) from U, Udot.
// pos / vel [2 * markers], indexed through `marker` like load_vector.
FEM_FN void update_markers(const Body &b, Lane l, double *pos, double *vel, const int *marker) {
	for (int i = l.tid; i < b.n_ibm; i += l.n) {
		const int e = b.pm_el[i], out = marker ? marker[i] : i;
		const double *T = b.T + 36 * e;
		double dU[6], dV[6], tU[6], tV[6], sU[2], sV[2];
		for (int k = 0; k < 6; k++) { dU[k] = b.U[3 * e + k]; dV[k] = b.Udot[3 * e + k]; }
		for (int n = 0; n < 2; n++) {
			dU[3 * n] += b.pos0[2 * (e + n)] - b.pos[2 * e];
			dU[3 * n + 1] += b.pos0[2 * (e + n) + 1] - b.pos[2 * e + 1];
			dU[3 * n + 2] = shift_angle(dU[3 * n + 2] + (b.angle0[e + n] - b.elangle[e]));
		}
		mat6_vec(T, dU, tU);
		mat6_vec(T, dV, tV);
		shape_funs(tU, b.pm_zeta[i], b.L[e], sU);
		shape_funs(tV, b.pm_zeta[i], b.L[e], sV);
		pos[2 * out] = b.pos[2 * e] + (T[0] * sU[0] + T[6] * sU[1]);
		pos[2 * out + 1] = b.pos[2 * e + 1] + (T[1] * sU[0] + T[7] * sU[1]);
		vel[2 * out] = T[0] * sV[0] + T[6] * sV[1];
		vel[2 * out + 1] = T[1] * sV[0] + T[7] * sV[1];
	}
	FEM_SYNC();
}

// IBMNodeClass::computeDs, src/IBMNode.cpp:182-204, for the markers of this body: distance to the nearest OTHER marker of the same
// body in lattice units (what scales the spread force and the epsilon matrix).  Every product and sum is rounded on its own on
// the device as well (no FMA contraction), so ds comes out bit-identical to the reference: it feeds paths that are bit-exact.
// pos / ds are the global marker arrays, indexed through `marker` (or directly when marker == nullptr).
FEM_FN void compute_ds(const Body &b, Lane l, const double *pos, double Dx, double *ds, const int *marker) {
	for (int i = l.tid; i < b.n_ibm; i += l.n) {
		const int gi = marker ? marker[i] : i;
		const double xi = pos[2 * gi], yi = pos[2 * gi + 1];
		double current = 10.0;
		for (int n = 0; n < b.n_ibm; n++) {
			if (n == i) continue;
			const int gn = marker ? marker[n] : n;
			const double dx = xi - pos[2 * gn], dy = yi - pos[2 * gn + 1];
#if defined(__CUDA_ARCH__)
			const double mag = __ddiv_rn(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), Dx);
#else
			const double mag = sqrt(dx * dx + dy * dy) / Dx;
#endif
			if (mag < current) current = mag;
		}
		ds[gi] = current;
	}
	FEM_SYNC();
}

// FEMBodyClass::subResidual, src/FEMBody.cpp:244-256 -> scal[1..3] = subRes, subNum, subDen
FEM_FN void sub_residual(const Body &b, Lane l) {
	for (int i = l.tid; i < b.n_dof; i += l.n) {
		b.R_km1[i] = b.R_k[i];
		b.R_k[i] = b.U[i] - b.U_km1[i];
	}
	FEM_SYNC();
	if (l.tid == 0) {
		double res = 0.0, num = 0.0, den = 0.0;
		for (int i = 0; i < b.n_dof; i++) {
			const double d = b.R_k[i] - b.R_km1[i];
			res += b.R_k[i] * b.R_k[i];
			num += b.R_km1[i] * d;
			den += d * d;
		}
		b.scal[1] = res; b.scal[2] = num; b.scal[3] = den;
	}
	FEM_SYNC();
}

// FEMBodyClass::dynamicFEM, src/FEMBody.cpp:26-68.  scal[0] = resNR, scal[4] = itNR on return.
FEM_FN void fem_dynamic(const Body &b, Lane l, const double *force, const double *epsilon, double *pos, double *vel, const int *marker) {
	const int dim = b.n_dof;
	for (int i = l.tid; i < dim; i += l.n) { b.U[i] = b.U_n[i]; b.Udot[i] = b.Udot_n[i]; b.Udotdot[i] = b.Udotdot_n[i]; }
	FEM_SYNC();
	update_geometry(b, l);
	load_vector(b, l, force, epsilon, marker);
	const double TOL = 1e-10;
	const int MAXIT = 20;
	int it = 0;
	double res;
	do {
		build_matrices(b, l);
		newmark_system(b, l);
		solve_system(b, l);
		for (int i = l.tid; i < dim; i += l.n) b.U[i] += b.delU[i];
		FEM_SYNC();
		update_geometry(b, l);
		if (l.tid == 0) {
			double s = 0.0;
			for (int i = 0; i < dim; i++) s += b.delU[i] * b.delU[i];
			b.scal[0] = sqrt(s) / (b.ref_L * sqrt((double)dim));
		}
		FEM_SYNC();
		res = b.scal[0];     // the same value in every thread: the loop condition is uniform across the CTA
		it++;
		FEM_SYNC();          // nobody overwrites scal[0] in the next iteration before everyone has read it
	} while (res > TOL && it < MAXIT);
	if (l.tid == 0) b.scal[4] = (double)it;
	finish_newmark(b, l);
	update_markers(b, l, pos, vel, marker);
	sub_residual(b, l);
}

// resetValues + predictor at time step t (src/FEMBody.cpp:341-349, :259-289; driven from src/Objects.cpp:160-174).
// The reference swaps vectors; here the values move (pointers are fixed views into device memory).
FEM_FN void fem_predict(const Body &b, Lane l, int t, double *pos, double *vel, const int *marker) {
	for (int i = l.tid; i < b.n_dof; i += l.n) {
		const double u = b.U[i], un = b.U_n[i], unm1 = b.U_nm1[i], unm2 = b.U_nm2[i];
		// after the swaps: U_nm2 = old U_nm1, U_nm1 = old U_n, U_n = old U, Udot_n = old Udot, Udotdot_n = old Udotdot;
		// U = old U_nm2; Udot, Udotdot hold leftovers that finishNewmark overwrites
		b.U_nm2[i] = unm1;
		b.U_nm1[i] = un;
		b.U_n[i] = u;
		b.Udot_n[i] = b.Udot[i];
		b.Udotdot_n[i] = b.Udotdot[i];
		double pred;
		if (t > 2) pred = 2.5 * u - 2.0 * un + 0.5 * unm1;
		else if (t == 2) pred = 2.0 * u - un;
		else if (t == 1) pred = u;
		else pred = unm2;   // t <= 0 (never reached by main()): the reference leaves U as the swaps left it
		b.U[i] = pred;
	}
	FEM_SYNC();
	update_geometry(b, l);
	finish_newmark(b, l);
	update_markers(b, l, pos, vel, marker);
	for (int i = l.tid; i < b.n_dof; i += l.n) { const double t0 = b.U_km1[i]; b.U_km1[i] = b.U[i]; b.U[i] = t0; }
	FEM_SYNC();
}

// the relaxed update of sub-iterations >= 1 (src/Objects.cpp:195-208)
FEM_FN void fem_relax(const Body &b, Lane l, double relax, double *pos, double *vel, const int *marker) {
	for (int i = l.tid; i < b.n_dof; i += l.n) b.U[i] = b.U_km1[i] + relax * (b.U[i] - b.U_km1[i]);
	FEM_SYNC();
	update_geometry(b, l);
	finish_newmark(b, l);
	update_markers(b, l, pos, vel, marker);
	for (int i = l.tid; i < b.n_dof; i += l.n) { const double t0 = b.U_km1[i]; b.U_km1[i] = b.U[i]; b.U[i] = t0; }
	FEM_SYNC();
}

}  // namespace life_fem
