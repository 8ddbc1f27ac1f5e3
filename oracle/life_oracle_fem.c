/*
 * life_oracle_fem.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Groundwork for SURVEY.md §8f row 3 (FEM on the device, not built yet — DESIGN.md §10): a plain-C restatement of the
 * structural solver the reference runs per flexible body and sub-iteration, FEMBodyClass::dynamicFEM (src/FEMBody.cpp:26-68) —
 * corotational 2-node beam elements (3 DOF per node), Newmark-beta time integration, Newton-Raphson with a dense LU solve —
 * together with the pieces of the sub-iteration loop that touch the same state: resetValues + predictor (src/FEMBody.cpp:259-289,
 * :341-349) and the Aitken relaxation update (src/Objects.cpp:191-210).  Every function cites the reference lines it follows and
 * performs the floating-point operations in the reference's order (compile with -ffp-contract=off), the linear solve goes
 * through the very dgetrf_/dgetrs_ the reference calls (Utils::solveLAPACK, src/Utils.cpp:288-311; bound at run time with
 * orc_fem_bind_lapack), so results are expected to agree with the compiled reference bit for bit.
 *
 * Pinned by tests/test_oracle_fem.py against oracle/_ref/libref_<case>.so (the unmodified reference): every dynamicFEM call of
 * live FSI runs of TurekHron, InvertedFlag, Honami and PELskin.  Geometry construction (src/FEMBody.cpp:500-578,
 * computeNodeMapping :292-338) is out of scope: bodies are created from the description the reference built
 * (ref_fem_geometry / ref_fem_maps of oracle/ref_harness.cpp).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SQ(x) ((x) * (x))
#define TH(x) ((x) * (x) * (x))
#define EL_DOFS 6
#define NODE_DOFS 3

typedef void (*getrf_fn)(int *, int *, double *, int *, int *, int *);
typedef void (*getrs_fn)(char *, int *, int *, double *, int *, int *, double *, int *, int *);
static getrf_fn p_dgetrf = NULL;
static getrs_fn p_dgetrs = NULL;

typedef struct orc_fem {
	int n_nodes, n_el, n_dof, n_bc, n_ibm;
	double alpha, delta, Dt, Dm, gravityX, gravityY, ref_L;
	/* constant description */
	double *pos0, *angle0;                     /* nodes: [2n], [n] */
	double *L0, *A, *I, *E, *rho;              /* elements */
	double *Mloc, *KLloc;                      /* element local mass / linear stiffness, 36 each (setLocalMatrices) */
	int *pm_el;  double *pm_zeta;              /* posMap per IBM node */
	int *fm_first, *fm_node;  double *fm_z1, *fm_z2;   /* forceMap per element */
	/* geometry derived from U (updateFEMValues) */
	double *pos, *angle;                       /* nodes */
	double *L, *elangle, *T, *Floc;            /* elements: length, angle, transformation (36), local internal forces (6) */
	/* system */
	double *M, *K, *R, *F, *delU, *work, *Kcopy;
	double *U, *Udot, *Udotdot, *U_n, *Udot_n, *Udotdot_n, *U_km1, *R_k, *R_km1, *U_nm1, *U_nm2;
	int itNR;
	double resNR, subRes, subNum, subDen;
} orc_fem;

/* Utils::shiftAngle, inc/Utils.h:235-246 */
static double shift_angle(double angle) {
	angle = fmod(angle + M_PI, 2.0 * M_PI);
	if (angle < 0.0) angle += 2.0 * M_PI;
	return angle - M_PI;
}

/* C = A * B for 6x6 row-major, sums started at 0.0 in k order (inc/Utils.h:393-409) */
static void mat6_mul(const double *A, const double *B, double *C) {
	for (int i = 0; i < 6; i++)
		for (int j = 0; j < 6; j++) {
			double s = 0.0;
			for (int k = 0; k < 6; k++) s += A[i * 6 + k] * B[k * 6 + j];
			C[i * 6 + j] = s;
		}
}
static void mat6_transpose(const double *A, double *B) {
	for (int i = 0; i < 6; i++)
		for (int j = 0; j < 6; j++) B[i * 6 + j] = A[j * 6 + i];
}
/* y = A * x, sums started at 0.0 in j order (inc/Utils.h:375-389) */
static void mat6_vec(const double *A, const double *x, double *y) {
	for (int i = 0; i < 6; i++) {
		double s = 0.0;
		for (int j = 0; j < 6; j++) s += A[i * 6 + j] * x[j];
		y[i] = s;
	}
}

/* FEMElementClass::setLocalMatrices, src/FEMElement.cpp:203-253 */
static void set_local_matrices(orc_fem *b, int e) {
	double *M = b->Mloc + 36 * e, *K = b->KLloc + 36 * e;
	const double L0 = b->L0[e], A = b->A[e], I = b->I[e], E = b->E[e], rho = b->rho[e];
	memset(M, 0, 36 * sizeof(double));
	memset(K, 0, 36 * sizeof(double));
	const double C1 = rho * A * L0 / 420.0;
	M[0 * 6 + 0] = C1 * 140.0;
	M[0 * 6 + 3] = C1 * 70.0;
	M[1 * 6 + 1] = C1 * 156.0;
	M[1 * 6 + 2] = C1 * 22.0 * L0;
	M[1 * 6 + 4] = C1 * 54;
	M[1 * 6 + 5] = C1 * (-13.0 * L0);
	M[2 * 6 + 2] = C1 * 4.0 * SQ(L0);
	M[2 * 6 + 4] = C1 * 13.0 * L0;
	M[2 * 6 + 5] = C1 * (-3.0 * SQ(L0));
	M[3 * 6 + 3] = C1 * 140.0;
	M[4 * 6 + 4] = C1 * 156.0;
	M[4 * 6 + 5] = C1 * (-22.0 * L0);
	M[5 * 6 + 5] = C1 * 4.0 * SQ(L0);
	K[0 * 6 + 0] = E * A / L0;
	K[0 * 6 + 3] = -E * A / L0;
	K[1 * 6 + 1] = 12.0 * E * I / TH(L0);
	K[1 * 6 + 2] = 6.0 * E * I / SQ(L0);
	K[1 * 6 + 4] = -12.0 * E * I / TH(L0);
	K[1 * 6 + 5] = 6.0 * E * I / SQ(L0);
	K[2 * 6 + 2] = 4.0 * E * I / L0;
	K[2 * 6 + 4] = -6.0 * E * I / SQ(L0);
	K[2 * 6 + 5] = 2.0 * E * I / L0;
	K[3 * 6 + 3] = E * A / L0;
	K[4 * 6 + 4] = 12.0 * E * I / TH(L0);
	K[4 * 6 + 5] = -6.0 * E * I / SQ(L0);
	K[5 * 6 + 5] = 4.0 * E * I / L0;
	for (int i = 1; i < 6; i++)
		for (int j = 0; j < i; j++) {
			M[i * 6 + j] = M[j * 6 + i];
			K[i * 6 + j] = K[j * 6 + i];
		}
}

/* FEMElementClass::setElementTransform, src/FEMElement.cpp:255-262 (the other entries stay 0 from the constructor) */
static void set_element_transform(orc_fem *b, int e) {
	double *T = b->T + 36 * e;
	const double c = cos(b->elangle[e]), s = sin(b->elangle[e]);
	T[0 * 6 + 0] = T[1 * 6 + 1] = T[3 * 6 + 3] = T[4 * 6 + 4] = c;
	T[0 * 6 + 1] = T[3 * 6 + 4] = s;
	T[1 * 6 + 0] = T[4 * 6 + 3] = -s;
	T[2 * 6 + 2] = T[5 * 6 + 5] = 1.0;
}

/* FEMBodyClass::updateFEMValues, src/FEMBody.cpp:198-223 */
static void update_fem_values(orc_fem *b) {
	for (int n = 0; n < b->n_nodes; n++) {
		for (int d = 0; d < 2; d++) b->pos[2 * n + d] = b->pos0[2 * n + d] + b->U[NODE_DOFS * n + d];
		b->angle[n] = b->angle0[n] + b->U[NODE_DOFS * n + 2];
	}
	for (int e = 0; e < b->n_el; e++) {
		const double vx = b->pos[2 * (e + 1)] - b->pos[2 * e], vy = b->pos[2 * (e + 1) + 1] - b->pos[2 * e + 1];
		b->elangle[e] = atan2(vy, vx);
		double dot = 0.0;
		dot += vx * vx;
		dot += vy * vy;
		b->L[e] = sqrt(dot);
		set_element_transform(b, e);
	}
}

/* FEMElementClass::loadVector, src/FEMElement.cpp:27-69, for every element (FEMBodyClass::constructRVector, src/FEMBody.cpp:226-234).
 * force [2*n_ibm] lattice units, epsilon [n_ibm]: the body's IBM nodes in body-local order */
static void construct_r_vector(orc_fem *b, const double *force, const double *epsilon) {
	memset(b->R, 0, sizeof(double) * b->n_dof);
	const double forceScale = b->Dm / SQ(b->Dt);
	for (int e = 0; e < b->n_el; e++) {
		const double *T = b->T + 36 * e;
		const double L = b->L[e];
		const double weight[2] = {b->rho[e] * b->A[e] * b->gravityX, b->rho[e] * b->A[e] * b->gravityY};
		double Tt[36];
		mat6_transpose(T, Tt);
		for (int k = b->fm_first[e]; k < b->fm_first[e + 1]; k++) {
			const int nd = b->fm_node[k];
			const double a = b->fm_z1[k], bb = b->fm_z2[k];
			const double sc = -epsilon[nd] * 1.0 * forceScale;
			const double g[2] = {sc * force[2 * nd] + weight[0], sc * force[2 * nd + 1] + weight[1]};
			double F[2];
			for (int i = 0; i < 2; i++) {
				double s = 0.0;
				for (int j = 0; j < 2; j++) s += T[i * 6 + j] * g[j];
				F[i] = s;
			}
			double R[6], RG[6];
			R[0] = F[0] * 0.5 * L * (0.5 * bb - 0.5 * a + 0.25 * SQ(a) - 0.25 * SQ(bb));
			R[1] = F[1] * 0.5 * L * (0.5 * bb - 0.5 * a - SQ(a) * SQ(a) / 16.0 + SQ(bb) * SQ(bb) / 16.0 + 3.0 * SQ(a) / 8.0 - 3.0 * SQ(bb) / 8.0);
			R[2] = F[1] * 0.5 * L * (L * (-SQ(a) * SQ(a) + SQ(bb) * SQ(bb)) / 32.0 - L * (-TH(a) + TH(bb)) / 24.0 - L * (-SQ(a) + SQ(bb)) / 16.0 + L * (bb - a) / 8.0);
			R[3] = F[0] * 0.5 * L * (-0.25 * SQ(a) + 0.25 * SQ(bb) + 0.5 * bb - 0.5 * a);
			R[4] = F[1] * 0.5 * L * (0.5 * bb - 0.5 * a + SQ(a) * SQ(a) / 16.0 - SQ(bb) * SQ(bb) / 16.0 - 3.0 * SQ(a) / 8.0 + 3.0 * SQ(bb) / 8.0);
			R[5] = F[1] * 0.5 * L * (L * (-SQ(a) * SQ(a) + SQ(bb) * SQ(bb)) / 32.0 + L * (-TH(a) + TH(bb)) / 24.0 - L * (-SQ(a) + SQ(bb)) / 16.0 - L * (bb - a) / 8.0);
			mat6_vec(Tt, R, RG);
			for (int i = 0; i < 6; i++) b->R[NODE_DOFS * e + i] += RG[i];
		}
	}
}

/* FEMBodyClass::buildGlobalMatrices, src/FEMBody.cpp:90-109: per element forceVector (src/FEMElement.cpp:72-97), massMatrix
 * (:100-107), stiffMatrix (:110-138), assembled with assembleGlobalMat (:163-186); element DOFs are 3e .. 3e+5 */
static void build_global_matrices(orc_fem *b) {
	const int dim = b->n_dof;
	memset(b->M, 0, sizeof(double) * dim * dim);
	memset(b->K, 0, sizeof(double) * dim * dim);
	memset(b->F, 0, sizeof(double) * dim);
	for (int e = 0; e < b->n_el; e++) {
		const double *T = b->T + 36 * e;
		const double L = b->L[e], L0 = b->L0[e], E = b->E[e], A = b->A[e], I = b->I[e];
		double Tt[36], tmp[36], G[36], G2[36];
		mat6_transpose(T, Tt);
		/* forceVector */
		const double u = (SQ(L) - SQ(L0)) / (L + L0);
		const double theta1 = shift_angle(b->angle[e] - b->elangle[e]);
		const double theta2 = shift_angle(b->angle[e + 1] - b->elangle[e]);
		const double F0 = (E * A / L0) * u;
		const double M1 = (2 * E * I / L0) * (2.0 * theta1 + theta2);
		const double M2 = (2 * E * I / L0) * (theta1 + 2.0 * theta2);
		double *F = b->Floc + 6 * e, FG[6];
		F[0] = -F0;
		F[1] = (1.0 / L0) * (M1 + M2);
		F[2] = M1;
		F[3] = F0;
		F[4] = -(1.0 / L0) * (M1 + M2);
		F[5] = M2;
		mat6_vec(Tt, F, FG);
		for (int i = 0; i < 6; i++) b->F[NODE_DOFS * e + i] += FG[i];
		/* massMatrix: Transpose(T) * M * T */
		mat6_mul(Tt, b->Mloc + 36 * e, tmp);
		mat6_mul(tmp, T, G);
		for (int i = 0; i < 6; i++)
			for (int j = 0; j < 6; j++) b->M[(NODE_DOFS * e + i) * dim + NODE_DOFS * e + j] += G[i * 6 + j];
		/* stiffMatrix: Transpose(T) * K_L * T + Transpose(T) * K_NL * T */
		mat6_mul(Tt, b->KLloc + 36 * e, tmp);
		mat6_mul(tmp, T, G);
		const double F0n = -F[0], V0 = F[4];
		double KNL[36];
		memset(KNL, 0, sizeof(KNL));
		KNL[0 * 6 + 1] = -V0 / L0;
		KNL[0 * 6 + 4] = V0 / L0;
		KNL[1 * 6 + 0] = -V0 / L0;
		KNL[1 * 6 + 1] = F0n / L0;
		KNL[1 * 6 + 3] = V0 / L0;
		KNL[1 * 6 + 4] = -F0n / L0;
		KNL[3 * 6 + 1] = V0 / L0;
		KNL[3 * 6 + 4] = -V0 / L0;
		KNL[4 * 6 + 0] = V0 / L0;
		KNL[4 * 6 + 1] = -F0n / L0;
		KNL[4 * 6 + 3] = -V0 / L0;
		KNL[4 * 6 + 4] = F0n / L0;
		mat6_mul(Tt, KNL, tmp);
		mat6_mul(tmp, T, G2);
		for (int i = 0; i < 6; i++)
			for (int j = 0; j < 6; j++) b->K[(NODE_DOFS * e + i) * dim + NODE_DOFS * e + j] += G[i * 6 + j] + G2[i * 6 + j];
	}
}

/* FEMBodyClass::setNewmark, src/FEMBody.cpp:112-126 */
static void set_newmark(orc_fem *b) {
	const int dim = b->n_dof;
	const double Dt = b->Dt;
	const double a0 = 1.0 / (b->alpha * SQ(Dt)), a2 = 1.0 / (b->alpha * Dt), a3 = 1.0 / (2.0 * b->alpha) - 1.0;
	double *v = b->work;
	for (int i = 0; i < dim; i++) v[i] = a0 * (b->U_n[i] - b->U[i]) + a2 * b->Udot[i] + a3 * b->Udotdot[i];
	for (int i = 0; i < dim; i++) {
		double s = 0.0;                                  /* Utils::MatMultiply, inc/Utils.h:286-307 */
		for (int j = 0; j < dim; j++) s += b->M[i * dim + j] * v[j];
		b->F[i] = b->R[i] - b->F[i] + s;
	}
	for (int i = 0; i < dim * dim; i++) b->K[i] = b->K[i] + a0 * b->M[i];
}

/* Utils::solveLAPACK(A, b, BC), src/Utils.cpp:288-311 */
static int solve_lapack(orc_fem *b) {
	if (!p_dgetrf || !p_dgetrs) return 1;
	const int dim = b->n_dof, BC = b->n_bc;
	char trans = 'T';
	int row = dim - BC, col = dim - BC, nrhs = 1, LDA = dim, LDB = dim, info = 0;
	int *ipiv = (int *)calloc((size_t)row, sizeof(int));
	memcpy(b->Kcopy, b->K, sizeof(double) * dim * dim);
	memcpy(b->delU, b->F, sizeof(double) * dim);
	p_dgetrf(&row, &col, b->Kcopy + BC * dim + BC, &LDA, ipiv, &info);
	p_dgetrs(&trans, &row, &nrhs, b->Kcopy + BC * dim + BC, &LDA, ipiv, b->delU + BC, &LDB, &info);
	for (int i = 0; i < BC; i++) b->delU[i] = 0.0;
	free(ipiv);
	return 0;
}

/* FEMBodyClass::finishNewmark, src/FEMBody.cpp:129-144 */
static void finish_newmark(orc_fem *b) {
	const double Dt = b->Dt;
	const double a6 = 1.0 / (b->alpha * SQ(Dt)), a7 = -1.0 / (b->alpha * Dt), a8 = -(1.0 / (2.0 * b->alpha) - 1.0);
	const double a9 = Dt * (1.0 - b->delta), a10 = b->delta * Dt;
	for (int i = 0; i < b->n_dof; i++) b->Udotdot[i] = a6 * (b->U[i] - b->U_n[i]) + a7 * b->Udot_n[i] + a8 * b->Udotdot_n[i];
	for (int i = 0; i < b->n_dof; i++) b->Udot[i] = b->Udot_n[i] + a9 * b->Udotdot_n[i] + a10 * b->Udotdot[i];
}

/* FEMElementClass::shapeFuns, src/FEMElement.cpp:141-160 */
static void shape_funs(const double *vec, double zeta, double L, double *res) {
	const double N0 = 1.0 - (zeta + 1.0) / 2.0;
	const double N1 = 1.0 - 3.0 * SQ((zeta + 1.0) / 2.0) + 2.0 * TH((zeta + 1.0) / 2.0);
	const double N2 = ((zeta + 1.0) / 2.0 - 2.0 * SQ((zeta + 1.0) / 2.0) + TH((zeta + 1.0) / 2.0)) * L;
	const double N3 = (zeta + 1.0) / 2.0;
	const double N4 = 3.0 * SQ((zeta + 1.0) / 2.0) - 2.0 * TH((zeta + 1.0) / 2.0);
	const double N5 = (-SQ((zeta + 1.0) / 2.0) + TH((zeta + 1.0) / 2.0)) * L;
	res[0] = vec[0] * N0 + vec[3] * N3;
	res[1] = vec[1] * N1 + vec[2] * N2 + vec[4] * N4 + vec[5] * N5;
}

/* FEMBodyClass::updateIBMValues, src/FEMBody.cpp:147-195: marker positions / velocities [2*n_ibm] from U, Udot */
static void update_ibm_values(orc_fem *b, double *pos_out, double *vel_out) {
	for (int i = 0; i < b->n_ibm; i++) {
		const int e = b->pm_el[i];
		const double zeta = b->pm_zeta[i];
		const double *T = b->T + 36 * e;
		double dU[6], dV[6], tU[6], tV[6], sU[2], sV[2];
		for (int k = 0; k < 6; k++) { dU[k] = b->U[NODE_DOFS * e + k]; dV[k] = b->Udot[NODE_DOFS * e + k]; }
		for (int n = 0; n < 2; n++) {
			dU[n * 3 + 0] += b->pos0[2 * (e + n)] - b->pos[2 * e];
			dU[n * 3 + 1] += b->pos0[2 * (e + n) + 1] - b->pos[2 * e + 1];
			dU[n * 3 + 2] += b->angle0[e + n] - b->elangle[e];
			dU[n * 3 + 2] = shift_angle(dU[n * 3 + 2]);
		}
		mat6_vec(T, dU, tU);
		mat6_vec(T, dV, tV);
		shape_funs(tU, zeta, b->L[e], sU);
		shape_funs(tV, zeta, b->L[e], sV);
		/* Transpose(Tsub) * v, Tsub = T[0..1][0..1] */
		for (int r = 0; r < 2; r++) {
			double su = 0.0, sv = 0.0;
			for (int c = 0; c < 2; c++) { su += T[c * 6 + r] * sU[c]; sv += T[c * 6 + r] * sV[c]; }
			pos_out[2 * i + r] = b->pos[2 * e + r] + su;
			vel_out[2 * i + r] = sv;
		}
	}
}

/* FEMBodyClass::subResidual, src/FEMBody.cpp:244-256 */
static void sub_residual(orc_fem *b) {
	const int dim = b->n_dof;
	memcpy(b->R_km1, b->R_k, sizeof(double) * dim);
	for (int i = 0; i < dim; i++) b->R_k[i] = b->U[i] - b->U_km1[i];
	double res = 0.0, num = 0.0, den = 0.0;
	for (int i = 0; i < dim; i++) res += b->R_k[i] * b->R_k[i];
	for (int i = 0; i < dim; i++) num += b->R_km1[i] * (b->R_k[i] - b->R_km1[i]);
	for (int i = 0; i < dim; i++) den += (b->R_k[i] - b->R_km1[i]) * (b->R_k[i] - b->R_km1[i]);
	b->subRes = res; b->subNum = num; b->subDen = den;
}

/* ---- exported --------------------------------------------------------------------------------------------------------------- */

/* dgetrf_ / dgetrs_ of the library at `path` (the OpenBLAS the compiled reference links, oracle/Makefile) */
int orc_fem_bind_lapack(const char *path) {
	void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return 1;
	p_dgetrf = (getrf_fn)dlsym(h, "dgetrf_");
	p_dgetrs = (getrs_fn)dlsym(h, "dgetrs_");
	return p_dgetrf && p_dgetrs ? 0 : 2;
}

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* consts: alpha, delta, Dt, Dm, gravityX, gravityY, ref_L; el [n_el*5] = L0, A, I, E, rho */
orc_fem *orc_fem_create(int n_nodes, int n_bc, int n_ibm, const double *consts, const double *pos0, const double *angle0,
                        const double *el, const int *pm_el, const double *pm_zeta, const int *fm_first, const int *fm_node,
                        const double *fm_z1, const double *fm_z2) {
	orc_fem *b = (orc_fem *)calloc(1, sizeof(orc_fem));
	const int ne = n_nodes - 1, dim = NODE_DOFS * n_nodes, nmap = fm_first[ne];
	b->n_nodes = n_nodes; b->n_el = ne; b->n_dof = dim; b->n_bc = n_bc; b->n_ibm = n_ibm;
	b->alpha = consts[0]; b->delta = consts[1]; b->Dt = consts[2]; b->Dm = consts[3];
	b->gravityX = consts[4]; b->gravityY = consts[5]; b->ref_L = consts[6];
	b->pos0 = dalloc(2 * n_nodes); b->angle0 = dalloc(n_nodes); b->pos = dalloc(2 * n_nodes); b->angle = dalloc(n_nodes);
	memcpy(b->pos0, pos0, sizeof(double) * 2 * n_nodes);
	memcpy(b->angle0, angle0, sizeof(double) * n_nodes);
	b->L0 = dalloc(ne); b->A = dalloc(ne); b->I = dalloc(ne); b->E = dalloc(ne); b->rho = dalloc(ne);
	b->L = dalloc(ne); b->elangle = dalloc(ne); b->T = dalloc(36 * ne); b->Floc = dalloc(6 * ne);
	b->Mloc = dalloc(36 * ne); b->KLloc = dalloc(36 * ne);
	for (int e = 0; e < ne; e++) {
		b->L0[e] = el[5 * e]; b->A[e] = el[5 * e + 1]; b->I[e] = el[5 * e + 2]; b->E[e] = el[5 * e + 3]; b->rho[e] = el[5 * e + 4];
		set_local_matrices(b, e);
	}
	b->pm_el = (int *)calloc(n_ibm ? n_ibm : 1, sizeof(int)); b->pm_zeta = dalloc(n_ibm);
	memcpy(b->pm_el, pm_el, sizeof(int) * n_ibm);
	memcpy(b->pm_zeta, pm_zeta, sizeof(double) * n_ibm);
	b->fm_first = (int *)calloc(ne + 1, sizeof(int)); b->fm_node = (int *)calloc(nmap ? nmap : 1, sizeof(int));
	b->fm_z1 = dalloc(nmap); b->fm_z2 = dalloc(nmap);
	memcpy(b->fm_first, fm_first, sizeof(int) * (ne + 1));
	memcpy(b->fm_node, fm_node, sizeof(int) * nmap);
	memcpy(b->fm_z1, fm_z1, sizeof(double) * nmap);
	memcpy(b->fm_z2, fm_z2, sizeof(double) * nmap);
	b->M = dalloc((size_t)dim * dim); b->K = dalloc((size_t)dim * dim); b->Kcopy = dalloc((size_t)dim * dim);
	b->R = dalloc(dim); b->F = dalloc(dim); b->delU = dalloc(dim); b->work = dalloc(dim);
	double **vecs[11] = {&b->U, &b->Udot, &b->Udotdot, &b->U_n, &b->Udot_n, &b->Udotdot_n, &b->U_km1, &b->R_k, &b->R_km1, &b->U_nm1, &b->U_nm2};
	for (int k = 0; k < 11; k++) *vecs[k] = dalloc(dim);
	update_fem_values(b);
	return b;
}

void orc_fem_destroy(orc_fem *b) {
	if (!b) return;
	double *d[] = {b->pos0, b->angle0, b->pos, b->angle, b->L0, b->A, b->I, b->E, b->rho, b->L, b->elangle, b->T, b->Floc, b->Mloc,
	               b->KLloc, b->pm_zeta, b->fm_z1, b->fm_z2, b->M, b->K, b->Kcopy, b->R, b->F, b->delU, b->work, b->U, b->Udot,
	               b->Udotdot, b->U_n, b->Udot_n, b->Udotdot_n, b->U_km1, b->R_k, b->R_km1, b->U_nm1, b->U_nm2};
	for (size_t k = 0; k < sizeof(d) / sizeof(d[0]); k++) free(d[k]);
	free(b->pm_el); free(b->fm_first); free(b->fm_node);
	free(b);
}

/* state [11 * n_dof] in the order of ref_fem_get_state: U, Udot, Udotdot, U_n, Udot_n, Udotdot_n, U_km1, R_k, R_km1, U_nm1, U_nm2 */
#define FEM_VECS 11
static double *state_vec(orc_fem *b, int k) {
	double *v[FEM_VECS] = {b->U, b->Udot, b->Udotdot, b->U_n, b->Udot_n, b->Udotdot_n, b->U_km1, b->R_k, b->R_km1, b->U_nm1, b->U_nm2};
	return v[k];
}
void orc_fem_set_state(orc_fem *b, const double *in) {
	for (int k = 0; k < FEM_VECS; k++) memcpy(state_vec(b, k), in + (size_t)k * b->n_dof, sizeof(double) * b->n_dof);
}
void orc_fem_get_state(orc_fem *b, double *out) {
	for (int k = 0; k < FEM_VECS; k++) memcpy(out + (size_t)k * b->n_dof, state_vec(b, k), sizeof(double) * b->n_dof);
}

/* FEMBodyClass::dynamicFEM, src/FEMBody.cpp:26-68 (newtonRaphsonDynamic :71-87, checkNRConvergence :237-241).
 * force / epsilon of the body's IBM nodes in; their new positions / velocities out; results [5] = subRes, subNum, subDen, resNR, itNR.
 * Returns non-zero if LAPACK has not been bound. */
int orc_fem_dynamic(orc_fem *b, const double *force, const double *epsilon, double *pos_out, double *vel_out, double *results) {
	const int dim = b->n_dof;
	memcpy(b->U, b->U_n, sizeof(double) * dim);
	memcpy(b->Udot, b->Udot_n, sizeof(double) * dim);
	memcpy(b->Udotdot, b->Udotdot_n, sizeof(double) * dim);
	update_fem_values(b);
	construct_r_vector(b, force, epsilon);
	const double TOL = 1e-10, MAXIT = 20;
	b->itNR = 0;
	do {
		build_global_matrices(b);
		set_newmark(b);
		if (solve_lapack(b)) return 1;
		for (int i = 0; i < dim; i++) b->U[i] = b->U[i] + b->delU[i];
		update_fem_values(b);
		double s = 0.0;
		for (int i = 0; i < dim; i++) s += b->delU[i] * b->delU[i];
		b->resNR = sqrt(s) / (b->ref_L * sqrt((double)dim));
		b->itNR++;
	} while (b->resNR > TOL && b->itNR < MAXIT);
	finish_newmark(b);
	update_ibm_values(b, pos_out, vel_out);
	sub_residual(b);
	results[0] = b->subRes; results[1] = b->subNum; results[2] = b->subDen; results[3] = b->resNR; results[4] = (double)b->itNR;
	return 0;
}

/* Start of a time step, sub-iteration 0 (src/Objects.cpp:160-174): resetValues (src/FEMBody.cpp:341-349) + predictor (:259-289) at
 * time step t */
void orc_fem_predict(orc_fem *b, int t, double *pos_out, double *vel_out) {
	const int dim = b->n_dof;
	double *tmp;
	tmp = b->U_nm2; b->U_nm2 = b->U_nm1; b->U_nm1 = tmp;      /* U_nm2.swap(U_nm1) */
	tmp = b->U_nm1; b->U_nm1 = b->U_n; b->U_n = tmp;          /* U_nm1.swap(U_n)   */
	tmp = b->U_n; b->U_n = b->U; b->U = tmp;                  /* U_n.swap(U)       */
	tmp = b->Udot_n; b->Udot_n = b->Udot; b->Udot = tmp;
	tmp = b->Udotdot_n; b->Udotdot_n = b->Udotdot; b->Udotdot = tmp;
	if (t > 2) for (int i = 0; i < dim; i++) b->U[i] = 2.5 * b->U_n[i] - 2.0 * b->U_nm1[i] + 0.5 * b->U_nm2[i];
	else if (t == 2) for (int i = 0; i < dim; i++) b->U[i] = 2.0 * b->U_n[i] - b->U_nm1[i];
	else if (t == 1) memcpy(b->U, b->U_n, sizeof(double) * dim);
	update_fem_values(b);
	finish_newmark(b);
	update_ibm_values(b, pos_out, vel_out);
	tmp = b->U_km1; b->U_km1 = b->U; b->U = tmp;              /* U_km1.swap(U) */
}

/* Aitken-relaxed update of a later sub-iteration (src/Objects.cpp:195-208) with the global relaxation factor `relax` */
void orc_fem_relax(orc_fem *b, double relax, double *pos_out, double *vel_out) {
	for (int i = 0; i < b->n_dof; i++) b->U[i] = b->U_km1[i] + relax * (b->U[i] - b->U_km1[i]);
	update_fem_values(b);
	finish_newmark(b);
	update_ibm_values(b, pos_out, vel_out);
	double *tmp = b->U_km1; b->U_km1 = b->U; b->U = tmp;
}
