/*
 * life_oracle.c — CPU restatement of LIFE's hot path.      *** TEST INFRASTRUCTURE, NOT PRODUCT CODE ***
 * See life_oracle.h for scope, citations and how this file is pinned against the compiled reference.
 *
 * Arithmetic is written in the operation order of the reference so that the BGK paths reproduce the reference's
 * doubles bit for bit when both are compiled without FMA contraction (baseline x86-64; this file is built with
 * -ffp-contract=off) — both collision operators.
 */
#define _USE_MATH_DEFINES
#include "life_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define NV 9
#define ND 2
#define SUPP 9                       /* suppSize, inc/defs.h:39 */
#define SQ(x) ((x) * (x))
#define TH(x) ((x) * (x) * (x))
#define QU(x) ((x) * (x) * (x) * (x))

/* D2Q9 constants, Grid.cpp:1247-1250 */
static const int CX[NV] = {0, 1, -1, 0, 0, 1, -1, 1, -1};
static const int CY[NV] = {0, 0, 0, 1, -1, 1, -1, -1, 1};
static const int OPP[NV] = {0, 2, 1, 4, 3, 6, 5, 8, 7};

typedef struct orc_marker {
	double pos[2], vel[2], force[2];
	double ds, epsilon;
	double interpRho, interpMom[2];
	int suppCount;
	int sidx[SUPP], sjdx[SUPP];
	double sdirac[SUPP];
} orc_marker;

struct orc_grid {
	orc_params p;
	int64_t Nx, Ny;
	int t;
	double c_s, w[NV];
	double tau, nu, Dx, Dt, Dm, Drho;
	double *f, *f_n, *u, *u_n, *rho, *rho_n, *force_xy, *force_ibm;
	int32_t *type;
	int64_t *BCVec;
	int64_t nBC;
	double *delU, *u_in, *rho_in;
	orc_marker *mk;
	int64_t nMarkers;
};

/* ------------------------------------------------------------------------------------------------------------------ */
/* equilibrium, Grid.cpp:249-264 (reads u_n / rho_n of the node)                                                      */
static inline double feq_bgk(const orc_grid *g, double rho, double ux, double uy, int v) {
	int cx = CX[v], cy = CY[v];
	return rho * g->w[v] * (1.0 + 3.0 * (cx * ux + cy * uy) + 4.5 * (SQ(ux) * (SQ(cx) - 1.0 / 3.0) + SQ(uy) * (SQ(cy) - 1.0 / 3.0)) + 9.0 * cx * cy * ux * uy);
}
static inline double feq_cm(const orc_grid *g, double rho, double ux, double uy, int v) {
	int cx = CX[v], cy = CY[v];
	return 0.25 * rho * g->w[v] * (9.0 * SQ(cx) * SQ(ux) + 6.0 * cx * ux - 3.0 * SQ(ux) + 2.0) * (9.0 * SQ(cy) * SQ(uy) + 6.0 * cy * uy - 3.0 * SQ(uy) + 2.0);
}
static inline double equilibrium(const orc_grid *g, int64_t id, int v) {
	double ux = g->u_n[id * ND], uy = g->u_n[id * ND + 1];
	return g->p.central_moments ? feq_cm(g, g->rho_n[id], ux, uy, v) : feq_bgk(g, g->rho_n[id], ux, uy, v);
}

/* Guo-type forcing projected on the lattice, Grid.cpp:267-279 (BGK only) */
static inline double lattice_force(const orc_grid *g, int64_t id, int v) {
	int cx = CX[v], cy = CY[v];
	double ux = g->u_n[id * ND], uy = g->u_n[id * ND + 1];
	double Fx = g->force_xy[id * ND] + g->force_ibm[id * ND];
	double Fy = g->force_xy[id * ND + 1] + g->force_ibm[id * ND + 1];
	return 3.0 * g->w[v] * (Fx * (cx - ux + cx * 3.0 * (cx * ux + cy * uy)) + (Fy * (cy - uy + cy * 3.0 * (cx * ux + cy * uy))));
}

/* push target with unconditional periodic wrap, Grid.cpp:229 / :240 */
static inline int64_t recv_id(const orc_grid *g, int64_t i, int64_t j, int v) {
	return ((i + CX[v] + g->Nx) % g->Nx) * g->Ny + ((j + CY[v] + g->Ny) % g->Ny);
}

/*
 * Central-moments collision, Grid.cpp:106-233.
 * Pre-collision moments k4Pre, k5Pre: the reference's loop (Grid.cpp:113-122), same order.
 * Post-collision central moments k0..k8: Grid.cpp:125-133.
 * Back-transform: the reference writes out nine 9-term polynomials in (ux, uy) (Grid.cpp:143-223): the expansion of
 *   central moments --(binomial shift by u)--> raw moments --(D2Q9 inverse moment matrix)--> populations.
 * They are restated term by term in the reference's association (round 2; round 1 evaluated the two stages as such, 1e-15 away),
 * so the central-moments path is bit-identical to the compiled reference as well (tests/test_oracle_vs_ref.py).
 */
static void cm_collide(const orc_grid *g, int64_t id, double fStar[NV]) {
	const double *fn = g->f_n + id * NV;
	double ux = g->u_n[id * ND], uy = g->u_n[id * ND + 1];
	double k4Pre = 0.0, k5Pre = 0.0;
	for (int v = 0; v < NV; v++) {
		double cx = CX[v] - ux;
		double cy = CY[v] - uy;
		k4Pre += fn[v] * (SQ(cx) - SQ(cy));
		k5Pre += fn[v] * cx * cy;
	}
	double Fx = g->force_xy[id * ND] + g->force_ibm[id * ND];
	double Fy = g->force_xy[id * ND + 1] + g->force_ibm[id * ND + 1];
	double rho = g->rho_n[id];
	double k0 = rho;
	double k1 = 0.5 * Fx;
	double k2 = 0.5 * Fy;
	double k3 = 2.0 * rho * SQ(g->c_s);
	double k4 = (1.0 - g->p.omega) * k4Pre;
	double k5 = (1.0 - g->p.omega) * k5Pre;
	double k6 = 0.5 * Fy * SQ(g->c_s);
	double k7 = 0.5 * Fx * SQ(g->c_s);
	double k8 = rho * QU(g->c_s);

	/* back-transform: the reference's nine expanded polynomials (Grid.cpp:143-223), every sum and product associated as the
	 * reference's expressions parse; x = ux, y = uy, xx = SQ(ux), yy = SQ(uy) */
	const double x = ux, y = uy, xx = ux * ux, yy = uy * uy;
	fStar[0] = (xx * yy - xx - yy + 1.0) * k0
	     + (2.0 * x * yy - 2.0 * x) * k1
	     + (2.0 * y * xx - 2.0 * y) * k2
	     + (0.5 * xx + 0.5 * yy - 1.0) * k3
	     + (0.5 * yy - 0.5 * xx) * k4
	     + 4.0 * x * y * k5
	     + 2.0 * y * k6
	     + 2.0 * x * k7
	     + k8;
	fStar[1] = (0.5 * xx - 0.5 * (xx * yy) - 0.5 * (x * yy) + 0.5 * x) * k0
	     + (x - x * yy - 0.5 * yy + 0.5) * k1
	     + (-y * xx - y * x) * k2
	     + (-0.25 * xx - 0.25 * x - 0.25 * yy + 0.25) * k3
	     + (0.25 * xx + 0.25 * x - 0.25 * yy + 0.25) * k4
	     + (-y - 2.0 * x * y) * k5
	     + (-y) * k6
	     + (-x - 0.5) * k7
	     - 0.5 * k8;
	fStar[2] = (-0.5 * (xx * yy) + 0.5 * xx + 0.5 * (x * yy) - 0.5 * x) * k0
	     + (x - x * yy + 0.5 * yy - 0.5) * k1
	     + (-y * xx + y * x) * k2
	     + (-0.25 * xx + 0.25 * x - 0.25 * yy + 0.25) * k3
	     + (0.25 * xx - 0.25 * x - 0.25 * yy + 0.25) * k4
	     + (y - 2.0 * x * y) * k5
	     + (-y) * k6
	     + (0.5 - x) * k7
	     - 0.5 * k8;
	fStar[3] = (0.5 * yy - 0.5 * (xx * y) - 0.5 * (xx * yy) + 0.5 * y) * k0
	     + (-x * yy - x * y) * k1
	     + (y - xx * y - 0.5 * xx + 0.5) * k2
	     + (-0.25 * xx - 0.25 * yy - 0.25 * y + 0.25) * k3
	     + (0.25 * xx - 0.25 * yy - 0.25 * y - 0.25) * k4
	     + (-x - 2.0 * x * y) * k5
	     + (-y - 0.5) * k6
	     + (-x) * k7
	     - 0.5 * k8;
	fStar[4] = (-0.5 * (xx * yy) + 0.5 * (xx * y) + 0.5 * yy - 0.5 * y) * k0
	     + (-x * yy + x * y) * k1
	     + (y - xx * y + 0.5 * xx - 0.5) * k2
	     + (-0.25 * xx - 0.25 * yy + 0.25 * y + 0.25) * k3
	     + (0.25 * xx - 0.25 * yy + 0.25 * y - 0.25) * k4
	     + (x - 2.0 * x * y) * k5
	     + (0.5 - y) * k6
	     + (-x) * k7
	     - 0.5 * k8;
	fStar[5] = (0.25 * (xx * yy) + 0.25 * (xx * y) + 0.25 * (x * yy) + 0.25 * (x * y)) * k0
	     + (0.25 * y + 0.5 * (x * y) + 0.5 * (x * yy) + 0.25 * yy) * k1
	     + (0.25 * x + 0.5 * (x * y) + 0.5 * (xx * y) + 0.25 * xx) * k2
	     + (0.125 * xx + 0.125 * x + 0.125 * yy + 0.125 * y) * k3
	     + (-0.125 * xx - 0.125 * x + 0.125 * yy + 0.125 * y) * k4
	     + (0.5 * x + 0.5 * y + x * y + 0.25) * k5
	     + (0.5 * y + 0.25) * k6
	     + (0.5 * x + 0.25) * k7
	     + 0.25 * k8;
	fStar[6] = (0.25 * (xx * yy) - 0.25 * (xx * y) - 0.25 * (x * yy) + 0.25 * (x * y)) * k0
	     + (0.25 * y - 0.5 * (x * y) + 0.5 * (x * yy) - 0.25 * yy) * k1
	     + (0.25 * x - 0.5 * (x * y) + 0.5 * (xx * y) - 0.25 * xx) * k2
	     + (0.125 * xx - 0.125 * x + 0.125 * yy - 0.125 * y) * k3
	     + (-0.125 * xx + 0.125 * x + 0.125 * yy - 0.125 * y) * k4
	     + (x * y - 0.5 * y - 0.5 * x + 0.25) * k5
	     + (0.5 * y - 0.25) * k6
	     + (0.5 * x - 0.25) * k7
	     + 0.25 * k8;
	fStar[7] = (0.25 * (xx * yy) - 0.25 * (xx * y) + 0.25 * (x * yy) - 0.25 * (x * y)) * k0
	     + (0.5 * (x * yy) - 0.5 * (x * y) - 0.25 * y + 0.25 * yy) * k1
	     + (0.5 * (x * y) - 0.25 * x + 0.5 * (xx * y) - 0.25 * xx) * k2
	     + (0.125 * xx + 0.125 * x + 0.125 * yy - 0.125 * y) * k3
	     + (-0.125 * xx - 0.125 * x + 0.125 * yy - 0.125 * y) * k4
	     + (0.5 * y - 0.5 * x + x * y - 0.25) * k5
	     + (0.5 * y - 0.25) * k6
	     + (0.5 * x + 0.25) * k7
	     + 0.25 * k8;
	fStar[8] = (0.25 * (xx * yy) + 0.25 * (xx * y) - 0.25 * (x * yy) - 0.25 * (x * y)) * k0
	     + (0.5 * (x * y) - 0.25 * y + 0.5 * (x * yy) - 0.25 * yy) * k1
	     + (0.5 * (xx * y) - 0.5 * (x * y) - 0.25 * x + 0.25 * xx) * k2
	     + (0.125 * xx - 0.125 * x + 0.125 * yy + 0.125 * y) * k3
	     + (-0.125 * xx + 0.125 * x + 0.125 * yy + 0.125 * y) * k4
	     + (0.5 * x - 0.5 * y + x * y - 0.25) * k5
	     + (0.5 * y + 0.25) * k6
	     + (0.5 * x - 0.25) * k7
	     + 0.25 * k8;
}

/* stream + collide of one node (push), Grid.cpp:103-246 */
static inline void stream_collide(orc_grid *g, int64_t i, int64_t j, int64_t id) {
	if (g->p.central_moments) {
		double fStar[NV];
		cm_collide(g, id, fStar);
		for (int v = 0; v < NV; v++)
			g->f[recv_id(g, i, j, v) * NV + v] = fStar[v];
	} else {
		for (int v = 0; v < NV; v++) {
			double fn = g->f_n[id * NV + v];
			g->f[recv_id(g, i, j, v) * NV + v] = fn + g->p.omega * (equilibrium(g, id, v) - fn) + (1.0 - 0.5 * g->p.omega) * lattice_force(g, id, v);
		}
	}
}

/* rho = sum f, u = (sum c f + F_xy/2)/rho, Grid.cpp:282-299 */
static inline void macroscopic(orc_grid *g, int64_t id) {
	double r = 0.0, mx = 0.0, my = 0.0;
	for (int v = 0; v < NV; v++) {
		r += g->f[id * NV + v];
		mx += CX[v] * g->f[id * NV + v];
		my += CY[v] * g->f[id * NV + v];
	}
	g->rho[id] = r;
	g->u[id * ND] = (mx + 0.5 * g->force_xy[id * ND]) / r;
	g->u[id * ND + 1] = (my + 0.5 * g->force_xy[id * ND + 1]) / r;
}

/* getNormalVector, Grid.cpp:498-545.  Returns normalDirection (0 = x, 1 = y), -1 for the reference's ERROR case. */
static int normal_vector(const orc_grid *g, int64_t i, int64_t j, int n[2]) {
	int dir = 0;
	n[0] = n[1] = 0;
	if (i == 0) { n[0] = 1; dir = 0; }
	else if (i == g->Nx - 1) { n[0] = -1; dir = 0; }
	if (j == 0) { n[1] = 1; dir = 1; }
	else if (j == g->Ny - 1) { n[1] = -1; dir = 1; }
	if (n[0] != 0 && n[1] != 0) {
		int tx = g->type[(i + n[0]) * g->Ny + j];
		int ty = g->type[i * g->Ny + j + n[1]];
		if (tx == ORC_FLUID && ty == ORC_FLUID) return -1;
		else if (tx == ORC_FLUID) { dir = 0; n[1] = 0; }
		else if (ty == ORC_FLUID) { dir = 1; n[0] = 0; }
	}
	return dir;
}

/* Utils::extrapolate order 1 (Utils.h:151-161) and Utils::zeroGradient order 2 (Utils.h:200-210) */
static inline double extrapolate1(const orc_grid *g, const double *vec, const int n[2], int64_t i, int64_t j, int d, int adims) {
	int64_t i1 = i + n[0], i2 = i + 2 * n[0], j1 = j + n[1], j2 = j + 2 * n[1];
	return 2.0 * vec[(i1 * g->Ny + j1) * adims + d] - vec[(i2 * g->Ny + j2) * adims + d];
}
static inline double zero_gradient2(const orc_grid *g, const double *vec, const int n[2], int64_t i, int64_t j, int d, int adims) {
	int64_t i1 = i + n[0], i2 = i + 2 * n[0], j1 = j + n[1], j2 = j + 2 * n[1];
	return (4.0 / 3.0) * vec[(i1 * g->Ny + j1) * adims + d] - (1.0 / 3.0) * vec[(i2 * g->Ny + j2) * adims + d];
}

/* regularisedBC, Grid.cpp:387-465 */
static void regularised_bc(orc_grid *g, int64_t i, int64_t j, int64_t id, const int n[2], int nd) {
	int ty = g->type[id];
	int corner = (n[0] != 0 && n[1] != 0);
	double *f = g->f + id * NV;
	if (corner) {
		if (ty == ORC_VELOCITY || ty == ORC_WALL || ty == ORC_FREESLIP)
			g->rho_n[id] = extrapolate1(g, g->rho, n, i, j, 0, 1);
	} else {
		double fplus = 0.0, fzero = 0.0;
		for (int v = 0; v < NV; v++) {
			int cn = nd == 0 ? CX[v] : CY[v];
			if (cn == -n[nd]) fplus += f[v];
			else if (cn == 0) fzero += f[v];
		}
		if (ty == ORC_VELOCITY || ty == ORC_WALL || ty == ORC_FREESLIP)
			g->rho_n[id] = (2.0 * fplus + fzero) / (1.0 - n[nd] * g->u_n[id * ND + nd]);
		else if (ty == ORC_PRESSURE)
			g->u_n[id * ND + nd] = n[nd] * (1.0 - (2.0 * fplus + fzero) / g->rho_n[id]);
	}
	double Sxx = 0.0, Syy = 0.0, Sxy = 0.0;
	for (int v = 0; v < NV; v++) {
		double feq = equilibrium(g, id, v);
		if (corner) {
			if (CX[v] == n[0] || CY[v] == n[1]) {
				if (n[0] * CX[v] + n[1] * CY[v] == 0) f[v] = feq;
				else f[v] = feq + (f[OPP[v]] - equilibrium(g, id, OPP[v]));
			}
		} else {
			int cn = nd == 0 ? CX[v] : CY[v];
			if (cn == n[nd]) f[v] = feq + (f[OPP[v]] - equilibrium(g, id, OPP[v]));
		}
		double fneq = f[v] - feq;
		Sxx += CX[v] * CX[v] * fneq;
		Syy += CY[v] * CY[v] * fneq;
		Sxy += CX[v] * CY[v] * fneq;
	}
	for (int v = 0; v < NV; v++)
		f[v] = equilibrium(g, id, v) + (g->w[v] / (2.0 * QU(g->c_s))) * (((SQ(CX[v]) - SQ(g->c_s)) * Sxx) + ((SQ(CY[v]) - SQ(g->c_s)) * Syy) + (2.0 * CX[v] * CY[v] * Sxy));
}

/* getRampCoefficient, Grid.cpp:548-556 */
static double ramp_coefficient(const orc_grid *g) {
	if (g->p.inlet_ramp > 0.0 && g->Dt * g->t <= g->p.inlet_ramp)
		return (1.0 - cos(M_PI * g->Dt * g->t / g->p.inlet_ramp)) / 2.0;
	return 1.0;
}

/* applyBCs, Grid.cpp:302-384 (+ convectiveBC :468-474) */
static void apply_bcs(orc_grid *g, int64_t i, int64_t j, int64_t id) {
	double ramp = ramp_coefficient(g);
	int n[2];
	int nd = normal_vector(g, i, j, n);
	switch (g->type[id]) {
	case ORC_WALL:
		g->u_n[id * ND] = 0.0;
		g->u_n[id * ND + 1] = 0.0;
		regularised_bc(g, i, j, id, n, nd);
		break;
	case ORC_VELOCITY:
		g->u_n[id * ND] = g->u_in[j * ND] * ramp;
		g->u_n[id * ND + 1] = g->u_in[j * ND + 1] * ramp;
		regularised_bc(g, i, j, id, n, nd);
		break;
	case ORC_FREESLIP:
		g->u_n[id * ND + nd] = 0.0;
		for (int d = 0; d < ND; d++)
			if (d != nd) g->u_n[id * ND + d] = zero_gradient2(g, g->u, n, i, j, d, ND);
		regularised_bc(g, i, j, id, n, nd);
		break;
	case ORC_PRESSURE:
		g->rho_n[id] = g->rho_in[j];
		for (int d = 0; d < ND; d++)
			if (d != nd) g->u_n[id * ND + d] = zero_gradient2(g, g->u, n, i, j, d, ND);
		regularised_bc(g, i, j, id, n, nd);
		break;
	case ORC_CONVECTIVE:
		for (int k = 0; k < 3; k++) {
			static const int vs[3] = {2, 6, 8};
			int v = vs[k];
			g->f[id * NV + v] = g->f_n[id * NV + v] + 3.0 * g->w[v] * (g->delU[j * ND] * CX[v] + g->delU[j * ND + 1] * CY[v]);
		}
		break;
	default:
		break;
	}
}

/* convectiveSpeed, Grid.cpp:477-495 */
static void convective_speed(orc_grid *g) {
	int64_t Nx = g->Nx, Ny = g->Ny;
	double uOut = 0.0;
	for (int64_t j = 0; j < Ny; j++) uOut += g->u[((Nx - 1) * Ny + j) * ND];
	uOut /= (double)Ny;
#pragma omp parallel for schedule(static)
	for (int64_t j = 0; j < Ny; j++) {
		for (int d = 0; d < ND; d++)
			g->delU[j * ND + d] = (-uOut / 2.0) * (3.0 * g->u[((Nx - 1) * Ny + j) * ND + d] - 4.0 * g->u[((Nx - 2) * Ny + j) * ND + d] + g->u[((Nx - 3) * Ny + j) * ND + d]);
	}
}

/* lbmKernel, Grid.cpp:36-100 */
void orc_lbm_kernel(orc_grid *g) {
	int64_t Nx = g->Nx, Ny = g->Ny;
	if (g->p.wall_right == ORC_CONVECTIVE) convective_speed(g);
	double *tmp;
	tmp = g->f_n; g->f_n = g->f; g->f = tmp;
	tmp = g->u_n; g->u_n = g->u; g->u = tmp;
	tmp = g->rho_n; g->rho_n = g->rho; g->rho = tmp;
#pragma omp parallel
	{
		if (g->p.womersley > 0.0) {
			double W = g->p.womersley;
#pragma omp for schedule(static)
			for (int64_t id = 0; id < Nx * Ny; id++) {
				g->force_xy[id * ND] = (g->rho_n[id] * g->Drho * g->p.gravityX + g->p.dpdx * cos(2.0 * M_PI * g->t * g->Dt / ((SQ(g->p.height_p) * M_PI) / (2.0 * SQ(W) * g->p.nu_p)))) * SQ(g->Dx * g->Dt) / g->Dm;
				g->force_xy[id * ND + 1] = (g->rho_n[id] * g->Drho * g->p.gravityY + g->p.dpdy * cos(2.0 * M_PI * g->t * g->Dt / ((SQ(g->p.height_p) * M_PI) / (2.0 * SQ(W) * g->p.nu_p)))) * SQ(g->Dx * g->Dt) / g->Dm;
			}
		}
#pragma omp for schedule(static)
		for (int64_t i = 0; i < Nx; i++)
			for (int64_t j = 0; j < Ny; j++)
				stream_collide(g, i, j, i * Ny + j);
#pragma omp for schedule(static)
		for (int64_t id = 0; id < Nx * Ny; id++)
			if (g->type[id] == ORC_FLUID) macroscopic(g, id);
#pragma omp for schedule(static)
		for (int64_t b = 0; b < g->nBC; b++) {
			int64_t id = g->BCVec[b];
			int64_t i = id / Ny, j = id - i * Ny;
			apply_bcs(g, i, j, id);
			macroscopic(g, id);
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* constructor + initialiseGrid, Grid.cpp:1232-1289 and 916-1062                                                      */
orc_grid *orc_create(const orc_params *p) {
	orc_grid *g = (orc_grid *)calloc(1, sizeof(orc_grid));
	g->p = *p;
	int64_t Nx = g->Nx = p->Nx, Ny = g->Ny = p->Ny, N = Nx * Ny;
	double rho0 = 1.0;
	g->c_s = 1.0 / sqrt(3.0);
	static const double W0[NV] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
	memcpy(g->w, W0, sizeof(W0));
	g->t = 0;
	g->tau = 1.0 / p->omega;
	g->nu = (g->tau - 0.5) * SQ(g->c_s);
	g->Dx = p->height_p / (Ny - 1);
	g->Dt = SQ(g->Dx) * g->nu / p->nu_p;
	g->Dm = (p->rho_p / rho0) * TH(g->Dx);
	g->Drho = (p->rho_p / rho0);

	g->u = (double *)calloc(N * ND, sizeof(double));
	g->u_n = (double *)calloc(N * ND, sizeof(double));
	g->rho = (double *)malloc(N * sizeof(double));
	g->rho_n = (double *)malloc(N * sizeof(double));
	g->force_xy = (double *)calloc(N * ND, sizeof(double));
	g->force_ibm = (double *)calloc(N * ND, sizeof(double));
	g->type = (int32_t *)calloc(N, sizeof(int32_t));
	g->f = (double *)calloc(N * NV, sizeof(double));
	g->f_n = (double *)calloc(N * NV, sizeof(double));
	g->u_in = (double *)calloc(Ny * ND, sizeof(double));
	g->rho_in = (double *)malloc(Ny * sizeof(double));
	g->delU = (double *)calloc(Ny * ND, sizeof(double));
	g->BCVec = (int64_t *)malloc((2 * (Nx + Ny)) * sizeof(int64_t));
	for (int64_t k = 0; k < N; k++) g->rho[k] = g->rho_n[k] = rho0;
	for (int64_t k = 0; k < Ny; k++) g->rho_in[k] = rho0;

	/* type matrix and BC list (Grid.cpp:925-951): left/right first, bottom/top override at the corners */
	g->nBC = 0;
	for (int64_t i = 0; i < Nx; i++)
		for (int64_t j = 0; j < Ny; j++) {
			int64_t id = i * Ny + j;
			if (i == 0) g->type[id] = p->wall_left;
			else if (i == Nx - 1) g->type[id] = p->wall_right;
			if (j == 0) g->type[id] = p->wall_bottom;
			else if (j == Ny - 1) g->type[id] = p->wall_top;
			if (g->type[id] != ORC_FLUID) g->BCVec[g->nBC++] = id;
		}

	/* inlet profile (Grid.cpp:954-996) */
	for (int64_t j = 0; j < Ny; j++) {
		if (p->profile == ORC_PROFILE_UNIFORM) {
			g->u_in[j * ND] = p->uxInlet_p * g->Dt / g->Dx;
			g->u_in[j * ND + 1] = p->uyInlet_p * g->Dt / g->Dx;
		} else if (p->profile == ORC_PROFILE_PARABOLIC) {
			double R = p->height_p / 2.0;
			double YPos = j * g->Dx - R;
			g->u_in[j * ND] = 1.5 * (p->uxInlet_p * g->Dt / g->Dx) * (1.0 - SQ(YPos / R));
			g->u_in[j * ND + 1] = 1.5 * (p->uyInlet_p * g->Dt / g->Dx) * (1.0 - SQ(YPos / R));
		} else if (p->profile == ORC_PROFILE_SHEAR) {
			double H = p->height_p;
			double YPos = j * g->Dx;
			g->u_in[j * ND] = (p->uxInlet_p * g->Dt / g->Dx) * (YPos / H);
			g->u_in[j * ND + 1] = (p->uyInlet_p * g->Dt / g->Dx) * (YPos / H);
		} else if (p->profile == ORC_PROFILE_BOUNDARYLAYER) {
			double H = p->height_p;
			double YPos = j * g->Dx;
			g->u_in[j * ND] = ((1.5 * p->uxInlet_p * g->Dt / g->Dx) / SQ(H)) * YPos * (2.0 * H - YPos);
			g->u_in[j * ND + 1] = ((1.5 * p->uyInlet_p * g->Dt / g->Dx) / SQ(H)) * YPos * (2.0 * H - YPos);
		}
	}

	/* initial velocity (Grid.cpp:999-1028) */
	for (int64_t i = 0; i < Nx; i++)
		for (int64_t j = 0; j < Ny; j++) {
			int64_t id = i * Ny + j;
			if (p->inlet_ramp > 0.0) {
				g->u[id * ND] = 0.0;
				g->u[id * ND + 1] = 0.0;
			} else if (p->profile != ORC_PROFILE_UNIFORM) {
				g->u[id * ND] = g->u_in[j * ND];
				g->u[id * ND + 1] = g->u_in[j * ND + 1];
			} else {
				g->u[id * ND] = p->ux0_p * g->Dt / g->Dx;
				g->u[id * ND + 1] = p->uy0_p * g->Dt / g->Dx;
			}
			if (g->type[id] == ORC_WALL) {
				g->u[id * ND] = 0.0;
				g->u[id * ND + 1] = 0.0;
			}
		}
	memcpy(g->u_n, g->u, N * ND * sizeof(double));
	memcpy(g->rho_n, g->rho, N * sizeof(double));

	/* Cartesian body force (Grid.cpp:1035-1045) */
	for (int64_t id = 0; id < N; id++) {
		g->force_xy[id * ND] = (g->rho[id] * g->Drho * p->gravityX + p->dpdx) * SQ(g->Dx * g->Dt) / g->Dm;
		g->force_xy[id * ND + 1] = (g->rho[id] * g->Drho * p->gravityY + p->dpdy) * SQ(g->Dx * g->Dt) / g->Dm;
	}

	/* populations at equilibrium (Grid.cpp:1048-1061) */
	for (int64_t id = 0; id < N; id++)
		for (int v = 0; v < NV; v++) g->f[id * NV + v] = equilibrium(g, id, v);
	memcpy(g->f_n, g->f, N * NV * sizeof(double));
	return g;
}

void orc_destroy(orc_grid *g) {
	if (!g) return;
	free(g->f); free(g->f_n); free(g->u); free(g->u_n); free(g->rho); free(g->rho_n);
	free(g->force_xy); free(g->force_ibm); free(g->type); free(g->BCVec);
	free(g->delU); free(g->u_in); free(g->rho_in); free(g->mk);
	free(g);
}

void orc_scalings(const orc_grid *g, double *o) {
	o[0] = g->Dx; o[1] = g->Dt; o[2] = g->Dm; o[3] = g->Drho; o[4] = g->tau; o[5] = g->nu;
}
int32_t orc_get_t(const orc_grid *g) { return g->t; }
void orc_set_t(orc_grid *g, int32_t t) { g->t = t; }

double *orc_array(orc_grid *g, int32_t which, int64_t *len) {
	int64_t N = g->Nx * g->Ny, n = 0;
	double *p = NULL;
	switch (which) {
	case 0: p = g->f; n = N * NV; break;
	case 1: p = g->f_n; n = N * NV; break;
	case 2: p = g->rho; n = N; break;
	case 3: p = g->rho_n; n = N; break;
	case 4: p = g->u; n = N * ND; break;
	case 5: p = g->u_n; n = N * ND; break;
	case 6: p = g->force_xy; n = N * ND; break;
	case 7: p = g->force_ibm; n = N * ND; break;
	case 8: p = g->u_in; n = g->Ny * ND; break;
	case 9: p = g->rho_in; n = g->Ny; break;
	case 10: p = g->delU; n = g->Ny * ND; break;
	default: break;
	}
	if (len) *len = n;
	return p;
}
int32_t *orc_types(orc_grid *g) { return g->type; }
int64_t orc_bc_count(const orc_grid *g) { return g->nBC; }
const int64_t *orc_bc_ids(const orc_grid *g) { return g->BCVec; }
int32_t orc_normal(const orc_grid *g, int64_t i, int64_t j, int32_t *nx, int32_t *ny) {
	int n[2];
	int d = normal_vector(g, i, j, n);
	*nx = n[0]; *ny = n[1];
	return d;
}
int64_t orc_stream_target(const orc_grid *g, int64_t i, int64_t j, int32_t v) { return recv_id(g, i, j, v); }

/* ------------------------------------------------------------------------------------------------------------------ */
/* markers                                                                                                            */

/* Utils::diracDelta, Utils.h:220-232 */
double orc_dirac_delta(double dist) {
	double a = fabs(dist);
	if (a > 1.5) return 0.0;
	else if (a > 0.5) return (5.0 - 3.0 * a - sqrt(-3.0 * SQ(1.0 - a) + 1.0)) / 6.0;
	else return (1.0 + sqrt(1.0 - 3.0 * SQ(a))) / 3.0;
}

void orc_set_markers(orc_grid *g, int64_t n, const double *pos, const double *vel, const double *ds, const double *eps) {
	if (n != g->nMarkers) {
		free(g->mk);
		g->mk = (orc_marker *)calloc(n > 0 ? n : 1, sizeof(orc_marker));
		g->nMarkers = n;
	}
	for (int64_t m = 0; m < n; m++) {
		if (pos) { g->mk[m].pos[0] = pos[2 * m]; g->mk[m].pos[1] = pos[2 * m + 1]; }
		if (vel) { g->mk[m].vel[0] = vel[2 * m]; g->mk[m].vel[1] = vel[2 * m + 1]; }
		if (ds) g->mk[m].ds = ds[m];
		if (eps) g->mk[m].epsilon = eps[m];
	}
}
void orc_get_marker_force(const orc_grid *g, double *force) {
	for (int64_t m = 0; m < g->nMarkers; m++) { force[2 * m] = g->mk[m].force[0]; force[2 * m + 1] = g->mk[m].force[1]; }
}
void orc_set_marker_force(orc_grid *g, const double *force) {
	for (int64_t m = 0; m < g->nMarkers; m++) { g->mk[m].force[0] = force[2 * m]; g->mk[m].force[1] = force[2 * m + 1]; }
}
void orc_get_interp(const orc_grid *g, double *rho, double *mom) {
	for (int64_t m = 0; m < g->nMarkers; m++) {
		rho[m] = g->mk[m].interpRho;
		mom[2 * m] = g->mk[m].interpMom[0]; mom[2 * m + 1] = g->mk[m].interpMom[1];
	}
}
void orc_get_ds_eps(const orc_grid *g, double *ds, double *eps) {
	for (int64_t m = 0; m < g->nMarkers; m++) { if (ds) ds[m] = g->mk[m].ds; if (eps) eps[m] = g->mk[m].epsilon; }
}

/* findSupport, IBMNode.cpp:139-179 */
int32_t orc_find_support(orc_grid *g) {
	int overflow = 0;
	double Dx = g->Dx;
	for (int64_t m = 0; m < g->nMarkers; m++) {
		orc_marker *k = &g->mk[m];
		memset(k->sidx, 0, sizeof(k->sidx)); memset(k->sjdx, 0, sizeof(k->sjdx)); memset(k->sdirac, 0, sizeof(k->sdirac));
		double stencilWidth = 1.5;
		int inear = (int)round(k->pos[0] / Dx);
		int jnear = (int)round(k->pos[1] / Dx);
		k->suppCount = 0;
		for (int i = inear - 2; i <= inear + 2; i++) {
			double distX = fabs(k->pos[0] / Dx - i);
			for (int j = jnear - 2; j <= jnear + 2; j++) {
				double distY = fabs(k->pos[1] / Dx - j);
				if (distX < stencilWidth && distY < stencilWidth && i >= 0 && i <= g->Nx - 1 && j >= 0 && j <= g->Ny - 1) {
					if (k->suppCount == SUPP) { overflow = 1; continue; }
					k->sidx[k->suppCount] = i;
					k->sjdx[k->suppCount] = j;
					k->sdirac[k->suppCount] = orc_dirac_delta(distX) * orc_dirac_delta(distY);
					k->suppCount++;
				}
			}
		}
	}
	return overflow;
}
void orc_get_supports(const orc_grid *g, int32_t *count, int32_t *idx, int32_t *jdx, double *dirac) {
	for (int64_t m = 0; m < g->nMarkers; m++) {
		count[m] = g->mk[m].suppCount;
		for (int s = 0; s < SUPP; s++) {
			idx[m * SUPP + s] = g->mk[m].sidx[s];
			jdx[m * SUPP + s] = g->mk[m].sjdx[s];
			dirac[m * SUPP + s] = g->mk[m].sdirac[s];
		}
	}
}

/* computeDs, IBMNode.cpp:182-204: distance (lattice units) to the closest other marker of the same body */
void orc_compute_ds(orc_grid *g, int64_t first, int64_t count) {
	for (int64_t a = first; a < first + count; a++) {
		double currentDs = 10.0;
		for (int64_t b = first; b < first + count; b++) {
			if (a == b) continue;
			double dx = g->mk[a].pos[0] - g->mk[b].pos[0], dy = g->mk[a].pos[1] - g->mk[b].pos[1];
			double dot = 0.0;
			dot += dx * dx;
			dot += dy * dy;
			double mag = sqrt(dot) / g->Dx;
			if (mag < currentDs) currentDs = mag;
		}
		g->mk[a].ds = currentDs;
	}
}

/* Dense solve A x = b (A row-major, destroyed), LU with partial pivoting — stands in for Utils::solveLAPACK
 * (Utils.cpp:288-311: dgetrf_ + dgetrs_('T') on the row-major matrix). */
static void lu_solve(double *A, double *b, int64_t n) {
	for (int64_t k = 0; k < n; k++) {
		int64_t piv = k;
		double best = fabs(A[k * n + k]);
		for (int64_t r = k + 1; r < n; r++)
			if (fabs(A[r * n + k]) > best) { best = fabs(A[r * n + k]); piv = r; }
		if (piv != k) {
			for (int64_t c = 0; c < n; c++) { double t = A[k * n + c]; A[k * n + c] = A[piv * n + c]; A[piv * n + c] = t; }
			double t = b[k]; b[k] = b[piv]; b[piv] = t;
		}
		for (int64_t r = k + 1; r < n; r++) {
			double l = A[r * n + k] / A[k * n + k];
			A[r * n + k] = l;
			for (int64_t c = k + 1; c < n; c++) A[r * n + c] -= l * A[k * n + c];
			b[r] -= l * b[k];
		}
	}
	for (int64_t r = n - 1; r >= 0; r--) {
		double s = b[r];
		for (int64_t c = r + 1; c < n; c++) s -= A[r * n + c] * b[c];
		b[r] = s / A[r * n + r];
	}
}

/* computeEpsilon, Objects.cpp:235-321 */
void orc_compute_epsilon(orc_grid *g, int64_t first, int64_t count) {
	int64_t dim = count;
	double Dx = g->Dx;
	double *A = (double *)calloc(dim * dim, sizeof(double));
	double *b = (double *)malloc(dim * sizeof(double));
	for (int64_t i = 0; i < dim; i++) {
		const orc_marker *ni = &g->mk[first + i];
		for (int64_t j = 0; j < dim; j++) {
			const orc_marker *nj = &g->mk[first + j];
			for (int s = 0; s < ni->suppCount; s++) {
				double di = ni->sdirac[s];
				double distX = fabs(nj->pos[0] / Dx - ni->sidx[s]);
				double distY = fabs(nj->pos[1] / Dx - ni->sjdx[s]);
				double dj = orc_dirac_delta(distX) * orc_dirac_delta(distY);
				A[i * dim + j] += di * dj;
			}
			A[i * dim + j] *= 1.0 * 1.0 * nj->ds;
		}
	}
	for (int64_t i = 0; i < dim; i++) b[i] = 1.0;
	lu_solve(A, b, dim);
	for (int64_t i = 0; i < dim; i++) g->mk[first + i].epsilon = b[i];
	free(A); free(b);
}

/* ibmKernelInterp, Objects.cpp:102-117 = interpolate (IBMNode.cpp:26-48) + forceCalc (:51-58) */
void orc_ibm_interp(orc_grid *g) {
	int64_t N = g->Nx * g->Ny;
	memset(g->force_ibm, 0, N * ND * sizeof(double));
	double velScale = g->Dt / g->Dx;
#pragma omp parallel for schedule(static)
	for (int64_t m = 0; m < g->nMarkers; m++) {
		orc_marker *k = &g->mk[m];
		k->interpRho = 0.0;
		k->interpMom[0] = k->interpMom[1] = 0.0;
		for (int s = 0; s < k->suppCount; s++) {
			int64_t id = (int64_t)k->sidx[s] * g->Ny + k->sjdx[s];
			k->interpRho += g->rho[id] * k->sdirac[s] * 1.0 * 1.0;
			for (int d = 0; d < ND; d++)
				k->interpMom[d] += g->rho[id] * g->u[id * ND + d] * k->sdirac[s] * 1.0 * 1.0;
		}
		double s = velScale * k->interpRho;
		for (int d = 0; d < ND; d++)
			k->force[d] = 2.0 * (s * k->vel[d] - k->interpMom[d]);
	}
}

/* ibmKernelSpread, Objects.cpp:120-149 = spread (IBMNode.cpp:61-94, marker order) + updateMacroscopic (:97-136) */
void orc_ibm_spread(orc_grid *g) {
	int64_t N = g->Nx * g->Ny;
	memset(g->force_ibm, 0, N * ND * sizeof(double));
	for (int64_t m = 0; m < g->nMarkers; m++) {
		const orc_marker *k = &g->mk[m];
		for (int s = 0; s < k->suppCount; s++) {
			int64_t id = (int64_t)k->sidx[s] * g->Ny + k->sjdx[s];
			double Fx = k->force[0] * k->epsilon * k->ds * 1.0 * k->sdirac[s];
			double Fy = k->force[1] * k->epsilon * k->ds * 1.0 * k->sdirac[s];
			g->force_ibm[id * ND] += Fx;
			g->force_ibm[id * ND + 1] += Fy;
		}
	}
	for (int64_t m = 0; m < g->nMarkers; m++) {
		const orc_marker *k = &g->mk[m];
		for (int s = 0; s < k->suppCount; s++) {
			int64_t id = (int64_t)k->sidx[s] * g->Ny + k->sjdx[s];
			double r = 0.0, mx = 0.0, my = 0.0;
			for (int v = 0; v < NV; v++) {
				r += g->f[id * NV + v];
				mx += CX[v] * g->f[id * NV + v];
				my += CY[v] * g->f[id * NV + v];
			}
			mx = (mx + 0.5 * (g->force_xy[id * ND] + g->force_ibm[id * ND])) / r;
			my = (my + 0.5 * (g->force_xy[id * ND + 1] + g->force_ibm[id * ND + 1])) / r;
			g->rho[id] = r;
			g->u[id * ND] = mx;
			g->u[id * ND + 1] = my;
		}
	}
}

/* main.cpp:70-74 for cases whose bodies are all rigid (objectKernel, Objects.cpp:26-60, without the FEM loop) */
void orc_step(orc_grid *g, int32_t n) {
	for (int s = 0; s < n; s++) {
		g->t++;
		orc_lbm_kernel(g);
		if (g->nMarkers > 0) {
			orc_ibm_interp(g);
			orc_ibm_spread(g);
		}
	}
}

double orc_equilibrium(int32_t cm, double rho, double ux, double uy, int32_t v) {
	orc_grid g;
	static const double W0[NV] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
	memcpy(g.w, W0, sizeof(W0));
	return cm ? feq_cm(&g, rho, ux, uy, v) : feq_bgk(&g, rho, ux, uy, v);
}

int32_t orc_num_threads(void) {
	int n = 0;
#pragma omp parallel reduction(+ : n)
	n += 1;
	return n;
}
