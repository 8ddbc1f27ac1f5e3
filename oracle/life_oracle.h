/*
 * life_oracle.h — CPU restatement of LIFE's hot path in plain C.   *** TEST INFRASTRUCTURE, NOT PRODUCT CODE ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * (life_b200/) never links, imports or executes anything under oracle/.
 *
 * What it restates (reference = joconnor22/LIFE v1.0.3, paths relative to /root/reference):
 *   GridClass ctor + initialiseGrid        src/Grid.cpp:1232-1289, 916-1062
 *   GridClass::lbmKernel and its callees   src/Grid.cpp:36-556
 *   IBMNodeClass hot members               src/IBMNode.cpp:26-204, inc/Utils.h:137-232
 *   ObjectsClass::ibmKernelInterp/Spread   src/Objects.cpp:102-149
 *   ObjectsClass::computeEpsilon           src/Objects.cpp:235-321 (dense LU with partial pivoting stands in for LAPACK
 *                                          dgetrf/dgetrs, which is un-vendored: README.md:43-45 `liblapack-dev`)
 * Differences from the reference, on purpose: the case is a run-time struct instead of compile-time params.h; all
 * linear indices are 64-bit (the reference's int overflows above 15446^2).
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks this file against the compiled, unmodified reference
 * (oracle/_ref/libref_<case>.so) for all seven shipped examples and the extra cases — bit for bit for both collision
 * operators (round 2: the central-moments back-transform follows the reference's association too) — and against the
 * fixtures under tests/golden/ that the same reference build generated (tests/golden/make_golden.py).
 */
#ifndef LIFE_ORACLE_H
#define LIFE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FLUID = 0, ORC_WALL = 1, ORC_VELOCITY = 2, ORC_FREESLIP = 3, ORC_PRESSURE = 4, ORC_CONVECTIVE = 5 };
enum { ORC_PROFILE_UNIFORM = -1, ORC_PROFILE_PARABOLIC = 0, ORC_PROFILE_SHEAR = 1, ORC_PROFILE_BOUNDARYLAYER = 2 };

/* run-time image of inc/params.h */
typedef struct orc_params {
	int64_t Nx, Ny;
	int32_t central_moments;   /* #define CENTRAL_MOMENTS */
	int32_t ordered;           /* #define ORDERED         */
	int32_t uni_epsilon;       /* #define UNI_EPSILON     */
	int32_t profile;           /* PROFILE or ORC_PROFILE_UNIFORM */
	int32_t wall_left, wall_right, wall_bottom, wall_top;
	double inlet_ramp;         /* INLET_RAMP, <= 0 off */
	double womersley;          /* WOMERSLEY,  <= 0 off */
	double omega;
	double height_p, rho_p, nu_p;
	double ux0_p, uy0_p;
	double gravityX, gravityY, dpdx, dpdy;
	double uxInlet_p, uyInlet_p;
} orc_params;

typedef struct orc_grid orc_grid;

orc_grid *orc_create(const orc_params *p);
void orc_destroy(orc_grid *g);

/* scalings computed by the constructor (Grid.cpp:1255-1260): out = {Dx, Dt, Dm, Drho, tau, nu} */
void orc_scalings(const orc_grid *g, double *out6);

int32_t orc_get_t(const orc_grid *g);
void orc_set_t(orc_grid *g, int32_t t);

/* GridClass::lbmKernel at the current t */
void orc_lbm_kernel(orc_grid *g);
/* main.cpp:70-74: n times { t++ ; lbmKernel ; objectKernel (rigid markers only) } */
void orc_step(orc_grid *g, int32_t n);

/* state access, reference layout.  which: 0 f, 1 f_n, 2 rho, 3 rho_n, 4 u, 5 u_n, 6 force_xy, 7 force_ibm, 8 u_in, 9 rho_in, 10 delU */
double *orc_array(orc_grid *g, int32_t which, int64_t *len);
int32_t *orc_types(orc_grid *g);
int64_t orc_bc_count(const orc_grid *g);
const int64_t *orc_bc_ids(const orc_grid *g);
/* normal of a boundary node as getNormalVector returns it; returns normalDirection */
int32_t orc_normal(const orc_grid *g, int64_t i, int64_t j, int32_t *nx, int32_t *ny);
/* push target of population v of node (i,j) (Grid.cpp:229 / :240) */
int64_t orc_stream_target(const orc_grid *g, int64_t i, int64_t j, int32_t v);

/* ---- markers ---- */
void orc_set_markers(orc_grid *g, int64_t n, const double *pos, const double *vel, const double *ds, const double *eps);
void orc_get_marker_force(const orc_grid *g, double *force);
void orc_set_marker_force(orc_grid *g, const double *force);
void orc_get_interp(const orc_grid *g, double *rho, double *mom);
/* IBMNodeClass::findSupport for every marker; returns 0, or 1 if some marker overflows the 9-entry buffer */
int32_t orc_find_support(orc_grid *g);
void orc_get_supports(const orc_grid *g, int32_t *count, int32_t *idx, int32_t *jdx, double *dirac);
/* IBMNodeClass::computeDs over markers [first, first+count) treated as one body */
void orc_compute_ds(orc_grid *g, int64_t first, int64_t count);
/* ObjectsClass::computeEpsilon over markers [first, first+count) treated as one body */
void orc_compute_epsilon(orc_grid *g, int64_t first, int64_t count);
void orc_get_ds_eps(const orc_grid *g, double *ds, double *eps);
void orc_ibm_interp(orc_grid *g);   /* ObjectsClass::ibmKernelInterp */
void orc_ibm_spread(orc_grid *g);   /* ObjectsClass::ibmKernelSpread */

/* stand-alone helpers exposed for unit tests */
double orc_dirac_delta(double dist);                                     /* Utils.h:220-232 */
double orc_equilibrium(int32_t cm, double rho, double ux, double uy, int32_t v);   /* Grid.cpp:249-264 */
int32_t orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
