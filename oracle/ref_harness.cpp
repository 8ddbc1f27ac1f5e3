// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference sources (/root/reference/src/*.cpp, compiled in place by
// oracle/Makefile with -Dprivate=public and the case's params.h force-included).  One shared object per
// compile-time case: oracle/_ref/libref_<case>.so.  It lets tests/ and bench.py's cpu_baseline leg
//   * run the reference's own time loop body (main.cpp:70-74: t++ ; grid.solver()),
//   * call the individual stages of the hot path (GridClass::lbmKernel Grid.cpp:36, and the pieces of
//     ObjectsClass::objectKernel Objects.cpp:26-60) one at a time, and
//   * read / write the reference's state arrays in their native layout (AoS, id = i*Ny + j).
// Nothing here computes anything on its own: every number comes out of reference code.
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this.

#include "Grid.h"
#include "Objects.h"
#include "FEMBody.h"
#include "Utils.h"
#include <unistd.h>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

static GridClass *g_grid = nullptr;
static ObjectsClass *g_obj = nullptr;

// Construct grid + objects exactly as main.cpp:42-45 does, inside a scratch working directory (the GridClass
// constructor deletes ./Results, Utils.cpp:56-58; ObjectsClass reads ./input/geometry.config, Objects.cpp:327).
REF_API int ref_create(const char *workdir) {
	if (g_grid != nullptr) return 1;
	if (workdir != nullptr && workdir[0] != 0 && chdir(workdir) != 0) return 2;
	g_grid = new GridClass();
	g_obj = new ObjectsClass(*g_grid);
	return 0;
}

REF_API void ref_destroy() {
	delete g_obj; g_obj = nullptr;
	delete g_grid; g_grid = nullptr;
}

// ---- compile-time case description ------------------------------------------------------------------------------
REF_API void ref_dims(int *out) {
	out[0] = Nx; out[1] = Ny;
	out[2] = g_obj ? static_cast<int>(g_obj->iNode.size()) : 0;
	out[3] = g_obj ? static_cast<int>(g_obj->iBody.size()) : 0;
	out[4] = g_grid ? static_cast<int>(g_grid->BCVec.size()) : 0;
}

REF_API void ref_walls(int *out) {
	out[0] = WALL_LEFT; out[1] = WALL_RIGHT; out[2] = WALL_BOTTOM; out[3] = WALL_TOP;
}

// bit0 CENTRAL_MOMENTS, bit1 ORDERED, bit2 UNI_EPSILON, bit3 WOMERSLEY, bit4 INLET_RAMP, bit5 hasIBM, bit6 hasFlex
REF_API int ref_flags() {
	int fl = 0;
#ifdef CENTRAL_MOMENTS
	fl |= 1;
#endif
#ifdef ORDERED
	fl |= 2;
#endif
#ifdef UNI_EPSILON
	fl |= 4;
#endif
#ifdef WOMERSLEY
	fl |= 8;
#endif
#ifdef INLET_RAMP
	fl |= 16;
#endif
	if (g_obj && g_obj->hasIBM) fl |= 32;
	if (g_obj && g_obj->hasFlex) fl |= 64;
	return fl;
}

// omega, Dx, Dt, Dm, Drho, inlet_ramp (or -1), womersley (or -1), dpdx, dpdy, gravityX, gravityY, height_p, nu_p, rho_p, subTol, uxInlet_p, uyInlet_p, ux0_p, uy0_p
REF_API void ref_scalars(double *out) {
	out[0] = omega; out[1] = g_grid->Dx; out[2] = g_grid->Dt; out[3] = g_grid->Dm; out[4] = g_grid->Drho;
#ifdef INLET_RAMP
	out[5] = INLET_RAMP;
#else
	out[5] = -1.0;
#endif
#ifdef WOMERSLEY
	out[6] = WOMERSLEY;
#else
	out[6] = -1.0;
#endif
	out[7] = dpdx; out[8] = dpdy; out[9] = gravityX; out[10] = gravityY;
	out[11] = height_p; out[12] = nu_p; out[13] = rho_p; out[14] = subTol;
	out[15] = uxInlet_p; out[16] = uyInlet_p; out[17] = ux0_p; out[18] = uy0_p;
}

// PROFILE (eProfileType) or -1 when the inlet is uniform
REF_API int ref_profile() {
#ifdef PROFILE
	return PROFILE;
#else
	return -1;
#endif
}

// ---- time loop pieces ----------------------------------------------------------------------------------------------
// OpenMP team size: what the reference's own counter observes in a parallel region (src/Utils.cpp:228-240), and a setter so a
// harness can undo an inherited OMP_NUM_THREADS=1 (torch.distributed.run exports that to every worker)
REF_API int ref_omp_threads() { return Utils::omp_thread_count(); }
REF_API void ref_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }

REF_API int ref_get_t() { return g_grid->t; }
REF_API void ref_set_t(int t) { g_grid->t = t; }

// main.cpp:70-74 without the I/O: n passes of { t++ ; solver() }
REF_API void ref_step(int n) {
	for (int s = 0; s < n; s++) {
		g_grid->t++;
		g_grid->solver();
	}
}

// Grid.cpp:36 alone (caller advances t with ref_set_t)
REF_API void ref_lbm_kernel() { g_grid->lbmKernel(); }
// Objects.cpp:26
REF_API void ref_object_kernel() { g_obj->objectKernel(); }
// The stages objectKernel() strings together (Objects.cpp:33-56), callable one at a time
REF_API void ref_recompute_object_vals() { g_obj->recomputeObjectVals(); }
REF_API void ref_ibm_interp() { g_obj->ibmKernelInterp(); }
REF_API void ref_fem_kernel() { g_obj->femKernel(); }
REF_API void ref_ibm_spread() { g_obj->ibmKernelSpread(); }
REF_API int ref_get_subit() { return g_obj->subIt; }
REF_API void ref_set_subit(int s) { g_obj->subIt = s; }
REF_API double ref_get_subres() { return g_obj->subRes; }
REF_API double ref_get_relax() { return g_obj->relax; }

// ---- lattice state, native layout --------------------------------------------------------------------------------
#define GETSET(name, member, T)                                                                                      \
	REF_API void ref_get_##name(T *out) { memcpy(out, g_grid->member.data(), g_grid->member.size() * sizeof(T)); }   \
	REF_API void ref_set_##name(const T *in) { memcpy(g_grid->member.data(), in, g_grid->member.size() * sizeof(T)); }
GETSET(f, f, double)
GETSET(f_n, f_n, double)
GETSET(rho, rho, double)
GETSET(rho_n, rho_n, double)
GETSET(u, u, double)
GETSET(u_n, u_n, double)
GETSET(force_xy, force_xy, double)
GETSET(force_ibm, force_ibm, double)
GETSET(u_in, u_in, double)
GETSET(rho_in, rho_in, double)

REF_API void ref_get_type(int *out) {
	for (size_t k = 0; k < g_grid->type.size(); k++) out[k] = static_cast<int>(g_grid->type[k]);
}
REF_API void ref_get_bcvec(int *out) { memcpy(out, g_grid->BCVec.data(), g_grid->BCVec.size() * sizeof(int)); }
REF_API int ref_get_delu(double *out) {
	memcpy(out, g_grid->delU.data(), g_grid->delU.size() * sizeof(double));
	return static_cast<int>(g_grid->delU.size());
}

// NOTE: GridClass::streamCollide / getNormalVector / equilibrium are declared `inline` inside Grid.cpp, so they have no
// external symbol and cannot be called from here.  The push map is recovered in tests/ from a full lbmKernel() pass
// over tagged populations instead (rho_n = 0 makes the BGK equilibrium exactly zero: f_new[recv] = tag - omega*tag).

// ---- markers ------------------------------------------------------------------------------------------------------------
REF_API void ref_get_markers(double *pos, double *vel, double *force, double *ds, double *eps, int *body, int *flex) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		IBMNodeClass &m = g_obj->iNode[n];
		if (pos) { pos[2 * n] = m.pos[eX]; pos[2 * n + 1] = m.pos[eY]; }
		if (vel) { vel[2 * n] = m.vel[eX]; vel[2 * n + 1] = m.vel[eY]; }
		if (force) { force[2 * n] = m.force[eX]; force[2 * n + 1] = m.force[eY]; }
		if (ds) ds[n] = m.ds;
		if (eps) eps[n] = m.epsilon;
		if (body) body[n] = m.iPtr->ID;
		if (flex) flex[n] = static_cast<int>(m.iPtr->flex);
	}
}
REF_API void ref_set_marker_force(const double *force) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		g_obj->iNode[n].force[eX] = force[2 * n];
		g_obj->iNode[n].force[eY] = force[2 * n + 1];
	}
}
REF_API void ref_set_marker_posvel(const double *pos, const double *vel) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		g_obj->iNode[n].pos[eX] = pos[2 * n]; g_obj->iNode[n].pos[eY] = pos[2 * n + 1];
		g_obj->iNode[n].vel[eX] = vel[2 * n]; g_obj->iNode[n].vel[eY] = vel[2 * n + 1];
	}
}
REF_API void ref_get_interp(double *rho, double *mom) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		rho[n] = g_obj->iNode[n].interpRho;
		mom[2 * n] = g_obj->iNode[n].interpMom[eX]; mom[2 * n + 1] = g_obj->iNode[n].interpMom[eY];
	}
}
// supports: count[n]; idx/jdx/dirac [n*9 + s]
REF_API void ref_get_supports(int *count, int *idx, int *jdx, double *dirac) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		IBMNodeClass &m = g_obj->iNode[n];
		count[n] = static_cast<int>(m.suppCount);
		for (int s = 0; s < suppSize; s++) {
			idx[n * suppSize + s] = m.supp[s].idx;
			jdx[n * suppSize + s] = m.supp[s].jdx;
			dirac[n * suppSize + s] = m.supp[s].diracVal;
		}
	}
}
// IBMNodeClass::findSupport (IBMNode.cpp:139) + computeDs (:182) on every marker, then ObjectsClass::computeEpsilon
// (Objects.cpp:235) — what initialiseObjects (Objects.cpp:941) does, re-runnable after ref_set_marker_posvel.
REF_API void ref_refresh_supports(int with_epsilon) {
	for (size_t n = 0; n < g_obj->iNode.size(); n++) {
		g_obj->iNode[n].findSupport();
		g_obj->iNode[n].computeDs();
	}
	if (with_epsilon) {
		int tKeep = g_grid->t;
		g_grid->t = 0;                 // computeEpsilon only touches rigid bodies when t == 0 (Objects.cpp:261)
		g_obj->computeEpsilon();
		g_grid->t = tKeep;
	}
}

// Restart files in the reference's own format (Grid.cpp:1163, Objects.cpp:1206) into ./Results/Restart
REF_API void ref_write_restart() { Utils::writeRestart(*g_grid); }

// Results/VTK/Fluid.<t>.vti by the reference's own writer (Grid.cpp:790-898), whether or not the case defines VTK
REF_API void ref_write_vtk() { g_grid->writeVTK(); }

// The reference's own reader (Grid.cpp:1072-1160) on ./Results/Restart/Fluid.restart; returns the time step it continued from.
// A header / index mismatch ends the process through ERROR() (exit 99), as in the reference.
REF_API int ref_read_restart() {
	g_grid->readRestart();
	return g_grid->tOffset;
}
REF_API double ref_ref_pressure() { return ref_P; }   // params.h: the constant writeVTK adds to the pressure

// ---- flexible bodies (FEMBodyClass, src/FEMBody.cpp) ------------------------------------------------------------------------------
// Read-only description + state vectors of the fb-th flexible body, and its dynamicFEM() alone: what the FEM restatement
// (oracle/life_oracle_fem.c, groundwork for SURVEY.md §8f row 3) is initialised from and compared with.
static IBMBodyClass *flex_body(int fb) {
	int k = 0;
	for (size_t ib = 0; ib < g_obj->iBody.size(); ib++)
		if (g_obj->iBody[ib].flex == eFlexible && k++ == fb) return &g_obj->iBody[ib];
	return nullptr;
}
REF_API int ref_fem_count() {
	int k = 0;
	for (size_t ib = 0; ib < g_obj->iBody.size(); ib++) k += g_obj->iBody[ib].flex == eFlexible;
	return k;
}
// out: nNodes, nElements, bodyDOFs, bcDOFs, nIBM (= posMap.size()), total forceMap entries, simDOFs
REF_API void ref_fem_dims(int fb, int *out) {
	FEMBodyClass *s = flex_body(fb)->sBody;
	int nmap = 0;
	for (size_t e = 0; e < s->element.size(); e++) nmap += static_cast<int>(s->element[e].forceMap.size());
	out[0] = static_cast<int>(s->node.size()); out[1] = static_cast<int>(s->element.size()); out[2] = s->bodyDOFs; out[3] = s->bcDOFs;
	out[4] = static_cast<int>(s->posMap.size()); out[5] = nmap; out[6] = g_obj->simDOFs;
}
// params.h / grid constants the solver reads: alpha, delta, Dt, Dm, Dx, gravityX, gravityY, ref_L
REF_API void ref_fem_constants(double *out) {
	out[0] = alpha; out[1] = delta; out[2] = g_grid->Dt; out[3] = g_grid->Dm; out[4] = g_grid->Dx;
	out[5] = gravityX; out[6] = gravityY; out[7] = ref_L;
}
// pos0 [2*nNodes], angle0 [nNodes]; el [nEl*5] = L0, A, I, E, rho per element
REF_API void ref_fem_geometry(int fb, double *pos0, double *angle0, double *el) {
	FEMBodyClass *s = flex_body(fb)->sBody;
	for (size_t n = 0; n < s->node.size(); n++) {
		pos0[2 * n] = s->node[n].pos0[eX]; pos0[2 * n + 1] = s->node[n].pos0[eY]; angle0[n] = s->node[n].angle0;
	}
	for (size_t e = 0; e < s->element.size(); e++) {
		FEMElementClass &E = s->element[e];
		el[5 * e] = E.L0; el[5 * e + 1] = E.A; el[5 * e + 2] = E.I; el[5 * e + 3] = E.E; el[5 * e + 4] = E.rho;
	}
}
// posMap (per IBM node of the body: element, zeta), forceMap (per element: entries fm_first[e] .. fm_first[e+1]: body-local IBM
// node, zeta1, zeta2), marker [nIBM] = index of the body's k-th IBM node in the global marker arrays
REF_API void ref_fem_maps(int fb, int *pm_el, double *pm_zeta, int *fm_first, int *fm_node, double *fm_z1, double *fm_z2, int *marker) {
	IBMBodyClass *b = flex_body(fb);
	FEMBodyClass *s = b->sBody;
	for (size_t i = 0; i < s->posMap.size(); i++) { pm_el[i] = s->posMap[i].elID; pm_zeta[i] = s->posMap[i].zeta; }
	int k = 0;
	for (size_t e = 0; e < s->element.size(); e++) {
		fm_first[e] = k;
		for (size_t n = 0; n < s->element[e].forceMap.size(); n++, k++) {
			fm_node[k] = s->element[e].forceMap[n].nodeID; fm_z1[k] = s->element[e].forceMap[n].zeta1; fm_z2[k] = s->element[e].forceMap[n].zeta2;
		}
	}
	fm_first[s->element.size()] = k;
	for (size_t i = 0; i < b->node.size(); i++) marker[i] = static_cast<int>(b->node[i] - &g_obj->iNode[0]);
}
// state [11 * bodyDOFs]: U, Udot, Udotdot, U_n, Udot_n, Udotdot_n, U_km1, R_k, R_km1, U_nm1, U_nm2
static const int FEM_VECS = 11;
static std::vector<double> *fem_vectors(FEMBodyClass *s, int k) {
	std::vector<double> *v[FEM_VECS] = {&s->U, &s->Udot, &s->Udotdot, &s->U_n, &s->Udot_n, &s->Udotdot_n, &s->U_km1, &s->R_k, &s->R_km1,
	                                    &s->U_nm1, &s->U_nm2};
	return v[k];
}
REF_API void ref_fem_get_state(int fb, double *out) {
	FEMBodyClass *s = flex_body(fb)->sBody;
	for (int k = 0; k < FEM_VECS; k++) memcpy(out + (size_t)k * s->bodyDOFs, fem_vectors(s, k)->data(), sizeof(double) * s->bodyDOFs);
}
REF_API void ref_fem_set_state(int fb, const double *in) {
	FEMBodyClass *s = flex_body(fb)->sBody;
	for (int k = 0; k < FEM_VECS; k++) memcpy(fem_vectors(s, k)->data(), in + (size_t)k * s->bodyDOFs, sizeof(double) * s->bodyDOFs);
}
// FEMBodyClass::dynamicFEM (src/FEMBody.cpp:26-68) of one body; out: subRes, subNum, subDen, resNR, itNR
REF_API void ref_fem_dynamic(int fb, double *out) {
	FEMBodyClass *s = flex_body(fb)->sBody;
	s->dynamicFEM();
	out[0] = s->subRes; out[1] = s->subNum; out[2] = s->subDen; out[3] = s->resNR; out[4] = s->itNR;
}
