// TEST INFRASTRUCTURE: the serial (one-thread CTA) instantiation of life_b200/csrc/fem_core.h behind a C interface, so that
// tests/test_fem_core.py can hold the device solver's logic against the compiled reference on the CPU.  Built by the test itself
// (g++ -O2 -ffp-contract=off).  Not part of the product.
#include "../../life_b200/csrc/fem_core.h"
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace life_fem;

struct HostBody {
	Body b{};
	std::vector<double> d;   // every double array of the body, back to back
	std::vector<int> ints;
};

static double *take(std::vector<double> &pool, size_t &off, size_t n) { double *p = pool.data() + off; off += n; return p; }

extern "C" {

// consts: alpha, delta, Dt, Dm, gravityX, gravityY, ref_L; el [n_el*5] = L0, A, I, E, rho
void *femc_create(int n_nodes, int n_bc, int n_ibm, const double *consts, const double *pos0, const double *angle0, const double *el,
                  const int *pm_el, const double *pm_zeta, const int *fm_first, const int *fm_node, const double *fm_z1, const double *fm_z2) {
	HostBody *h = new HostBody();
	Body &b = h->b;
	const int ne = n_nodes - 1, dim = 3 * n_nodes, nmap = fm_first[ne];
	b.n_nodes = n_nodes; b.n_el = ne; b.n_dof = dim; b.n_bc = n_bc; b.n_ibm = n_ibm;
	b.alpha = consts[0]; b.delta = consts[1]; b.Dt = consts[2]; b.Dm = consts[3]; b.gravityX = consts[4]; b.gravityY = consts[5]; b.ref_L = consts[6];
	const size_t total = 3 * n_nodes * 2 + 5 * ne + 72 * ne + n_ibm + 2 * nmap + 2 * ne + 36 * ne + 6 * ne + 6 * ne + 2 * (size_t)dim * dim + 4 * dim + 8 + 11 * dim;
	h->d.assign(total + 64, 0.0);
	size_t o = 0;
	double *p0 = take(h->d, o, 2 * n_nodes), *a0 = take(h->d, o, n_nodes);
	memcpy(p0, pos0, sizeof(double) * 2 * n_nodes); memcpy(a0, angle0, sizeof(double) * n_nodes);
	b.pos0 = p0; b.angle0 = a0;
	b.pos = take(h->d, o, 2 * n_nodes); b.angle = take(h->d, o, n_nodes);
	double *L0 = take(h->d, o, ne), *A = take(h->d, o, ne), *I = take(h->d, o, ne), *E = take(h->d, o, ne), *rho = take(h->d, o, ne);
	double *Ml = take(h->d, o, 36 * ne), *Kl = take(h->d, o, 36 * ne);
	for (int e = 0; e < ne; e++) {
		L0[e] = el[5 * e]; A[e] = el[5 * e + 1]; I[e] = el[5 * e + 2]; E[e] = el[5 * e + 3]; rho[e] = el[5 * e + 4];
		fem_local_matrices(L0[e], A[e], I[e], E[e], rho[e], Ml + 36 * e, Kl + 36 * e);
	}
	b.L0 = L0; b.A = A; b.I = I; b.E = E; b.rho = rho; b.Mloc = Ml; b.KLloc = Kl;
	double *pz = take(h->d, o, n_ibm), *z1 = take(h->d, o, nmap), *z2 = take(h->d, o, nmap);
	memcpy(pz, pm_zeta, sizeof(double) * n_ibm); memcpy(z1, fm_z1, sizeof(double) * nmap); memcpy(z2, fm_z2, sizeof(double) * nmap);
	b.pm_zeta = pz; b.fm_z1 = z1; b.fm_z2 = z2;
	h->ints.assign(n_ibm + (ne + 1) + nmap + dim + 8, 0);
	int *ip = h->ints.data();
	memcpy(ip, pm_el, sizeof(int) * n_ibm); b.pm_el = ip; ip += n_ibm;
	memcpy(ip, fm_first, sizeof(int) * (ne + 1)); b.fm_first = ip; ip += ne + 1;
	memcpy(ip, fm_node, sizeof(int) * nmap); b.fm_node = ip; ip += nmap;
	b.piv = ip;
	b.L = take(h->d, o, ne); b.elangle = take(h->d, o, ne); b.T = take(h->d, o, 36 * ne); b.Floc = take(h->d, o, 6 * ne); b.Rel = take(h->d, o, 6 * ne);
	b.M = take(h->d, o, (size_t)dim * dim); b.K = take(h->d, o, (size_t)dim * dim);
	b.R = take(h->d, o, dim); b.F = take(h->d, o, dim); b.delU = take(h->d, o, dim); b.work = take(h->d, o, dim);
	b.scal = take(h->d, o, 8);
	double **vecs[11] = {&b.U, &b.Udot, &b.Udotdot, &b.U_n, &b.Udot_n, &b.Udotdot_n, &b.U_km1, &b.R_k, &b.R_km1, &b.U_nm1, &b.U_nm2};
	for (int k = 0; k < 11; k++) *vecs[k] = take(h->d, o, dim);
	if (o > h->d.size()) abort();
	update_geometry(b, Lane{0, 1});
	return h;
}

void femc_destroy(void *p) { delete static_cast<HostBody *>(p); }

static double *vec_of(Body &b, int k) {
	double *v[11] = {b.U, b.Udot, b.Udotdot, b.U_n, b.Udot_n, b.Udotdot_n, b.U_km1, b.R_k, b.R_km1, b.U_nm1, b.U_nm2};
	return v[k];
}
void femc_set_state(void *p, const double *in) {
	Body &b = static_cast<HostBody *>(p)->b;
	for (int k = 0; k < 11; k++) memcpy(vec_of(b, k), in + (size_t)k * b.n_dof, sizeof(double) * b.n_dof);
}
void femc_get_state(void *p, double *out) {
	Body &b = static_cast<HostBody *>(p)->b;
	for (int k = 0; k < 11; k++) memcpy(out + (size_t)k * b.n_dof, vec_of(b, k), sizeof(double) * b.n_dof);
}
// results [5] = subRes, subNum, subDen, resNR, itNR
void femc_dynamic(void *p, const double *force, const double *epsilon, double *pos, double *vel, double *results) {
	Body &b = static_cast<HostBody *>(p)->b;
	fem_dynamic(b, Lane{0, 1}, force, epsilon, pos, vel, nullptr);
	results[0] = b.scal[1]; results[1] = b.scal[2]; results[2] = b.scal[3]; results[3] = b.scal[0]; results[4] = b.scal[4];
}
// ds [n_ibm] of the body's markers from their positions pos [2 * n_ibm] (body-local order), lattice spacing Dx
void femc_compute_ds(void *p, const double *pos, double Dx, double *ds) { compute_ds(static_cast<HostBody *>(p)->b, Lane{0, 1}, pos, Dx, ds, nullptr); }
void femc_predict(void *p, int t, double *pos, double *vel) { fem_predict(static_cast<HostBody *>(p)->b, Lane{0, 1}, t, pos, vel, nullptr); }
void femc_relax(void *p, double relax, double *pos, double *vel) { fem_relax(static_cast<HostBody *>(p)->b, Lane{0, 1}, relax, pos, vel, nullptr); }

}  // extern "C"
