// TEST INFRASTRUCTURE: barrier placement of life_b200/csrc/fem_core.h, checked on the CPU.
//
// The device code runs one CTA per body; its threads cooperate through __syncthreads().  Here the CTA is emulated with real
// threads and FEM_SYNC() is a pthread barrier, the program is built with -fsanitize=thread, and every recorded call (predictor,
// relaxed update, dynamicFEM with the state and marker forces of a live run of the compiled reference; written by
// tests/test_fem_core.py) is executed twice from the same state: by one thread and by NTHREADS threads.  A missing barrier shows up
// as a ThreadSanitizer data-race report (two threads touching the same word with no barrier in between), a misplaced one as a
// result that differs from the serial run — the two must agree bit for bit, since every value is produced by exactly one thread
// with the same arithmetic whatever the CTA size.
//
//   fem_core_race <calls.bin> [nthreads]      exit 0 = all calls identical (ThreadSanitizer reports make the exit code 66)
#include <pthread.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static thread_local pthread_barrier_t *tl_barrier = nullptr;
static inline void host_cta_sync() {
	if (tl_barrier) pthread_barrier_wait(tl_barrier);
}
#define FEM_SYNC() host_cta_sync()
#include "../../life_b200/csrc/fem_core.h"

using namespace life_fem;

struct Pool {
	std::vector<double> d;
	std::vector<int> i;
	size_t od = 0, oi = 0;
	double *D(size_t n) { double *p = d.data() + od; od += n; return p; }
	int *I(size_t n) { int *p = i.data() + oi; oi += n; return p; }
};

template <typename T>
static void rd(FILE *f, T *p, size_t n) {
	if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(3); }
}

template <typename Fn>
static void run_cta(int nthreads, Fn fn) {
	if (nthreads == 1) { fn(Lane{0, 1}); return; }
	pthread_barrier_t bar;
	pthread_barrier_init(&bar, nullptr, (unsigned)nthreads);
	std::vector<std::thread> th;
	for (int t = 0; t < nthreads; t++)
		th.emplace_back([&, t] { tl_barrier = &bar; fn(Lane{t, nthreads}); tl_barrier = nullptr; });
	for (auto &x : th) x.join();
	pthread_barrier_destroy(&bar);
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	const int nthreads = argc > 2 ? atoi(argv[2]) : 7;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	int32_t hdr[5];
	rd(f, hdr, 5);
	const int n_nodes = hdr[0], n_bc = hdr[1], n_ibm = hdr[2], nmap = hdr[3], n_calls = hdr[4];
	const int ne = n_nodes - 1, dim = 3 * n_nodes;
	double consts[7];
	rd(f, consts, 7);
	Pool P;
	P.d.assign(3 * n_nodes * 2 + 5 * ne + 72 * ne + n_ibm + 2 * nmap + 2 * ne + 48 * ne + 2 * (size_t)dim * dim + 4 * dim + 8 + 11 * dim + 64, 0.0);
	P.i.assign(n_ibm + ne + 1 + nmap + dim + 8, 0);
	Body b{};
	b.n_nodes = n_nodes; b.n_el = ne; b.n_dof = dim; b.n_bc = n_bc; b.n_ibm = n_ibm;
	b.alpha = consts[0]; b.delta = consts[1]; b.Dt = consts[2]; b.Dm = consts[3]; b.gravityX = consts[4]; b.gravityY = consts[5]; b.ref_L = consts[6];
	double *pos0 = P.D(2 * n_nodes), *angle0 = P.D(n_nodes);
	rd(f, pos0, 2 * n_nodes); rd(f, angle0, n_nodes);
	std::vector<double> el(5 * ne);
	rd(f, el.data(), el.size());
	double *L0 = P.D(ne), *A = P.D(ne), *I = P.D(ne), *E = P.D(ne), *rho = P.D(ne), *Ml = P.D(36 * ne), *Kl = P.D(36 * ne);
	for (int e = 0; e < ne; e++) {
		L0[e] = el[5 * e]; A[e] = el[5 * e + 1]; I[e] = el[5 * e + 2]; E[e] = el[5 * e + 3]; rho[e] = el[5 * e + 4];
		fem_local_matrices(L0[e], A[e], I[e], E[e], rho[e], Ml + 36 * e, Kl + 36 * e);
	}
	int *pm_el = P.I(n_ibm); double *pm_zeta = P.D(n_ibm);
	rd(f, pm_el, n_ibm); rd(f, pm_zeta, n_ibm);
	int *fm_first = P.I(ne + 1), *fm_node = P.I(nmap);
	double *fm_z1 = P.D(nmap), *fm_z2 = P.D(nmap);
	rd(f, fm_first, ne + 1); rd(f, fm_node, nmap); rd(f, fm_z1, nmap); rd(f, fm_z2, nmap);
	b.pos0 = pos0; b.angle0 = angle0; b.L0 = L0; b.A = A; b.I = I; b.E = E; b.rho = rho; b.Mloc = Ml; b.KLloc = Kl;
	b.pm_el = pm_el; b.pm_zeta = pm_zeta; b.fm_first = fm_first; b.fm_node = fm_node; b.fm_z1 = fm_z1; b.fm_z2 = fm_z2;
	b.pos = P.D(2 * n_nodes); b.angle = P.D(n_nodes); b.L = P.D(ne); b.elangle = P.D(ne); b.T = P.D(36 * ne); b.Floc = P.D(6 * ne); b.Rel = P.D(6 * ne);
	b.M = P.D((size_t)dim * dim); b.K = P.D((size_t)dim * dim); b.R = P.D(dim); b.F = P.D(dim); b.delU = P.D(dim); b.work = P.D(dim);
	b.scal = P.D(8); b.piv = P.I(dim);
	double *vec[11];
	for (int k = 0; k < 11; k++) vec[k] = P.D(dim);
	b.U = vec[0]; b.Udot = vec[1]; b.Udotdot = vec[2]; b.U_n = vec[3]; b.Udot_n = vec[4]; b.Udotdot_n = vec[5]; b.U_km1 = vec[6];
	b.R_k = vec[7]; b.R_km1 = vec[8]; b.U_nm1 = vec[9]; b.U_nm2 = vec[10];
	if (P.od > P.d.size() || P.oi > P.i.size()) return 4;

	std::vector<double> state(11 * (size_t)dim), force(2 * n_ibm), eps(n_ibm);
	std::vector<double> out[2], pos(2 * n_ibm), vel(2 * n_ibm);
	int bad = 0;
	for (int c = 0; c < n_calls; c++) {
		int32_t kind_t[2];
		double relax;
		rd(f, kind_t, 2); rd(f, &relax, 1);
		rd(f, state.data(), state.size()); rd(f, force.data(), force.size()); rd(f, eps.data(), eps.size());
		for (int pass = 0; pass < 2; pass++) {
			for (int k = 0; k < 11; k++) memcpy(vec[k], state.data() + (size_t)k * dim, sizeof(double) * dim);
			std::fill(pos.begin(), pos.end(), 0.0); std::fill(vel.begin(), vel.end(), 0.0);
			for (int k = 0; k < 8; k++) b.scal[k] = 0.0;
			run_cta(pass == 0 ? 1 : nthreads, [&](Lane l) {
				update_geometry(b, l);      // a body freshly pointed at this state: geometry follows U
				if (kind_t[0] == 0) fem_dynamic(b, l, force.data(), eps.data(), pos.data(), vel.data(), nullptr);
				else if (kind_t[0] == 1) fem_predict(b, l, kind_t[1], pos.data(), vel.data(), nullptr);
				else fem_relax(b, l, relax, pos.data(), vel.data(), nullptr);
			});
			out[pass].clear();
			for (int k = 0; k < 11; k++) out[pass].insert(out[pass].end(), vec[k], vec[k] + dim);
			out[pass].insert(out[pass].end(), pos.begin(), pos.end());
			out[pass].insert(out[pass].end(), vel.begin(), vel.end());
			out[pass].insert(out[pass].end(), b.scal, b.scal + 5);
		}
		if (out[0].size() != out[1].size() || memcmp(out[0].data(), out[1].data(), sizeof(double) * out[0].size()) != 0) {
			fprintf(stderr, "call %d (kind %d): %d-thread CTA differs from the serial run\n", c, kind_t[0], nthreads);
			bad++;
		}
	}
	fclose(f);
	printf("%d calls, CTA of %d threads vs serial: %d differ\n", n_calls, nthreads, bad);
	return bad ? 1 : 0;
}
