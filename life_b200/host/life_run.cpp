// life_run.cpp — run-time front end of liblife_b200: LIFE's time loop for body-free cases with the case
// read from a FILE instead of being compiled in (the reference needs a rebuild per case: inc/params.h is compile time, SURVEY.md
// F10), on 1..8 B200s of one box, one host thread per GPU, through the C ABI only.
//
//     life_run CASEFILE [key=value ...]
//
// The case file holds `key = value` lines with the names of inc/params.h (`#` starts a comment); later command-line pairs
// override it.  Keys (defaults in brackets):
//     Nx, Ny                         lattice size (64-bit; no 15446^2 limit)
//     height_p [1] rho_p [1] nu_p    physical domain height, density, kinematic viscosity
//     omega | tStep                  relaxation frequency directly, or from the time step (params.h:83-90; exactly one)
//     CENTRAL_MOMENTS [0]            collision operator
//     WALL_LEFT/RIGHT/BOTTOM/TOP     eFluid | eWall | eVelocity | eFreeSlip | ePressure | eConvective  [eWall]
//     PROFILE [uniform]              uniform | eParabolic | eShear | eBoundaryLayer
//     uxInlet_p uyInlet_p ux0_p uy0_p gravityX gravityY dpdx dpdy INLET_RAMP WOMERSLEY   [0 / off]
//     nSteps tinfo tVTK tRestart     time loop and output cadence of src/main.cpp:70-86 (tVTK / tRestart <= 0: never)
//     ref_P [0] ref_L [height_p] ref_nu [nu_p]     reference values of the printed report / the pressure block
//     gpus [1]                       x-slabs = GPUs = host threads
//     exact [0]                      cfg.exact: the reference's operation order, bitwise its results
//     inplace [0]                    cfg.inplace: one population buffer (72 B/node resident)
//     results [Results]              output directory (VTK/ and Restart/ below it, as the reference lays them out)
//
// What it does is src/main.cpp:25-94 + GridClass::GridClass / initialiseGrid (src/Grid.cpp:916-1062, :1232-1289) restated for
// run-time values: scalings, inlet profile, initial u / rho / force_xy, f = f_eq — built slab by slab in column chunks, so a
// 16384^2 lattice never needs the reference's 59 GB of host arrays — then the time loop: life_step; every tinfo the report of
// GridClass::writeInfo from life_max_speed; every tVTK / tRestart the device-fed Fluid.<t>.vti / Fluid.restart (byte-identical to
// the reference's writers); a restart file found at start is read back (src/main.cpp:48-52).  Bodies (geometry.config, FEM, the
// coupling loop) are NOT here: those cases run LIFE's own host code around the library (life_host.cpp, LIFE_B200_GPUS=N).
// There is no CPU path: without a B200 life_create fails and the program exits 99 like the reference's ERROR().
#include "life_b200.h"
#include "rank_team.h"
#include <sys/stat.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

[[noreturn]] void fatal(const std::string &msg) {
	std::printf("\n\nERROR: %s\n\n", msg.c_str());      // the reference's convention (inc/Utils.h:72-77): message, exit(99)
	std::fflush(stdout);
	std::_Exit(99);
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Case {
	std::map<std::string, std::string> kv;
	bool has(const std::string &k) const { return kv.count(k) != 0; }
	double num(const std::string &k, double dflt) const {
		auto it = kv.find(k);
		if (it == kv.end()) return dflt;
		char *end = nullptr;
		const double v = std::strtod(it->second.c_str(), &end);      // decimal or hex-float
		if (end == it->second.c_str()) fatal("case file: '" + k + "' is not a number: " + it->second);
		return v;
	}
	std::string str(const std::string &k, const std::string &dflt) const {
		auto it = kv.find(k);
		return it == kv.end() ? dflt : it->second;
	}
	void set(const std::string &line) {
		std::string l = line.substr(0, line.find('#'));
		const size_t eq = l.find('=');
		if (eq == std::string::npos) return;
		auto trim = [](std::string s) {
			const char *ws = " \t\r\n;";
			const size_t a = s.find_first_not_of(ws), b = s.find_last_not_of(ws);
			return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
		};
		const std::string k = trim(l.substr(0, eq)), v = trim(l.substr(eq + 1));
		if (!k.empty() && !v.empty()) kv[k] = v;
	}
};

int wall_type(const Case &c, const std::string &key) {
	const std::string v = c.str(key, "eWall");
	const char *names[6] = {"eFluid", "eWall", "eVelocity", "eFreeSlip", "ePressure", "eConvective"};      // inc/defs.h:52
	for (int k = 0; k < 6; k++)
		if (v == names[k] || v == std::to_string(k)) return k;
	fatal("case file: " + key + " = " + v + " is not a lattice-site type");
}

// D2Q9 constants, src/Grid.cpp:1247-1250
const int CX[9] = {0, 1, -1, 0, 0, 1, -1, 1, -1}, CY[9] = {0, 0, 0, 1, -1, 1, -1, -1, 1};
const double W[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};

// GridClass::equilibrium, src/Grid.cpp:249-264
double equilibrium(bool cm, double rho, double ux, double uy, int v) {
	const int cx = CX[v], cy = CY[v];
	if (cm)
		return 0.25 * rho * W[v] * (9.0 * (cx * cx) * (ux * ux) + 6.0 * cx * ux - 3.0 * (ux * ux) + 2.0) *
		       (9.0 * (cy * cy) * (uy * uy) + 6.0 * cy * uy - 3.0 * (uy * uy) + 2.0);
	return rho * W[v] * (1.0 + 3.0 * (cx * ux + cy * uy) + 4.5 * ((ux * ux) * ((cx * cx) - 1.0 / 3.0) + (uy * uy) * ((cy * cy) - 1.0 / 3.0)) + 9.0 * cx * cy * ux * uy);
}

}  // namespace

int main(int argc, char **argv) {
	if (argc < 2) {
		std::printf("usage: %s CASEFILE [key=value ...]\n", argv[0]);
		return 2;
	}
	const double t_start = now();
	Case c;
	{
		std::ifstream in(argv[1]);
		if (!in) fatal(std::string("cannot open case file ") + argv[1]);
		std::string line;
		while (std::getline(in, line)) c.set(line);
		for (int a = 2; a < argc; a++) c.set(argv[a]);
	}
	if (!c.has("Nx") || !c.has("Ny") || !c.has("nu_p")) fatal("case file: Nx, Ny and nu_p are required");
	if (c.has("omega") == c.has("tStep")) fatal("case file: give exactly one of omega and tStep (inc/params.h:83-90)");

	// ---- GridClass::GridClass, src/Grid.cpp:1232-1289: scalings ------------------------------------------------------------
	const int64_t Nx = (int64_t)c.num("Nx", 0), Ny = (int64_t)c.num("Ny", 0);
	const double height_p = c.num("height_p", 1.0), rho_p = c.num("rho_p", 1.0), nu_p = c.num("nu_p", 0.0);
	const double c_s = 1.0 / std::sqrt(3.0);
	double omega;
	if (c.has("omega")) omega = c.num("omega", 1.0);
	else omega = 1.0 / (nu_p * c.num("tStep", 0.0) / (std::pow(1.0 / std::sqrt(3.0), 2.0) * std::pow(height_p / (Ny - 1), 2.0)) + 0.5);      // params.h:89-90
	const double rho0 = 1.0;
	const double tau = 1.0 / omega;
	const double nu = (tau - 0.5) * (c_s * c_s);
	const double Dx = height_p / (Ny - 1);
	const double Dt = (Dx * Dx) * nu / nu_p;
	const double Dm = (rho_p / rho0) * (Dx * Dx * Dx);
	const double Drho = (rho_p / rho0);
	const bool cm = c.num("CENTRAL_MOMENTS", 0) != 0;
	const double ramp = c.num("INLET_RAMP", -1.0), womersley = c.num("WOMERSLEY", -1.0);
	const double uxIn = c.num("uxInlet_p", 0.0), uyIn = c.num("uyInlet_p", 0.0), ux0 = c.num("ux0_p", 0.0), uy0 = c.num("uy0_p", 0.0);
	const double gX = c.num("gravityX", 0.0), gY = c.num("gravityY", 0.0), dpdx = c.num("dpdx", 0.0), dpdy = c.num("dpdy", 0.0);
	const int nSteps = (int)c.num("nSteps", 100), tinfo = std::max(1, (int)c.num("tinfo", 10)), tVTK = (int)c.num("tVTK", 0), tRestart = (int)c.num("tRestart", 0);
	const double ref_P = c.num("ref_P", 0.0), ref_L = c.num("ref_L", height_p), ref_nu = c.num("ref_nu", nu_p);
	const int gpus = std::max(1, (int)c.num("gpus", 1));
	const std::string results = c.str("results", "Results"), profile = c.str("PROFILE", "uniform");
	const int walls[4] = {wall_type(c, "WALL_LEFT"), wall_type(c, "WALL_RIGHT"), wall_type(c, "WALL_BOTTOM"), wall_type(c, "WALL_TOP")};

	std::printf("\n*** life_run: %lld x %lld, %s, omega = %.10g, Dx = %.6g, Dt = %.6g, %d step(s), %d GPU(s)%s ***\n", (long long)Nx, (long long)Ny,
	            cm ? "central moments" : "BGK", omega, Dx, Dt, nSteps, gpus, c.num("exact", 0) != 0 ? ", exact mode" : "");

	// ---- contexts: one per GPU, one host thread each -------------------------------------------------------------------------
	life_config cfg{};
	cfg.abi_version = LIFE_ABI_VERSION;
	cfg.collision = cm ? LIFE_CENTRAL_MOMENTS : LIFE_BGK;
	cfg.Nx = Nx; cfg.Ny = Ny; cfg.omega = omega;
	cfg.wall_left = walls[0]; cfg.wall_right = walls[1]; cfg.wall_bottom = walls[2]; cfg.wall_top = walls[3];
	cfg.inlet_ramp = ramp; cfg.Dx = Dx; cfg.Dt = Dt; cfg.Dm = Dm; cfg.Drho = Drho;
	cfg.womersley = womersley; cfg.height_p = height_p; cfg.nu_p = nu_p;
	cfg.gravity_x = gX; cfg.gravity_y = gY; cfg.dpdx = dpdx; cfg.dpdy = dpdy;
	cfg.device = -1; cfg.nranks = 1;
	cfg.exact = c.num("exact", 0) != 0;
	cfg.kernel = (int)c.num("kernel", 0);
	cfg.inplace = c.num("inplace", 0) != 0;
	unsigned char nccl_id[128] = {0};
	if (gpus > 1 && life_nccl_unique_id(nccl_id) != LIFE_OK) fatal(std::string("life_nccl_unique_id: ") + life_last_error(nullptr));
	std::vector<life_ctx *> ctx((size_t)gpus, nullptr);
	std::vector<std::string> errs((size_t)gpus);
	RankTeam team;
	team.start(gpus);
	auto all = [&](const std::function<void(int)> &fn) { team.run(fn); };
	auto ck = [&](int r, int rc, const char *what) {
		if (rc != LIFE_OK) { team.abandon = true; fatal(std::string("liblife_b200: ") + what + " failed on rank " + std::to_string(r) + ": " + life_last_error(ctx[(size_t)r])); }
	};
	all([&](int r) {
		life_config k = cfg;
		if (gpus > 1) { k.rank = r; k.nranks = gpus; k.device = r; k.nccl_id = nccl_id; }
		if (life_create(&k, &ctx[(size_t)r]) != LIFE_OK) errs[(size_t)r] = life_last_error(nullptr);
	});
	for (int r = 0; r < gpus; r++)
		if (!ctx[(size_t)r]) { team.abandon = true; fatal("liblife_b200: life_create failed on rank " + std::to_string(r) + ": " + errs[(size_t)r]); }

	// ---- initialiseGrid, src/Grid.cpp:954-1058 --------------------------------------------------------------------------------
	std::vector<double> u_in((size_t)Ny * 2), rho_in((size_t)Ny, rho0);
	for (int64_t j = 0; j < Ny; j++) {
		double sx, sy;
		if (profile == "uniform") { sx = uxIn * Dt / Dx; sy = uyIn * Dt / Dx; }
		else if (profile == "eParabolic") {
			const double R = height_p / 2.0, YPos = j * Dx - R;
			sx = 1.5 * (uxIn * Dt / Dx) * (1.0 - (YPos / R) * (YPos / R)); sy = 1.5 * (uyIn * Dt / Dx) * (1.0 - (YPos / R) * (YPos / R));
		} else if (profile == "eShear") {
			const double H = height_p, YPos = j * Dx;
			sx = (uxIn * Dt / Dx) * (YPos / H); sy = (uyIn * Dt / Dx) * (YPos / H);
		} else if (profile == "eBoundaryLayer") {
			const double H = height_p, YPos = j * Dx;
			sx = ((1.5 * uxIn * Dt / Dx) / (H * H)) * YPos * (2.0 * H - YPos); sy = ((1.5 * uyIn * Dt / Dx) / (H * H)) * YPos * (2.0 * H - YPos);
		} else fatal("case file: unknown PROFILE " + profile);
		u_in[2 * j] = sx; u_in[2 * j + 1] = sy;
	}
	const double fxy0[2] = {(rho0 * Drho * gX + dpdx) * ((Dx * Dt) * (Dx * Dt)) / Dm, (rho0 * Drho * gY + dpdy) * ((Dx * Dt) * (Dx * Dt)) / Dm};
	mkdir(results.c_str(), 0777);
	mkdir((results + "/VTK").c_str(), 0777);
	mkdir((results + "/Restart").c_str(), 0777);
	const std::string restart_path = results + "/Restart/Fluid.restart";
	int tOffset = 0;
	struct stat sb;
	const bool restart = stat(restart_path.c_str(), &sb) == 0;      // src/main.cpp:48-52, src/Utils.cpp:52
	if (restart) {
		std::vector<int32_t> t_file((size_t)gpus, 0);
		all([&](int r) { ck(r, life_read_restart(ctx[(size_t)r], restart_path.c_str(), fxy0, u_in.data(), rho_in.data(), &t_file[(size_t)r]), "life_read_restart"); });
		tOffset = t_file[0];
		std::printf("\nRestart file found: continuing from time step %d\n", tOffset);
	} else {
		all([&](int r) {
			life_ctx *x = ctx[(size_t)r];
			int64_t i0, i1;
			life_slab(x, &i0, &i1);
			const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(i1 - i0, (int64_t)(64 << 20) / (Ny * 14)));      // ~0.5 GB of host arrays per rank
			std::vector<double> f((size_t)(chunk * Ny * 9)), rho((size_t)(chunk * Ny), rho0), u((size_t)(chunk * Ny * 2)), fxy((size_t)(chunk * Ny * 2));
			ck(r, life_upload_begin(x, u_in.data(), rho_in.data()), "life_upload_begin");
			for (int64_t a = i0; a < i1; a += chunk) {
				const int64_t nc = std::min(chunk, i1 - a);
				for (int64_t i = a; i < a + nc; i++)
					for (int64_t j = 0; j < Ny; j++) {
						const size_t id = (size_t)((i - a) * Ny + j);
						int type = LIFE_FLUID;
						if (i == 0) type = walls[0]; else if (i == Nx - 1) type = walls[1];
						if (j == 0) type = walls[2]; else if (j == Ny - 1) type = walls[3];
						double ux, uy;
						if (ramp > 0.0) { ux = 0.0; uy = 0.0; }
						else if (profile != "uniform") { ux = u_in[2 * j]; uy = u_in[2 * j + 1]; }
						else { ux = ux0 * Dt / Dx; uy = uy0 * Dt / Dx; }
						if (type == LIFE_WALL) { ux = 0.0; uy = 0.0; }
						u[2 * id] = ux; u[2 * id + 1] = uy;
						fxy[2 * id] = fxy0[0]; fxy[2 * id + 1] = fxy0[1];
						for (int v = 0; v < 9; v++) f[9 * id + v] = equilibrium(cm, rho0, ux, uy, v);
					}
				ck(r, life_upload_columns(x, a - i0, nc, f.data(), rho.data(), u.data(), fxy.data(), nullptr), "life_upload_columns");
			}
			ck(r, life_upload_end(x), "life_upload_end");
		});
	}
	const double t_ready = now();

	// ---- the report of GridClass::writeInfo (src/Grid.cpp:559-616) and the files ------------------------------------------------
	double clock0 = now();
	auto write_info = [&](int t) {
		std::vector<double> vmax((size_t)gpus, 0.0);
		std::vector<int32_t> blown((size_t)gpus, 0);
		std::vector<int64_t> bi((size_t)gpus, -1), bj((size_t)gpus, -1);
		all([&](int r) { ck(r, life_max_speed(ctx[(size_t)r], &vmax[(size_t)r], &blown[(size_t)r], &bi[(size_t)r], &bj[(size_t)r]), "life_max_speed"); });
		if (blown[0]) { team.abandon = true; fatal("Simulation blew up (t = " + std::to_string(t) + ") at i = " + std::to_string(bi[0]) + ", j = " + std::to_string(bj[0]) + "...exiting"); }
		const double loop = t == tOffset ? 0.0 : (now() - clock0) / (t - tOffset);
		const double vphys = vmax[0] * Dx / Dt;
		std::printf("\n\nTime step %d of %d\nSimulation has done %.4g of %.4g seconds\nMLUPS = %.4g\nMax Velocity = %.5g\nMax Velocity (m/s) = %.5g\nMax Reynolds number = %.5g",
		            t, tOffset + nSteps, t * Dt, (tOffset + nSteps) * Dt, loop > 0.0 ? Nx * (double)Ny / (1000000.0 * loop) : 0.0, vmax[0], vphys, vphys * ref_L / ref_nu);
		std::fflush(stdout);
	};
	auto write_vtk = [&](int t) {
		const std::string name = results + "/VTK/Fluid." + std::to_string(t) + ".vti";
		all([&](int r) { ck(r, life_write_vtk(ctx[(size_t)r], name.c_str(), rho_p, ref_P, LIFE_IO_ASYNC), "life_write_vtk"); });
	};
	write_info(tOffset);
	if (tVTK > 0) write_vtk(tOffset);
	clock0 = now();

	// ---- the time loop, src/main.cpp:70-86 ---------------------------------------------------------------------------------------
	for (int t = tOffset + 1; t <= tOffset + nSteps; t++) {
		all([&](int r) { ck(r, life_step(ctx[(size_t)r], t), "life_step"); });
		if (t % tinfo == 0) write_info(t);
		if (tVTK > 0 && t % tVTK == 0) write_vtk(t);
		if (tRestart > 0 && t % tRestart == 0)
			all([&](int r) { ck(r, life_write_restart(ctx[(size_t)r], restart_path.c_str(), t, LIFE_IO_ASYNC), "life_write_restart"); });
	}
	all([&](int r) { ck(r, life_sync(ctx[(size_t)r]), "life_sync"); });
	const double t_loop = now();
	all([&](int r) { ck(r, life_io_wait(ctx[(size_t)r]), "life_io_wait"); });
	const long long launches = (long long)life_launch_count(ctx[0]);
	all([&](int r) { life_destroy(ctx[(size_t)r]); });
	const double t_end = now();
	std::printf("\n\n*** FINISHED ***\n\nSimulation took %.2f seconds (context + initial state %.2f s, time loop %.3f s = %.1f MLUPS, last file writes %.2f s); "
	            "%lld kernel launches on rank 0\n\n",
	            t_end - t_start, t_ready - t_start, t_loop - clock0, Nx * (double)Ny * nSteps / (t_loop - clock0) / 1e6, t_end - t_loop, launches);
	return 0;
}
