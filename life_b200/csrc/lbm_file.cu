// Device-fed fluid files: GridClass::writeVTK (src/Grid.cpp:790-898), GridClass::writeRestart
// (src/Grid.cpp:1163-1229) and GridClass::readRestart (src/Grid.cpp:1072-1160) without the host mirrors.
//
// The reference walks its host arrays and calls ofstream::write once per 8-byte value.  Here the bytes of the file are produced
// on the device in FILE ORDER by a pack kernel, cross PCIe through two pinned staging buffers and land with pwrite() at their
// final offsets:
//   .vti      three appended blocks (Density, Pressure, Velocity), each j-major / i-fastest = the transpose of the lattice's
//             i-major / j-fastest storage -> k_vtk_pack transposes 32 x 32 tiles through shared memory and applies the
//             reference's scalings in its operation order (explicit _rn intrinsics: no FMA contraction, bit-exact);
//   .restart  one 120-byte record per node in lattice order (int i, int j, rho, u[2], force_ibm[2], f[9]) = 15 eight-byte
//             words -> k_restart_pack turns 15 SoA planes into AoS records through shared memory, 256 nodes per CTA, so that both
//             the plane reads and the record writes are coalesced; k_restart_unpack is the inverse for reading.
// Chunks alternate between the two staging slots: while the host writes chunk k, the device packs and copies chunk k+1.
//
// LIFE_IO_ASYNC: the state is frozen first (macroscopic planes, plus populations and force_ibm for a restart) with kernels /
// device-to-device copies on the compute stream, then a worker thread drives the chunks on a separate low-priority stream while
// the caller goes on stepping.  If the snapshot does not fit in HBM the call runs synchronously from the live arrays instead.
// With several ranks every rank writes its own byte ranges of the shared file (restart: its contiguous run of records; .vti:
// its segment of every row), rank 0 adds the framing, and life_io_wait() closes with a barrier before the rename.
#include "ctx.h"
#include "d2q9.cuh"
#include "macro.cuh"
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <sstream>
#include <sys/stat.h>
#include <sys/types.h>
#include <system_error>
#include <thread>
#include <unistd.h>

namespace life {

namespace {

// both formats are written in the host's byte order and the reference swaps on big-endian hosts (src/Grid.cpp:793, :1172): this
// library exists for little-endian B200 hosts only, and the .vti head says so
static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, "lbm_file.cu writes little-endian files");

constexpr int RW = 15;   // 8-byte words per restart record: (i | j << 32), rho, ux, uy, fx, fy, f0..f8
constexpr int64_t RESTART_HEAD = 44;   // int t, Nx, Ny; double omega, Dx, Dt, Dm (src/Grid.cpp:1183-1189), unpadded
enum { JOB_VTK = 0, JOB_RESTART = 1 };

// scalings of writeVTK, src/Grid.cpp:861-887
struct VtkScale {
	double Drho;      // density  = rho * Drho
	double rho_off;   // pressure = ref_P + (rho - rho_off) * cs2 * Dm / den, rho_off = rho_p / Drho, den = Dx * SQ(Dt)
	double cs2, Dm, den, ref_P;
	double vel;       // velocity = u * (Dx / Dt)
};

// ---- kernels ----------------------------------------------------------------------------------------------------------------

// One block of the .vti appended data for rows [j0, j0 + nj) of this slab: out[(jj * nxl + il) * comp + k].
// kind 0 = Density, 1 = Pressure (comp 1), 2 = Velocity (comp 3, z = 0).  32 x 8 threads per 32 x 32 tile.
__global__ void __launch_bounds__(256) k_vtk_pack(const MacroArgs a, const double *__restrict__ stored, const VtkScale s,
                                                  const int kind, const int64_t j0, const int64_t nj, double *__restrict__ out) {
	__shared__ double t0[32][33], t1[32][33];
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const int64_t ilb = (int64_t)blockIdx.x * 32, jb = (int64_t)blockIdx.y * 32;
	const int64_t nxl = a.L.nxl;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int64_t il = ilb + ty + 8 * k, jj = jb + tx;   // tx runs along j: coalesced plane reads
		if (il < nxl && jj < nj) {
			const int64_t idx = a.L.node(il, j0 + jj);
			double rho = 0.0, ux = 0.0, uy = 0.0;
			if (stored) {
				if (kind < 2) rho = stored[idx];
				else { ux = stored[a.L.S + idx]; uy = stored[2 * a.L.S + idx]; }
			} else {
				node_macro(a, idx, rho, ux, uy);
			}
			if (kind == 0) {
				t0[ty + 8 * k][tx] = __dmul_rn(rho, s.Drho);
			} else if (kind == 1) {
				// ref_P + (((rho - rho_p / Drho) * SQ(c_s)) * Dm) / (Dx * SQ(Dt)), left to right as C++ parses src/Grid.cpp:873
				t0[ty + 8 * k][tx] = __dadd_rn(s.ref_P, __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(rho, s.rho_off), s.cs2), s.Dm), s.den));
			} else {
				t0[ty + 8 * k][tx] = __dmul_rn(ux, s.vel);
				t1[ty + 8 * k][tx] = __dmul_rn(uy, s.vel);
			}
		}
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int64_t il = ilb + tx, jj = jb + ty + 8 * k;   // tx runs along i: coalesced file-order writes
		if (il < nxl && jj < nj) {
			if (kind < 2) {
				out[jj * nxl + il] = t0[tx][ty + 8 * k];
			} else {
				double *o = out + (jj * nxl + il) * 3;
				o[0] = t0[tx][ty + 8 * k];
				o[1] = t1[tx][ty + 8 * k];
				o[2] = 0.0;
			}
		}
	}
}

// Restart records of local columns [c0, c0 + ncols): out[((c - c0) * Ny + j) * 15 + w].  One CTA per (column, 256-row tile).
__global__ void __launch_bounds__(256) k_restart_pack(const MacroArgs a, const double *__restrict__ stored,
                                                      const double *__restrict__ fibm, const int64_t i_begin, const int64_t c0,
                                                      unsigned long long *__restrict__ out) {
	__shared__ unsigned long long s[RW][257];
	const int64_t Ny = a.L.Ny;
	const int64_t tiles = (Ny + 255) / 256;
	const int64_t colc = blockIdx.x / tiles;
	const int64_t jb = (int64_t)(blockIdx.x % tiles) * 256;
	const int nrows = (int)(Ny - jb < 256 ? Ny - jb : 256);
	const int t = threadIdx.x;
	if (t < nrows) {
		const int64_t il = c0 + colc, j = jb + t, idx = a.L.node(il, j);
		double p[NV], rho, ux, uy;
		if (stored) {
#pragma unroll
			for (int v = 0; v < NV; v++) p[v] = __ldg(a.f + a.ps.at(v, idx, a.L.S));
			rho = stored[idx]; ux = stored[a.L.S + idx]; uy = stored[2 * a.L.S + idx];
		} else {
			node_macro_p(a, idx, p, rho, ux, uy);
		}
		// int i at byte 0, int j at byte 4 (little endian)
		s[0][t] = (unsigned long long)(uint32_t)(i_begin + il) | ((unsigned long long)(uint32_t)j << 32);
		s[1][t] = (unsigned long long)__double_as_longlong(rho);
		s[2][t] = (unsigned long long)__double_as_longlong(ux);
		s[3][t] = (unsigned long long)__double_as_longlong(uy);
		s[4][t] = fibm ? (unsigned long long)__double_as_longlong(fibm[idx]) : 0ull;
		s[5][t] = fibm ? (unsigned long long)__double_as_longlong(fibm[a.L.S + idx]) : 0ull;
#pragma unroll
		for (int v = 0; v < NV; v++) s[6 + v][t] = (unsigned long long)__double_as_longlong(p[v]);
	}
	__syncthreads();
	unsigned long long *o = out + (colc * Ny + jb) * RW;
	for (int e = t; e < nrows * RW; e += 256) o[e] = s[e % RW][e / RW];
}

// Inverse of k_restart_pack: records of local columns [c0, c0 + ncols) -> f, (rho, ux, uy) and force_ibm planes.
__global__ void __launch_bounds__(256) k_restart_unpack(const unsigned long long *__restrict__ in, const Layout L, const int64_t c0,
                                                        double *__restrict__ f, double *__restrict__ macro, double *__restrict__ fibm) {
	__shared__ unsigned long long s[RW][257];
	const int64_t Ny = L.Ny;
	const int64_t tiles = (Ny + 255) / 256;
	const int64_t colc = blockIdx.x / tiles;
	const int64_t jb = (int64_t)(blockIdx.x % tiles) * 256;
	const int nrows = (int)(Ny - jb < 256 ? Ny - jb : 256);
	const int t = threadIdx.x;
	const unsigned long long *src = in + (colc * Ny + jb) * RW;
	for (int e = t; e < nrows * RW; e += 256) s[e % RW][e / RW] = src[e];
	__syncthreads();
	if (t < nrows) {
		const int64_t idx = L.node(c0 + colc, jb + t);
		macro[idx] = __longlong_as_double((long long)s[1][t]);
		macro[L.S + idx] = __longlong_as_double((long long)s[2][t]);
		macro[2 * L.S + idx] = __longlong_as_double((long long)s[3][t]);
		if (fibm) {
			fibm[idx] = __longlong_as_double((long long)s[4][t]);
			fibm[L.S + idx] = __longlong_as_double((long long)s[5][t]);
		}
#pragma unroll
		for (int v = 0; v < NV; v++) f[v * L.S + idx] = __longlong_as_double((long long)s[6 + v][t]);
	}
}

// ---- host side --------------------------------------------------------------------------------------------------------------

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// head and tail of the .vti file exactly as the reference's stream insertions produce them (src/Grid.cpp:796-855, :891-898):
// default ostream formatting (6 significant digits, %g style) for the doubles 0.0 and Dx, integers for Nx-1, Ny-1 and offsets.
void vtk_frame(int64_t Nx, int64_t Ny, double Dx, std::string &head, std::string &tail) {
	const unsigned long long n8 = (unsigned long long)Nx * (unsigned long long)Ny * sizeof(double);
	std::ostringstream o;
	o << "<?xml version=\"1.0\"?>\n";
	o << "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
	o << std::string(1, '\t') << "<ImageData "
	  << "WholeExtent=\"" << 0.0 << " " << Nx - 1 << " " << 0.0 << " " << Ny - 1 << " " << 0.0 << " " << 0.0 << "\" "
	  << "Origin=\"" << 0.0 << " " << 0.0 << " " << 0.0 << "\" "
	  << "Spacing=\"" << Dx << " " << Dx << " " << Dx << "\">\n";
	o << std::string(2, '\t') << "<Piece Extent=\"" << 0.0 << " " << Nx - 1 << " " << 0.0 << " " << Ny - 1 << " " << 0.0 << " " << 0.0 << "\">\n";
	o << std::string(3, '\t') << "<PointData>\n";
	o << std::string(4, '\t') << "<DataArray type=\"Float64\" Name=\"Density\" format=\"appended\" offset=\"" << 0 << "\"/>\n";
	o << std::string(4, '\t') << "<DataArray type=\"Float64\" Name=\"Pressure\" format=\"appended\" offset=\"" << 1 * (n8 + sizeof(unsigned long long)) << "\"/>\n";
	o << std::string(4, '\t') << "<DataArray type=\"Float64\" Name=\"Velocity\" NumberOfComponents=\"3\" format=\"appended\" offset=\"" << 2 * (n8 + sizeof(unsigned long long)) << "\"/>\n";
	o << std::string(3, '\t') << "</PointData>\n";
	o << std::string(2, '\t') << "</Piece>\n";
	o << std::string(1, '\t') << "</ImageData>\n";
	o << std::string(1, '\t') << "<AppendedData encoding=\"raw\">\n";
	o << std::string(2, '\t') << "_";
	head = o.str();
	tail = "\n" + std::string(1, '\t') + "</AppendedData>\n" + "</VTKFile>\n";
}

bool write_at(int fd, const void *buf, size_t n, int64_t off, std::string &err) {
	const char *p = static_cast<const char *>(buf);
	while (n > 0) {
		const ssize_t w = pwrite(fd, p, n, (off_t)off);
		if (w < 0) {
			if (errno == EINTR) continue;
			err = std::string("pwrite: ") + strerror(errno);
			return false;
		}
		p += w; off += w; n -= (size_t)w;
	}
	return true;
}

bool read_at(int fd, void *buf, size_t n, int64_t off, std::string &err) {
	char *p = static_cast<char *>(buf);
	while (n > 0) {
		const ssize_t r = pread(fd, p, n, (off_t)off);
		if (r < 0) {
			if (errno == EINTR) continue;
			err = std::string("pread: ") + strerror(errno);
			return false;
		}
		if (r == 0) { err = "unexpected end of file"; return false; }
		p += r; off += r; n -= (size_t)r;
	}
	return true;
}

// One contiguous run of a chunk <-> the file, split over a few host threads: the copy between the pinned buffer and the page
// cache (and, for a fresh file, the page and block allocation behind it) is what bounds these paths, and it scales with cores.
int io_threads() {
	static const int n = [] {
		const unsigned hc = std::thread::hardware_concurrency();
		return (int)std::max(1u, std::min(8u, hc / 2));
	}();
	return n;
}

template <bool WRITE>
bool transfer_at(int fd, char *buf, size_t n, int64_t off, std::string &err) {
	constexpr size_t MIN_PIECE = (size_t)4 << 20;
	const int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)io_threads(), n / MIN_PIECE));
	if (parts == 1) return WRITE ? write_at(fd, buf, n, off, err) : read_at(fd, buf, n, off, err);
	const size_t piece = ((n / parts) + 4095) & ~(size_t)4095;
	std::vector<std::string> errs((size_t)parts);
	std::vector<char> okv((size_t)parts, 1);
	std::vector<std::thread> th;
	auto work = [&](int k) {
		const size_t b = (size_t)k * piece, e = std::min(n, b + piece);
		if (b >= e) return;
		okv[(size_t)k] = (WRITE ? write_at(fd, buf + b, e - b, off + (int64_t)b, errs[(size_t)k]) : read_at(fd, buf + b, e - b, off + (int64_t)b, errs[(size_t)k])) ? 1 : 0;
	};
	for (int k = 1; k < parts; k++) {
		try { th.emplace_back(work, k); }
		catch (const std::system_error &) { work(k); }   // no thread to be had: do that piece here
	}
	work(0);
	for (auto &t : th) t.join();
	for (int k = 0; k < parts; k++)
		if (!okv[(size_t)k]) { err = errs[(size_t)k]; return false; }
	return true;
}

// Restart records of `ncols` whole columns starting at global column i0: false if a record does not carry its own (i, j);
// any_force = some record holds a non-zero force_ibm.  Columns are split over the host threads.
bool scan_records(const char *h, int64_t ncols, int64_t Ny, int64_t i0, bool &any_force) {
	const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(io_threads(), ncols * Ny / 65536));
	std::vector<char> good((size_t)parts, 1), force((size_t)parts, 0);
	auto work = [&](int k) {
		const int64_t cb = ncols * k / parts, ce = ncols * (k + 1) / parts;
		bool g = true, f = false;
		for (int64_t c = cb; c < ce && g; c++) {
			const char *rec = h + c * Ny * 8 * RW;
			const int32_t i = (int32_t)(i0 + c);
			for (int64_t j = 0; j < Ny; j++, rec += 8 * RW) {
				int32_t ij[2];
				double fxy[2];
				memcpy(ij, rec, 8);
				memcpy(fxy, rec + 32, 16);
				if (ij[0] != i || ij[1] != (int32_t)j) { g = false; break; }
				f = f || fxy[0] != 0.0 || fxy[1] != 0.0;
			}
		}
		good[(size_t)k] = g;
		force[(size_t)k] = f;
	};
	std::vector<std::thread> th;
	for (int k = 1; k < parts; k++) {
		try { th.emplace_back(work, k); }
		catch (const std::system_error &) { work(k); }
	}
	work(0);
	for (auto &t : th) t.join();
	any_force = false;
	for (int k = 0; k < parts; k++) {
		if (!good[(size_t)k]) return false;
		any_force = any_force || force[(size_t)k];
	}
	return true;
}

struct FileJob {
	int kind = JOB_VTK;
	std::string path;                 // where the bytes go (restart: the .temp name)
	std::string final_path;           // restart: renamed onto this once complete
	MacroArgs src{};                  // populations / forces the values are evaluated from (live arrays or the snapshot)
	const double *stored = nullptr;   // rho, ux, uy planes, or null: evaluate from src
	const double *fibm = nullptr;     // force_ibm planes for the restart records, or null: zeros
	int64_t Nx = 0, i_begin = 0;
	int rank = 0, nranks = 1;
	double Dx = 0.0;
	VtkScale vs{};
	char head[RESTART_HEAD] = {0};    // restart header bytes
	cudaStream_t st = nullptr;
	int device = 0;
	bool async = false;
};

struct Chunk {
	int block;          // .vti: 0 / 1 / 2 ; restart: unused
	int64_t a0, an;     // .vti: rows [a0, a0 + an) ; restart: local columns [a0, a0 + an)
};

}  // namespace

struct IoState {
	cudaStream_t stream = nullptr;               // pack kernels + D2H copies of an asynchronous job
	cudaEvent_t ev_snap = nullptr;               // snapshot complete (recorded on the compute stream)
	cudaEvent_t ev_slot[2] = {nullptr, nullptr};
	void *d_stage[2] = {nullptr, nullptr};       // device staging, file byte layout
	void *h_stage[2] = {nullptr, nullptr};       // pinned host staging
	size_t stage_bytes = 0;                      // capacity of each of the four
	size_t want_bytes = (size_t)32 << 20;
	double *snap = nullptr;                      // frozen copy of what an asynchronous job reads
	size_t snap_bytes = 0;
	std::thread worker;
	bool pending = false;                        // a job has been started and not yet completed by io_wait
	std::atomic<bool> finished{true};            // the thread running the job has written its result fields
	FileJob job;
	// result of the job (written by the thread that runs it, read after the join)
	int rc = LIFE_OK;
	std::string err;
	int64_t launches = 0, bytes = 0;
	double seconds = 0.0;
	bool was_async = false;
};

namespace {

// Runs on the caller's thread (synchronous) or on the worker (asynchronous).  Touches only the job, the staging buffers and the
// result fields of IoState.
void run_job_body(IoState *io);
void run_job(IoState *io) {
	run_job_body(io);
	io->finished.store(true, std::memory_order_release);
}

void run_job_body(IoState *io) {
	const FileJob &job = io->job;
	const double t_begin = now_s();
	io->rc = LIFE_OK;
	io->err.clear();
	io->launches = 0;
	io->bytes = 0;
	auto cuda_fail = [&](const char *what, cudaError_t e) {
		io->rc = LIFE_E_CUDA;
		io->err = std::string(what) + ": " + cudaGetErrorString(e);
	};
	cudaError_t ce = cudaSetDevice(job.device);
	if (ce != cudaSuccess) { cuda_fail("cudaSetDevice", ce); return; }

	const Layout &L = job.src.L;
	const int64_t Nx = job.Nx, Ny = L.Ny, nxl = L.nxl;
	const int fd = open(job.path.c_str(), O_WRONLY | O_CREAT, 0666);
	if (fd < 0) {
		io->rc = LIFE_E_IO;
		io->err = "cannot open " + job.path + " for writing: " + strerror(errno);
		return;
	}

	// framing (rank 0) and the list of chunks
	std::vector<Chunk> chunks;
	int64_t data_off[3] = {0, 0, 0};
	bool ok = true;
	if (job.kind == JOB_VTK) {
		std::string head, tail;
		vtk_frame(Nx, Ny, job.Dx, head, tail);
		const int64_t n8 = Nx * Ny * 8;
		data_off[0] = (int64_t)head.size() + 8;
		data_off[1] = data_off[0] + n8 + 8;
		data_off[2] = data_off[1] + n8 + 8;
		const int64_t end = data_off[2] + 3 * n8;
		if (job.rank == 0) {
			const unsigned long long sz1 = (unsigned long long)n8, sz3 = 3ull * (unsigned long long)n8;
			ok = ftruncate(fd, (off_t)(end + (int64_t)tail.size())) == 0;
			if (!ok) io->err = std::string("ftruncate: ") + strerror(errno);
			ok = ok && write_at(fd, head.data(), head.size(), 0, io->err) && write_at(fd, &sz1, 8, data_off[0] - 8, io->err) &&
			     write_at(fd, &sz1, 8, data_off[1] - 8, io->err) && write_at(fd, &sz3, 8, data_off[2] - 8, io->err) &&
			     write_at(fd, tail.data(), tail.size(), end, io->err);
			io->bytes += (int64_t)head.size() + 24 + (int64_t)tail.size();
		}
		for (int b = 0; b < 3; b++) {
			const int64_t row_bytes = nxl * 8 * (b == 2 ? 3 : 1);
			int64_t rows = (int64_t)io->stage_bytes / row_bytes;
			if (rows >= 32) rows -= rows % 32;
			rows = std::max<int64_t>(1, std::min<int64_t>(rows, 32 * 65535));
			for (int64_t j0 = 0; j0 < Ny; j0 += rows) chunks.push_back({b, j0, std::min(rows, Ny - j0)});
		}
	} else {
		if (job.rank == 0) {
			ok = ftruncate(fd, (off_t)(RESTART_HEAD + Nx * Ny * 8 * RW)) == 0;
			if (!ok) io->err = std::string("ftruncate: ") + strerror(errno);
			ok = ok && write_at(fd, job.head, RESTART_HEAD, 0, io->err);
			io->bytes += RESTART_HEAD;
		}
		const int64_t col_bytes = Ny * 8 * RW;
		const int64_t cols = std::max<int64_t>(1, (int64_t)io->stage_bytes / col_bytes);
		for (int64_t c0 = 0; c0 < nxl; c0 += cols) chunks.push_back({0, c0, std::min(cols, nxl - c0)});
	}
	if (!ok) { io->rc = LIFE_E_IO; close(fd); return; }

	auto chunk_bytes = [&](const Chunk &c) -> size_t {
		return job.kind == JOB_VTK ? (size_t)(c.an * nxl * 8 * (c.block == 2 ? 3 : 1)) : (size_t)(c.an * Ny * 8 * RW);
	};
	// pack chunk c into device slot s, copy it to pinned slot s, mark completion
	auto enqueue = [&](const Chunk &c, int s) -> bool {
		if (job.kind == JOB_VTK) {
			const dim3 grid((unsigned)((nxl + 31) / 32), (unsigned)((c.an + 31) / 32));
			k_vtk_pack<<<grid, 256, 0, job.st>>>(job.src, job.stored, job.vs, c.block, c.a0, c.an, static_cast<double *>(io->d_stage[s]));
		} else {
			const int64_t blocks = ((Ny + 255) / 256) * c.an;
			k_restart_pack<<<(unsigned)blocks, 256, 0, job.st>>>(job.src, job.stored, job.fibm, job.i_begin, c.a0,
			                                                    static_cast<unsigned long long *>(io->d_stage[s]));
		}
		io->launches++;
		cudaError_t e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaMemcpyAsync(io->h_stage[s], io->d_stage[s], chunk_bytes(c), cudaMemcpyDeviceToHost, job.st);
		if (e == cudaSuccess) e = cudaEventRecord(io->ev_slot[s], job.st);
		if (e != cudaSuccess) { cuda_fail("file chunk", e); return false; }
		return true;
	};
	// wait for slot s and put its bytes where they belong in the file
	auto flush = [&](const Chunk &c, int s) -> bool {
		const cudaError_t e = cudaEventSynchronize(io->ev_slot[s]);
		if (e != cudaSuccess) { cuda_fail("file chunk", e); return false; }
		char *h = static_cast<char *>(io->h_stage[s]);
		bool w = true;
		if (job.kind == JOB_RESTART) {
			w = transfer_at<true>(fd, h, chunk_bytes(c), RESTART_HEAD + ((job.i_begin + c.a0) * Ny) * 8 * RW, io->err);
		} else {
			const int64_t comp8 = 8 * (c.block == 2 ? 3 : 1);
			if (nxl == Nx) {   // whole rows: one contiguous run
				w = transfer_at<true>(fd, h, chunk_bytes(c), data_off[c.block] + c.a0 * Nx * comp8, io->err);
			} else {           // this slab's segment of every row
				for (int64_t r = 0; r < c.an && w; r++)
					w = write_at(fd, h + r * nxl * comp8, (size_t)(nxl * comp8), data_off[c.block] + ((c.a0 + r) * Nx + job.i_begin) * comp8, io->err);
			}
		}
		if (!w) { io->rc = LIFE_E_IO; return false; }
		io->bytes += (int64_t)chunk_bytes(c);
		return true;
	};

	const size_t nc = chunks.size();
	for (size_t k = 0; k < nc && ok; k++) {
		ok = enqueue(chunks[k], (int)(k & 1));
		if (ok && k >= 1) ok = flush(chunks[k - 1], (int)((k - 1) & 1));
	}
	if (ok && nc >= 1) ok = flush(chunks[nc - 1], (int)((nc - 1) & 1));
	if (!ok) cudaStreamSynchronize(job.st);   // nothing may still be writing into the staging buffers when we return
	if (close(fd) != 0 && ok) {
		ok = false;
		io->rc = LIFE_E_IO;
		io->err = std::string("close: ") + strerror(errno);
	}
	// single rank: the file is complete, put it in place (src/Grid.cpp:1228); several ranks: life_io_wait does it after a barrier
	if (ok && job.kind == JOB_RESTART && job.nranks <= 1 && rename(job.path.c_str(), job.final_path.c_str()) != 0) {
		io->rc = LIFE_E_IO;
		io->err = "rename " + job.path + ": " + strerror(errno);
	}
	io->seconds = now_s() - t_begin;
}

int ensure_io(life_ctx *ctx) {
	if (ctx->io) return LIFE_OK;
	IoState *io = new (std::nothrow) IoState();
	if (!io) return fail(ctx, LIFE_E_NOMEM, "file path: out of host memory");
	ctx->io = io;
	int lo = 0, hi = 0;
	LIFE_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = numerically largest = lowest priority
	LIFE_CUDA(ctx, cudaStreamCreateWithPriority(&io->stream, cudaStreamNonBlocking, lo));
	LIFE_CUDA(ctx, cudaEventCreateWithFlags(&io->ev_snap, cudaEventDisableTiming));
	for (int s = 0; s < 2; s++) LIFE_CUDA(ctx, cudaEventCreateWithFlags(&io->ev_slot[s], cudaEventDisableTiming | cudaEventBlockingSync));
	return LIFE_OK;
}

// two device + two pinned staging buffers, each able to hold at least one .vti velocity row and one restart column
int ensure_staging(life_ctx *ctx) {
	IoState *io = ctx->io;
	size_t need = std::min(io->want_bytes, (size_t)(ctx->L.nxl * ctx->L.Ny * 8 * RW));   // never more than this slab's restart records
	need = std::max(need, (size_t)(ctx->L.nxl * 24));
	need = std::max(need, (size_t)(ctx->L.Ny * 8 * RW));
	need = (need + 255) & ~(size_t)255;
	if (io->stage_bytes == need) return LIFE_OK;
	for (int s = 0; s < 2; s++) {
		if (io->d_stage[s]) cudaFree(io->d_stage[s]);
		if (io->h_stage[s]) cudaFreeHost(io->h_stage[s]);
		io->d_stage[s] = io->h_stage[s] = nullptr;
	}
	io->stage_bytes = 0;
	for (int s = 0; s < 2; s++) {
		LIFE_CUDA(ctx, cudaMalloc(&io->d_stage[s], need));
		LIFE_CUDA(ctx, cudaMallocHost(&io->h_stage[s], need));
	}
	io->stage_bytes = need;
	return LIFE_OK;
}

// room for the frozen copy; false (and no error) if HBM cannot hold it
bool ensure_snapshot(life_ctx *ctx, size_t bytes) {
	IoState *io = ctx->io;
	if (io->snap_bytes >= bytes) return true;
	if (io->snap) cudaFree(io->snap);
	io->snap = nullptr;
	io->snap_bytes = 0;
	if (cudaMalloc(&io->snap, bytes) != cudaSuccess) {
		cudaGetLastError();   // clear the allocation failure: the caller falls back to the synchronous path
		io->snap = nullptr;
		return false;
	}
	io->snap_bytes = bytes;
	return true;
}

// common front half of life_write_vtk / life_write_restart; `job` arrives with kind, paths and scalings filled in
int start_job(life_ctx *ctx, FileJob &job, int mode) {
	if (!ctx->have_state) return fail(ctx, LIFE_E_STATE, "file output: no state uploaded");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_io(ctx);
	if (rc) return rc;
	if ((rc = io_wait(ctx))) return rc;   // completes (and reports) the previous asynchronous write
	IoState *io = ctx->io;
	// A rank that fails before its job exists must still take part in the collective completion (io_wait's status exchange), or the
	// other ranks would wait for it forever: it enters the exchange with its error and everybody returns one.
	auto early = [&](int code) {
		if (!ctx->comm) return code;
		io->job = job;
		io->rc = code;
		io->err = ctx->err;
		io->pending = true;
		io->finished.store(true, std::memory_order_relaxed);
		return io_wait(ctx);
	};
	if ((rc = ensure_staging(ctx))) return early(rc);
	const Layout &L = ctx->L;
	const size_t S = (size_t)L.S;

	job.src = macro_args(ctx);
	job.Nx = ctx->cfg.Nx;
	job.i_begin = ctx->i_begin;
	job.rank = ctx->cfg.rank;
	job.nranks = ctx->cfg.nranks;
	job.Dx = ctx->cfg.Dx;
	job.device = ctx->device;
	const bool restart = job.kind == JOB_RESTART;
	const bool with_fibm = restart && ctx->fibm != nullptr;

	bool async = mode == LIFE_IO_ASYNC;
	if (async) async = ensure_snapshot(ctx, sizeof(double) * S * (restart ? (with_fibm ? 14 : 12) : 3));
	if (async) {
		// freeze the state on the compute stream: [rho ux uy | f0..f8 | fx fy]
		double *snap = io->snap;
		if (ctx->stored_macro_valid) {
			LIFE_CUDA(ctx, cudaMemcpyAsync(snap, ctx->macro, sizeof(double) * 3 * S, cudaMemcpyDeviceToDevice, ctx->stream));
		} else if ((rc = launch_macro(ctx, snap, 0, L.nxl))) {
			return early(rc);
		}
		job.stored = snap;
		if (restart) {
			LIFE_CUDA(ctx, cudaMemcpyAsync(snap + 3 * S, ctx->fA, sizeof(double) * 9 * S, cudaMemcpyDeviceToDevice, ctx->stream));
			job.src.f = snap + 3 * S;
			if (with_fibm) {
				LIFE_CUDA(ctx, cudaMemcpyAsync(snap + 12 * S, ctx->fibm, sizeof(double) * 2 * S, cudaMemcpyDeviceToDevice, ctx->stream));
				job.fibm = snap + 12 * S;
			}
		}
		LIFE_CUDA(ctx, cudaEventRecord(io->ev_snap, ctx->stream));
		LIFE_CUDA(ctx, cudaStreamWaitEvent(io->stream, io->ev_snap, 0));
		job.st = io->stream;
	} else {
		job.stored = ctx->stored_macro_valid ? ctx->macro : nullptr;
		job.fibm = with_fibm ? ctx->fibm : nullptr;
		job.st = ctx->stream;
	}
	job.async = async;
	io->job = job;
	io->was_async = async;
	io->pending = true;
	io->finished.store(false, std::memory_order_relaxed);
	if (async) {
		try {
			io->worker = std::thread(run_job, io);
			return LIFE_OK;
		} catch (const std::system_error &) {
			// no worker thread to be had: the job (already pointed at the snapshot and the side stream) runs here instead
		}
	}
	run_job(io);
	return io_wait(ctx);
}

}  // namespace

// All ranks have reached this point AND know whether any of them failed: a one-element max all-reduce of each rank's status on the
// compute stream, then a host wait.  *any receives the largest status (0 = every rank is fine).
static int comm_agree(life_ctx *ctx, int mine, int *any) {
	if (!ctx->d_red) LIFE_CUDA(ctx, cudaMalloc(&ctx->d_red, 64));
	double *h = reinterpret_cast<double *>(ctx->h_pin) + 8;
	*h = mine != LIFE_OK ? 1.0 : 0.0;
	LIFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_red + 4, h, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	LIFE_NCCL(ctx, ncclAllReduce(ctx->d_red + 4, ctx->d_red + 4, 1, ncclDouble, ncclMax, ctx->comm, ctx->stream));
	LIFE_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_red + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	LIFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	*any = *h != 0.0 ? 1 : 0;
	return LIFE_OK;
}

int io_wait(life_ctx *ctx) {
	IoState *io = ctx->io;
	if (!io || !io->pending) return LIFE_OK;
	if (io->worker.joinable()) io->worker.join();
	io->pending = false;
	ctx->launches += io->launches;
	int rc = io->rc;
	std::string err = io->err;
	if (ctx->comm) {
		// Every rank's bytes must be in the file before rank 0 renames it — and the rename must not happen at all if ANY rank failed
		// (a truncated .temp must never replace the last good restart file): the first exchange carries every rank's status.
		// The rename must have happened before any rank goes on (to read the file back, or to open the next .temp): second exchange,
		// which also spreads a failed rename.
		int any = 0;
		int brc = comm_agree(ctx, rc, &any);
		if (brc) return brc;
		if (any == 0 && io->job.kind == JOB_RESTART && ctx->cfg.rank == 0 &&
		    rename(io->job.path.c_str(), io->job.final_path.c_str()) != 0) {
			rc = LIFE_E_IO;
			err = "rename " + io->job.path + ": " + strerror(errno);
		}
		int any2 = 0;
		if ((brc = comm_agree(ctx, rc, &any2))) return brc;
		if (rc == LIFE_OK && (any || any2)) {
			rc = LIFE_E_IO;
			err = "file output failed on another rank; " + (io->job.kind == JOB_RESTART ? "the previous " + io->job.final_path + " was kept" : io->job.path + " is incomplete");
		}
	}
	if (rc) return fail(ctx, rc, err);
	return LIFE_OK;
}

void io_free(life_ctx *ctx) {
	IoState *io = ctx->io;
	if (!io) return;
	if (io->worker.joinable()) io->worker.join();
	for (int s = 0; s < 2; s++) {
		if (io->d_stage[s]) cudaFree(io->d_stage[s]);
		if (io->h_stage[s]) cudaFreeHost(io->h_stage[s]);
		if (io->ev_slot[s]) cudaEventDestroy(io->ev_slot[s]);
	}
	if (io->snap) cudaFree(io->snap);
	if (io->ev_snap) cudaEventDestroy(io->ev_snap);
	if (io->stream) cudaStreamDestroy(io->stream);
	delete io;
	ctx->io = nullptr;
}

}  // namespace life

using namespace life;

extern "C" {

int life_vtk_frame(int64_t Nx, int64_t Ny, double Dx, char *head, int64_t head_cap, int64_t *head_len, char *tail,
                   int64_t tail_cap, int64_t *tail_len) {
	if (Nx < 1 || Ny < 1) return LIFE_E_ARG;
	std::string h, t;
	vtk_frame(Nx, Ny, Dx, h, t);
	if (head_len) *head_len = (int64_t)h.size();
	if (tail_len) *tail_len = (int64_t)t.size();
	if (head) {
		if (head_cap < (int64_t)h.size()) return LIFE_E_ARG;
		memcpy(head, h.data(), h.size());
	}
	if (tail) {
		if (tail_cap < (int64_t)t.size()) return LIFE_E_ARG;
		memcpy(tail, t.data(), t.size());
	}
	return LIFE_OK;
}

int life_write_vtk(life_ctx *ctx, const char *path, double rho_p, double ref_P, int32_t mode) {
	if (!ctx) return LIFE_E_ARG;
	if (!path || !*path) return fail(ctx, LIFE_E_ARG, "life_write_vtk: no path");
	if (mode != LIFE_IO_SYNC && mode != LIFE_IO_ASYNC) return fail(ctx, LIFE_E_ARG, "life_write_vtk: unknown mode");
	const life_config &c = ctx->cfg;
	FileJob job;
	job.kind = JOB_VTK;
	job.path = path;
	// the reference's expressions, evaluated once on the host exactly as it evaluates them per node (src/Grid.cpp:863, :873, :883)
	const double c_s = 1.0 / sqrt(3.0);   // src/Grid.cpp:1246
	job.vs.Drho = c.Drho;
	job.vs.rho_off = rho_p / c.Drho;
	job.vs.cs2 = c_s * c_s;
	job.vs.Dm = c.Dm;
	job.vs.den = c.Dx * (c.Dt * c.Dt);
	job.vs.ref_P = ref_P;
	job.vs.vel = c.Dx / c.Dt;
	return start_job(ctx, job, mode);
}

int life_write_restart(life_ctx *ctx, const char *path, int32_t t, int32_t mode) {
	if (!ctx) return LIFE_E_ARG;
	if (!path || !*path) return fail(ctx, LIFE_E_ARG, "life_write_restart: no path");
	if (mode != LIFE_IO_SYNC && mode != LIFE_IO_ASYNC) return fail(ctx, LIFE_E_ARG, "life_write_restart: unknown mode");
	const life_config &c = ctx->cfg;
	if (c.Nx > INT32_MAX || c.Ny > INT32_MAX) return fail(ctx, LIFE_E_ARG, "life_write_restart: the file format holds 32-bit indices");
	FileJob job;
	job.kind = JOB_RESTART;
	job.final_path = path;
	job.path = std::string(path) + ".temp";   // src/Grid.cpp:1167
	const int32_t hi[3] = {t, (int32_t)c.Nx, (int32_t)c.Ny};
	const double hd[4] = {c.omega, c.Dx, c.Dt, c.Dm};
	memcpy(job.head, hi, 12);
	memcpy(job.head + 12, hd, 32);
	return start_job(ctx, job, mode);
}

int life_io_wait(life_ctx *ctx) {
	if (!ctx) return LIFE_E_ARG;
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	return io_wait(ctx);
}

int life_io_busy(life_ctx *ctx, int32_t *busy) {
	if (!ctx || !busy) return LIFE_E_ARG;
	const IoState *io = ctx->io;
	*busy = io && io->pending && !io->finished.load(std::memory_order_acquire) ? 1 : 0;
	return LIFE_OK;
}

int life_io_stats(life_ctx *ctx, double *seconds, int64_t *bytes, int32_t *was_async) {
	if (!ctx) return LIFE_E_ARG;
	const IoState *io = ctx->io;
	const bool done = io && !io->pending;
	if (seconds) *seconds = done ? io->seconds : 0.0;
	if (bytes) *bytes = done ? io->bytes : 0;
	if (was_async) *was_async = done && io->was_async ? 1 : 0;
	return LIFE_OK;
}

int life_io_set_staging(life_ctx *ctx, int64_t bytes) {
	if (!ctx) return LIFE_E_ARG;
	if (bytes < 256) return fail(ctx, LIFE_E_ARG, "life_io_set_staging: at least 256 bytes");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_io(ctx);
	if (rc) return rc;
	if ((rc = io_wait(ctx))) return rc;
	ctx->io->want_bytes = (size_t)bytes;
	return LIFE_OK;
}

int life_read_restart(life_ctx *ctx, const char *path, const double *force_xy, const double *u_in, const double *rho_in,
                      int32_t *t_out) {
	if (!ctx) return LIFE_E_ARG;
	if (!path || !*path) return fail(ctx, LIFE_E_ARG, "life_read_restart: no path");
	LIFE_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = ensure_io(ctx);
	if (rc) return rc;
	if ((rc = io_wait(ctx))) return rc;
	if ((rc = ensure_staging(ctx))) return rc;
	IoState *io = ctx->io;
	const life_config &c = ctx->cfg;
	const Layout &L = ctx->L;

	const int fd = open(path, O_RDONLY);
	if (fd < 0) return fail(ctx, LIFE_E_IO, "Error opening Fluid.restart file...exiting (" + std::string(path) + ": " + strerror(errno) + ")");
	struct Closer { int fd; ~Closer() { close(fd); } } closer{fd};
	std::string err;
	char head[RESTART_HEAD];
	if (!read_at(fd, head, RESTART_HEAD, 0, err)) return fail(ctx, LIFE_E_IO, "life_read_restart: header: " + err);
	int32_t hi[3];
	double hd[4];
	memcpy(hi, head, 12);
	memcpy(hd, head + 12, 32);
	// src/Grid.cpp:1103-1104
	if (c.Nx != hi[1] || c.Ny != hi[2] || c.omega != hd[0] || c.Dx != hd[1] || c.Dt != hd[2] || c.Dm != hd[3])
		return fail(ctx, LIFE_E_ARG, "Grid size/scaling has changed between runs...this is not supported");
	struct stat sb;
	if (fstat(fd, &sb) != 0) return fail(ctx, LIFE_E_IO, std::string("fstat: ") + strerror(errno));
	if ((int64_t)sb.st_size < RESTART_HEAD + c.Nx * c.Ny * 8 * RW)
		return fail(ctx, LIFE_E_IO, "life_read_restart: " + std::string(path) + " is shorter than its header says");

	if ((rc = life_upload_begin(ctx, u_in, rho_in))) return rc;
	if ((rc = ensure_macro(ctx))) return rc;
	// Womersley with gravity keeps force_xy as a field the sweep recomputes before its first use (src/Grid.cpp:55-61); until then it
	// holds what initialiseGrid set, as in the reference after readRestart
	if (ctx->wom_field && force_xy && (rc = fill_field(ctx, ctx->fxyf, 2, 0, L.nxl, force_xy[0], force_xy[1]))) return rc;
	const int64_t col_bytes = L.Ny * 8 * RW;
	const int64_t cols = std::max<int64_t>(1, (int64_t)io->stage_bytes / col_bytes);
	bool any_fibm = false;
	int64_t k = 0;
	for (int64_t c0 = 0; c0 < L.nxl; c0 += cols, k++) {
		const int s = (int)(k & 1);
		const int64_t nc = std::min(cols, L.nxl - c0);
		const size_t bytes = (size_t)(nc * col_bytes);
		if (k >= 2) LIFE_CUDA(ctx, cudaEventSynchronize(io->ev_slot[s]));   // the copy out of this slot two chunks ago
		char *h = static_cast<char *>(io->h_stage[s]);
		if (!transfer_at<false>(fd, h, bytes, RESTART_HEAD + (ctx->i_begin + c0) * col_bytes, err)) {
			cudaStreamSynchronize(ctx->stream);
			return fail(ctx, LIFE_E_IO, "life_read_restart: " + err);
		}
		// src/Grid.cpp:1137-1138: every record must sit at its own (i, j); and is there any IBM force at all?
		bool chunk_fibm = false;
		if (!scan_records(h, nc, L.Ny, ctx->i_begin + c0, chunk_fibm)) {
			cudaStreamSynchronize(ctx->stream);
			return fail(ctx, LIFE_E_ARG, "Grid indices do not match Fluid.restart file...exiting");
		}
		if (chunk_fibm && !any_fibm) {
			if ((rc = ensure_fibm(ctx))) return rc;   // zero-filled: the chunks before this one held no force
			any_fibm = true;
		}
		LIFE_CUDA(ctx, cudaMemcpyAsync(io->d_stage[s], h, bytes, cudaMemcpyHostToDevice, ctx->stream));
		LIFE_CUDA(ctx, cudaEventRecord(io->ev_slot[s], ctx->stream));
		const int64_t blocks = ((L.Ny + 255) / 256) * nc;
		k_restart_unpack<<<(unsigned)blocks, 256, 0, ctx->stream>>>(static_cast<const unsigned long long *>(io->d_stage[s]), L, c0, ctx->fA,
		                                                          ctx->macro, any_fibm ? ctx->fibm : nullptr);
		ctx->launches++;
		LIFE_CUDA(ctx, cudaGetLastError());
	}
	// what life_upload_columns would have recorded for (f, rho, u, uniform force_xy, force_ibm) over all columns
	ctx->up_macro = 1;
	ctx->up_cols = L.nxl;
	ctx->up_fxy_seen = true;
	ctx->up_fxy_uniform = true;
	ctx->up_fxy0[0] = force_xy ? force_xy[0] : 0.0;
	ctx->up_fxy0[1] = force_xy ? force_xy[1] : 0.0;
	if (any_fibm) {
		ctx->fibm_any = true;
		ctx->fibm_full_dirty = true;
		ctx->fibm_consumed = false;
	}
	if ((rc = life_upload_end(ctx))) return rc;
	if (t_out) *t_out = hi[0];
	return LIFE_OK;
}

}  // extern "C"
