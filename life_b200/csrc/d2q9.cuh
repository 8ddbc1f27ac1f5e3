// D2Q9 lattice constants and the two collision operators, as device functions over register-resident populations.
//
// Numbering, weights and opposites follow the reference (src/Grid.cpp:1247-1250):
//   c = (0,0) (1,0) (-1,0) (0,1) (0,-1) (1,1) (-1,-1) (1,-1) (-1,1),  w = 4/9, 1/9 x4, 1/36 x4,  opp = 0 2 1 4 3 6 5 8 7.
// c_s is 1/sqrt(3) evaluated in double, so c_s^2 and c_s^4 are NOT exactly 1/3 and 1/9; the constants below are the
// doubles the reference's SQ(c_s) / QU(c_s) produce (src/Grid.cpp:1246, inc/defs.h:55-57).
//
// The collision arithmetic is algebraically factored (shared squares, +/- pairs, one reciprocal of rho) instead of
// evaluating nine independent polynomials as the reference does (src/Grid.cpp:143-223, :262, :278): ~150 fp64
// operations per node instead of ~450-700, which keeps the sweep on the HBM roof rather than the fp64 roof
// (SURVEY.md §7 "hard parts").  Results differ from the reference's evaluation order only in rounding.
#pragma once
#include <cstdint>

namespace life {

constexpr int NV = 9;
__device__ __constant__ const int kCx[NV] = {0, 1, -1, 0, 0, 1, -1, 1, -1};
__device__ __constant__ const int kCy[NV] = {0, 0, 0, 1, -1, 1, -1, -1, 1};
#define LIFE_CX(v) ((v) == 1 || (v) == 5 || (v) == 7 ? 1 : ((v) == 2 || (v) == 6 || (v) == 8 ? -1 : 0))
#define LIFE_CY(v) ((v) == 3 || (v) == 5 || (v) == 8 ? 1 : ((v) == 4 || (v) == 6 || (v) == 7 ? -1 : 0))
#define LIFE_OPP(v) ((v) == 0 ? 0 : (((v) & 1) ? (v) + 1 : (v) - 1))

constexpr double W0 = 4.0 / 9.0, W1 = 1.0 / 9.0, W2 = 1.0 / 36.0;
#define LIFE_W(v) ((v) == 0 ? life::W0 : ((v) < 5 ? life::W1 : life::W2))
constexpr double CS2 = 0x1.5555555555557p-2;   // SQ(1/sqrt(3))   = 0.3333333333333334
constexpr double CS4 = 0x1.c71c71c71c721p-4;   // QU(1/sqrt(3))   = 0.11111111111111117

enum { COLL_BGK = 0, COLL_CM = 1 };

// rho = sum f ; (mx,my) = sum c f, accumulated in ascending v like GridClass::macroscopic (src/Grid.cpp:290-294)
__device__ __forceinline__ void moments(const double (&f)[NV], double &rho, double &mx, double &my) {
	rho = f[0];
	rho += f[1]; rho += f[2]; rho += f[3]; rho += f[4]; rho += f[5]; rho += f[6]; rho += f[7]; rho += f[8];
	mx = f[1];
	mx -= f[2]; mx += f[5]; mx -= f[6]; mx += f[7]; mx -= f[8];
	my = f[3];
	my -= f[4]; my += f[5]; my -= f[6]; my -= f[7]; my += f[8];
}

// Equilibrium of one direction: BGK form (src/Grid.cpp:262) or the 4th-order product form used with central moments
// (src/Grid.cpp:259).  Only the boundary kernel calls this (a few thousand nodes), so it is written for clarity.
template <int COLL>
__device__ __forceinline__ double equilibrium(double rho, double ux, double uy, int v) {
	const double cx = (double)kCx[v], cy = (double)kCy[v];
	const double w = v == 0 ? W0 : (v < 5 ? W1 : W2);
	if (COLL == COLL_CM) {
		// association of the reference's 9.0 * SQ(cx) * SQ(ux) + 6.0 * cx * ux - 3.0 * SQ(ux) + 2.0
		return 0.25 * rho * w * (9.0 * (cx * cx) * (ux * ux) + 6.0 * cx * ux - 3.0 * (ux * ux) + 2.0) *
		       (9.0 * (cy * cy) * (uy * uy) + 6.0 * cy * uy - 3.0 * (uy * uy) + 2.0);
	} else {
		return rho * w * (1.0 + 3.0 * (cx * ux + cy * uy) +
		                  4.5 * (ux * ux * (cx * cx - 1.0 / 3.0) + uy * uy * (cy * cy - 1.0 / 3.0)) +
		                  9.0 * cx * cy * ux * uy);
	}
}

// BGK with Guo forcing:  f* = f + omega (feq - f) + (1 - omega/2) F_v            (src/Grid.cpp:237-244, :262, :267-279)
//   feq_v = rho w_v (a + 4.5 (c.u)^2 + 3 c.u),  a = 1 - 1.5 u.u
//   F_v   = 3 w_v ((F.c)(1 + 3 c.u) - F.u)
template <bool HASF>
__device__ __forceinline__ void collide_bgk(const double (&f)[NV], double rho, double ux, double uy, double Fx, double Fy,
                                            double omega, double (&o)[NV]) {
	const double omc = 1.0 - omega;
	const double ux2 = ux * ux, uy2 = uy * uy;
	const double a = 1.0 - 1.5 * (ux2 + uy2);
	const double r0 = W0 * rho * omega, r1 = W1 * rho * omega, r2 = W2 * rho * omega;
	const double ax = a + 4.5 * ux2, ay = a + 4.5 * uy2;
	const double s = ux + uy, d = ux - uy;
	const double as = a + 4.5 * s * s, ad = a + 4.5 * d * d;
	const double tx = 3.0 * ux, ty = 3.0 * uy, ts = 3.0 * s, td = 3.0 * d;
	o[0] = omc * f[0] + r0 * a;
	o[1] = omc * f[1] + r1 * (ax + tx);
	o[2] = omc * f[2] + r1 * (ax - tx);
	o[3] = omc * f[3] + r1 * (ay + ty);
	o[4] = omc * f[4] + r1 * (ay - ty);
	o[5] = omc * f[5] + r2 * (as + ts);
	o[6] = omc * f[6] + r2 * (as - ts);
	o[7] = omc * f[7] + r2 * (ad + td);
	o[8] = omc * f[8] + r2 * (ad - td);
	if (HASF) {
		const double g = 3.0 * (1.0 - 0.5 * omega);
		const double g0 = g * W0, g1 = g * W1, g2 = g * W2;
		const double A = Fx * ux + Fy * uy;
		const double Fs = Fx + Fy, Fd = Fx - Fy;
		o[0] -= g0 * A;
		o[1] += g1 * (Fx * (1.0 + tx) - A);
		o[2] -= g1 * (Fx * (1.0 - tx) + A);
		o[3] += g1 * (Fy * (1.0 + ty) - A);
		o[4] -= g1 * (Fy * (1.0 - ty) + A);
		o[5] += g2 * (Fs * (1.0 + ts) - A);
		o[6] -= g2 * (Fs * (1.0 - ts) + A);
		o[7] += g2 * (Fd * (1.0 + td) - A);
		o[8] -= g2 * (Fd * (1.0 - td) + A);
	}
}

// BGK in the REFERENCE'S OPERATION ORDER (cfg.exact): every direction on its own, nothing shared, nothing factored —
//   f*_v = f_v + omega (feq_v - f_v) + (1 - omega/2) F_v                                              src/Grid.cpp:243
//   feq_v = rho w_v (1 + 3 (cx ux + cy uy) + 4.5 (ux^2 (cx^2 - 1/3) + uy^2 (cy^2 - 1/3)) + 9 cx cy ux uy)   src/Grid.cpp:262
//   F_v   = 3 w_v (Fx (cx - ux + cx 3 (cx ux + cy uy)) + (Fy (cy - uy + cy 3 (cx ux + cy uy))))           src/Grid.cpp:278
// with C++'s left-to-right association of each sum and product.  Only meaningful in a translation unit compiled with
// -fmad=false (life_b200/build.py builds lbm_bulk.cu / lbm_boundary.cu a second time that way, -DLIFE_EXACT): every
// operation below is then one IEEE double operation, as in the reference's g++ build for baseline x86-64 (no FMA), and the
// result is the reference's double bit for bit.  The lattice velocities are compile-time constants after unrolling; the
// terms that vanish with them are +-0 and drop out of every sum without changing it.
template <bool HASF>
__device__ __forceinline__ void collide_bgk_ref(const double (&f)[NV], double rho, double ux, double uy, double Fx, double Fy,
                                                double omega, double (&o)[NV]) {
#pragma unroll
	for (int v = 0; v < NV; v++) {
		const int cx = LIFE_CX(v), cy = LIFE_CY(v);
		const double w = LIFE_W(v);
		const double feq = rho * w * (1.0 + 3.0 * (cx * ux + cy * uy) +
		                              4.5 * ((ux * ux) * ((cx * cx) - 1.0 / 3.0) + (uy * uy) * ((cy * cy) - 1.0 / 3.0)) +
		                              9.0 * cx * cy * ux * uy);
		double r = f[v] + omega * (feq - f[v]);
		if (HASF) {
			const double F = 3.0 * w * (Fx * (cx - ux + cx * 3.0 * (cx * ux + cy * uy)) + (Fy * (cy - uy + cy * 3.0 * (cx * ux + cy * uy))));
			r = r + (1.0 - 0.5 * omega) * F;
		}
		o[v] = r;
	}
}

// Central-moments collision (src/Grid.cpp:106-233).
//   pre-collision   k4Pre = sum f ((cx-ux)^2 - (cy-uy)^2),  k5Pre = sum f (cx-ux)(cy-uy)            (:113-122)
//   post-collision  k0 = rho, k1 = Fx/2, k2 = Fy/2, k3 = 2 rho cs^2, k4 = (1-w) k4Pre, k5 = (1-w) k5Pre,
//                   k6 = Fy cs^2/2, k7 = Fx cs^2/2, k8 = rho cs^4                                   (:125-133)
//   back-transform  central -> raw moments by the binomial shift with u, raw moments -> populations by the D2Q9
//                   inverse moment matrix; the reference's nine polynomials (:143-223) are this product expanded.
// `sum`, `mx`, `my` are the actual zeroth/first raw moments of f (they enter the pre-collision moments), `rho`, `ux`,
// `uy` are the start-of-step macroscopics (rho_n, u_n) — equal to the moments except on the first step after an upload.
template <bool HASF>
__device__ __forceinline__ void collide_cm(const double (&f)[NV], double sum, double mx, double my, double rho, double ux,
                                           double uy, double Fx, double Fy, double omega, double (&o)[NV]) {
	const double ux2 = ux * ux, uy2 = uy * uy, uxy = ux * uy;
	const double diag = (f[5] + f[6]) + (f[7] + f[8]);
	const double m20p = (f[1] + f[2]) + diag;
	const double m02p = (f[3] + f[4]) + diag;
	const double m11p = (f[5] + f[6]) - (f[7] + f[8]);
	const double k4Pre = (m20p - m02p) - 2.0 * (ux * mx - uy * my) + (ux2 - uy2) * sum;
	const double k5Pre = m11p - ux * my - uy * mx + uxy * sum;
	const double omc = 1.0 - omega;
	const double k3 = 2.0 * rho * CS2;
	const double k4 = omc * k4Pre, k5 = omc * k5Pre;
	const double k8 = rho * CS4;
	const double k20 = 0.5 * (k3 + k4), k02 = 0.5 * (k3 - k4);
	double m10 = ux * rho, m01 = uy * rho;
	double m20 = k20 + ux2 * rho, m02 = k02 + uy2 * rho;
	double m11 = k5 + uxy * rho;
	double m21 = 2.0 * ux * k5 + uy * k20 + ux2 * m01;
	double m12 = 2.0 * uy * k5 + ux * k02 + uy2 * m10;
	double m22 = k8 + ux2 * k02 + uy2 * k20 + 4.0 * uxy * k5 + ux2 * uy2 * rho;
	if (HASF) {
		const double k1 = 0.5 * Fx, k2 = 0.5 * Fy;
		const double k6 = k2 * CS2, k7 = k1 * CS2;
		m10 += k1;
		m01 += k2;
		m20 += 2.0 * ux * k1;
		m02 += 2.0 * uy * k2;
		m11 += ux * k2 + uy * k1;
		m21 += k6 + ux2 * k2 + 2.0 * uxy * k1;
		m12 += k7 + uy2 * k1 + 2.0 * uxy * k2;
		m22 += 2.0 * ux * k7 + 2.0 * uy * k6 + 2.0 * ux * uy2 * k1 + 2.0 * ux2 * uy * k2;
	}
	o[0] = rho - m20 - m02 + m22;
	const double hx = 0.5 * (m20 - m22), gx = 0.5 * (m10 - m12);
	o[1] = hx + gx;
	o[2] = hx - gx;
	const double hy = 0.5 * (m02 - m22), gy = 0.5 * (m01 - m21);
	o[3] = hy + gy;
	o[4] = hy - gy;
	const double p = 0.25 * (m22 + m11), q = 0.25 * (m12 + m21);
	const double r = 0.25 * (m22 - m11), t = 0.25 * (m12 - m21);
	o[5] = p + q;
	o[6] = p - q;
	o[7] = r + t;
	o[8] = r - t;
}


// Central moments in the REFERENCE'S OPERATION ORDER (cfg.exact; same compilation rules as collide_bgk_ref): the pre-collision
// moments as the v-ascending sums of src/Grid.cpp:113-122, the post-collision moments of :125-133, and the nine back-transform
// polynomials of :143-223 with every sum and product associated as the reference's expression parses (left to right, its
// parentheses kept).  x = ux, y = uy, xx = ux*ux, yy = uy*uy; k0..k8 as in the reference.
__device__ __forceinline__ void collide_cm_ref(const double (&f)[NV], double rho, double x, double y, double Fx, double Fy, double omega,
                                               double (&o)[NV]) {
	double k4Pre = 0.0, k5Pre = 0.0;
#pragma unroll
	for (int v = 0; v < NV; v++) {
		const double cx = LIFE_CX(v) - x, cy = LIFE_CY(v) - y;
		k4Pre += f[v] * ((cx * cx) - (cy * cy));
		k5Pre += f[v] * cx * cy;
	}
	const double k0 = rho;
	const double k1 = 0.5 * Fx;
	const double k2 = 0.5 * Fy;
	const double k3 = 2.0 * rho * CS2;
	const double k4 = (1.0 - omega) * k4Pre;
	const double k5 = (1.0 - omega) * k5Pre;
	const double k6 = 0.5 * Fy * CS2;
	const double k7 = 0.5 * Fx * CS2;
	const double k8 = rho * CS4;
	const double xx = x * x, yy = y * y;
	o[0] = (xx * yy - xx - yy + 1.0) * k0
	     + (2.0 * x * yy - 2.0 * x) * k1
	     + (2.0 * y * xx - 2.0 * y) * k2
	     + (0.5 * xx + 0.5 * yy - 1.0) * k3
	     + (0.5 * yy - 0.5 * xx) * k4
	     + 4.0 * x * y * k5
	     + 2.0 * y * k6
	     + 2.0 * x * k7
	     + k8;
	o[1] = (0.5 * xx - 0.5 * (xx * yy) - 0.5 * (x * yy) + 0.5 * x) * k0
	     + (x - x * yy - 0.5 * yy + 0.5) * k1
	     + (-y * xx - y * x) * k2
	     + (-0.25 * xx - 0.25 * x - 0.25 * yy + 0.25) * k3
	     + (0.25 * xx + 0.25 * x - 0.25 * yy + 0.25) * k4
	     + (-y - 2.0 * x * y) * k5
	     + (-y) * k6
	     + (-x - 0.5) * k7
	     - 0.5 * k8;
	o[2] = (-0.5 * (xx * yy) + 0.5 * xx + 0.5 * (x * yy) - 0.5 * x) * k0
	     + (x - x * yy + 0.5 * yy - 0.5) * k1
	     + (-y * xx + y * x) * k2
	     + (-0.25 * xx + 0.25 * x - 0.25 * yy + 0.25) * k3
	     + (0.25 * xx - 0.25 * x - 0.25 * yy + 0.25) * k4
	     + (y - 2.0 * x * y) * k5
	     + (-y) * k6
	     + (0.5 - x) * k7
	     - 0.5 * k8;
	o[3] = (0.5 * yy - 0.5 * (xx * y) - 0.5 * (xx * yy) + 0.5 * y) * k0
	     + (-x * yy - x * y) * k1
	     + (y - xx * y - 0.5 * xx + 0.5) * k2
	     + (-0.25 * xx - 0.25 * yy - 0.25 * y + 0.25) * k3
	     + (0.25 * xx - 0.25 * yy - 0.25 * y - 0.25) * k4
	     + (-x - 2.0 * x * y) * k5
	     + (-y - 0.5) * k6
	     + (-x) * k7
	     - 0.5 * k8;
	o[4] = (-0.5 * (xx * yy) + 0.5 * (xx * y) + 0.5 * yy - 0.5 * y) * k0
	     + (-x * yy + x * y) * k1
	     + (y - xx * y + 0.5 * xx - 0.5) * k2
	     + (-0.25 * xx - 0.25 * yy + 0.25 * y + 0.25) * k3
	     + (0.25 * xx - 0.25 * yy + 0.25 * y - 0.25) * k4
	     + (x - 2.0 * x * y) * k5
	     + (0.5 - y) * k6
	     + (-x) * k7
	     - 0.5 * k8;
	o[5] = (0.25 * (xx * yy) + 0.25 * (xx * y) + 0.25 * (x * yy) + 0.25 * (x * y)) * k0
	     + (0.25 * y + 0.5 * (x * y) + 0.5 * (x * yy) + 0.25 * yy) * k1
	     + (0.25 * x + 0.5 * (x * y) + 0.5 * (xx * y) + 0.25 * xx) * k2
	     + (0.125 * xx + 0.125 * x + 0.125 * yy + 0.125 * y) * k3
	     + (-0.125 * xx - 0.125 * x + 0.125 * yy + 0.125 * y) * k4
	     + (0.5 * x + 0.5 * y + x * y + 0.25) * k5
	     + (0.5 * y + 0.25) * k6
	     + (0.5 * x + 0.25) * k7
	     + 0.25 * k8;
	o[6] = (0.25 * (xx * yy) - 0.25 * (xx * y) - 0.25 * (x * yy) + 0.25 * (x * y)) * k0
	     + (0.25 * y - 0.5 * (x * y) + 0.5 * (x * yy) - 0.25 * yy) * k1
	     + (0.25 * x - 0.5 * (x * y) + 0.5 * (xx * y) - 0.25 * xx) * k2
	     + (0.125 * xx - 0.125 * x + 0.125 * yy - 0.125 * y) * k3
	     + (-0.125 * xx + 0.125 * x + 0.125 * yy - 0.125 * y) * k4
	     + (x * y - 0.5 * y - 0.5 * x + 0.25) * k5
	     + (0.5 * y - 0.25) * k6
	     + (0.5 * x - 0.25) * k7
	     + 0.25 * k8;
	o[7] = (0.25 * (xx * yy) - 0.25 * (xx * y) + 0.25 * (x * yy) - 0.25 * (x * y)) * k0
	     + (0.5 * (x * yy) - 0.5 * (x * y) - 0.25 * y + 0.25 * yy) * k1
	     + (0.5 * (x * y) - 0.25 * x + 0.5 * (xx * y) - 0.25 * xx) * k2
	     + (0.125 * xx + 0.125 * x + 0.125 * yy - 0.125 * y) * k3
	     + (-0.125 * xx - 0.125 * x + 0.125 * yy - 0.125 * y) * k4
	     + (0.5 * y - 0.5 * x + x * y - 0.25) * k5
	     + (0.5 * y - 0.25) * k6
	     + (0.5 * x + 0.25) * k7
	     + 0.25 * k8;
	o[8] = (0.25 * (xx * yy) + 0.25 * (xx * y) - 0.25 * (x * yy) - 0.25 * (x * y)) * k0
	     + (0.5 * (x * y) - 0.25 * y + 0.5 * (x * yy) - 0.25 * yy) * k1
	     + (0.5 * (xx * y) - 0.5 * (x * y) - 0.25 * x + 0.25 * xx) * k2
	     + (0.125 * xx - 0.125 * x + 0.125 * yy + 0.125 * y) * k3
	     + (-0.125 * xx + 0.125 * x + 0.125 * yy + 0.125 * y) * k4
	     + (0.5 * x - 0.5 * y + x * y - 0.25) * k5
	     + (0.5 * y + 0.25) * k6
	     + (0.5 * x - 0.25) * k7
	     + 0.25 * k8;
}

}  // namespace life
