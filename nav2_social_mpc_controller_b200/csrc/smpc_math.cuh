// smpc_math.cuh — elementary functions of the social-force loops (solve kernel pair loop, SFM people projection):
// constant-memory coefficient tables and branch-free exp / rsqrt / atan2 for the argument ranges those loops produce.
#pragma once
#include <cuda_runtime.h>

namespace smpc {

// ---------------------------------------------------------------------------------------------------
// Elementary functions of the pair loop. A 64-bit literal costs two move instructions (UMOV / IMAD.MOV.U32) every time
// it is used — DFMA takes only 32-bit immediates — and the libm expansions carry a slow-path branch each; in the
// 20-agent kernel those moves were 22 % of all executed instructions (profiles/r02_a20_before_consts_opmix.txt). The
// coefficient tables below live in constant memory instead (LDCU.128: one uniform load per TWO doubles, hoistable),
// and the three functions are branch-free for the argument range the social force can produce.
// ---------------------------------------------------------------------------------------------------
static __constant__ double kAtanTab[16] = {
    0.41421356237309503,  // tan(pi/8)
    -0.01917688711906226, 0.03923165829558719, -0.0508544973794026,  0.0585814891280221,
    -0.06664511447381948, 0.07692183190826087, -0.09090904578123903, 0.11111111015256361,
    -0.14285714284666542, 0.1999999999999552,  -0.3333333333333333,
    0.7853981633974483,   1.5707963267948966,  3.141592653589793,    0.0};
static __constant__ double kExpTab[14] = {
    1.4426950408889634,      // log2(e)
    -0.6931471805599453,     // -ln2 (high part)
    -2.3190468138462996e-17, // -ln2 (low part)
    // exp(r) = 1 + r + r^2 h(r), |r| <= ln2/2: h = degree-9 Chebyshev interpolant (tools/fit_exp.py), highest first
    2.5100395159429243e-08, 2.7620101012098e-07,    2.7557268439678e-06,  2.4801521269532122e-05,
    0.00019841269863066696, 0.0013888888917230724,  0.00833333333333006,  0.041666666666624094,
    0.16666666666666669,    0.5000000000000001,     -708.0};

// exp(x) for x <= 0 (the social force only calls it with -(d/B) - (n B theta)^2): Cody-Waite reduction x = n ln2 + r,
// degree-11 polynomial, scaling by 2^n as a multiplication (NaN propagates). 0.62 ulp against 60-digit mpmath over
// [-708, 0] (tools/fit_exp.py; CUDA's exp: 1 ulp). Arguments below -708 (results below 2^-1022) are clamped: the
// function returns exp(-708) = 3.3e-308 there instead of a denormal or 0 — 1e-300 times smaller than anything the force
// sums can resolve.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double xc = (x < kExpTab[13]) ? kExpTab[13] : x;  // NaN stays NaN
  const double t = fma(xc, kExpTab[0], 6755399441055744.0);  // 1.5 * 2^52: the low word of t is rint(x log2 e)
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, kExpTab[1], xc);
  r = fma(nf, kExpTab[2], r);
  double q = kExpTab[3];  // Horner: the two exp chains of a pair interleave with each other and with the gradient
#pragma unroll            // terms that do not depend on them (Estrin's scheme measured 3 % slower: +3 instructions each)
  for (int c = 4; c <= 12; ++c) q = fma(q, r, kExpTab[c]);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const double scale = __hiloint2double((__double2loint(t) + 1023) << 20, 0);  // 2^n, n in [-1021, 0]: normal
  return q * scale;
}

// 1 / sqrt(x) for normal positive x: hardware seed (2^-22) + one third-order step y (1 + e/2 + 3 e^2 / 8),
// e = 1 - x y^2 (< 1 ulp, the class of CUDA's rsqrt, without its exponent-range slow path). x = 0 gives NaN.
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(y * e, p, y);
}

// atan2(s, c) for a point (c, s) near the unit circle (the social force calls it with the cross / dot product of two
// unit vectors, i.e. s = sin(theta), c = cos(theta) up to rounding). CUDA's general atan2 is ~135 instructions with
// its scaling and special-case handling and was 19 % of all instructions of a 20-agent solve; this one is ~40:
// one division (reciprocal seed + two Newton steps + residual correction) of the octant-reduced argument
//   t = min / max                  (min <= tan(pi/8) max)        atan = P(t)
//   t = (min - max) / (min + max)  (otherwise: |t| <= tan(pi/8)) atan = pi/4 + P(t)
// and the odd degree-23 polynomial P(t) = t + t z Q(z), z = t^2 (Chebyshev fit on [0, tan^2(pi/8)], 2.4e-16 relative).
// Measured against 60-digit mpmath over 2e5 angles incl. |theta| down to 1e-9: 3.3e-16 relative (1.5 ulp) — the
// accuracy class of libm's atan2 (CUDA: 2 ulp). Inputs must be finite, not both zero and of comparable magnitude.
__device__ __forceinline__ double atan2_unit(double s, double c) {
  const double ax = fabs(c), ay = fabs(s);
  const bool steep = ay > ax;
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  const bool hi = mn > kAtanTab[0] * mx;
  const double num = hi ? mn - mx : mn, den = hi ? mn + mx : mx;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(fma(-den, r, 1.0), r, r);  // 2^-23 -> 2^-46; the quotient's own correction below squares that again
  double t = num * r;
  t = fma(fma(-den, t, num), r, t);
  const double z = t * t;
  // Q(z), degree 10, by Estrin's scheme: dependent depth 5 instead of 10 (this polynomial sits on the pair's critical
  // path: nothing downstream can start before theta is known). kAtanTab[11 - k] is the coefficient of z^k.
  const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  const double a0 = fma(kAtanTab[10], z, kAtanTab[11]), a1 = fma(kAtanTab[8], z, kAtanTab[9]);
  const double a2 = fma(kAtanTab[6], z, kAtanTab[7]), a3 = fma(kAtanTab[4], z, kAtanTab[5]);
  const double a4 = fma(kAtanTab[2], z, kAtanTab[3]);
  const double b0 = fma(a1, z2, a0), b1 = fma(a3, z2, a2), b2 = fma(kAtanTab[1], z2, a4);
  const double q = fma(b2, z8, fma(b1, z4, b0));
  double a = fma(t, z * q, t);
  if (hi) a += kAtanTab[12];
  if (steep) a = kAtanTab[13] - a;
  if (c < 0.0) a = kAtanTab[14] - a;
  return copysign(a, s);
}

}  // namespace smpc
