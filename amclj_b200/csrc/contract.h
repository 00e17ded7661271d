// contract.h — the fp32 / integer arithmetic contract of libamclj_b200, shared by host
// and device code of the PRODUCT.  Everything here is written with explicit fmaf and is
// compiled without implicit contraction (nvcc --fmad=false), so every integer that leaves
// the library (cells, steps, fixed-point weights, resample indices) is reproducible bit
// for bit by an independent CPU restatement.
//
// Reference functions restated with this arithmetic (jvia/amclj, paths from its root):
//   map.clj:36-46 world->map, sensor.clj:99-103 project-pose, motion.clj:24-45.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define AMCLJ_HD __host__ __device__ __forceinline__
#else
#define AMCLJ_HD inline
#endif

namespace amclj {

// Deterministic sincos: 3-constant Cody-Waite reduction by pi/2 and the Cephes minimax
// polynomials; only IEEE mul/add/fma/rint, so host and device agree bit for bit.
AMCLJ_HD void sincos_det(float a, float &sn, float &cs) {
  float j = rintf(a * 0.636619772f);
  j = fminf(fmaxf(j, -1.0e9f), 1.0e9f);
  float r = fmaf(j, -1.5707962513e+00f, a);
  r = fmaf(j, -7.5497894159e-08f, r);
  r = fmaf(j, -5.3903029534e-15f, r);
  int q = ((int)j) & 3;
  float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  float s = fmaf(ps * z, r, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  float so = (q & 1) ? c : s, co = (q & 1) ? s : c;
  if (q & 2) so = -so;
  if ((q + 1) & 2) co = -co;
  sn = so;
  cs = co;
}

// Heading kept in (-pi, pi] (as quat/to-heading returns it, quaternions.clj:82-97): an unbounded fp32
// angle would lose resolution with uptime (ulp 1e-3 rad at |theta| ~ 1e4) and drag the ray end cells
// with it.  Values already inside the interval pass through untouched; IEEE ops only.
AMCLJ_HD float wrap_heading(float t) {
  if (fabsf(t) <= 3.14159274f) return t;
  float k = rintf(t * 0.159154937f);
  k = fminf(fmaxf(k, -1.0e9f), 1.0e9f);
  float r = fmaf(k, -6.28318548f, t);
  return fmaf(k, 1.74845553e-07f, r); /* 2 pi = 6.28318548 - 1.74845553e-07 */
}

// Deterministic exp for x <= 0 (used to turn log-weights into linear ones before the
// fixed-point conversion): Cody-Waite by ln2, Cephes degree-5 polynomial, exact 2^n scale.
// Anything below -87 (and NaN) flushes to 0.
AMCLJ_HD float exp_det(float x) {
  if (!(x > -87.0f)) return 0.0f;
  if (x > 0.0f) x = 0.0f;
  float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = fmaf(r, 1.9875691500e-4f, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float y = fmaf(p, r * r, r) + 1.0f;
  int e = (int)n + 127; /* n >= -126 here */
  union { uint32_t u; float f; } sc;
  sc.u = (uint32_t)e << 23;
  return y * sc.f;
}

// map.clj:36-46: Math/round(v / resolution) = floor(v/res + 0.5), IEEE divide.
AMCLJ_HD int32_t round_cell(float v, float res) {
  float t = floorf(v / res + 0.5f);
  t = fminf(fmaxf(t, -1.0e9f), 1.0e9f); /* NaN -> -1e9 */
  return (int32_t)t;
}

// Deterministic double-precision sincos for |a| up to a few pi (headings are kept in (-pi, pi]): Cody-Waite
// reduction by pi/2 with fdlibm's split of pi/2, fdlibm's kernel polynomials, explicit fma — host and device
// agree bit for bit, and with the C library to ~1e-16.  Used by the near-tie refinement of end cells only.
AMCLJ_HD void sincos_det_d(double a, double &sn, double &cs) {
  double j = rint(a * 6.36619772367581382433e-01);
  j = fmin(fmax(j, -1.0e6), 1.0e6);
  double r = fma(-j, 1.57079632673412561417e+00, a);
  r = fma(-j, 6.07710050650619224932e-11, r);
  const int q = ((int)j) & 3;
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(ps * z, r, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double c = fma(pc * z, z, fma(-0.5, z, 1.0));
  double so = (q & 1) ? c : s, co = (q & 1) ? s : c;
  if (q & 2) so = -so;
  if ((q + 1) & 2) co = -co;
  sn = so;
  cs = co;
}

// END cells (sensor.clj:134-135 + map.clj:36-46), relative to the start cell, with a near-tie refinement.
// The reference rounds (x + R cos(..)) / res in double.  Here the start cell x0 and the particle's offset
// inside it, f = x / res - x0 in [-0.5, 0.5), are taken ONCE per particle in double (start_cell_frac); per ray
// the end cell's offset from the start cell is round(f + rc * c) with rc = R / res (cells), one float FMA:
//     u = fma(rc, c, fK),   fK = float(f + K),   K = 1.5 * 2^(23-fb) + 0.5 + 2^(1-fb)
// which leaves round((f + rc c + 0.5) * 2^fb) + 2 in the low mantissa bits of u: offset = (bits >> fb) - coff,
// and "within two units of 2^-fb of a .5 boundary" <=> (bits & (2^fb - 4)) == 0.  Only there could the float
// result differ from the reference's double computation; those rays (about 1 in 2,000 at fb = 14) are
// recomputed in double from the same float32 inputs (end_cells_exact in sensor_common.cuh).  Error budget, in
// units of 2^-fb: rounding of fK 0.5, rounding of u 0.5, and (error of the direction c + rounding of rc) times
// the ray length, where |err c| <= 2.5e-7 (float angle addition of sincos_det, measured max error 9.3e-8, and
// of the float beam table, 3e-8; one product and one FMA rounding) and rc carries 6e-8 relative: 3.1e-7 per
// cell of ray.  A cell can only come out wrong unflagged when the sum reaches 2; fb is the largest that keeps
// the last term below 0.9 and the offset inside the mantissa window: lgfloor's 112-cell rays get fb = 14,
// C5's 600-cell rays fb = 12.  No division per ray, and no dependence on the size of the map: a ray's geometry
// is relative to its start cell.
struct CellRefine {
  float K;        /* 0: no refinement (rays of 4,000 cells or more, non-finite geometry): plain float rounding */
  int32_t fb;
  int32_t coff;   /* bits(1.5 * 2^(23-fb)) >> fb */
  uint32_t mask;  /* 2^fb - 4 */
  float rc;       /* ray length in cells, R / res (signed: REF casts rays of -1 m) */
};
inline CellRefine make_cell_refine(float R, float res, bool enabled) {
  CellRefine c = {0.0f, 0, 0, 0u, 0.0f};
  const double rcd = (double)R / (double)res, ar = fabs(rcd);
  if (!(ar < 4000.0) || !(res > 1.0e-6f && res < 1.0e6f)) return c; /* NaN fails too */
  int fb = 14;
  while (fb >= 8 && (ar + 3.0 >= ldexp(1.0, 22 - fb) || 3.1e-7 * (ar + 1.0) * ldexp(1.0, fb) > 0.9)) fb--;
  if (fb < 8) return c;
  const float big = (float)ldexp(1.5, 23 - fb);
  c.K = big + 0.5f + (float)ldexp(1.0, 1 - fb);
  c.fb = fb;
  uint32_t bits;
  memcpy(&bits, &big, 4);
  c.coff = (int32_t)(bits >> fb);
  /* disabled (AMCLJ_EXACT_CELLS=0, timing experiments): a mask bit that is always set, i.e. never "near" */
  c.mask = enabled ? (1u << fb) - 4u : 0x40000000u;
  c.rc = (float)rcd;
  return c;
}

// The START cell of a particle's rays (map.clj:36-46 on the pose itself): once per particle, so it is taken
// the way the JVM takes it from the same float32 inputs — a double divide — and no particle starts its 360
// rays from a cell that differs from the reference's because v / res was rounded to float first.
AMCLJ_HD int32_t round_cell_start(float v, float res) {
  double t = floor((double)v / (double)res + 0.5);
  t = fmin(fmax(t, -1.0e9), 1.0e9); /* NaN -> -1e9 */
  return (int32_t)t;
}
// ... together with fK = float((v / res - cell) + K) for the end cells of its rays (CellRefine above); the cell
// is round_cell_start's.  Meaningful for a start on the map only (a far-away or NaN start has no use for it:
// its rays end at step 0).
AMCLJ_HD int32_t start_cell_frac(float v, float res, float K, float &fK) {
  const double q = (double)v / (double)res;
  double t = floor(q + 0.5);
  t = fmin(fmax(t, -1.0e9), 1.0e9); /* NaN -> -1e9 */
  fK = (float)((q - t) + (double)K);
  return (int32_t)t;
}

// Fixed-point weight: q = floor(clamp(w / wmax, 0, 1) * 2^32)  (NaN -> 0).
AMCLJ_HD uint64_t fixed_weight(float w, float wmax) {
  if (!(wmax > 0.0f)) return 0;
  float r = w / wmax;
  r = fminf(fmaxf(r, 0.0f), 1.0f);
  return (uint64_t)(r * 4294967296.0f); /* exact power-of-two scale, truncation = floor */
}

// Philox4x32-10 (Salmon et al., SC'11).
AMCLJ_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                            uint32_t k1, uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr uint32_t TAG_MOTION = 0x4d4f5449u;
constexpr uint32_t TAG_ROULETTE = 0x524f554cu;
constexpr uint32_t TAG_TOURNAMENT = 0x544f5552u;
constexpr uint32_t TAG_INIT = 0x494e4954u;
constexpr uint32_t TAG_GAUSS = 0x47415553u;

AMCLJ_HD float u24(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Systematic-resampling arithmetic on the integer CDF (Probabilistic Robotics table 4.4):
// pointer m sits at U_m = floor((m*T + r) / N).  With T = a*N + b this is
// m*a + floor((m*b + r) / N); m*b + r < 2^63 for N <= 2^31, T < 2^62.
struct SysPlan {
  uint64_t total, a, b, r, n;
};
AMCLJ_HD uint64_t sys_threshold(const SysPlan &s, uint64_t m) {
  return m * s.a + (m * s.b + s.r) / s.n;
}

// roulette acceptance (pf.clj:165): wn[idx] > u  <=>  q * 2^32 > u32 * T
AMCLJ_HD bool roulette_accept(uint64_t q, uint32_t u, uint64_t total) {
#if defined(__CUDA_ARCH__)
  uint64_t lo = (uint64_t)u * total; /* may wrap: compute 96-bit product */
  uint64_t hi = __umul64hi((uint64_t)u, total);
  /* q*2^32 as (qhi, qlo) */
  uint64_t qhi = q >> 32, qlo = q << 32;
  return qhi > hi || (qhi == hi && qlo > lo);
#else
  return ((unsigned __int128)q << 32) > (unsigned __int128)u * total;
#endif
}

#if defined(__CUDACC__)
// Particle bounding box kept on the device as four ordered-uint maxima (x, -x, y, -y): is the
// cloud's box at most `compact_m` metres on a side?  (NaN coordinates poison the box: not compact.)
__device__ __forceinline__ bool cloud_is_compact(const uint32_t *bbox, float compact_m) {
  auto dec = [](uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); };
  float wx = dec(bbox[0]) + dec(bbox[1]), wy = dec(bbox[2]) + dec(bbox[3]); /* max - min */
  return wx <= compact_m && wy <= compact_m;
}
#endif

}  // namespace amclj
