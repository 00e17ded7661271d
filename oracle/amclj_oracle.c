/*
 * amclj_oracle.c — CPU restatement of the reference's particle-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under amclj_b200/ (the product) may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker and as the timed CPU stand-in.
 *
 * HOW IT IS PINNED.  jvia/amclj ships no golden vectors or known-answer tests for this path
 * (test/amclj/pf_test.clj pins only add-pose; one fact references an undefined symbol) and it
 * cannot run here (Clojure on the JVM; no JVM in this image), so there is no oracle/_ref.  The
 * pin is the reference's own SOURCE TEXT instead: tests/golden/clj_eval.py evaluates
 * src/amclj/{quaternions,map,motion,sensor,pf}.clj verbatim (the hot path uses a small subset of
 * Clojure), tests/golden/make_golden_from_reference.py runs the reference's functions on its own
 * dev fixtures (sensor.clj:26-74, core.clj:56-63, launch/lgfloor.*) and on seeded random inputs
 * with the noise streams injected, and tests/test_reference_text.py holds this file to those
 * vectors: 7,740 rays cell for cell, weights, motion, quaternions, resamplers to 1e-12.
 * Further witnesses: hand-derived known answers on tiny maps (tests/test_oracle_kat.py) and an
 * independent Python transliteration that also covers the north-star switches, which have no
 * reference text (tests/golden/make_golden.py).
 *
 * Citations are file:line in the reference repository (jvia/amclj).
 *
 * Two numeric modes:
 *   *_f64  — mirrors the JVM: doubles everywhere, float32 message fields widened.
 *   *_f32  — the fp32 contract the CUDA kernels implement: same algorithm, float
 *            arithmetic in a fixed operation order (explicit fmaf, IEEE divide, a
 *            specified sincos polynomial) so that every integer output (cells, steps,
 *            resample indices) is bit-reproducible.  Compile with -ffp-contract=off.
 *
 * Third-party arithmetic the reference takes from incanter 1.5.4 (project.clj:8; source
 * not under /root/reference): stats/sample-normal, stats/sample-uniform and Math/random
 * are replaced by injected streams (north_star); core/pow on scalars = Math/pow;
 * core/sel m y x = element (row y, col x).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/amclj_b200.h" /* POD amclj_params only */

#define ORACLE_API __attribute__((visibility("default")))

typedef struct oracle_ray {
  int32_t x0, y0;       /* start cell (un-swapped) */
  int32_t x1, y1;       /* projected end cell (un-swapped, before S3) */
  int32_t steep;
  int32_t hit;          /* 1: a blocked cell stopped the ray */
  int32_t hit_x, hit_y; /* last tested cell (un-swapped) */
  int32_t steps;        /* |x - x0| along the major axis at termination */
  int32_t reserved;
  int64_t cells;        /* cells tested (start cell included) */
  double range;         /* metres (or the miss value) */
} oracle_ray;

typedef struct oracle_map {
  const int8_t *cells; /* [H][W] row-major, data[y*W + x]  (nav_msgs.clj:34-38) */
  int32_t W, H;
  float res; /* ROS float32 resolution, widened wherever used */
} oracle_map;

/* ------------------------------------------------------------------------------------
 * map.clj:19-24 get-cell, :48-57 inhabitable?/uninhabitable?
 * blocked <=> cell != 0 (unknown -1 and occupied 100 both stop rays); out of range -> -1.
 * Documented deviation: the reference tests `(> x width)`, letting
 * x == width fall through into incanter's sel (exception / unpinned wrap).  Contract here:
 * x < 0 or x >= W or y < 0 or y >= H => blocked.
 * ---------------------------------------------------------------------------------- */
/* The Clojure `def` constants (sensor.clj:11-21, motion.clj:12-16) are double literals — 0.95, 0.002 —
 * while amclj_params carries them as float (the ABI).  The double restatement widens a parameter to the
 * double its shortest round-tripping decimal names (0.95f -> 0.95), i.e. back to the reference's own
 * literal, not to 0.949999988...; message fields (ranges, angles, resolution) are float32 on the wire and
 * stay float-widened. */
static double decimal_exact(float f) {
  if (!(f == f) || f == 0.0f || f - f != 0.0f) return (double)f;
  char buf[40];
  for (int prec = 1; prec <= 9; prec++) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
    if (strtof(buf, NULL) == f) return strtod(buf, NULL);
  }
  return (double)f;
}

static inline int blocked(const oracle_map *m, int64_t x, int64_t y) {
  if (x < 0 || x >= m->W || y < 0 || y >= m->H) return 1;
  return m->cells[(size_t)y * (size_t)m->W + (size_t)x] != 0;
}

ORACLE_API int amclj_oracle_blocked(const int8_t *cells, int32_t W, int32_t H, int64_t x,
                                    int64_t y) {
  oracle_map m = {cells, W, H, 1.0f};
  return blocked(&m, x, y);
}

/* ------------------------------------------------------------------------------------
 * quaternions.clj:82-97 to-heading, :7-16 mult, :19-36 rotate (pitch = bank = 0)
 * ---------------------------------------------------------------------------------- */
ORACLE_API double amclj_oracle_to_heading(double qx, double qy, double qz, double qw) {
  double test = qx * qy + qz * qw;
  if (test > 0.4999) return 2.0 * atan2(qx, qw);
  if (test < -0.499) return -2.0 * atan2(qx, qw);
  return atan2(2.0 * qz * qw - 2.0 * qx * qy, 1.0 - 2.0 * qz * qz - 2.0 * qy * qy);
}

/* q = (x, y, z, w) */
ORACLE_API void amclj_oracle_quat_mult(const double a[4], const double b[4], double out[4]) {
  double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  out[0] = ax * bw + aw * bx + ay * bz + (-(az * by));
  out[1] = aw * by + (-(ax * bz)) + ay * bw + az * bx;
  out[2] = aw * bz + ax * by + (-(ay * bx)) + az * bw;
  out[3] = aw * bw - ax * bx - ay * by - az * bz;
}

ORACLE_API void amclj_oracle_quat_rotate(const double q[4], double heading, double out[4]) {
  double c1 = cos(heading / 2), s1 = sin(heading / 2);
  double c2 = cos(0.0 / 2), s2 = sin(0.0 / 2);
  double c3 = cos(0.0 / 2), s3 = sin(0.0 / 2);
  double c1c2 = c1 * c2, s1s2 = s1 * s2;
  double r[4];
  r[3] = c1c2 * c3 - s1s2 * s3;
  r[0] = c1c2 * s3 + s1s2 * c3;
  r[1] = c1 * s2 * c3 - s1 * c2 * s3;
  r[2] = s1 * c2 * c3 + c1 * s2 * s3;
  amclj_oracle_quat_mult(r, q, out);
}

/* ------------------------------------------------------------------------------------
 * sensor.clj:144-145 beam subsampling: step = int((n-1)/(max-beams-1)), range(0, n, step).
 * max_beams <= 0 => every beam.  A zero step (n-1 < max_beams-1) would make Clojure's
 * `range` infinite; contract: step = max(step, 1).
 * ---------------------------------------------------------------------------------- */
ORACLE_API int32_t amclj_oracle_beam_step(int32_t n, int32_t max_beams) {
  if (max_beams <= 1) return 1;
  int32_t step = (n - 1) / (max_beams - 1);
  return step < 1 ? 1 : step;
}

ORACLE_API int32_t amclj_oracle_beams_used(int32_t n, int32_t max_beams) {
  int32_t step = amclj_oracle_beam_step(n, max_beams);
  return n <= 0 ? 0 : (n + step - 1) / step;
}

/* ------------------------------------------------------------------------------------
 * sensor.clj:112-128 extract-line, in the swapped frame of map-range.
 * s4 == 0: the reference's `(> x max-x)` test AFTER the cell test.
 * s4 == 1: stop after testing the endpoint cell.
 * ---------------------------------------------------------------------------------- */
static void extract_line(const oracle_map *m, int64_t x0, int64_t y0, int64_t mx, int64_t my,
                         int64_t xs, int64_t ys, int steep, int s4, double miss_value,
                         oracle_ray *out) {
  int64_t dx = llabs(x0 - mx), dy = llabs(y0 - my), derr = dy;
  int64_t x = x0, y = y0, err = 0, cells = 0;
  for (;;) {
    cells++;
    int b = steep ? blocked(m, y, x) : blocked(m, x, y);
    if (b) {
      out->hit = 1;
      out->range = hypot((double)(x - x0), (double)(y - y0)) * (double)m->res;
      break;
    }
    if (s4 ? (x == mx) : (x > mx)) {
      out->hit = 0;
      out->range = miss_value;
      break;
    }
    x += xs;
    err += derr;
    if (2 * err >= dx) {
      y += ys;
      err -= dx;
    }
  }
  out->hit_x = (int32_t)(steep ? y : x);
  out->hit_y = (int32_t)(steep ? x : y);
  out->steps = (int32_t)llabs(x - x0);
  out->cells = cells;
}

/* sensor.clj:130-139 map-range on integer cells (shared by both numeric modes). */
static void map_range_cells(const amclj_params *p, const oracle_map *m, int64_t cx0,
                            int64_t cy0, int64_t cx1, int64_t cy1, double miss_value,
                            oracle_ray *out) {
  out->x0 = (int32_t)cx0;
  out->y0 = (int32_t)cy0;
  out->x1 = (int32_t)cx1;
  out->y1 = (int32_t)cy1;
  /* sensor.clj:87-91 steep? [[x1 y1] [x0 y0]] */
  int steep = llabs(cy0 - cy1) > llabs(cx0 - cx1);
  out->steep = steep;
  int64_t x0 = steep ? cy0 : cx0, y0 = steep ? cx0 : cy0;
  int64_t x1, y1;
  if (steep) {
    x1 = cy1;
    y1 = cx1;
  } else if (p->s3_proper_endpoint) {
    x1 = cx1;
    y1 = cy1;
  } else { /* sensor.clj:138: `(if steep [y1 x1] [x0 y0])` => end = start */
    x1 = x0;
    y1 = y0;
  }
  /* sensor.clj:108-110 calc-step: ties -> -1 */
  int64_t xs = x0 < x1 ? 1 : -1, ys = y0 < y1 ? 1 : -1;
  extract_line(m, x0, y0, x1, y1, xs, ys, steep, p->s4_endpoint_termination, miss_value, out);
}

/* map.clj:36-46 world->map: Math/round(v / resolution) = floor(v/res + 0.5) */
static inline int64_t round_cell_f64(double v, double res) {
  double t = floor(v / res + 0.5);
  if (!(t > -1e9)) t = -1e9; /* also NaN */
  if (t > 1e9) t = 1e9;
  return (int64_t)t;
}

ORACLE_API void amclj_oracle_map_range_f64(const amclj_params *p, const int8_t *cells,
                                           int32_t W, int32_t H, float res, double ox,
                                           double oy, double theta, double obs_bearing,
                                           double scan_range_max, oracle_ray *out) {
  oracle_map m = {cells, W, H, res};
  double rres = (double)res;
  double dir = p->s1_use_heading ? theta + obs_bearing : obs_bearing; /* sensor.clj:131,133 */
  double R = p->s2_use_scan_range_max ? scan_range_max : decimal_exact(p->max_range);
  if (p->s10_apply_laser_offset) {
    double c = cos(theta), s = sin(theta);
    double lx = decimal_exact(p->laser_x), ly = decimal_exact(p->laser_y);
    double nx = ox + (lx * c - ly * s), ny = oy + (lx * s + ly * c);
    ox = nx;
    oy = ny;
  }
  int64_t x0 = round_cell_f64(ox, rres), y0 = round_cell_f64(oy, rres);
  /* sensor.clj:99-103 project-pose */
  double ex = ox + R * cos(dir), ey = oy + R * sin(dir);
  int64_t x1 = round_cell_f64(ex, rres), y1 = round_cell_f64(ey, rres);
  map_range_cells(p, &m, x0, y0, x1, y1, R, out);
}

/* ------------------------------------------------------------------------------------
 * fp32 contract pieces
 * ---------------------------------------------------------------------------------- */
/* Deterministic sincos: 3-constant Cody-Waite reduction by pi/2 (explicit fmaf) and the
 * Cephes single-precision minimax polynomials.  Bit-identical to the device routine. */
ORACLE_API void amclj_oracle_sincosf(float a, float *sn, float *cs) {
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
  *sn = so;
  *cs = co;
}

static inline int32_t round_cell_f32(float v, float res) {
  float t = floorf(v / res + 0.5f);
  t = fminf(fmaxf(t, -1.0e9f), 1.0e9f); /* also NaN -> -1e9 */
  return (int32_t)t;
}

/* Per-beam direction table: a_k = angleMin + k*angleIncrement in double (sensor.clj:93-97,
 * float32 message fields widened), cos/sin in double, rounded once to float. */
ORACLE_API void amclj_oracle_beam_table(int32_t n, float angle_min, float angle_inc,
                                        int32_t max_beams, float *cb, float *sb,
                                        int32_t *beam_index) {
  int32_t step = amclj_oracle_beam_step(n, max_beams), j = 0;
  for (int32_t i = 0; i < n; i += step, j++) {
    double a = (double)angle_min + (double)i * (double)angle_inc;
    cb[j] = (float)cos(a);
    sb[j] = (float)sin(a);
    if (beam_index) beam_index[j] = i;
  }
}

/* fp32 contract, END cells: float arithmetic, except next to a .5 boundary, where the float quotient could
 * round the other way than the double computation of the reference — there the cell is taken in double
 * (end point by the angle-addition identity on double sines and cosines, double divide).
 * "Next to" is decided on the bits of q + K, K = 1.5 * 2^(23-fb) + 0.5 + 2^(1-fb): the low fb mantissa bits
 * then hold round((q + 0.5) * 2^fb) + 2, and the quotient is near a boundary when that value mod 2^fb is
 * below 4.  fb = the most fraction bits that keep every quotient of the map inside the mantissa window. */
static void sincos_det_f64(double a, double *sn, double *cs) {
  double j = rint(a * 6.36619772367581382433e-01);
  j = fmin(fmax(j, -1.0e6), 1.0e6);
  double r = fma(-j, 1.57079632673412561417e+00, a);
  r = fma(-j, 6.07710050650619224932e-11, r);
  int q = ((int)j) & 3;
  double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  double s = fma(ps * z, r, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  double c = fma(pc * z, z, fma(-0.5, z, 1.0));
  double so = (q & 1) ? c : s, co = (q & 1) ? s : c;
  if (q & 2) so = -so;
  if ((q + 1) & 2) co = -co;
  *sn = so;
  *cs = co;
}

ORACLE_API void amclj_oracle_sincos_f64(double a, double *sn, double *cs) { sincos_det_f64(a, sn, cs); }

/* contract (csrc/contract.h, CellRefine): fraction bits of the relative end-cell arithmetic for a ray of R / res
 * cells; 0 = no refinement (rays of 4,000 cells or more, odd resolutions): plain float rounding */
static int refine_fraction_bits(float R, float res) {
  double ar = fabs((double)R / (double)res);
  if (!(ar < 4000.0) || !(res > 1.0e-6f && res < 1.0e6f)) return 0;
  int fb = 14;
  while (fb >= 8 && (ar + 3.0 >= ldexp(1.0, 22 - fb) || 3.1e-7 * (ar + 1.0) * ldexp(1.0, fb) > 0.9)) fb--;
  return fb < 8 ? 0 : fb;
}

/* offset of the end cell from the start cell along one axis: u = fma(rc, c, fK); *near = the double computation
 * must decide (u within two units of 2^-fb of a .5 boundary) */
static int32_t offset_from_fma(float rc, float c, float fK, int fb, int *near) {
  float big = (float)ldexp(1.5, 23 - fb);
  float u = fmaf(rc, c, fK);
  uint32_t bits, bbig;
  memcpy(&bits, &u, 4);
  memcpy(&bbig, &big, 4);
  if ((bits & ((1u << fb) - 4u)) == 0u) *near = 1;
  return (int32_t)((int32_t)bits >> fb) - (int32_t)(bbig >> fb);
}

/* start cell (the JVM's own double divide on the float32 input) and fK = float((v / res - cell) + K) */
static int32_t start_cell_frac(float v, float res, float K, float *fK) {
  double q = (double)v / (double)res;
  double t = floor(q + 0.5);
  t = fmin(fmax(t, -1.0e9), 1.0e9); /* NaN -> -1e9 */
  *fK = (float)((q - t) + (double)K);
  return (int32_t)t;
}

ORACLE_API void amclj_oracle_map_range_f32(const amclj_params *p, const int8_t *cells,
                                           int32_t W, int32_t H, float res, float ox, float oy,
                                           float ct, float st, float cb, float sb,
                                           float scan_range_max, float theta, double cbd, double sbd,
                                           oracle_ray *out) {
  oracle_map m = {cells, W, H, res};
  float c = cb, s = sb;
  if (p->s1_use_heading) {
    c = fmaf(ct, cb, -(st * sb));
    s = fmaf(st, cb, ct * sb);
  }
  float R = p->s2_use_scan_range_max ? scan_range_max : p->max_range;
  if (p->s10_apply_laser_offset) {
    float nx = ox + fmaf(p->laser_x, ct, -(p->laser_y * st));
    float ny = oy + fmaf(p->laser_x, st, p->laser_y * ct);
    ox = nx;
    oy = ny;
  }
  /* contract: the START cell is taken once per particle with the JVM's own double divide on the float32 inputs
   * (Math/round((double)x / (double)res), map.clj:36-46); the END cell relative to it by one float FMA per axis,
   * u = fma(R / res, c, (x / res - x0) + K), and by the reference's own double computation where u lies next to
   * a .5 boundary.  A start off the map, a non-finite heading or a geometry without refinement keeps the plain
   * float rounding of the absolute end point. */
  int fb = refine_fraction_bits(R, res);
  float K = fb ? (float)ldexp(1.5, 23 - fb) + 0.5f + (float)ldexp(1.0, 1 - fb) : 0.0f;
  float fKx, fKy;
  int32_t x0 = start_cell_frac(ox, res, K, &fKx), y0 = start_cell_frac(oy, res, K, &fKy);
  int32_t x1, y1;
  int fast = fb && x0 >= 0 && x0 < W && y0 >= 0 && y0 < H &&
             (!p->s1_use_heading || (fabsf(st) <= 1.001f && fabsf(ct) <= 1.001f)); /* a NaN heading fails */
  if (fast) {
    int near = 0;
    float rc = (float)((double)R / (double)res);
    x1 = x0 + offset_from_fma(rc, c, fKx, fb, &near);
    y1 = y0 + offset_from_fma(rc, s, fKy, fb, &near);
    if (near) { /* the reference's own computation from the same float32 inputs */
      double cd = cbd, sd = sbd;
      if (p->s1_use_heading) {
        double std_, ctd;
        sincos_det_f64((double)theta, &std_, &ctd);
        cd = fma(ctd, cbd, -(std_ * sbd));
        sd = fma(std_, cbd, ctd * sbd);
      }
      double exd = fma((double)R, cd, (double)ox), eyd = fma((double)R, sd, (double)oy);
      x1 = (int32_t)round_cell_f64(exd, (double)res);
      y1 = (int32_t)round_cell_f64(eyd, (double)res);
    }
  } else { /* plain float rounding of the end point */
    float ex = fmaf(R, c, ox), ey = fmaf(R, s, oy);
    x1 = round_cell_f32(ex, res);
    y1 = round_cell_f32(ey, res);
  }
  map_range_cells(p, &m, x0, y0, x1, y1, (double)R, out);
  if (out->hit) { /* contract: sqrtf of the exact integer d^2, times res, in float */
    int64_t ddx = (int64_t)out->hit_x - x0, ddy = (int64_t)out->hit_y - y0;
    out->range = (double)(sqrtf((float)(ddx * ddx + ddy * ddy)) * res);
  }
}

/* ------------------------------------------------------------------------------------
 * sensor.clj:141-173 likelihood-field-range-finder-model(*): beam mixture and weight
 * ---------------------------------------------------------------------------------- */
typedef struct oracle_sensor_stats {
  uint64_t rays, cells, hits;
} oracle_sensor_stats;

ORACLE_API void amclj_oracle_sensor_f64(const amclj_params *p, const int8_t *cells, int32_t W,
                                        int32_t H, float res, const float *ranges, int32_t n,
                                        float angle_min, float angle_inc, float range_max,
                                        const double *x, const double *y, const double *theta,
                                        int64_t np, double *raw, oracle_sensor_stats *stats,
                                        int32_t threads) {
  int32_t step = amclj_oracle_beam_step(n, p->s6_max_beams);
  double rmax = (double)range_max;
  double zhit = decimal_exact(p->z_hit), zshort = decimal_exact(p->z_short), zmax = decimal_exact(p->z_max),
         zrand = decimal_exact(p->z_rand);
  double lam = decimal_exact(p->lambda_short);
  double sig = p->s5_use_sigma_hit ? decimal_exact(p->sigma_hit) : zhit; /* sensor.clj:151-153 */
  uint64_t t_rays = 0, t_cells = 0, t_hits = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) \
    reduction(+ : t_rays, t_cells, t_hits)
#endif
  for (int64_t ip = 0; ip < np; ip++) {
    double acc = 0.0;
    for (int32_t i = 0; i < n; i += step) {
      double obs = (double)ranges[i];
      double bearing = (double)angle_min + (double)i * (double)angle_inc;
      oracle_ray ray;
      amclj_oracle_map_range_f64(p, cells, W, H, res, x[ip], y[ip], theta[ip], bearing, rmax,
                                 &ray);
      double z = obs - ray.range;
      double pz1 = zhit * exp((-(z * z)) / (2.0 * sig * sig));
      double pz2 = z < 0 ? zshort * lam * exp((-lam) * obs) : 0.0;
      double pz3 = obs == rmax ? 1.0 * zmax : 0.0;
      double pz4 = obs < rmax ? zrand * (1.0 / rmax) : 0.0;
      double s = pz1 + pz2 + pz3 + pz4;
      /* north-star log-weights only (no reference counterpart): a return beyond rangeMax, inf or
       * NaN has probability 0 under every term above; such a beam is left out (factor 1) */
      if (p->s7_log_weights && !(obs <= rmax && obs >= -3.0e38)) s = 1.0;
      acc += p->s7_log_weights ? log(s) : pow(s, 3.0);
      t_rays++;
      t_cells += (uint64_t)ray.cells;
      t_hits += (uint64_t)ray.hit;
    }
    raw[ip] = acc;
  }
  if (stats) {
    stats->rays = t_rays;
    stats->cells = t_cells;
    stats->hits = t_hits;
  }
}

/* Per-ray view of the double restatement (sensor.clj:130-139 as the JVM computes it from float32-valued poses):
 * hit flag, steps and range of every used beam of every particle, [np][beams_used].  For the tests that hold the
 * fp32 contract (and the CUDA path) to "casts the same rays as the double path" over millions of rays. */
ORACLE_API void amclj_oracle_rays_f64(const amclj_params *p, const int8_t *cells, int32_t W, int32_t H,
                                      float res, int32_t n, float angle_min, float angle_inc,
                                      float range_max, const double *x, const double *y,
                                      const double *theta, int64_t np, int32_t *steps, uint8_t *hit,
                                      double *range_m, int32_t threads) {
  int32_t step = amclj_oracle_beam_step(n, p->s6_max_beams);
  int32_t nb = amclj_oracle_beams_used(n, p->s6_max_beams);
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
#endif
  for (int64_t ip = 0; ip < np; ip++) {
    int32_t j = 0;
    for (int32_t i = 0; i < n; i += step, j++) {
      double bearing = (double)angle_min + (double)i * (double)angle_inc;
      oracle_ray ray;
      amclj_oracle_map_range_f64(p, cells, W, H, res, x[ip], y[ip], theta[ip], bearing, (double)range_max, &ray);
      int64_t o = ip * nb + j;
      if (steps) steps[o] = ray.steps;
      if (hit) hit[o] = (uint8_t)ray.hit;
      if (range_m) range_m[o] = ray.range;
    }
  }
}

/* fp32 contract.  Optional per-ray outputs, [np][beams_used]. */
ORACLE_API void amclj_oracle_sensor_f32(const amclj_params *p, const int8_t *cells, int32_t W,
                                        int32_t H, float res, const float *ranges, int32_t n,
                                        float angle_min, float angle_inc, float range_max,
                                        const float *x, const float *y, const float *theta,
                                        int64_t np, float *raw, int32_t *hit_x, int32_t *hit_y,
                                        int32_t *steps, uint8_t *hit, float *range_m,
                                        oracle_sensor_stats *stats, int32_t threads) {
  int32_t nb = amclj_oracle_beams_used(n, p->s6_max_beams);
  float *cb = (float *)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
  float *sb = (float *)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
  int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nb > 0 ? nb : 1));
  amclj_oracle_beam_table(n, angle_min, angle_inc, p->s6_max_beams, cb, sb, bi);
  float zhit = p->z_hit, zshort = p->z_short, zmax = p->z_max, zrand = p->z_rand;
  float lam = p->lambda_short;
  float sig = p->s5_use_sigma_hit ? p->sigma_hit : zhit;
  float inv2s2 = 1.0f / (2.0f * sig * sig);
  uint64_t t_rays = 0, t_cells = 0, t_hits = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) \
    reduction(+ : t_rays, t_cells, t_hits)
#endif
  for (int64_t ip = 0; ip < np; ip++) {
    float st, ct;
    amclj_oracle_sincosf(theta[ip], &st, &ct);
    double acc = 0.0; /* float terms, wide accumulator: the best fp32-term estimate */
    for (int32_t j = 0; j < nb; j++) {
      float obs = ranges[bi[j]];
      oracle_ray ray;
      double abeam = (double)angle_min + (double)bi[j] * (double)angle_inc; /* sensor.clj:93-97 */
      amclj_oracle_map_range_f32(p, cells, W, H, res, x[ip], y[ip], ct, st, cb[j], sb[j],
                                 range_max, theta[ip], cos(abeam), sin(abeam), &ray);
      float rng = (float)ray.range;
      float z = obs - rng;
      float pz1 = zhit * expf(-(z * z) * inv2s2);
      float pz2 = z < 0 ? zshort * lam * expf(-lam * obs) : 0.0f;
      float pz3 = obs == range_max ? zmax : 0.0f;
      float pz4 = obs < range_max ? zrand / range_max : 0.0f;
      float s = pz1 + pz2 + pz3 + pz4;
      if (p->s7_log_weights && !(obs <= range_max && obs >= -3.0e38f)) s = 1.0f; /* beam left out, as above */
      acc += p->s7_log_weights ? (double)logf(s) : (double)(s * s * s);
      size_t o = (size_t)ip * (size_t)nb + (size_t)j;
      if (hit_x) hit_x[o] = ray.hit_x;
      if (hit_y) hit_y[o] = ray.hit_y;
      if (steps) steps[o] = ray.steps;
      if (hit) hit[o] = (uint8_t)ray.hit;
      if (range_m) range_m[o] = rng;
      t_rays++;
      t_cells += (uint64_t)ray.cells;
      t_hits += (uint64_t)ray.hit;
    }
    if (raw) raw[ip] = (float)acc;
  }
  if (stats) {
    stats->rays = t_rays;
    stats->cells = t_cells;
    stats->hits = t_hits;
  }
  free(cb);
  free(sb);
  free(bi);
}

/* ------------------------------------------------------------------------------------
 * motion.clj:24-45 sample-motion-model-odometry*  (control unpacked at :68-79)
 * normals: [np][3] standard normals in the reference's draw order rot1, trans, rot2
 * (motion.clj:35-38; `sample` = sd * z, motion.clj:21-22).
 * ---------------------------------------------------------------------------------- */
typedef struct oracle_control {
  int32_t is_static; /* motion.clj:26-27: no motion => pose returned untouched */
  int32_t reserved;
  double rot1, trans, rot2;
  double sd_rot1, sd_trans, sd_rot2;
} oracle_control;

ORACLE_API void amclj_oracle_control(const amclj_params *p, const double curr[3],
                                     const double prev[3], oracle_control *c) {
  double xn = curr[0], yn = curr[1], tn = curr[2];
  double x = prev[0], y = prev[1], t = prev[2];
  memset(c, 0, sizeof(*c));
  c->is_static = ((x - xn) == 0.0 && (y - yn) == 0.0 && (t - tn) == 0.0);
  if (c->is_static) return;
  double rot1 = atan2(yn - y, xn - x) - t;
  double trans = hypot(x - xn, y - yn);
  double rot2 = tn - t - rot1;
  double a1 = decimal_exact(p->alpha1), a2 = decimal_exact(p->alpha2), a3 = decimal_exact(p->alpha3),
         a4 = decimal_exact(p->alpha4);
  c->rot1 = rot1;
  c->trans = trans;
  c->rot2 = rot2;
  double v1 = p->s9_pr_rot1_variance ? a1 * pow(rot1, 2) + a2 * pow(trans, 2)
                                     : a1 * pow(rot1, 2) + a2 * pow(rot2, 2);
  c->sd_rot1 = sqrt(v1);
  c->sd_trans = sqrt(a3 * pow(trans, 2) + a4 * pow(rot1, 2) + a4 * pow(rot2, 2));
  c->sd_rot2 = sqrt(a1 * pow(rot2, 2) + a2 * pow(trans, 2));
}

ORACLE_API void amclj_oracle_motion_f64(const amclj_params *p, const double curr[3],
                                        const double prev[3], const float *normals, double *x,
                                        double *y, double *theta, int64_t np) {
  oracle_control c;
  amclj_oracle_control(p, curr, prev, &c);
  if (c.is_static) return;
  for (int64_t i = 0; i < np; i++) {
    double n1 = normals[3 * i + 0], n2 = normals[3 * i + 1], n3 = normals[3 * i + 2];
    double ptheta = theta[i];
    double r1 = c.rot1 - c.sd_rot1 * n1;
    double tr = c.trans == 0.0 ? 0.0 : c.trans - c.sd_trans * n2;
    double r2 = c.rot2 - c.sd_rot2 * n3;
    x[i] = x[i] + tr * cos(ptheta + r1);
    y[i] = y[i] + tr * sin(ptheta + r1);
    theta[i] = ptheta + (r1 + r2); /* quat/rotate by (+ rot1* rot2*) */
  }
}

/* fp32 contract: the stored heading stays in (-pi, pi], as quat/to-heading (quaternions.clj:82-97)
 * returns it every update; values already inside pass through untouched */
static inline float wrap_heading_f32(float t) {
  if (fabsf(t) <= 3.14159274f) return t;
  float k = rintf(t * 0.159154937f);
  k = fminf(fmaxf(k, -1.0e9f), 1.0e9f);
  float r = fmaf(k, -6.28318548f, t);
  return fmaf(k, 1.74845553e-07f, r);
}

ORACLE_API void amclj_oracle_motion_f32(const amclj_params *p, const double curr[3],
                                        const double prev[3], const float *normals, float *x,
                                        float *y, float *theta, int64_t np) {
  oracle_control c;
  amclj_oracle_control(p, curr, prev, &c);
  if (c.is_static) return;
  float rot1 = (float)c.rot1, trans = (float)c.trans, rot2 = (float)c.rot2;
  float sd1 = (float)c.sd_rot1, sdt = (float)c.sd_trans, sd2 = (float)c.sd_rot2;
  for (int64_t i = 0; i < np; i++) {
    float n1 = normals[3 * i + 0], n2 = normals[3 * i + 1], n3 = normals[3 * i + 2];
    float r1 = fmaf(-n1, sd1, rot1);
    float tr = c.trans == 0.0 ? 0.0f : fmaf(-n2, sdt, trans);
    float r2 = fmaf(-n3, sd2, rot2);
    float sn, cs;
    amclj_oracle_sincosf(theta[i] + r1, &sn, &cs);
    x[i] = fmaf(tr, cs, x[i]);
    y[i] = fmaf(tr, sn, y[i]);
    theta[i] = wrap_heading_f32(theta[i] + (r1 + r2));
  }
}

/* ------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11): the injected-stream generator shared, as a
 * published algorithm, by the oracle and the kernels.
 * ---------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

ORACLE_API void amclj_oracle_philox(const uint32_t ctr[4], const uint32_t key[2],
                                    uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

#define TAG_MOTION 0x4d4f5449u
#define TAG_ROULETTE 0x524f554cu
#define TAG_TOURNAMENT 0x544f5552u
#define TAG_INIT 0x494e4954u

static inline float u24(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

/* The in-kernel motion draw: counter (gid lo, gid hi, 0, TAG_MOTION), key = seed.
 * Box-Muller on (r0,r1) -> n1,n2 and (r2,r3) -> n3. */
ORACLE_API void amclj_oracle_motion_normals(uint64_t seed, int64_t gid_first, int64_t np,
                                            float *normals) {
  for (int64_t i = 0; i < np; i++) {
    uint64_t gid = (uint64_t)(gid_first + i);
    uint32_t r[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0u, TAG_MOTION, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float ra = sqrtf(-2.0f * logf(u24(r[0]))), rb = sqrtf(-2.0f * logf(u24(r[2])));
    float a = 6.283185307179586f * u24(r[1]), b = 6.283185307179586f * u24(r[3]);
    normals[3 * i + 0] = ra * cosf(a);
    normals[3 * i + 1] = ra * sinf(a);
    normals[3 * i + 2] = rb * cosf(b);
  }
}

/* ------------------------------------------------------------------------------------
 * Normalisation contract (SURVEY 7 "exact, partition-independent resampling"):
 * q_i = floor(w_i / w_max * 2^32) elementwise in IEEE float, integer sums from there on.
 * Replaces the float `(reduce + (map :weight poses))` normaliser of pf.clj:159-160 so that
 * CPU, 1 GPU and G GPUs produce identical index lists.
 * ---------------------------------------------------------------------------------- */
ORACLE_API float amclj_oracle_wmax(const float *w, int64_t n) {
  float m = 0.0f;
  for (int64_t i = 0; i < n; i++) m = fmaxf(m, w[i]);
  return m;
}

static inline uint64_t fixed_weight(float w, float wmax) {
  float r = w / wmax;
  r = fminf(fmaxf(r, 0.0f), 1.0f); /* NaN -> 0 */
  return (uint64_t)floorf(ldexpf(r, 32));
}

ORACLE_API uint64_t amclj_oracle_fixed_weights(const float *w, int64_t n, float wmax,
                                               uint64_t *q) {
  uint64_t total = 0;
  for (int64_t i = 0; i < n; i++) {
    uint64_t v = wmax > 0.0f ? fixed_weight(w[i], wmax) : 0;
    if (q) q[i] = v;
    total += v;
  }
  return total;
}

ORACLE_API uint64_t amclj_oracle_systematic_offset(double u0, uint64_t total) {
  uint64_t bits = (uint64_t)ldexp(u0, 64); /* u0 in [0,1) */
  return (uint64_t)(((unsigned __int128)bits * total) >> 64);
}

/* Low-variance / systematic resampling (Probabilistic Robotics table 4.4) on the integer
 * CDF.  Pointer m sits at U_m = floor((m*T + r) / N); it selects the particle i with
 * C_{i-1} <= U_m < C_i.  This call emits global slots [m_lo, m_lo+count) against the local
 * CDF segment that starts at cdf_offset; idx_out is LOCAL source index (caller guarantees
 * the slots fall in this segment; -1 otherwise). */
ORACLE_API void amclj_oracle_systematic(const uint64_t *q, int64_t n_local, uint64_t total,
                                        uint64_t cdf_offset, uint64_t r, int64_t n_global,
                                        int64_t m_lo, int64_t count, int64_t *idx_out) {
  int64_t i = 0;
  uint64_t c = cdf_offset + (n_local > 0 ? q[0] : 0); /* inclusive CDF at i */
  for (int64_t k = 0; k < count; k++) {
    unsigned __int128 num = (unsigned __int128)(uint64_t)(m_lo + k) * total + r;
    uint64_t U = (uint64_t)(num / (uint64_t)n_global);
    while (i < n_local && c <= U) {
      i++;
      if (i < n_local) c += q[i];
    }
    idx_out[k] = (i < n_local && U >= cdf_offset) ? i : -1;
  }
}

/* #{m : U_m < C} = ceil((C*N - r) / T) clamped to [0, N] */
static int64_t slots_below(uint64_t C, uint64_t total, uint64_t r, int64_t n_global) {
  unsigned __int128 cn = (unsigned __int128)C * (uint64_t)n_global;
  if (cn <= r) return 0;
  unsigned __int128 v = (cn - r + total - 1) / total;
  return v > (unsigned __int128)n_global ? n_global : (int64_t)v;
}

/* Independent restatement of the planner the product exports (amclj_plan_systematic). */
ORACLE_API void amclj_oracle_plan(const uint64_t *totals, int32_t world, int64_t n_global,
                                  double u0, uint64_t *r_out, uint64_t *offsets,
                                  int64_t *m_lo, int64_t *m_cnt) {
  uint64_t T = 0;
  for (int g = 0; g < world; g++) {
    offsets[g] = T;
    T += totals[g];
  }
  uint64_t r = T ? amclj_oracle_systematic_offset(u0, T) : 0;
  *r_out = r;
  for (int g = 0; g < world; g++) {
    if (T == 0) {
      m_lo[g] = 0;
      m_cnt[g] = 0;
      continue;
    }
    int64_t lo = slots_below(offsets[g], T, r, n_global);
    int64_t hi = slots_below(offsets[g] + totals[g], T, r, n_global);
    m_lo[g] = lo;
    m_cnt[g] = hi - lo;
  }
}

/* ------------------------------------------------------------------------------------
 * pf.clj:156-167 roulette-selection.  Attempt t draws idx ~ U{0..N-1} and u ~ U[0,1);
 * accepted when wn[idx] > u, i.e. q_idx / T > u32 / 2^32 in the integer contract; output
 * order = acceptance order.  Returns the number of attempts consumed, or -1 if the stream
 * ran out before N acceptances.
 * ---------------------------------------------------------------------------------- */
static inline int roulette_accept(uint64_t q, uint32_t u, uint64_t total) {
  return ((unsigned __int128)q << 32) > (unsigned __int128)u * total;
}

ORACLE_API int64_t amclj_oracle_roulette_stream(const uint64_t *q, int64_t n, uint64_t total,
                                                const int32_t *sidx, const uint32_t *su,
                                                int64_t len, int32_t *idx_out) {
  int64_t got = 0;
  for (int64_t t = 0; t < len; t++) {
    if (got == n) return t;
    int32_t i = sidx[t];
    if (i >= 0 && i < n && roulette_accept(q[i], su[t], total)) idx_out[got++] = i;
  }
  return got == n ? len : -1;
}

/* Philox attempt stream: block b = t >> 1 with counter (b lo, b hi, 0, TAG), key = seed;
 * attempt t uses words (2(t&1), 2(t&1)+1): idx = (word * N) >> 32, u = next word. */
static inline void roulette_attempt(uint64_t seed, uint32_t tag, uint64_t t, int64_t n,
                                    int32_t *idx, uint32_t *u) {
  uint64_t b = t >> 1;
  uint32_t r[4];
  philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), 0u, tag, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  int k = (int)(t & 1) * 2;
  *idx = (int32_t)(((uint64_t)r[k] * (uint64_t)n) >> 32);
  *u = r[k + 1];
}

ORACLE_API void amclj_oracle_roulette_attempts(uint64_t seed, int64_t t0, int64_t len,
                                               int64_t n, int32_t *sidx, uint32_t *su) {
  for (int64_t t = 0; t < len; t++)
    roulette_attempt(seed, TAG_ROULETTE, (uint64_t)(t0 + t), n, &sidx[t], &su[t]);
}

ORACLE_API int64_t amclj_oracle_roulette_philox(const uint64_t *q, int64_t n, uint64_t total,
                                                uint64_t seed, int32_t *idx_out,
                                                int32_t threads) {
  /* Chunked so the expected N^2 attempts can use every host core; the accepted attempts
   * of each chunk are concatenated in attempt order, so the result equals the sequential
   * loop of pf.clj:162-167. */
  const int64_t CH = 1 << 16;
  int64_t got = 0, t0 = 0;
  int nth = 1;
#ifdef _OPENMP
  nth = threads > 0 ? threads : omp_get_max_threads();
#endif
  int64_t nchunks = (int64_t)nth * 64;
  int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nchunks * CH));
  int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)nchunks);
  if (total == 0) {
    free(buf);
    free(cnt);
    return -1;
  }
  while (got < n) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nth)
#endif
    for (int64_t c = 0; c < nchunks; c++) {
      int64_t k = 0, base = t0 + c * CH;
      for (int64_t t = 0; t < CH; t++) {
        int32_t i;
        uint32_t u;
        roulette_attempt(seed, TAG_ROULETTE, (uint64_t)(base + t), n, &i, &u);
        if (roulette_accept(q[i], u, total)) buf[c * CH + k++] = i;
      }
      cnt[c] = k;
    }
    for (int64_t c = 0; c < nchunks && got < n; c++)
      for (int64_t k = 0; k < cnt[c] && got < n; k++) idx_out[got++] = buf[c * CH + k];
    t0 += nchunks * CH;
  }
  free(buf);
  free(cnt);
  return t0;
}

/* pf.clj:182-195 tournament-selection: N rounds, two uniform draws, keep the heavier
 * (`(> w1 w2)` ? p1 : p2 — ties keep the second).  Draw j uses stream entry j. */
ORACLE_API void amclj_oracle_tournament_stream(const float *w, int64_t n, const int32_t *sidx,
                                               int32_t *idx_out) {
  for (int64_t k = 0; k < n; k++) {
    int32_t a = sidx[2 * k], b = sidx[2 * k + 1];
    idx_out[k] = (w[a] > w[b]) ? a : b;
  }
}

ORACLE_API void amclj_oracle_tournament_philox(const float *w, int64_t n, uint64_t seed,
                                               int32_t *idx_out) {
  for (int64_t k = 0; k < n; k++) {
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)((uint64_t)k >> 32), 0u, TAG_TOURNAMENT,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    int32_t a = (int32_t)(((uint64_t)r[0] * (uint64_t)n) >> 32);
    int32_t b = (int32_t)(((uint64_t)r[1] * (uint64_t)n) >> 32);
    idx_out[k] = (w[a] > w[b]) ? a : b;
  }
}

/* ------------------------------------------------------------------------------------
 * pf.clj:67-84 pose-estimate: unweighted mean of x, y and of the quaternion components.
 * Planar unit quaternion of heading theta: (0, 0, sin(theta/2), cos(theta/2))
 * (quaternions.clj:19-36 applied to the identity).  out = mean x, y, qz, qw.
 * ---------------------------------------------------------------------------------- */
ORACLE_API void amclj_oracle_pose_estimate(const float *x, const float *y, const float *theta,
                                           int64_t n, double out[4]) {
  double sx = 0, sy = 0, sz = 0, sw = 0;
  for (int64_t i = 0; i < n; i++) {
    sx += x[i];
    sy += y[i];
    sz += sin((double)theta[i] / 2);
    sw += cos((double)theta[i] / 2);
  }
  double inv = n > 0 ? 1.0 / (double)n : 0.0;
  out[0] = sx * inv;
  out[1] = sy * inv;
  out[2] = sz * inv;
  out[3] = sw * inv;
}

/* ------------------------------------------------------------------------------------
 * map.clj:59-84 uniform-initialization.  A proposal is (x, y) ~ U[0,W) x U[0,H) in CELL
 * units, accepted when cell (int x, int y) == 0 (map.clj:52-55 truncates; no rounding),
 * then scaled by the resolution (map->world, map.clj:29-34).  Heading: N(0,1) rad in the
 * reference (map.clj:67-68); heading_mode 1 = U(-pi, pi).
 * Contract for the device path: particle gid uses its own Philox stream, attempt a has
 * counter (gid lo, gid hi, a, TAG_INIT): x = u24(r0)*W, y = u24(r1)*H; heading from
 * (r2, r3) of the accepted attempt.
 * ---------------------------------------------------------------------------------- */
ORACLE_API void amclj_oracle_uniform_init(const int8_t *cells, int32_t W, int32_t H, float res,
                                          int64_t gid_first, int64_t np, int32_t heading_mode,
                                          uint64_t seed, float *x, float *y, float *theta) {
  oracle_map m = {cells, W, H, res};
  for (int64_t i = 0; i < np; i++) {
    uint64_t gid = (uint64_t)(gid_first + i);
    for (uint32_t a = 0;; a++) {
      uint32_t r[4];
      philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), a, TAG_INIT, (uint32_t)seed,
                    (uint32_t)(seed >> 32), r);
      float px = u24(r[0]) * (float)W, py = u24(r[1]) * (float)H;
      int32_t cx = (int32_t)px, cy = (int32_t)py;
      if (a < 100000u && (cx >= W || cy >= H || blocked(&m, cx, cy))) continue;
      x[i] = px * res;
      y[i] = py * res;
      if (heading_mode == 0)
        theta[i] = sqrtf(-2.0f * logf(u24(r[2]))) * cosf(6.283185307179586f * u24(r[3]));
      else
        theta[i] = (u24(r[2]) * 2.0f - 1.0f) * 3.14159265358979f;
      break;
    }
  }
}

/* ------------------------------------------------------------------------------------
 * Log-weights -> linear weights before the fixed-point conversion (S7 = 1; no reference
 * counterpart: the reference's weights are linear, sensor.clj:162).  Contract: w_i =
 * exp(raw_i - max raw) with a specified exp (Cody-Waite by ln2, Cephes degree-5
 * polynomial, explicit fmaf, flush below -87), so raw -> q_i is bit-reproducible.
 * ---------------------------------------------------------------------------------- */
ORACLE_API float amclj_oracle_expf_contract(float x) {
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
  return ldexpf(y, (int)n);
}

ORACLE_API void amclj_oracle_linear_weights(const float *raw, int64_t n, float rawmax,
                                            float *w) {
  for (int64_t i = 0; i < n; i++) w[i] = amclj_oracle_expf_contract(raw[i] - rawmax);
}

ORACLE_API float amclj_oracle_max(const float *v, int64_t n) {
  float m = -INFINITY;
  for (int64_t i = 0; i < n; i++) m = fmaxf(m, v[i]);
  return m;
}

/* map.clj:87-111 gaussian-initialization / gaussian-noise: x, y ~ N(pose, sd_xy); heading
 * rotated by N(rot-mean, sd) where the reference passes sd = rot-mean = 0 (map.clj:106).
 * Device contract: Philox counter (gid lo, gid hi, 0, TAG_GAUSS), Box-Muller as for motion. */
#define TAG_GAUSS 0x47415553u
ORACLE_API void amclj_oracle_gaussian_init(float mx, float my, float mth, float sd_xy,
                                           float sd_th, int64_t gid_first, int64_t np,
                                           uint64_t seed, float *x, float *y, float *theta) {
  for (int64_t i = 0; i < np; i++) {
    uint64_t gid = (uint64_t)(gid_first + i);
    uint32_t r[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0u, TAG_GAUSS, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float ra = sqrtf(-2.0f * logf(u24(r[0]))), rb = sqrtf(-2.0f * logf(u24(r[2])));
    float a = 6.283185307179586f * u24(r[1]), b = 6.283185307179586f * u24(r[3]);
    x[i] = fmaf(sd_xy, ra * cosf(a), mx);
    y[i] = fmaf(sd_xy, ra * sinf(a), my);
    theta[i] = fmaf(sd_th, rb * cosf(b), mth);
  }
}

/* ------------------------------------------------------------------------------------
 * pf.clj:243-279 augmented MCL.  w-avg = incanter stats/median of the weights (third party,
 * incanter 1.5.4: restated as the textbook median, mean of the two middle order statistics for
 * an even count); slow/fast are the w-slow/w-fast atoms (pf.clj:243-244, :249-250).  Slot m
 * draws Math/random; `(> random (max 0.0 (- 1.0 (/ fast slow))))` injects a particle from
 * uniform-initialization (map.clj:71-84; heading N(0,1), map.clj:67-68), otherwise roulette-step
 * (pf.clj:169-180) rejection-samples against the UN-normalised weights.
 * Injected streams: slot m owns Philox counters (m lo, m hi, attempt, TAG_AUGMENT); attempt 0
 * word 0 is the inject draw; attempts 1.. are the proposals of whichever loop runs.
 * ---------------------------------------------------------------------------------- */
#define TAG_AUGMENT 0x4155474du

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

ORACLE_API double amclj_oracle_median(const float *w, int64_t n) {
  float *t = (float *)malloc(sizeof(float) * (size_t)n);
  memcpy(t, w, sizeof(float) * (size_t)n);
  qsort(t, (size_t)n, sizeof(float), cmp_float);
  double m = (n % 2) ? (double)t[n / 2] : ((double)t[n / 2] + (double)t[n / 2 - 1]) / 2.0;
  free(t);
  return m;
}

ORACLE_API void amclj_oracle_augmented(const float *w, const float *x, const float *y,
                                       const float *th, int64_t n, double *slow, double *fast,
                                       double alpha_slow, double alpha_fast, const int8_t *cells,
                                       int32_t W, int32_t H, float res, uint64_t seed, float *ox,
                                       float *oy, float *oth, float *ow, int32_t *oidx) {
  oracle_map mp = {cells, W, H, res};
  double w_avg = amclj_oracle_median(w, n);
  *slow += alpha_slow * (w_avg - *slow);
  *fast += alpha_fast * (w_avg - *fast);
  double thr = 1.0 - *fast / *slow;
  if (!(thr > 0.0)) thr = 0.0;
  float fthr = (float)thr;
  for (int64_t m = 0; m < n; m++) {
    uint32_t r[4];
#define DRAW(at) philox4x32_10((uint32_t)m, (uint32_t)((uint64_t)m >> 32), (at), TAG_AUGMENT, \
                               (uint32_t)seed, (uint32_t)(seed >> 32), r)
    DRAW(0u);
    if (u24(r[0]) > fthr) {
      for (uint32_t at = 1;; at++) {
        DRAW(at);
        float px = u24(r[0]) * (float)W, py = u24(r[1]) * (float)H;
        int32_t cx = (int32_t)px, cy = (int32_t)py;
        if (at < 100000u && (cx >= W || cy >= H || blocked(&mp, cx, cy))) continue;
        ox[m] = px * res;
        oy[m] = py * res;
        oth[m] = sqrtf(-2.0f * logf(u24(r[2]))) * cosf(6.283185307179586f * u24(r[3]));
        ow[m] = 0.0f;
        oidx[m] = -1;
        break;
      }
      continue;
    }
    for (uint32_t at = 1;; at++) {
      DRAW(at);
      int32_t i = (int32_t)(((uint64_t)r[0] * (uint64_t)n) >> 32);
      if (w[i] > u24(r[1]) || at >= 100000u) {
        ox[m] = x[i];
        oy[m] = y[i];
        oth[m] = th[i];
        ow[m] = w[i];
        oidx[m] = i;
        break;
      }
    }
#undef DRAW
  }
}

ORACLE_API int32_t amclj_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
