// api.cu — the C ABI of libamclj_b200.so (include/amclj_b200.h): context, staging of the
// launch-uniform tables, and the stream choreography of one filter update.
//
// Boundary replaced (jvia/amclj): the three calls of one mcl* iteration, pf.clj:229-233 —
// apply-motion-model (pf.clj:139-145), apply-sensor-model (:148-154) and
// roulette-selection (:156-167) — plus pose-estimate (:67-84, :234).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "internal.h"

using namespace amclj;

namespace {

int fail(amclj_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? AMCLJ_ERR_NOMEM : AMCLJ_ERR_CUDA, \
                  "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
  } while (0)

#define NEED(cond, code, ...)                          \
  do {                                                 \
    if (!(cond)) return fail(ctx, code, __VA_ARGS__);  \
  } while (0)

#define ENTER()                                                             \
  do {                                                                      \
    if (!ctx) return AMCLJ_ERR_INVALID;                                     \
    cudaError_t e0_ = cudaSetDevice(ctx->device);                           \
    if (e0_ != cudaSuccess)                                                 \
      return fail(ctx, AMCLJ_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e0_)); \
  } while (0)

template <class T>
cudaError_t dalloc(T **p, size_t count) {
  return cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
}
template <class T>
void dfree(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void preset(amclj_params *p, int which) {
  memset(p, 0, sizeof(*p));
  p->max_range = -1.0f; /* sensor.clj:12 */
  p->z_hit = 0.95f;     /* sensor.clj:14-19 */
  p->z_short = 0.1f;
  p->z_max = 0.05f;
  p->z_rand = 0.05f;
  p->sigma_hit = 0.2f;
  p->lambda_short = 0.1f;
  p->laser_x = 0.135f; /* sensor.clj:80 */
  p->laser_y = 0.0f;
  p->alpha1 = 0.002f; /* motion.clj:12-15 */
  p->alpha2 = 0.002f;
  p->alpha3 = 0.02f;
  p->alpha4 = 0.02f;
  if (which == AMCLJ_PRESET_REF) {
    p->s6_max_beams = 30; /* sensor.clj:13 */
    p->s8_resampler = AMCLJ_RESAMPLE_ROULETTE;
  } else {
    p->s1_use_heading = 1;
    p->s2_use_scan_range_max = 1;
    p->s3_proper_endpoint = 1;
    p->s4_endpoint_termination = 1;
    p->s5_use_sigma_hit = 1;
    p->s6_max_beams = 0;
    p->s7_log_weights = 1;
    p->s8_resampler = AMCLJ_RESAMPLE_SYSTEMATIC;
  }
}

/* sensor.clj:144-145: step = int((n-1)/(max-beams-1)); range(0, n, step) */
int beam_step(int n, int max_beams) {
  if (max_beams <= 1) return 1;
  int step = (n - 1) / (max_beams - 1);
  return step < 1 ? 1 : step;
}
int beams_used(int n, int max_beams) {
  if (n <= 0) return 0;
  int step = beam_step(n, max_beams);
  return (n + step - 1) / step;
}

int ensure_particles(amclj_ctx *ctx, int64_t n) {
  if (n <= ctx->cap) return AMCLJ_OK;
  NEED(!ctx->frozen, AMCLJ_ERR_STATE, "particle buffers are exported to peers and cannot grow");
  int64_t cap = (n + 1023) / 1024 * 1024;
  float **bufs[] = {&ctx->x, &ctx->y, &ctx->th, &ctx->w, &ctx->raw, &ctx->x2, &ctx->y2, &ctx->th2, &ctx->w2};
  /* growing keeps the live particles */
  for (float **b : bufs) {
    float *nb = nullptr;
    CK(dalloc(&nb, (size_t)cap));
    if (*b && ctx->n > 0)
      CK(cudaMemcpyAsync(nb, *b, sizeof(float) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(*b);
    *b = nb;
  }
  dfree(ctx->cdf);
  CK(dalloc(&ctx->cdf, (size_t)cap));
  dfree(ctx->idx);
  CK(dalloc(&ctx->idx, (size_t)cap));
  dfree(ctx->perm);
  CK(dalloc(&ctx->perm, (size_t)cap));
  dfree(ctx->normals);
  CK(dalloc(&ctx->normals, (size_t)cap * 3));
  NEED(scan_tiles(cap) <= ctx->scan_state_cap, AMCLJ_ERR_INVALID, "too many particles for one ctx");
  ctx->cap = cap;
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int ensure_stage(amclj_ctx *ctx, int64_t rows) {
  if (rows <= ctx->stage_cap) return AMCLJ_OK;
  int64_t cap = (rows + 1023) / 1024 * 1024;
  dfree(ctx->stage_out);
  dfree(ctx->stage_in);
  dfree(ctx->stage_idx);
  CK(dalloc(&ctx->stage_out, (size_t)cap));
  CK(dalloc(&ctx->stage_in, (size_t)cap));
  CK(dalloc(&ctx->stage_idx, (size_t)cap));
  ctx->stage_cap = cap;
  return AMCLJ_OK;
}

/* The Clojure `def` constants are double literals (motion.clj:12-15: 0.002, 0.02) while amclj_params
 * carries floats: widen a parameter to the double its shortest round-tripping decimal names
 * (0.002f -> 0.002), i.e. back to the reference's literal. */
double decimal_exact(float f) {
  if (!(f == f) || f == 0.0f || f - f != 0.0f) return (double)f;
  char buf[40];
  for (int prec = 1; prec <= 9; prec++) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
    if (strtof(buf, nullptr) == f) return strtod(buf, nullptr);
  }
  return (double)f;
}

/* motion.clj:68-79 + :24-34: the control is launch-uniform; doubles as on the JVM. */
void make_control(const amclj_params &p, const double curr[3], const double prev[3],
                  MotionLaunch &m) {
  double xn = curr[0], yn = curr[1], tn = curr[2], x = prev[0], y = prev[1], t = prev[2];
  m.is_static = ((x - xn) == 0.0 && (y - yn) == 0.0 && (t - tn) == 0.0); /* motion.clj:26 */
  m.trans_zero = 0;
  m.rot1 = m.trans = m.rot2 = m.sd1 = m.sdt = m.sd2 = 0.f;
  if (m.is_static) return;
  double rot1 = atan2(yn - y, xn - x) - t;
  double trans = hypot(x - xn, y - yn);
  double rot2 = tn - t - rot1;
  double a1 = decimal_exact(p.alpha1), a2 = decimal_exact(p.alpha2), a3 = decimal_exact(p.alpha3), a4 = decimal_exact(p.alpha4);
  double v1 = p.s9_pr_rot1_variance ? a1 * rot1 * rot1 + a2 * trans * trans
                                    : a1 * rot1 * rot1 + a2 * rot2 * rot2; /* motion.clj:35 */
  m.trans_zero = trans == 0.0; /* motion.clj:36 */
  m.rot1 = (float)rot1;
  m.trans = (float)trans;
  m.rot2 = (float)rot2;
  m.sd1 = (float)sqrt(v1);
  m.sdt = (float)sqrt(a3 * trans * trans + a4 * rot1 * rot1 + a4 * rot2 * rot2);
  m.sd2 = (float)sqrt(a1 * rot2 * rot2 + a2 * trans * trans);
}

/* df_mode 0 and 8 pick the field placement per update; 8 leaves the persistent ray tables out */
bool auto_mode(const amclj_ctx *ctx) { return ctx->p.df_mode == 0 || ctx->p.df_mode == 8; }

int df_mode_of(const amclj_ctx *ctx) {
  switch (ctx->p.df_mode) {
    case 1: return ctx->map.fits_smem ? DF_SMEM4 : -1;
    case 2: return ctx->map.df8 ? DF_GLOBAL8 : -1;
    case 3: return DF_GLOBAL4;
    case 4: return ctx->map.wedge ? DF_WEDGE8 : -1;
    case 5: return ctx->map.wedge ? DF_WEDGE8 : -1; /* directional planes, particles walked in tile order */
    case 6: return ctx->map.wedge ? DF_WEDGE8 : (ctx->map.df8 ? DF_GLOBAL8 : DF_GLOBAL4); /* ray tables, filled over these */
    case 7: return ctx->map.wedge ? DF_WEDGE8 : (ctx->map.df8 ? DF_GLOBAL8 : DF_GLOBAL4); /* persistent ray tables */
    default: return ctx->map.fits_smem ? DF_SMEM4 : (ctx->map.df8 ? DF_GLOBAL8 : DF_GLOBAL4);
  }
}

/* End-cell ring of a ray of r cells (raytable.cu): a start and an end cell are rounded
 * independently (map.clj:36-46), so (x1 - x0, y1 - y0) = r (cos, sin) + delta with |delta| < 1 per
 * axis: an offset (ex, ey) is reachable only when the square of half-side 1 around it meets the
 * circle of radius r.  The ring keeps every cell whose square of half-side 1.05 does (the 0.05
 * covers the fp32 arithmetic of the end point, ~1e-4 cells); per row the admissible |ex| form one
 * interval.  An end cell outside the ring (never seen) makes the kernel walk that ray instead.
 * Returns E (rows are ey = -E .. E), or -1 when the geometry is out of range for the tables. */
int build_ring(double r, std::vector<uint4> &rows, std::vector<short2> &ents) {
  rows.clear();
  ents.clear();
  const double half = 1.05;
  if (!(r >= 0.0) || !(r + 2.0 * half < 30000.0)) return -1; /* offsets are stored as 16-bit values */
  const int E = (int)ceil(r + half);
  rows.reserve((size_t)2 * E + 1);
  uint32_t base = 0;
  for (int ey = -E; ey <= E; ey++) {
    const double ay = fabs((double)ey), ny = fmax(0.0, ay - half), fy = ay + half;
    int lo = -1, hi = -1;
    for (int a = 0; a <= E + 1; a++) {
      double nearest = hypot(fmax(0.0, (double)a - half), ny), farthest = hypot((double)a + half, fy);
      if (nearest <= r && r <= farthest) {
        if (lo < 0) lo = a;
        hi = a;
      }
    }
    uint32_t cnt = lo < 0 ? 0u : (uint32_t)(hi - lo + 1);
    if (lo < 0) lo = 0;
    rows.push_back(make_uint4(base, (uint32_t)lo, cnt, base + cnt));
    for (uint32_t k = 0; k < cnt; k++) ents.push_back(make_short2((short)(lo + (int)k), (short)ey));
    for (uint32_t k = 0; k < cnt; k++) {
      int a = lo + (int)k;
      ents.push_back(make_short2(a == 0 ? (short)-32768 : (short)-a, (short)ey));
    }
    base += 2 * cnt;
    if (base > (1u << 20)) return -1; /* a table per cell would not pay */
  }
  if (ents.empty()) return -1;
  ents.push_back(make_short2((short)-32767, 0)); /* last entry: the NaN the fast path reads for an end cell off the ring */
  return E;
}

void fill_sensor_args(const amclj_ctx *ctx, SensorLaunch &a, int mode);

/* stage the ring for this scan geometry (raytable.cu, and the set-up records of sensor.cu) */
int stage_ring(amclj_ctx *ctx, double r) {
  RayTables &t = ctx->rt;
  const bool tables_wanted = ctx->rt_enable && (auto_mode(ctx) || ctx->p.df_mode == 6 || ctx->p.df_mode == 7);
  t.available = false;
  if (t.r_cells == r && t.rows) {
    t.have_ring = t.n_entries > 0;
    t.available = t.have_ring && tables_wanted;
    return AMCLJ_OK;
  }
  t.r_cells = r;
  t.n_entries = 0;
  t.have_ring = false;
  t.recs_mode = -1;
  std::vector<uint4> rows;
  std::vector<short2> ents;
  const int E = build_ring(r, rows, ents);
  if (E < 0) return AMCLJ_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  if ((int)rows.size() > t.rows_cap) {
    dfree(t.rows);
    CK(dalloc(&t.rows, rows.size()));
    t.rows_cap = (int)rows.size();
  }
  if ((int)ents.size() > t.ents_cap) {
    dfree(t.ents);
    CK(dalloc(&t.ents, ents.size()));
    t.ents_cap = (int)ents.size();
  }
  CK(cudaMemcpyAsync(t.rows, rows.data(), sizeof(uint4) * rows.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(t.ents, ents.data(), sizeof(short2) * ents.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  t.nrows = (int)rows.size();
  t.E = E;
  t.n_entries = (int)ents.size();
  t.have_ring = true;
  t.available = tables_wanted;
  return AMCLJ_OK;
}

/* particles per start cell from which filling a table beats walking the cell's rays: a table is
 * n_entries rays, a looked-up ray costs about a fifth of a walked one */
int rt_threshold(const amclj_ctx *ctx, bool forced) {
  if (forced) return 1;
  if (ctx->rt.thr > 0) return ctx->rt.thr;
  int nb = ctx->beams.nb > 0 ? ctx->beams.nb : 1;
  int thr = (int)ceil(1.75 * (double)ctx->rt.n_entries / (double)nb);
  return thr < 4 ? 4 : thr;
}

/* scratch of the ray-table path, sized for this particle count */
int ensure_rt_scratch(amclj_ctx *ctx, int64_t n, bool diag, int thr, int *max_slots, int *max_items) {
  RayTables &t = ctx->rt;
  if (!t.bins) {
    CK(rt_prepare_device());
    CK(dalloc(&t.bins, (size_t)RT_BIN_CAP));
    CK(cudaMemsetAsync(t.bins, 0, sizeof(uint32_t) * RT_BIN_CAP, ctx->stream)); /* rt_finish_kernel re-zeroes them after every update */
    CK(dalloc(&t.slots, (size_t)RT_BIN_CAP));
    CK(dalloc(&t.item0, (size_t)RT_BIN_CAP + 1));
    CK(dalloc(&t.ctl, 1));
    CK(cudaMemsetAsync(t.ctl, 0, sizeof(RtCtl), ctx->stream));
  }
  int64_t items = (n + RT_PI - 1) / RT_PI + RT_BIN_CAP + 1;
  if (n > t.pstate_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(t.pstate);
    t.pstate_cap = 0;
    CK(dalloc(&t.pstate, (size_t)n));
    t.pstate_cap = n;
  }
  size_t table_bytes = (size_t)t.n_entries * (diag ? 16 : 4);
  int64_t slots = (int64_t)(t.budget / table_bytes);
  if (slots > RT_BIN_CAP) slots = RT_BIN_CAP;
  if (slots > n / thr + 1) slots = n / thr + 1;
  if (slots < 1) slots = 1;
  size_t need = (size_t)slots * table_bytes;
  if (need > t.tables_bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (t.tables) cudaFree(t.tables);
    t.tables = nullptr;
    t.tables_bytes = 0;
    CK(cudaMalloc(&t.tables, need));
    t.tables_bytes = need;
  }
  *max_slots = (int)slots;
  *max_items = (int)(items > 0x7fffffff ? 0x7fffffff : items);
  return AMCLJ_OK;
}


/* ---- persistent map-wide ray tables (ptable.cu) -------------------------------------------- */
void pt_free(amclj_ctx *ctx) {
  PersistentTables &t = ctx->pt;
  dfree(t.slot_of_cell); dfree(t.cell_of_slot); dfree(t.bin_count); dfree(t.bin_start); dfree(t.cursor); dfree(t.scan_state);
  if (t.tables) cudaFree(t.tables);
  t.tables = nullptr;
  t.tables_bytes = 0;
  t.n_slots = 0;
  t.built = false;
  t.key_map_gen = -1;
}

/* slot maps of a new map: one table slot per free cell (map.clj:48-57: free = cell 0) */
int pt_set_map(amclj_ctx *ctx, const int8_t *cells, int W, int H) {
  pt_free(ctx);
  PersistentTables &t = ctx->pt;
  const size_t ncell = (size_t)W * H;
  if (!t.enable || ncell > (1u << 24)) return AMCLJ_OK; /* larger maps: no table per cell would fit anyway */
  std::vector<int32_t> soc(ncell);
  std::vector<int2> cos;
  for (size_t i = 0; i < ncell; i++)
    if (cells[i] == 0) {
      soc[i] = (int32_t)cos.size();
      cos.push_back(make_int2((int)(i % W), (int)(i / W)));
    }
  const int32_t n_slots = (int32_t)cos.size();
  if (n_slots == 0) return AMCLJ_OK;
  for (size_t i = 0; i < ncell; i++)
    if (cells[i] != 0) soc[i] = n_slots; /* the zero table */
  CK(dalloc(&t.slot_of_cell, ncell));
  CK(dalloc(&t.cell_of_slot, (size_t)n_slots));
  CK(dalloc(&t.bin_count, (size_t)n_slots + 2));
  CK(dalloc(&t.bin_start, (size_t)n_slots + 2));
  CK(dalloc(&t.cursor, (size_t)n_slots + 2));
  t.scan_words = (n_slots + 1 + 2047) / 2048 + 2;
  CK(dalloc(&t.scan_state, (size_t)t.scan_words));
  CK(cudaMemsetAsync(t.scan_state, 0, sizeof(unsigned long long) * (size_t)t.scan_words, ctx->stream));
  CK(cudaMemcpyAsync(t.slot_of_cell, soc.data(), sizeof(int32_t) * ncell, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(t.cell_of_slot, cos.data(), sizeof(int2) * (size_t)n_slots, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(t.bin_count, 0, sizeof(uint32_t) * ((size_t)n_slots + 2), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  t.n_slots = n_slots;
  return AMCLJ_OK;
}

int pt_fill_mode(const amclj_ctx *ctx) {
  return ctx->map.wedge ? DF_WEDGE8 : (ctx->map.df8 ? DF_GLOBAL8 : DF_GLOBAL4);
}

void fill_pt_args(const amclj_ctx *ctx, PtLaunch &L) {
  const PersistentTables &t = ctx->pt;
  const RayTables &r = ctx->rt;
  memset(&L, 0, sizeof(L));
  fill_sensor_args(ctx, L.s, pt_fill_mode(ctx));
  L.rows = r.rows; L.ents = r.ents; L.nrows = r.nrows; L.E = r.E; L.n_entries = r.n_entries;
  L.stride = t.stride;
  L.n_slots = t.n_slots;
  L.slot_of_cell = t.slot_of_cell; L.cell_of_slot = t.cell_of_slot; L.tables = t.tables;
  L.bin_count = t.bin_count; L.bin_start = t.bin_start; L.cursor = t.cursor;
  L.pkey = t.pkey;
  L.pstate_d = t.pstate_d;
  L.rinv_d = 1.0 / (double)ctx->map.res;
  L.perm = ctx->perm;
  L.pstate = r.pstate;
  L.qacc = ctx->partials;
  L.warps = t.warps;
  L.nbuf = t.nbuf_used;
  L.scan_state = t.scan_state;
  L.scan_words = t.scan_words;
  L.tail_pct = t.tail_pct;
  L.tail_div = t.tail_div;
  L.tail_min = t.tail_min;
  L.tail_ctr = reinterpret_cast<uint32_t *>(t.scan_state) + 1;
  L.rows32 = (t.rows32 && r.n_entries < 65536 && r.E < 250) ? 1 : 0; /* base : 16 | lo : 8 | cnt : 8 (cnt <= ~sqrt(8 r) + 4) */
  L.fuse_finish = (t.fuse_finish && (ctx->beams.nb + 29) / 30 <= 15) ? 1 : 0;
}

/* Called at the end of stage_scan_impl: are the persistent tables usable for the staged scan, and do
 * they have to be (re)filled?  They depend on the map, the ray length in cells, S3 / S4, the miss value
 * and the resolution — not on the scan's ranges, the beam count or the particles. */
int pt_prepare(amclj_ctx *ctx, double rcells) {
  PersistentTables &t = ctx->pt;
  const RayTables &r = ctx->rt;
  const amclj_params &p = ctx->p;
  t.built = false;
  const bool wanted = t.enable && (p.df_mode == 0 || p.df_mode == 7) && t.n_slots > 0 && r.have_ring && r.floor_trick_ok &&
                      ctx->beams.cr.K != 0.0f &&
                      ctx->map.res > 1.0e-6f && ctx->map.res < 1.0e6f;
  if (!wanted) return AMCLJ_OK;
  const int32_t stride = (r.n_entries + 3) & ~3;
  const size_t bytes = ((size_t)t.n_slots + 1) * (size_t)stride * sizeof(float);
  if (bytes > t.budget) return AMCLJ_OK;
  PtLaunch probe;
  memset(&probe, 0, sizeof(probe));
  probe.s.nb = ctx->beams.nb; probe.nrows = r.nrows; probe.stride = stride;
  /* warps per CTA and table buffers per warp: as many warps as the shared memory takes (they hide one
   * another's latencies: 8 -> 14 warps took the look-up from 1.86 to 1.40 ms); with one buffer per warp
   * more of them fit */
  int warps = 0, nbuf = 0;
  const int wcap = t.warps_req > 0 && t.warps_req <= pt_max_warps() ? t.warps_req : pt_max_warps();
  for (int nb_try = (t.nbuf == 2 ? 2 : 1); nb_try <= (t.nbuf == 1 ? 1 : 2) && !warps; nb_try++) {
    probe.nbuf = nb_try;
    for (int w = wcap; w >= 4; w--)
      if (pt_lookup_smem(probe, w) + 1024 <= (size_t)ctx->max_smem_optin) { warps = w; nbuf = nb_try; break; }
  }
  if (warps < 4) return AMCLJ_OK; /* the tables do not fit the shared memory: the other paths take it */
  const float R = p.s2_use_scan_range_max ? ctx->beams.range_max : p.max_range;
  uint32_t Rb, resb;
  memcpy(&Rb, &R, 4);
  memcpy(&resb, &ctx->map.res, 4);
  const bool same = t.tables && t.key_map_gen == ctx->map_gen && t.key_r == rcells && t.key_s3 == p.s3_proper_endpoint &&
                    t.key_s4 == p.s4_endpoint_termination && t.key_entries == r.n_entries && t.key_R_bits == Rb &&
                    t.key_res_bits == resb && t.fill_mode == pt_fill_mode(ctx);
  t.stride = stride;
  if (!same) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (bytes > t.tables_bytes) {
      if (t.tables) cudaFree(t.tables);
      t.tables = nullptr;
      t.tables_bytes = 0;
      CK(cudaMalloc(reinterpret_cast<void **>(&t.tables), bytes));
      t.tables_bytes = bytes;
    }
    PtLaunch L;
    fill_pt_args(ctx, L);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    CK(launch_pt_fill(ctx, L, pt_fill_mode(ctx), ctx->stream));
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    t.ms_fill = ms;
    t.key_map_gen = ctx->map_gen; t.key_r = rcells; t.key_s3 = p.s3_proper_endpoint; t.key_s4 = p.s4_endpoint_termination;
    t.key_entries = r.n_entries; t.key_R_bits = Rb; t.key_res_bits = resb; t.fill_mode = pt_fill_mode(ctx);
  }
  t.warps = warps;
  t.nbuf_used = nbuf;
  t.built = true;
  return AMCLJ_OK;
}

/* scratch of the persistent-table update for this particle count */
int ensure_pt_scratch(amclj_ctx *ctx, int64_t n) {
  PersistentTables &t = ctx->pt;
  RayTables &r = ctx->rt;
  if (n > t.pkey_cap || n > r.pstate_cap || n > ctx->partials_cap) CK(cudaStreamSynchronize(ctx->stream));
  if (n > t.pkey_cap) {
    dfree(t.pkey);
    dfree(t.pstate_d);
    t.pkey_cap = 0;
    CK(dalloc(&t.pkey, (size_t)n));
    CK(dalloc(&t.pstate_d, 2 * (size_t)n)); /* {(sin, cos), ray origin} per particle */
    t.pkey_cap = n;
  }
  if (n > r.pstate_cap) {
    dfree(r.pstate);
    r.pstate_cap = 0;
    CK(dalloc(&r.pstate, (size_t)n));
    r.pstate_cap = n;
  }
  if (n > ctx->partials_cap) {
    dfree(ctx->partials);
    ctx->partials_cap = 0;
    CK(dalloc(&ctx->partials, (size_t)n));
    ctx->partials_cap = n;
  }
  return AMCLJ_OK;
}

/* the persistent-table sensor pass over the particles of `base`; results land in base.raw */
int enqueue_ptable(amclj_ctx *ctx, const SensorLaunch &base, bool dbg) {
  int rc = ensure_pt_scratch(ctx, base.n);
  if (rc) return rc;
  PtLaunch L;
  fill_pt_args(ctx, L);
  L.s.x = base.x; L.s.y = base.y; L.s.th = base.th; L.s.raw = base.raw; L.s.n = base.n;
  L.s.rawmax_ord = base.rawmax_ord; L.s.stats = base.stats;
  L.s.dbg_first = base.dbg_first; L.s.dbg_count = base.dbg_count; L.s.dbg_range = base.dbg_range;
  const bool hist_done = ctx->pt.hist_done && !dbg && base.x == ctx->x && base.n == ctx->n;
  ctx->pt.hist_done = false;
  CK(launch_pt_update(ctx, L, pt_fill_mode(ctx), dbg, hist_done, ctx->stream));
  ctx->launches += (L.fuse_finish ? 4 : 5) - (hist_done ? 1 : 0); /* histogram, scan, scatter, look-up (, finish) */
  return AMCLJ_OK;
}

/* the per-message part of the beam table: (obs, e2, p34, 0) of one beam (sensor.clj:154-161) */
static inline void fill_mix(const amclj_params &p, float o, float range_max, float *mix4) {
  mix4[0] = o;
  mix4[1] = p.z_short * p.lambda_short * expf(-p.lambda_short * o);      /* sensor.clj:154-157 */
  float pz3 = o == range_max ? p.z_max : 0.0f;               /* sensor.clj:158-159 */
  float pz4 = o < range_max ? p.z_rand / range_max : 0.0f;   /* sensor.clj:160-161 */
  mix4[2] = pz3 + pz4;
  mix4[3] = 0.0f;
  if (p.s7_log_weights && !(o <= range_max && o >= -3.0e38f)) {
    /* Log-weights only (the north-star aggregation; no reference counterpart): a return beyond
     * rangeMax, an inf or a NaN — common in ROS LaserScans — has probability 0 under every term
     * of sensor.clj:151-161, and log 0 would wipe out every particle alike.  Such a beam is left
     * out: z = +inf makes pz1 = pz2 = 0 and the constant term 1 makes the beam's factor exactly 1.
     * The reference's sum of cubes (S7 off) just adds 0 for it and keeps doing so here. */
    mix4[0] = INFINITY;
    mix4[1] = 0.0f;
    mix4[2] = 1.0f;
  }
}

int stage_scan_impl(amclj_ctx *ctx, const float *ranges, int n, float angle_min, float angle_inc,
                    float range_max) {
  NEED(n > 0 && n <= (1 << 20), AMCLJ_ERR_INVALID, "scan: n = %d out of range", n);
  NEED(ctx->map.df4, AMCLJ_ERR_STATE, "scan: no map set");
  const amclj_params &p = ctx->p;
  float R = p.s2_use_scan_range_max ? range_max : p.max_range; /* S2 */
  NEED(isfinite(R) && isfinite(range_max) && isfinite(angle_min) && isfinite(angle_inc),
       AMCLJ_ERR_INVALID, "scan: non-finite geometry");
  double rcells = fabs((double)R) / (double)ctx->map.res;
  NEED(rcells < 4.0e6, AMCLJ_ERR_INVALID, "scan: ray of %.0f cells is too long", rcells);
  int step = beam_step(n, p.s6_max_beams);
  int nb = beams_used(n, p.s6_max_beams);
  /* host staging in pinned memory, guarded by an event, so the upload needs no synchronisation */
  size_t tab_bytes = sizeof(float) * (size_t)nb * 6;
  const size_t dbl_off = (tab_bytes + 15) & ~(size_t)15, dbl_bytes = sizeof(double) * (size_t)nb * 2;
  if (dbl_off + dbl_bytes > ctx->pin_tab_bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->pin_tab) cudaFreeHost(ctx->pin_tab);
    ctx->pin_tab = nullptr;
    CK(cudaMallocHost(&ctx->pin_tab, dbl_off + dbl_bytes));
    ctx->pin_tab_bytes = dbl_off + dbl_bytes;
    ctx->beams.dirs_valid = false; /* a fresh staging buffer holds no directions */
  }
  CK(cudaEventSynchronize(ctx->ev[4])); /* the previous upload from this buffer has finished */
  /* float4 (obs, e2, p34, 0) first so that it stays 16-byte aligned for any beam count | float2 (cb, sb) */
  float *mix = reinterpret_cast<float *>(ctx->pin_tab), *dir = mix + 4 * (size_t)nb;
  double *dird = reinterpret_cast<double *>(reinterpret_cast<char *>(ctx->pin_tab) + dbl_off); /* near-tie refinement */
  /* the beam directions depend on the scan geometry only: a laser sends the same one with every message, so the
   * 2 x nb cos / sin calls and the upload of the double table happen when it changes (the pinned staging buffer
   * and the device tables still hold the previous directions otherwise) */
  BeamTable &bt = ctx->beams;
  const bool same_dirs = bt.dirs_valid && bt.dirs_pin == ctx->pin_tab && bt.dirs_n == n && bt.dirs_step == step &&
                         bt.dirs_nb == nb && memcmp(&bt.dirs_amin, &angle_min, 4) == 0 &&
                         memcmp(&bt.dirs_ainc, &angle_inc, 4) == 0 && nb <= bt.cap;
  int j = 0;
  for (int i = 0; i < n; i += step, j++) {
    if (!same_dirs) {
      /* sensor.clj:93-97 bearing: float32 message fields, double arithmetic */
      double a = (double)angle_min + (double)i * (double)angle_inc;
      dird[2 * j] = cos(a);
      dird[2 * j + 1] = sin(a);
      dir[2 * j] = (float)dird[2 * j];
      dir[2 * j + 1] = (float)dird[2 * j + 1];
    }
    fill_mix(p, ranges ? ranges[i] : 0.0f, range_max, mix + 4 * j);
  }
  BeamTable &b = ctx->beams;
  if (nb > b.cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(b.dev);
    dfree(b.dev_d);
    CK(dalloc(&b.dev, (size_t)nb * 6));
    CK(dalloc(&b.dev_d, (size_t)nb * 2));
    b.cap = nb;
  }
  if (same_dirs) { /* only the per-message part: (obs, e2, p34, 0) per beam */
    CK(cudaMemcpyAsync(b.dev, ctx->pin_tab, sizeof(float) * (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CK(cudaMemcpyAsync(b.dev, ctx->pin_tab, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.dev_d, dird, dbl_bytes, cudaMemcpyHostToDevice, ctx->stream));
    b.dirs_valid = true; b.dirs_pin = ctx->pin_tab; b.dirs_n = n; b.dirs_step = step; b.dirs_nb = nb;
    b.dirs_amin = angle_min; b.dirs_ainc = angle_inc;
  }
  b.cr = make_cell_refine(R, ctx->map.res, ctx->exact_cells);
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  b.nb = nb;
  b.n_scan = n;
  b.angle_min = angle_min;
  b.angle_inc = angle_inc;
  b.range_max = range_max;
  /* magic reciprocals for the minor-axis step count jd = floor(t / D), D = 2*dx, t < (dcap+1)*D
   * (dcap = largest jump the field can return).  Shift-free form M = ceil(2^32 / D) is exact
   * while (dcap+1) * D^2 <= 2^32; longer rays use M = ceil(2^(31+L) / D) >> (L-1), L = floor(log2 D). */
  int ndiv = (int)ceil(rcells) + 8;
  int mode = df_mode_of(ctx);
  NEED(mode >= 0, AMCLJ_ERR_INVALID, "scan: df_mode %d unavailable for this map", ctx->p.df_mode);
  /* the directional planes (also when the auto policy may pick them) and the byte field jump up to 255 */
  int dcap = (mode == DF_GLOBAL8 || mode == DF_WEDGE8 || ((auto_mode(ctx) || ctx->p.df_mode == 7) && ctx->map.wedge)) ? 255 : 15;
  uint64_t Dmax = 2ull * (uint64_t)(ndiv - 1);
  int wide = (uint64_t)(dcap + 1) * Dmax * Dmax > (1ull << 32);
  DivTable &d = ctx->divs;
  if (ndiv > d.n || wide != d.wide || dcap != d.dcap) {
    std::vector<uint2> t((size_t)ndiv);
    t[0] = make_uint2(0u, 0u);
    for (int dx = 1; dx < ndiv; dx++) {
      uint64_t D = 2ull * (uint64_t)dx;
      int L = 63 - __builtin_clzll(D);
      if (wide) t[dx] = make_uint2((uint32_t)(((1ull << (31 + L)) + D - 1) / D), (uint32_t)(L - 1));
      else t[dx] = make_uint2((uint32_t)(((1ull << 32) + D - 1) / D), 0u);
    }
    if (ndiv > d.cap) {
      dfree(d.dev);
      CK(dalloc(&d.dev, (size_t)ndiv));
      d.cap = ndiv;
    }
    CK(cudaMemcpyAsync(d.dev, t.data(), sizeof(uint2) * t.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    d.n = ndiv;
    d.wide = wide;
    d.dcap = dcap;
  }
  int rrc = stage_ring(ctx, rcells);
  if (rrc) return rrc;
  /* line set-up records of the walk kernel: one per ring cell, for the field placement it will run on */
  RayTables &rt = ctx->rt;
  rt.floor_trick_ok = (double)(ctx->map.W > ctx->map.H ? ctx->map.W : ctx->map.H) + rcells + 4.0 < 4.0e6;
  if (rt.have_ring && ctx->ring_setup && rt.n_entries <= 8192 && rt.floor_trick_ok) {
    int walk_mode = (auto_mode(ctx) && ctx->map.wedge) ? DF_WEDGE8 : mode;
    /* what the records depend on; rebuilt only when any of it changed */
    const long long key[6] = {(long long)walk_mode, (long long)p.s3_proper_endpoint, (long long)p.s4_endpoint_termination,
                              (long long)ctx->map_gen, (long long)d.wide * 1000 + d.dcap, (long long)d.n};
    if (rt.recs_mode != walk_mode || memcmp(key, rt.recs_key, sizeof(key)) != 0 || rt.recs_r != rcells) {
      rt.recs_mode = -1;
      if (2 * rt.n_entries > rt.recs_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        dfree(rt.recs);
        rt.recs_cap = 0;
        CK(dalloc(&rt.recs, (size_t)2 * rt.n_entries));
        rt.recs_cap = 2 * rt.n_entries;
      }
      b.staged = true; /* fill_sensor_args reads the staged tables */
      SensorLaunch a;
      fill_sensor_args(ctx, a, walk_mode);
      CK(launch_ring_records(a, walk_mode, rt.ents, rt.n_entries, rt.recs, ctx->stream));
      memcpy(rt.recs_key, key, sizeof(key));
      rt.recs_r = rcells;
      rt.recs_mode = walk_mode;
    }
  } else {
    rt.recs_mode = -1;
  }
  b.staged = true;
  return pt_prepare(ctx, rcells);
}

/* Is this scan's geometry (beam count, angles, range_max) the one that is staged?  set_map / set_params clear
 * `staged`, so everything else the staging derived (directions, ring, dividers, ray tables) still holds. */
bool scan_geometry_staged(const amclj_ctx *ctx, int n, float angle_min, float angle_inc, float range_max) {
  const BeamTable &b = ctx->beams;
  if (!b.staged || !b.dirs_valid || b.dirs_pin != ctx->pin_tab || n <= 0) return false;
  const int step = beam_step(n, ctx->p.s6_max_beams), nb = beams_used(n, ctx->p.s6_max_beams);
  return b.dirs_n == n && b.dirs_step == step && b.dirs_nb == nb && b.nb == nb && nb <= b.cap &&
         memcmp(&b.dirs_amin, &angle_min, 4) == 0 && memcmp(&b.dirs_ainc, &angle_inc, 4) == 0 &&
         memcmp(&b.range_max, &range_max, 4) == 0;
}

/* The per-message part of a scan whose geometry is staged already: the nb (obs, e2, p34, 0) entries, built in the
 * pinned staging buffer and uploaded on the ctx stream.  Same values as stage_scan_impl writes. */
int restage_ranges(amclj_ctx *ctx, const float *ranges, int n, float range_max) {
  const amclj_params &p = ctx->p;
  const int step = beam_step(n, p.s6_max_beams), nb = ctx->beams.nb;
  CK(cudaEventSynchronize(ctx->ev[4])); /* the previous upload from this buffer has finished */
  float *mix = reinterpret_cast<float *>(ctx->pin_tab);
  int j = 0;
  for (int i = 0; i < n; i += step, j++) fill_mix(p, ranges ? ranges[i] : 0.0f, range_max, mix + 4 * j);
  CK(cudaMemcpyAsync(ctx->beams.dev, ctx->pin_tab, sizeof(float) * (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  ctx->beams.n_scan = n;
  return AMCLJ_OK;
}

void fill_sensor_args(const amclj_ctx *ctx, SensorLaunch &a, int mode) {
  const amclj_params &p = ctx->p;
  const DeviceMap &m = ctx->map;
  memset(&a, 0, sizeof(a));
  a.W = m.W;
  a.H = m.H;
  a.res = m.res;
  if (mode == DF_WEDGE8) {
    a.df = m.wedge;
    a.row_nib = m.pitch8;
    a.df_bytes = m.wedge_plane * WEDGE_CLASSES;
    a.wedge_plane = (int32_t)m.wedge_plane;
  } else if (mode == DF_GLOBAL8) {
    a.df = m.df8;
    a.row_nib = m.pitch8;
    a.df_bytes = m.bytes8;
  } else {
    a.df = m.df4;
    a.row_nib = m.row_bytes4 * 2;
    a.df_bytes = m.bytes4;
  }
  a.beam = ctx->beams.dev;
  a.beam_d = ctx->beams.dev_d;
  a.cr = ctx->beams.cr;
  a.rinv_d = 1.0 / (double)ctx->map.res;
  a.nb = ctx->beams.nb;
  a.divtab = ctx->divs.dev;
  a.ndiv = ctx->divs.n;
  a.div_wide = ctx->divs.wide;
  a.refill = mode == DF_WEDGE8 ? ctx->refill_wedge : (mode == DF_SMEM4 ? ctx->refill : ctx->refill_global);
  a.chunks = 1;
  a.R = p.s2_use_scan_range_max ? ctx->beams.range_max : p.max_range;
  a.miss_value = a.R; /* sensor.clj:121 returns max-range on a miss */
  a.s1 = p.s1_use_heading;
  a.s3 = p.s3_proper_endpoint;
  a.s4 = p.s4_endpoint_termination;
  a.s7 = p.s7_log_weights;
  a.s10 = p.s10_apply_laser_offset;
  a.z_hit = p.z_hit;
  float sig = p.s5_use_sigma_hit ? p.sigma_hit : p.z_hit; /* S5, sensor.clj:151-153 */
  a.inv2s2 = 1.0f / (2.0f * sig * sig);
  a.k2 = (float)((double)a.inv2s2 * 1.4426950408889634);
  a.laser_x = p.laser_x;
  a.laser_y = p.laser_y;
  a.walk_ctr = ctx->walk_ctr;
  a.walk_tail_pct = (mode == DF_GLOBAL8 || mode == DF_GLOBAL4) ? ctx->walk_tail_pct_global : ctx->walk_tail_pct;
  const RayTables &rt = ctx->rt;
  if (rt.have_ring && rt.recs_mode == mode && rt.recs) {
    a.ring_rows = rt.rows;
    a.ring_recs = rt.recs;
    a.ring_nrows = rt.nrows;
    a.ring_E = rt.E;
    a.ring_entries = rt.n_entries;
    a.ring_ok = rt.floor_trick_ok ? 1 : 0;
  }
}

/* enqueue motion + sensor on the ctx stream */
int enqueue_motion(amclj_ctx *ctx, const double curr[3], const double prev[3], bool injected,
                   uint64_t seed) {
  MotionLaunch m;
  memset(&m, 0, sizeof(m));
  make_control(ctx->p, curr, prev, m);
  m.x = ctx->x;
  m.y = ctx->y;
  m.th = ctx->th;
  m.n = ctx->n;
  m.gid0 = ctx->global_offset;
  m.normals = injected ? ctx->normals : nullptr;
  m.seed = seed;
  m.bbox = nullptr;
  if (ctx->ctrl_fresh && !m.is_static &&
      ((auto_mode(ctx) && (ctx->map.wedge || ctx->order_large_maps || ctx->rt_enable)) || ctx->p.df_mode == 6)) {
    m.bbox = ctx->bbox; /* the moved cloud's bounding box comes for free */
    ctx->bbox_done = true;
  }
  /* the persistent-table sort's histogram pass rides along when the sensor pass of this update will run on
   * the tables (same rounding of the same moved pose: ptable.cu pt_hist_kernel) */
  ctx->pt.hist_done = false;
  if (ctx->pt.fuse_motion && ctx->pt.built && ctx->ctrl_fresh && !m.is_static && ctx->beams.staged &&
      (ctx->p.df_mode == 7 || (ctx->p.df_mode == 0 && !(ctx->p.collect_stats & 1)))) {
    int rc = ensure_pt_scratch(ctx, ctx->n);
    if (rc) return rc;
    const PersistentTables &t = ctx->pt;
    m.hist_slot_of_cell = t.slot_of_cell;
    m.hist_W = ctx->map.W; m.hist_H = ctx->map.H; m.hist_n_slots = t.n_slots;
    m.hist_s10 = ctx->p.s10_apply_laser_offset;
    m.hist_res = ctx->map.res; m.hist_laser_x = ctx->p.laser_x; m.hist_laser_y = ctx->p.laser_y;
    m.hist_bin_count = t.bin_count; m.hist_pkey = t.pkey; m.hist_qacc = ctx->partials;
    m.hist_scan_state = t.scan_state; m.hist_scan_words = t.scan_words;
    ctx->pt.hist_done = true;
  }
  CK(launch_motion(m, ctx->stream));
  ctx->launches += 1;
  return AMCLJ_OK;
}

/* clear the control block: everything (with_max) or only what the scan accumulates into */
cudaError_t reset_ctrl(amclj_ctx *ctx, bool with_max) {
  /* layout: [0..1] particle bounding box | [2] max | [3] total | [4] tile counter | [5..] tiles */
  size_t words = (size_t)scan_tiles(ctx->n) + 2;
  ctx->scan_clean = true;
  if (with_max) return cudaMemsetAsync(ctx->ctrl, 0, sizeof(uint64_t) * (words + 3), ctx->stream);
  return cudaMemsetAsync(ctx->ctrl + 3, 0, sizeof(uint64_t) * words, ctx->stream);
}

/* the ray-table path over the particles of `base` (raytable.cu); results land in base.raw */
int enqueue_raytable(amclj_ctx *ctx, const SensorLaunch &base, bool diag, int must_run) {
  const bool forced = must_run == 2;
  int fill_mode = ctx->map.wedge ? DF_WEDGE8 : (df_mode_of(ctx) == DF_GLOBAL8 ? DF_GLOBAL8 : DF_GLOBAL4);
  RtLaunch L;
  memset(&L, 0, sizeof(L));
  fill_sensor_args(ctx, L.s, fill_mode);
  L.s.x = base.x; L.s.y = base.y; L.s.th = base.th; L.s.raw = base.raw; L.s.n = base.n;
  L.s.rawmax_ord = base.rawmax_ord; L.s.stats = base.stats;
  L.s.bbox = base.bbox; L.s.compact_m = base.compact_m; L.s.decision = base.decision;
  L.s.dbg_first = base.dbg_first; L.s.dbg_count = base.dbg_count;
  L.s.dbg_hit_x = base.dbg_hit_x; L.s.dbg_hit_y = base.dbg_hit_y; L.s.dbg_steps = base.dbg_steps;
  L.s.dbg_hit = base.dbg_hit; L.s.dbg_range = base.dbg_range;
  const RayTables &t = ctx->rt;
  L.thr = rt_threshold(ctx, forced);
  int rc = ensure_rt_scratch(ctx, base.n, diag, L.thr, &L.max_slots, &L.max_items);
  if (rc) return rc;
  L.rows = t.rows; L.ents = t.ents; L.nrows = t.nrows; L.E = t.E; L.n_entries = t.n_entries;
  L.bins = t.bins; L.slots = t.slots; L.item0 = t.item0; L.tables = t.tables; L.ctl = t.ctl;
  if (base.n > ctx->partials_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx->partials);
    CK(dalloc(&ctx->partials, (size_t)base.n));
    ctx->partials_cap = base.n;
  }
  L.qacc = ctx->partials;
  L.perm = ctx->perm;
  L.pstate = t.pstate;
  L.must_run = must_run;
  L.tail_pct = t.tail_pct;
  CK(launch_raytable(ctx, L, fill_mode, diag, ctx->stream));
  ctx->launches += 4; /* histogram + plan, scatter + fill, look-up, finish */
  return AMCLJ_OK;
}

/* decode_max: also materialise the max as a float (AMCLJ_BUF_RAW_MAX) for a cross-rank all-reduce */
int enqueue_sensor(amclj_ctx *ctx, bool decode_max) {
  NEED(ctx->beams.staged, AMCLJ_ERR_STATE, "sensor: no scan staged");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "sensor: no particles");
  int mode = df_mode_of(ctx);
  NEED(mode >= 0, AMCLJ_ERR_INVALID, "sensor: df_mode %d unavailable for this map", ctx->p.df_mode);
  SensorLaunch a;
  fill_sensor_args(ctx, a, mode);
  a.x = ctx->x;
  a.y = ctx->y;
  a.th = ctx->th;
  a.raw = ctx->raw;
  a.n = ctx->n;
  a.rawmax_ord = ctx->rawmax_ord;
  a.chunks = ctx->force_chunks > 0 ? ctx->force_chunks : sensor_chunks(ctx, ctx->n, a.nb);
  if (a.chunks > a.nb) a.chunks = a.nb;
  if (a.chunks > 1) {
    int64_t need = (int64_t)a.chunks * ctx->n;
    if (need > ctx->partials_cap) {
      CK(cudaStreamSynchronize(ctx->stream));
      dfree(ctx->partials);
      CK(dalloc(&ctx->partials, (size_t)need));
      ctx->partials_cap = need;
    }
    a.partials = ctx->partials;
  }
  bool diag = (ctx->p.collect_stats & 1) != 0;
  ctx->time_main = ctx->time_main_env || (ctx->p.collect_stats & 2) != 0;
  a.stats = diag ? ctx->stats : nullptr;
  if (!ctx->ctrl_fresh) {
    CK(reset_ctrl(ctx, true));
    ctx->bbox_done = false;
  }
  ctx->ctrl_fresh = false;
  if (diag) {
    CK(cudaMemsetAsync(ctx->stats, 0, 7 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->stats + 8, 0, 4 * sizeof(unsigned long long), ctx->stream));
  }
  /* Persistent map-wide ray tables (ptable.cu): when the staged scan has them (df_mode 0 on a map small
   * enough for one table per free cell, or df_mode 7), every cloud runs on them — no policy, nothing the
   * host has to read back.  Diagnostic passes (cell / probe counters) keep the walking paths, which visit
   * the cells the counters count. */
  NEED(ctx->p.df_mode != 7 || ctx->pt.built, AMCLJ_ERR_INVALID,
       "sensor: df_mode 7 (persistent ray tables) unavailable for this map / rays of %.0f cells", ctx->rt.r_cells);
  if (ctx->pt.built && (ctx->p.df_mode == 7 || !diag)) {
    int prc = enqueue_ptable(ctx, a, false);
    if (prc) return prc;
    if (decode_max) {
      CK(launch_decode_max(ctx->rawmax_ord, ctx->rawmax, ctx->stream));
      ctx->launches += 1;
    }
    ctx->bbox_done = false;
    ctx->max_in_ord = !decode_max;
    ctx->last_diag = diag;
    ctx->have_raw = true;
    ctx->have_cdf = false;
    ctx->st.map_in_smem = 0;
    ctx->st.beams_used = ctx->beams.nb;
    ctx->st.n_particles = ctx->n;
    *ctx->pin_decision = 3;
    return AMCLJ_OK;
  }
  /* df_mode 0 on a map that has the directional planes: they are used for every cloud.  A compact
   * cloud is local as it is; a spread-out one is walked in map-tile order so that each CTA's probes
   * share an L1-sized neighbourhood.  A 5 us bounding-box kernel leaves the decision on the device
   * (the tile-order kernels return at once for a compact cloud, and the sensor kernel ignores the
   * permutation); either way the results are identical. */
  bool directional = auto_mode(ctx) && ctx->map.wedge != nullptr;
  /* large maps (byte field through L2) get the tile-order walk too: it is what keeps a CTA's
   * probes in L1 */
  bool ordered = directional || (auto_mode(ctx) && mode == DF_GLOBAL8 && ctx->order_large_maps);
  /* dense clouds (pose tracking: thousands of particles per start cell) go through per-cell ray
   * tables (raytable.cu).  Whether the cloud is dense is decided on the device, every update; the
   * verdict of the previous update, left in pinned host memory, only picks which kernels are
   * launched at all: 0 spread out -> walk kernel only; 1 compact / -1 unknown -> both, the device flag
   * picks one; 2 dense -> ray tables only (they take any cloud, a sparse one slowly). */
  const bool rt_forced = ctx->p.df_mode == 6;
  const bool rt_auto = auto_mode(ctx) && ctx->rt_enable && ctx->rt.available && !ctx->p.s10_apply_laser_offset;
  NEED(!rt_forced || ctx->rt.available, AMCLJ_ERR_INVALID, "sensor: df_mode 6 (ray tables) unavailable for rays of %.0f cells",
       ctx->rt.r_cells);
  const int hint = *ctx->pin_decision;
  const bool use_rt = rt_forced || (rt_auto && hint != 0);
  const bool only_rt = rt_forced || (rt_auto && hint == 2);
  if (directional) {
    mode = DF_WEDGE8;
    SensorLaunch b;
    fill_sensor_args(ctx, b, DF_WEDGE8);
    b.x = a.x; b.y = a.y; b.th = a.th; b.raw = a.raw; b.n = a.n; b.rawmax_ord = a.rawmax_ord;
    b.chunks = a.chunks; b.partials = a.partials; b.stats = a.stats;
    a = b;
  }
  /* the per-ring-cell set-up records pay when neighbouring lanes read neighbouring records (a compact
   * cloud: C2 on the walk kernel 0.411 -> 0.373 ms); a spread-out cloud scatters the 48 bytes per ray
   * over the banks (C3: 3.95 -> 4.00 ms), so it keeps computing its set-up */
  if (hint == 0 || ctx->p.df_mode == 5) a.ring_ok = 0;
  /* task distribution: a spread-out cloud walked in tile order gains from CTA-contiguous runs drawn
   * through a shared counter (C3: 3.95 -> 3.82 ms); a compact cloud split into beam chunks does not
   * (C2 on the walk kernel: 0.372 -> 0.381 ms) and keeps the fixed interleaved shares */
  if (!(hint == 0 || ctx->p.df_mode == 5)) a.walk_tail_pct = -1;
  if (ordered || use_rt) {
    if (!ctx->bbox_done) {
      CK(launch_bbox(ctx->x, ctx->y, ctx->n, ctx->bbox, ctx->sm_count, ctx->stream));
      ctx->launches += 1;
    }
    ctx->bbox_done = false;
    a.bbox = ctx->bbox;
    a.compact_m = ctx->compact_cells * ctx->map.res;
    a.decision = const_cast<int32_t *>(ctx->pin_decision); /* pinned host memory: the kernel stores straight into it */
  }
  if (ordered && !only_rt) {
    if (!ctx->tile_counters) CK(dalloc(&ctx->tile_counters, (size_t)tile_count(ctx->map.W, ctx->map.H) + 1));
    /* the previous update's verdict is only a hint: when the cloud was compact the tile-order
     * kernels are not even launched (a wrong guess costs speed for one update, never a result) */
    if (hint != 1) {
      CK(launch_tile_order(ctx->x, ctx->y, ctx->th, ctx->n, ctx->map.res, ctx->map.W, ctx->map.H,
                           ctx->tile_counters, ctx->perm, ctx->bbox, a.compact_m,
                           ctx->tile_shift > 0 ? ctx->tile_shift : tile_order_shift(ctx->n, ctx->map.free_cells), ctx->stream));
      ctx->launches += 3;
      a.perm = ctx->perm;
    }
  }
  if (use_rt) {
    int rrc = enqueue_raytable(ctx, a, diag, rt_forced ? 2 : (only_rt ? 1 : 0));
    if (rrc) return rrc;
    a.skip_flag = &ctx->rt.ctl->dense;
  }
  if (!only_rt) {
    CK(launch_sensor(ctx, a, mode, diag, ctx->stream));
    ctx->launches += a.chunks > 1 ? 2 : 1;
  }
  if (decode_max) {
    CK(launch_decode_max(ctx->rawmax_ord, ctx->rawmax, ctx->stream));
    ctx->launches += 1;
  }
  ctx->max_in_ord = !decode_max;
  ctx->last_diag = diag;
  ctx->have_raw = true;
  ctx->have_cdf = false;
  ctx->st.map_in_smem = mode == DF_SMEM4 && !only_rt;
  ctx->st.beams_used = ctx->beams.nb;
  ctx->st.n_particles = ctx->n;
  return AMCLJ_OK;
}

/* raw (or w) -> max -> fixed point -> CDF, all on the stream */
int enqueue_cdf(amclj_ctx *ctx, bool from_raw, bool max_known) {
  const float *v = from_raw ? ctx->raw : ctx->w;
  int is_log = from_raw && ctx->p.s7_log_weights;
  if (!max_known) { /* launch_max clears and refills the max (ordered uint and float) */
    CK(launch_max(v, ctx->n, ctx->rawmax_ord, ctx->rawmax, ctx->sm_count, ctx->stream));
    ctx->launches += 2;
    ctx->max_in_ord = false;
  }
  if (!ctx->scan_clean || !max_known) CK(reset_ctrl(ctx, false));
  CK(launch_scan(v, ctx->rawmax, ctx->max_in_ord ? ctx->rawmax_ord : nullptr, is_log, ctx->n, ctx->cdf,
                 ctx->scan_state, ctx->scan_counter, ctx->total, ctx->stream));
  ctx->launches += 1;
  ctx->scan_clean = false;
  ctx->have_cdf = true;
  ctx->cdf_from_raw = from_raw;
  return AMCLJ_OK;
}

void swap_particles(amclj_ctx *ctx) {
  std::swap(ctx->x, ctx->x2);
  std::swap(ctx->y, ctx->y2);
  std::swap(ctx->th, ctx->th2);
  std::swap(ctx->w, ctx->w2);
  ctx->flip ^= 1; /* peers that mapped both buffer sets follow the alternation */
}

int enqueue_systematic(amclj_ctx *ctx, double u0) {
  GatherArgs g;
  memset(&g, 0, sizeof(g));
  g.cdf = ctx->cdf;
  g.n_local = ctx->n;
  g.cdf_offset = 0;
  g.plan = nullptr; /* derived in the kernel from the device-resident total */
  g.total_ptr = ctx->total;
  g.u0_bits = (uint64_t)ldexp(u0, 64); /* u0 in [0,1) as a 0.64 fraction */
  g.n_global = ctx->n;
  g.m_lo = 0;
  g.count = ctx->n;
  g.x = ctx->x;
  g.y = ctx->y;
  g.th = ctx->th;
  g.ox = ctx->x2;
  g.oy = ctx->y2;
  g.oth = ctx->th2;
  g.ow = ctx->w2;
  g.oidx = ctx->idx;
  g.wout = 1.0f / (float)ctx->n;
  g.zero_flag = ctx->stats + 12;
  CK(launch_systematic(g, ctx->stream));
  ctx->launches += 1;
  swap_particles(ctx);
  ctx->have_raw = false;
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int read_scalars(amclj_ctx *ctx, uint64_t *total, float *wmax) {
  uint64_t *pt = reinterpret_cast<uint64_t *>(ctx->pinned);
  float *pm = reinterpret_cast<float *>(pt + 1);
  CK(cudaMemcpyAsync(pt, ctx->total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(pm, ctx->rawmax, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->last_total = *pt;
  ctx->last_wmax = *pm;
  if (total) *total = *pt;
  if (wmax) *wmax = *pm;
  return AMCLJ_OK;
}

int run_roulette(amclj_ctx *ctx, const int32_t *h_sidx, const uint32_t *h_su, int64_t stream_len,
                 uint64_t seed) {
  int rc = read_scalars(ctx, nullptr, nullptr);
  if (rc) return rc;
  uint64_t T = ctx->last_total;
  NEED(T > 0, AMCLJ_ERR_STATE, "roulette: all weights are zero");
  int64_t n = ctx->n;
  int32_t *d_sidx = nullptr;
  uint32_t *d_su = nullptr;
  if (h_sidx) {
    NEED(h_su && stream_len > 0, AMCLJ_ERR_INVALID, "roulette: incomplete injected stream");
    CK(dalloc(&d_sidx, (size_t)stream_len));
    CK(dalloc(&d_su, (size_t)stream_len));
    CK(cudaMemcpyAsync(d_sidx, h_sidx, sizeof(int32_t) * (size_t)stream_len, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_su, h_su, sizeof(uint32_t) * (size_t)stream_len, cudaMemcpyHostToDevice, ctx->stream));
  }
  unsigned long long *cnt = ctx->stats + 4; /* accepted, last attempt */
  CK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), ctx->stream));
  int64_t got = 0, t0 = 0;
  int status = AMCLJ_OK;
  unsigned long long *h = reinterpret_cast<unsigned long long *>(ctx->pinned);
  while (got < n) {
    /* expected attempts per acceptance = n (pf.clj:161-166): size the super-chunk for what is
     * still missing, within [2^16, 2^30] */
    double want = (double)(n - got) * (double)n * 1.25 + 65536.0;
    int64_t len = want > 1073741824.0 ? 1073741824ll : (int64_t)want;
    if (h_sidx) {
      len = stream_len - t0;
      if (len <= 0) { status = fail(ctx, AMCLJ_ERR_INVALID, "roulette: injected stream exhausted after %lld of %lld", (long long)got, (long long)n); break; }
    }
    int per_thread = len >= (1 << 24) ? 64 : (len >= (1 << 18) ? 16 : 4);
    int64_t tiles = (len + 256ll * per_thread - 1) / (256ll * per_thread);
    if (tiles > ctx->scan_state_cap) { /* cap the chunk to the descriptor array */
      tiles = ctx->scan_state_cap;
      len = tiles * 256ll * per_thread;
    }
    RouletteArgs a;
    memset(&a, 0, sizeof(a));
    a.cdf = ctx->cdf;
    a.n = n;
    a.total = T;
    a.seed = seed;
    a.sidx = d_sidx;
    a.su = d_su;
    a.t0 = t0;
    a.t_end = t0 + len;
    a.per_thread = per_thread;
    a.got_before = got;
    a.idx_out = ctx->idx;
    a.tile_state = ctx->scan_state;
    a.tile_counter = ctx->scan_counter;
    a.accepted_out = cnt;
    a.last_attempt = cnt + 1;
    ctx->scan_clean = false;
    cudaError_t e = launch_roulette(a, tiles, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { status = fail(ctx, AMCLJ_ERR_CUDA, "roulette: %s", cudaGetErrorString(e)); break; }
    got += (int64_t)h[0];
    t0 += len;
  }
  dfree(d_sidx);
  dfree(d_su);
  if (status) return status;
  /* carried weight = the normalised weight (pf.clj:160) */
  CK(launch_gather_idx(ctx->idx, n, ctx->x, ctx->y, ctx->th, 1, ctx->cdf, T, nullptr, 0.f, ctx->x2,
                       ctx->y2, ctx->th2, ctx->w2, ctx->stream));
  swap_particles(ctx);
  ctx->have_raw = false;
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int run_tournament(amclj_ctx *ctx, const int32_t *h_sidx, int64_t stream_len, uint64_t seed) {
  int64_t n = ctx->n;
  const float *wv = ctx->have_raw ? ctx->raw : ctx->w;
  int32_t *d_sidx = nullptr;
  if (h_sidx) {
    NEED(stream_len >= 2 * n, AMCLJ_ERR_INVALID, "tournament: stream needs 2n entries");
    CK(dalloc(&d_sidx, (size_t)(2 * n)));
    CK(cudaMemcpyAsync(d_sidx, h_sidx, sizeof(int32_t) * (size_t)(2 * n), cudaMemcpyHostToDevice, ctx->stream));
  }
  cudaError_t e = launch_tournament(wv, n, d_sidx, seed, ctx->idx, ctx->stream);
  if (e == cudaSuccess)
    e = launch_gather_idx(ctx->idx, n, ctx->x, ctx->y, ctx->th, 2, nullptr, 0, wv, 0.f, ctx->x2, ctx->y2,
                          ctx->th2, ctx->w2, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dfree(d_sidx);
  CK(e);
  swap_particles(ctx);
  ctx->have_raw = false;
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int upload_normals(amclj_ctx *ctx, const float *normals) {
  if (!normals) return AMCLJ_OK;
  CK(cudaMemcpyAsync(ctx->normals, normals, sizeof(float) * 3 * (size_t)ctx->n, cudaMemcpyHostToDevice,
                     ctx->stream));
  return AMCLJ_OK;
}


/* after a synchronisation: did an update since the last check find every weight zero?  (the device
 * kept the particles as they were; the caller hears about it) */
int check_zero_total(amclj_ctx *ctx, const char *who) {
  unsigned long long *h = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(ctx->pinned) + 1024);
  CK(cudaMemcpyAsync(h, ctx->stats + 12, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (h[1] != ctx->mbox_errors_seen) {
    ctx->mbox_errors_seen = h[1];
    return fail(ctx, AMCLJ_ERR_STATE, "%s: a rank never published its share of the exchange (peers not all updating?)", who);
  }
  const uint64_t seen = *h;
  const bool fresh = seen != ctx->zero_seen;
  ctx->zero_seen = seen;
  ctx->st.zero_weight_updates = seen;
  if (fresh)
    return fail(ctx, AMCLJ_ERR_STATE, "%s: all weights are zero (a zero-probability scan?): particles kept, not resampled", who);
  return AMCLJ_OK;
}

void stamp(amclj_ctx *ctx, int k) {
  cudaEventRecord(ctx->ev[k], ctx->stream);
  if (k == 0) {
    ctx->launches = 0; /* a new update */
    ctx->stages_timed = true;
  }
  ctx->st.kernels_launched = ctx->launches;
}

void collect_times(amclj_ctx *ctx) {
  float a = 0, b = 0, c = 0, t = 0;
  if (cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]) != cudaSuccess || a < 0) a = 0;
  if (cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]) != cudaSuccess || b < 0) b = 0;
  if (cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]) != cudaSuccess || c < 0) c = 0;
  if (cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[3]) != cudaSuccess || t < 0) t = 0;
  (void)cudaGetLastError();
  if (!ctx->stages_timed) a = b = c = 0; /* the fused update ran without its inner stamps */
  ctx->st.ms_motion = a;
  ctx->st.ms_sensor = b;
  ctx->st.ms_resample = c;
  ctx->st.ms_total = t;
  float m = 0;
  if (cudaEventElapsedTime(&m, ctx->ev[5], ctx->ev[6]) != cudaSuccess || m < 0) m = 0;
  (void)cudaGetLastError();
  ctx->st.ms_sensor_main = m;
  ctx->st.sensor_path = *ctx->pin_decision;
}

}  // namespace

extern "C" {

int32_t amclj_abi_version(void) { return AMCLJ_ABI_VERSION; }

int32_t amclj_params_default(amclj_params *p, int32_t which) {
  if (!p || (which != AMCLJ_PRESET_REF && which != AMCLJ_PRESET_NS)) return AMCLJ_ERR_INVALID;
  preset(p, which);
  return AMCLJ_OK;
}

int32_t amclj_create(int32_t device, amclj_ctx **out) {
  if (!out) return AMCLJ_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) return AMCLJ_ERR_NODEVICE; /* no CPU fallback */
  if (device < 0 || device >= count) return AMCLJ_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return AMCLJ_ERR_NODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AMCLJ_ERR_NODEVICE;
  amclj_ctx *ctx = new amclj_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  preset(&ctx->p, AMCLJ_PRESET_REF);
  if (const char *e = getenv("AMCLJ_REFILL")) { /* tuning knob, 1..32 */
    int v = atoi(e);
    if (v >= 1 && v <= 32) ctx->refill = v;
  }
  if (const char *e = getenv("AMCLJ_CHUNKS")) ctx->force_chunks = atoi(e);
  if (const char *e = getenv("AMCLJ_ORDER_LARGE")) ctx->order_large_maps = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_WEDGE_BUDGET_MB")) ctx->wedge_budget = (size_t)atoll(e) << 20;
  if (const char *e = getenv("AMCLJ_REFILL_WEDGE")) ctx->refill_wedge = atoi(e);
  if (const char *e = getenv("AMCLJ_REFILL_GLOBAL")) ctx->refill_global = atoi(e);
  if (const char *e = getenv("AMCLJ_COMPACT_CELLS")) ctx->compact_cells = (float)atof(e);
  if (const char *e = getenv("AMCLJ_RAYTABLE")) ctx->rt_enable = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_EXACT_CELLS")) ctx->exact_cells = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_RING_SETUP")) ctx->ring_setup = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_TILE_SHIFT")) ctx->tile_shift = atoi(e);
  if (const char *e = getenv("AMCLJ_WALK_CARVEOUT")) ctx->walk_carveout = atoi(e);
  if (const char *e = getenv("AMCLJ_WALK_TAIL_PCT")) ctx->walk_tail_pct = ctx->walk_tail_pct_global = atoi(e) > 100 ? 100 : atoi(e);
  if (const char *e = getenv("AMCLJ_TIME_MAIN")) ctx->time_main_env = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_RT_THR")) ctx->rt.thr = atoi(e);
  if (const char *e = getenv("AMCLJ_RT_WAVES")) ctx->rt.waves = atoi(e);
  if (const char *e = getenv("AMCLJ_RT_TAIL_PCT")) ctx->rt.tail_pct = atoi(e) < 0 ? 0 : (atoi(e) > 100 ? 100 : atoi(e));
  if (const char *e = getenv("AMCLJ_RT_BUDGET_MB")) ctx->rt.budget = (size_t)atoll(e) << 20;
  if (const char *e = getenv("AMCLJ_PTABLE")) ctx->pt.enable = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_PT_BUDGET_MB")) ctx->pt.budget = (size_t)atoll(e) << 20;
  if (const char *e = getenv("AMCLJ_PT_WARPS")) ctx->pt.warps_req = atoi(e);
  if (const char *e = getenv("AMCLJ_PT_NBUF")) ctx->pt.nbuf = atoi(e);
  if (const char *e = getenv("AMCLJ_PT_FUSE_MOTION")) ctx->pt.fuse_motion = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_PT_TAIL_DIV")) ctx->pt.tail_div = atoi(e) < 1 ? 1 : atoi(e);
  if (const char *e = getenv("AMCLJ_PT_TAIL_MIN")) ctx->pt.tail_min = atoi(e) < 1 ? 1 : atoi(e);
  if (const char *e = getenv("AMCLJ_PT_TAIL_PCT")) ctx->pt.tail_pct = atoi(e) < 0 ? 0 : (atoi(e) > 100 ? 100 : atoi(e));
  if (const char *e = getenv("AMCLJ_PT_ROWS32")) ctx->pt.rows32 = atoi(e) != 0;
  if (const char *e = getenv("AMCLJ_PT_FUSE")) ctx->pt.fuse_finish = atoi(e) != 0;
  memset(&ctx->st, 0, sizeof(ctx->st));
  bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->stream = ctx->own_stream;
  for (int i = 0; ok && i < 7; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  for (int i = 0; ok && i < 2; i++) ok = cudaEventCreateWithFlags(&ctx->ev_pub[i], cudaEventDisableTiming) == cudaSuccess;
  /* control block: [0..1] particle bounding box (4 ordered-uint maxima) | [2] ordered-uint max
   * weight | [3] fixed-point total | [4] tile counter | [5..] tile descriptors of the scan (2^20
   * tiles = 2^31 particles; also serves the roulette compaction) */
  ctx->scan_state_cap = 1 << 20;
  ok = ok && dalloc(&ctx->ctrl, (size_t)ctx->scan_state_cap + 5) == cudaSuccess;
  if (ok) {
    ctx->bbox = reinterpret_cast<uint32_t *>(ctx->ctrl);
    ctx->rawmax_ord = reinterpret_cast<uint32_t *>(ctx->ctrl + 2);
    ctx->total = ctx->ctrl + 3;
    ctx->scan_counter = reinterpret_cast<uint32_t *>(ctx->ctrl + 4);
    ctx->scan_state = ctx->ctrl + 5;
    ok = cudaMemset(ctx->ctrl, 0, sizeof(uint64_t) * ((size_t)ctx->scan_state_cap + 5)) == cudaSuccess;
  }
  ok = ok && dalloc(&ctx->rawmax, 1) == cudaSuccess && dalloc(&ctx->stats, 16) == cudaSuccess &&
       dalloc(&ctx->plan, 1) == cudaSuccess && dalloc(&ctx->decision, 1) == cudaSuccess && dalloc(&ctx->walk_ctr, 2) == cudaSuccess && dalloc(&ctx->totals_all, MAX_WORLD) == cudaSuccess && dalloc(&ctx->mbox, (size_t)MBOX_WORDS) == cudaSuccess &&
       dalloc(&ctx->pose_part, (size_t)pose_blocks() * 4) == cudaSuccess &&
       cudaMallocHost(&ctx->pinned, 4096) == cudaSuccess;
  if (ok) {
    ctx->pin_decision = reinterpret_cast<volatile int32_t *>(reinterpret_cast<char *>(ctx->pinned) + 2048);
    *ctx->pin_decision = -1;
    ok = cudaMemset(ctx->decision, 0xff, sizeof(int32_t)) == cudaSuccess;
  }
  if (ok) ok = cudaMemset(ctx->stats, 0, 16 * sizeof(unsigned long long)) == cudaSuccess;
  if (ok) ok = cudaMemset(ctx->walk_ctr, 0, 2 * sizeof(uint32_t)) == cudaSuccess;
  if (ok) ok = cudaMemset(ctx->mbox, 0, sizeof(unsigned long long) * MBOX_WORDS) == cudaSuccess;
  if (!ok) {
    amclj_destroy(ctx);
    return AMCLJ_ERR_CUDA;
  }
  for (int i = 0; i < 7; i++) cudaEventRecord(ctx->ev[i], ctx->stream);
  cudaDeviceSynchronize(); /* the clears above ran on the legacy stream: done before anything uses ctx->stream */
  *out = ctx;
  return AMCLJ_OK;
}

int32_t amclj_destroy(amclj_ctx *ctx) {
  if (!ctx) return AMCLJ_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  dfree(ctx->map.cells); dfree(ctx->map.df4); dfree(ctx->map.df8); dfree(ctx->map.wedge);
  dfree(ctx->beams.dev); dfree(ctx->beams.dev_d); dfree(ctx->divs.dev);
  dfree(ctx->x); dfree(ctx->y); dfree(ctx->th); dfree(ctx->w); dfree(ctx->raw);
  dfree(ctx->x2); dfree(ctx->y2); dfree(ctx->th2); dfree(ctx->w2);
  dfree(ctx->cdf); dfree(ctx->idx); dfree(ctx->normals); dfree(ctx->partials); dfree(ctx->perm); dfree(ctx->tile_counters);
  dfree(ctx->ctrl); dfree(ctx->rawmax); dfree(ctx->stats); dfree(ctx->walk_ctr);
  dfree(ctx->pose_part); dfree(ctx->plan); dfree(ctx->hist); dfree(ctx->totals_all); dfree(ctx->decision); dfree(ctx->mbox);
  for (int g = 0; g < MAX_WORLD; g++)
    for (int k = 0; k < AMCLJ_IPC_BUFFERS; k++)
      if (ctx->ipc_opened[g][k]) cudaIpcCloseMemHandle(ctx->ipc_opened[g][k]);
  dfree(ctx->stage_out); dfree(ctx->stage_in); dfree(ctx->stage_idx);
  dfree(ctx->rt.rows); dfree(ctx->rt.ents); dfree(ctx->rt.recs); dfree(ctx->rt.pstate); dfree(ctx->rt.bins); dfree(ctx->rt.item0); dfree(ctx->rt.slots); dfree(ctx->rt.ctl);
  if (ctx->rt.tables) cudaFree(ctx->rt.tables);
  pt_free(ctx);
  dfree(ctx->pt.pkey);
  dfree(ctx->pt.pstate_d);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->pin_tab) cudaFreeHost(ctx->pin_tab);
  for (int i = 0; i < 7; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; i++)
    if (ctx->ev_pub[i]) cudaEventDestroy(ctx->ev_pub[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return AMCLJ_OK;
}

const char *amclj_last_error(const amclj_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int32_t amclj_set_stream(amclj_ctx *ctx, void *cuda_stream) {
  ENTER();
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return AMCLJ_OK;
}

int32_t amclj_synchronize(amclj_ctx *ctx) {
  ENTER();
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_set_params(amclj_ctx *ctx, const amclj_params *p) {
  ENTER();
  NEED(p, AMCLJ_ERR_INVALID, "set_params: null");
  NEED(p->df_mode >= 0 && p->df_mode <= 8, AMCLJ_ERR_INVALID, "set_params: df_mode %d", p->df_mode);
  if ((p->df_mode == 4 || p->df_mode == 5) && ctx->map.cells && !ctx->map.wedge) { /* directional fields are built on demand */
    NEED((size_t)ctx->map.pitch8 * (ctx->map.H + 2) * WEDGE_CLASSES <= (8ull << 30), AMCLJ_ERR_INVALID,
         "set_params: map too large for directional fields");
    CK(build_wedge_fields(ctx, ctx->stream));
  }
  NEED(p->s8_resampler >= 0 && p->s8_resampler <= 2, AMCLJ_ERR_INVALID, "set_params: resampler %d", p->s8_resampler);
  ctx->p = *p;
  ctx->beams.staged = false; /* the beam table bakes S2/S5/S6 and the mixture constants */
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int32_t amclj_get_params(const amclj_ctx *ctx, amclj_params *p) {
  if (!ctx || !p) return AMCLJ_ERR_INVALID;
  *p = ctx->p;
  return AMCLJ_OK;
}

int32_t amclj_set_shard(amclj_ctx *ctx, int32_t rank, int32_t world, int64_t global_offset,
                        int64_t n_global) {
  ENTER();
  NEED(world >= 1 && rank >= 0 && rank < world && global_offset >= 0 && n_global >= 0,
       AMCLJ_ERR_INVALID, "set_shard: bad shard (%d of %d)", rank, world);
  ctx->rank = rank;
  ctx->world = world;
  ctx->global_offset = global_offset;
  ctx->n_global = n_global;
  return AMCLJ_OK;
}

int32_t amclj_set_map(amclj_ctx *ctx, const int8_t *cells, int32_t width, int32_t height,
                      float resolution) {
  ENTER();
  NEED(cells && width > 0 && height > 0 && width <= 32768 && height <= 32768, AMCLJ_ERR_INVALID,
       "set_map: bad dimensions %d x %d", width, height);
  NEED(isfinite(resolution) && resolution > 0.f, AMCLJ_ERR_INVALID, "set_map: bad resolution");
  CK(cudaStreamSynchronize(ctx->stream));
  DeviceMap &m = ctx->map;
  dfree(m.cells); dfree(m.df4); dfree(m.df8); dfree(m.wedge); dfree(ctx->tile_counters);
  m.W = width;
  m.H = height;
  m.res = resolution;
  size_t ncell = (size_t)width * height;
  m.free_cells = 0;
  for (size_t i = 0; i < ncell; i++) m.free_cells += cells[i] == 0;
  CK(dalloc(&m.cells, ncell));
  CK(cudaMemcpyAsync(m.cells, cells, ncell, cudaMemcpyHostToDevice, ctx->stream));
  m.row_bytes4 = ((width + 2 + 1) / 2 + 15) / 16 * 16;
  m.bytes4 = (size_t)m.row_bytes4 * (height + 2);
  m.fits_smem = m.bytes4 + 2048 <= (size_t)ctx->max_smem_optin;
  CK(dalloc(&m.df4, m.bytes4));
  bool want8 = !m.fits_smem || ctx->p.df_mode == 2;
  m.pitch8 = (width + 2 + 3) / 4 * 4;
  m.bytes8 = (size_t)m.pitch8 * (height + 2);
  if (want8) CK(dalloc(&m.df8, m.bytes8));
  CK(build_distance_fields(ctx, want8, ctx->stream));
  if (ctx->p.df_mode == 4 || ctx->p.df_mode == 5) {
    NEED(m.bytes8 * WEDGE_CLASSES <= (8ull << 30), AMCLJ_ERR_INVALID, "set_map: map too large for directional fields");
    CK(build_wedge_fields(ctx, ctx->stream));
  } else if ((auto_mode(ctx) || ctx->p.df_mode == 6 || ctx->p.df_mode == 7) && m.bytes8 * WEDGE_CLASSES <= ctx->wedge_budget &&
             m.bytes8 * WEDGE_CLASSES < (1ull << 32)) {
    CK(build_wedge_fields(ctx, ctx->stream)); /* auto: both field sets, chosen per update */
  }
  ctx->beams.staged = false;
  ctx->divs.n = 0;
  *ctx->pin_decision = -1; /* a new map: no hint about the cloud */
  ctx->map_gen++;
  {
    int prc = pt_set_map(ctx, cells, width, height);
    if (prc) return prc;
  }
  ctx->rt.recs_mode = -1;
  ctx->st.map_in_smem = m.fits_smem;
  return AMCLJ_OK;
}

int32_t amclj_set_particles(amclj_ctx *ctx, const float *x, const float *y, const float *theta,
                            const float *w, int64_t n) {
  ENTER();
  NEED(x && y && theta && n > 0 && n <= (1ll << 31) - 1, AMCLJ_ERR_INVALID, "set_particles: bad arguments");
  int64_t keep = ctx->n;
  ctx->n = 0; /* nothing to preserve when the set is replaced */
  int rc = ensure_particles(ctx, n);
  if (rc) { ctx->n = keep; return rc; }
  size_t b = sizeof(float) * (size_t)n;
  CK(cudaMemcpyAsync(ctx->x, x, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->y, y, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->th, theta, b, cudaMemcpyHostToDevice, ctx->stream));
  if (w) CK(cudaMemcpyAsync(ctx->w, w, b, cudaMemcpyHostToDevice, ctx->stream));
  else CK(launch_fill(ctx->w, n, 1.0f / (float)n, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n = n;
  if (ctx->world == 1) { ctx->n_global = n; ctx->global_offset = 0; }
  ctx->have_raw = false;
  ctx->have_cdf = false;
  return AMCLJ_OK;
}

int32_t amclj_get_particles(amclj_ctx *ctx, float *x, float *y, float *theta, float *w) {
  ENTER();
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "get_particles: no particles");
  size_t b = sizeof(float) * (size_t)ctx->n;
  if (x) CK(cudaMemcpyAsync(x, ctx->x, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (y) CK(cudaMemcpyAsync(y, ctx->y, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (theta) CK(cudaMemcpyAsync(theta, ctx->th, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (w) CK(cudaMemcpyAsync(w, ctx->w, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int64_t amclj_num_particles(const amclj_ctx *ctx) { return ctx ? ctx->n : 0; }

/* quaternions.clj:82-97 to-heading, in double as on the JVM */
static double to_heading(double qx, double qy, double qz, double qw) {
  const double test = qx * qy + qz * qw;
  if (test > 0.4999) return 2.0 * atan2(qx, qw);   /* singularity at the north pole */
  if (test < -0.499) return -2.0 * atan2(qx, qw);  /* ... and at the south pole (the reference's own constant) */
  return atan2(2.0 * qz * qw - 2.0 * qx * qy, 1.0 - 2.0 * qz * qz - 2.0 * qy * qy);
}

int32_t amclj_set_poses_quat(amclj_ctx *ctx, const double *poses7, const float *w, int64_t n) {
  ENTER();
  NEED(poses7 && n > 0 && n <= (1ll << 31) - 1, AMCLJ_ERR_INVALID, "set_poses_quat: bad arguments");
  std::vector<float> x((size_t)n), y((size_t)n), th((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const double *p = poses7 + 7 * i;
    x[(size_t)i] = (float)p[0];
    y[(size_t)i] = (float)p[1];
    th[(size_t)i] = (float)to_heading(p[3], p[4], p[5], p[6]);
  }
  return amclj_set_particles(ctx, x.data(), y.data(), th.data(), w, n);
}

int32_t amclj_get_poses_quat(amclj_ctx *ctx, double *poses7, float *w) {
  ENTER();
  NEED(poses7, AMCLJ_ERR_INVALID, "get_poses_quat: null");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "get_poses_quat: no particles");
  const size_t n = (size_t)ctx->n;
  std::vector<float> x(n), y(n), th(n);
  int rc = amclj_get_particles(ctx, x.data(), y.data(), th.data(), w);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    double *p = poses7 + 7 * i;
    const double h = (double)th[i];
    p[0] = (double)x[i]; p[1] = (double)y[i]; p[2] = 0.0;
    /* quaternions.clj:57-62 from-heading = rotate (0 0 0 1) by the heading (:19-36, pitch = bank = 0) */
    p[3] = 0.0; p[4] = 0.0; p[5] = sin(h / 2.0); p[6] = cos(h / 2.0);
  }
  return AMCLJ_OK;
}

int32_t amclj_init_uniform(amclj_ctx *ctx, int64_t n, int32_t heading_mode, uint64_t seed) {
  ENTER();
  NEED(ctx->map.cells, AMCLJ_ERR_STATE, "init_uniform: no map set");
  NEED(n > 0 && n <= (1ll << 31) - 1, AMCLJ_ERR_INVALID, "init_uniform: bad n");
  ctx->n = 0;
  int rc = ensure_particles(ctx, n);
  if (rc) return rc;
  if (ctx->world == 1) { ctx->n_global = n; ctx->global_offset = 0; }
  float w0 = 1.0f / (float)(ctx->n_global > 0 ? ctx->n_global : n);
  CK(launch_init_uniform(ctx->map.cells, ctx->map.W, ctx->map.H, ctx->map.res, ctx->global_offset, n,
                         heading_mode, seed, ctx->x, ctx->y, ctx->th, ctx->w, w0, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n = n;
  ctx->have_raw = ctx->have_cdf = false;
  *ctx->pin_decision = -1; /* a freshly initialised cloud: forget what the last one looked like */
  return AMCLJ_OK;
}

int32_t amclj_init_gaussian(amclj_ctx *ctx, int64_t n, double x, double y, double theta,
                            double sd_xy, double sd_theta, uint64_t seed) {
  ENTER();
  NEED(n > 0 && n <= (1ll << 31) - 1, AMCLJ_ERR_INVALID, "init_gaussian: bad n");
  ctx->n = 0;
  int rc = ensure_particles(ctx, n);
  if (rc) return rc;
  if (ctx->world == 1) { ctx->n_global = n; ctx->global_offset = 0; }
  float w0 = 1.0f / (float)(ctx->n_global > 0 ? ctx->n_global : n);
  CK(launch_init_gaussian((float)x, (float)y, (float)theta, (float)sd_xy, (float)sd_theta,
                          ctx->global_offset, n, seed, ctx->x, ctx->y, ctx->th, ctx->w, w0, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n = n;
  ctx->have_raw = ctx->have_cdf = false;
  *ctx->pin_decision = -1; /* a freshly initialised cloud: forget what the last one looked like */
  return AMCLJ_OK;
}

int32_t amclj_motion_update(amclj_ctx *ctx, const double curr[3], const double prev[3],
                            const float *normals, uint64_t seed) {
  ENTER();
  NEED(curr && prev, AMCLJ_ERR_INVALID, "motion: null control");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "motion: no particles");
  int rc = upload_normals(ctx, normals);
  if (rc) return rc;
  stamp(ctx, 0);
  rc = enqueue_motion(ctx, curr, prev, normals != nullptr, seed);
  if (rc) return rc;
  stamp(ctx, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->st.ms_motion = ms;
  return AMCLJ_OK;
}

int32_t amclj_stage_scan(amclj_ctx *ctx, const float *ranges, int32_t n, float angle_min,
                         float angle_increment, float range_max) {
  ENTER();
  NEED(ranges, AMCLJ_ERR_INVALID, "scan: null ranges");
  return stage_scan_impl(ctx, ranges, n, angle_min, angle_increment, range_max);
}

int32_t amclj_sensor_update(amclj_ctx *ctx, const float *ranges, int32_t n, float angle_min,
                            float angle_increment, float range_max) {
  ENTER();
  NEED(ranges, AMCLJ_ERR_INVALID, "scan: null ranges");
  int rc = stage_scan_impl(ctx, ranges, n, angle_min, angle_increment, range_max);
  if (rc) return rc;
  stamp(ctx, 1);
  rc = enqueue_sensor(ctx, true);
  if (rc) return rc;
  stamp(ctx, 2);
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
  ctx->st.ms_sensor = ms;
  return AMCLJ_OK;
}

int32_t amclj_get_raw_weights(amclj_ctx *ctx, float *raw) {
  ENTER();
  NEED(raw, AMCLJ_ERR_INVALID, "get_raw_weights: null");
  NEED(ctx->have_raw, AMCLJ_ERR_STATE, "get_raw_weights: no sensor pass since the last resample");
  CK(cudaMemcpyAsync(raw, ctx->raw, sizeof(float) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_ray_ring(double r_cells, int32_t max_rows, int32_t *lo, int32_t *cnt, int32_t *half_rows) {
  std::vector<uint4> rows;
  std::vector<short2> ents;
  const int E = build_ring(r_cells, rows, ents);
  if (E < 0) return 0;
  if (half_rows) *half_rows = E;
  for (int i = 0; i < (int)rows.size() && i < max_rows; i++) {
    if (lo) lo[i] = (int32_t)rows[i].y;
    if (cnt) cnt[i] = (int32_t)rows[i].z;
  }
  return (int32_t)ents.size();
}

int32_t amclj_ray_table_entries(const amclj_ctx *ctx) {
  return ctx && ctx->rt.available ? ctx->rt.n_entries : 0;
}

int32_t amclj_beams_used(const amclj_ctx *ctx, int32_t n) {
  if (!ctx) return AMCLJ_ERR_INVALID;
  return beams_used(n, ctx->p.s6_max_beams);
}

int32_t amclj_raycast_debug(amclj_ctx *ctx, int32_t n, float angle_min, float angle_increment,
                            float range_max, int64_t first, int64_t count, int32_t *hit_x,
                            int32_t *hit_y, int32_t *steps, uint8_t *hit, float *range_m) {
  ENTER();
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "raycast_debug: no particles");
  NEED(first >= 0 && count > 0 && first + count <= ctx->n, AMCLJ_ERR_INVALID, "raycast_debug: bad range");
  int rc = stage_scan_impl(ctx, nullptr, n, angle_min, angle_increment, range_max);
  if (rc) return rc;
  ctx->beams.staged = false; /* the staged table carries no ranges */
  int mode = df_mode_of(ctx);
  NEED(mode >= 0, AMCLJ_ERR_INVALID, "raycast_debug: df_mode unavailable");
  size_t rays = (size_t)count * ctx->beams.nb;
  int32_t *d_hx = nullptr, *d_hy = nullptr, *d_st = nullptr;
  uint8_t *d_hit = nullptr;
  float *d_rng = nullptr;
  int status = AMCLJ_OK;
  cudaError_t e = cudaSuccess;
  if (hit_x) e = dalloc(&d_hx, rays);
  if (e == cudaSuccess && hit_y) e = dalloc(&d_hy, rays);
  if (e == cudaSuccess && steps) e = dalloc(&d_st, rays);
  if (e == cudaSuccess && hit) e = dalloc(&d_hit, rays);
  if (e == cudaSuccess && range_m) e = dalloc(&d_rng, rays);
  if (e == cudaSuccess) {
    SensorLaunch a;
    fill_sensor_args(ctx, a, mode);
    a.x = ctx->x + first;
    a.y = ctx->y + first;
    a.th = ctx->th + first;
    a.n = count;
    a.dbg_first = 0;
    a.dbg_count = count;
    a.dbg_hit_x = d_hx;
    a.dbg_hit_y = d_hy;
    a.dbg_steps = d_st;
    a.dbg_hit = d_hit;
    a.dbg_range = d_rng;
    a.stats = ctx->stats;
    ctx->st.map_in_smem = mode == DF_SMEM4;
    ctx->last_diag = true;
    ctx->st.beams_used = ctx->beams.nb;
    cudaMemsetAsync(ctx->stats, 0, 12 * sizeof(unsigned long long), ctx->stream);
    if (ctx->p.df_mode == 7) { /* parity probe through the persistent tables' production look-up (ranges only) */
      if (!ctx->pt.built) {
        status = fail(ctx, AMCLJ_ERR_INVALID, "raycast_debug: df_mode 7 unavailable for this map / rays of %.0f cells", ctx->rt.r_cells);
      } else {
        if (d_hx) cudaMemsetAsync(d_hx, 0xff, rays * 4, ctx->stream);
        if (d_hy) cudaMemsetAsync(d_hy, 0xff, rays * 4, ctx->stream);
        if (d_st) cudaMemsetAsync(d_st, 0xff, rays * 4, ctx->stream);
        if (d_hit) cudaMemsetAsync(d_hit, 0xff, rays, ctx->stream);
        ctx->st.map_in_smem = 0;
        status = enqueue_ptable(ctx, a, true);
      }
    } else if (ctx->p.df_mode == 6) { /* parity probe through the ray tables: every occupied cell gets a table */
      if (!ctx->rt.available) {
        status = fail(ctx, AMCLJ_ERR_INVALID, "raycast_debug: df_mode 6 unavailable for rays of %.0f cells", ctx->rt.r_cells);
      } else {
        cudaMemsetAsync(ctx->bbox, 0, 4 * sizeof(uint32_t), ctx->stream);
        e = launch_bbox(a.x, a.y, a.n, ctx->bbox, ctx->sm_count, ctx->stream);
        a.bbox = ctx->bbox;
        a.compact_m = ctx->compact_cells * ctx->map.res;
        ctx->ctrl_fresh = false;
        ctx->bbox_done = false;
        ctx->st.map_in_smem = 0;
        if (e == cudaSuccess) status = enqueue_raytable(ctx, a, true, 2);
      }
    } else {
      e = launch_sensor(ctx, a, mode, true, ctx->stream);
    }
  }
  auto back = [&](void *dst, const void *src, size_t bytes) {
    if (e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  back(hit_x, d_hx, rays * 4);
  back(hit_y, d_hy, rays * 4);
  back(steps, d_st, rays * 4);
  back(hit, d_hit, rays);
  back(range_m, d_rng, rays * 4);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess && status == AMCLJ_OK) status = fail(ctx, AMCLJ_ERR_CUDA, "raycast_debug: %s", cudaGetErrorString(e));
  dfree(d_hx); dfree(d_hy); dfree(d_st); dfree(d_hit); dfree(d_rng);
  return status;
}

int32_t amclj_normalize(amclj_ctx *ctx, int32_t from_raw, uint64_t *total_out, float *wmax_out) {
  ENTER();
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "normalize: no particles");
  NEED(!from_raw || ctx->have_raw, AMCLJ_ERR_STATE, "normalize: no sensor pass to normalise");
  int rc = enqueue_cdf(ctx, from_raw != 0, false);
  if (rc) return rc;
  return read_scalars(ctx, total_out, wmax_out);
}

int32_t amclj_resample(amclj_ctx *ctx, int32_t mode, double u0, const int32_t *stream_idx,
                       const uint32_t *stream_u, int64_t stream_len, uint64_t seed,
                       int32_t *idx_out) {
  ENTER();
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "resample: no particles");
  NEED(mode >= 0 && mode <= 2, AMCLJ_ERR_INVALID, "resample: mode %d", mode);
  int rc = AMCLJ_OK;
  stamp(ctx, 2);
  if (mode != AMCLJ_RESAMPLE_TOURNAMENT && !ctx->have_cdf) {
    rc = enqueue_cdf(ctx, ctx->have_raw, false);
    if (rc) return rc;
  }
  if (mode == AMCLJ_RESAMPLE_SYSTEMATIC) {
    NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "resample: u0 must be in [0,1)");
    rc = read_scalars(ctx, nullptr, nullptr);
    if (rc) return rc;
    NEED(ctx->last_total > 0, AMCLJ_ERR_STATE, "resample: all weights are zero");
    rc = enqueue_systematic(ctx, u0);
  } else if (mode == AMCLJ_RESAMPLE_ROULETTE) {
    rc = run_roulette(ctx, stream_idx, stream_u, stream_len, seed);
  } else {
    rc = run_tournament(ctx, stream_idx, stream_len, seed);
  }
  if (rc) return rc;
  stamp(ctx, 3);
  if (idx_out)
    CK(cudaMemcpyAsync(idx_out, ctx->idx, sizeof(int32_t) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
  ctx->st.ms_resample = ms;
  return AMCLJ_OK;
}

int32_t amclj_get_resample_indices(amclj_ctx *ctx, int32_t *idx_out) {
  ENTER();
  NEED(idx_out && ctx->n > 0, AMCLJ_ERR_INVALID, "get_resample_indices: bad arguments");
  CK(cudaMemcpyAsync(idx_out, ctx->idx, sizeof(int32_t) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_filter_update_async(amclj_ctx *ctx, const double curr[3], const double prev[3],
                                  uint64_t seed, double u0) {
  ENTER();
  NEED(curr && prev, AMCLJ_ERR_INVALID, "filter_update: null control");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "filter_update: no particles");
  NEED(ctx->beams.staged, AMCLJ_ERR_STATE, "filter_update: no scan staged");
  NEED(ctx->p.s8_resampler == AMCLJ_RESAMPLE_SYSTEMATIC, AMCLJ_ERR_INVALID,
       "filter_update_async: only the systematic resampler runs without a host round trip");
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "filter_update: u0 must be in [0,1)");
  stamp(ctx, 0);
  CK(reset_ctrl(ctx, true)); /* one clear per update: bounding box, max, total, scan tiles */
  ctx->ctrl_fresh = true;
  ctx->bbox_done = false;
  /* stage boundaries are timed only on request (collect_stats bit 1): an event record between two
   * kernels keeps the second from being staged behind the first (~2.5 us each on a 170 us update) */
  ctx->time_main = ctx->time_main_env || (ctx->p.collect_stats & 2) != 0;
  ctx->stages_timed = ctx->time_main;
  int rc = enqueue_motion(ctx, curr, prev, ctx->normals_staged, seed);
  ctx->normals_staged = false;
  if (rc) return rc;
  if (ctx->stages_timed) stamp(ctx, 1);
  rc = enqueue_sensor(ctx, false);
  if (rc) return rc;
  if (ctx->stages_timed) stamp(ctx, 2);
  rc = enqueue_cdf(ctx, true, true);
  if (rc) return rc;
  rc = enqueue_systematic(ctx, u0);
  if (rc) return rc;
  stamp(ctx, 3);
  return AMCLJ_OK;
}

/* stage_scan + filter_update_async in ONE call.  When the scan's geometry is the staged one (every message of a
 * laser after the first), the motion kernel — which needs nothing of the scan — is enqueued FIRST and the host builds
 * and uploads the message's beam table while it runs, instead of the GPU waiting for the host to do so. */
int32_t amclj_filter_update_scan_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                       double u0, const float *ranges, int32_t n, float angle_min,
                                       float angle_increment, float range_max) {
  ENTER();
  NEED(curr && prev && ranges, AMCLJ_ERR_INVALID, "filter_update_scan: null argument");
  const bool fast = scan_geometry_staged(ctx, n, angle_min, angle_increment, range_max) && ctx->n > 0 &&
                    ctx->p.s8_resampler == AMCLJ_RESAMPLE_SYSTEMATIC && u0 >= 0.0 && u0 < 1.0;
  if (!fast) {
    int rc = stage_scan_impl(ctx, ranges, n, angle_min, angle_increment, range_max);
    if (rc) return rc;
    return amclj_filter_update_async(ctx, curr, prev, seed, u0);
  }
  stamp(ctx, 0);
  CK(reset_ctrl(ctx, true));
  ctx->ctrl_fresh = true;
  ctx->bbox_done = false;
  ctx->time_main = ctx->time_main_env || (ctx->p.collect_stats & 2) != 0;
  ctx->stages_timed = ctx->time_main;
  int rc = enqueue_motion(ctx, curr, prev, ctx->normals_staged, seed);
  ctx->normals_staged = false;
  if (rc) return rc;
  rc = restage_ranges(ctx, ranges, n, range_max);
  if (rc) return rc;
  if (ctx->stages_timed) stamp(ctx, 1);
  rc = enqueue_sensor(ctx, false);
  if (rc) return rc;
  if (ctx->stages_timed) stamp(ctx, 2);
  rc = enqueue_cdf(ctx, true, true);
  if (rc) return rc;
  rc = enqueue_systematic(ctx, u0);
  if (rc) return rc;
  stamp(ctx, 3);
  return AMCLJ_OK;
}

int32_t amclj_filter_update(amclj_ctx *ctx, const double curr[3], const double prev[3],
                            const float *normals, uint64_t seed, const float *ranges, int32_t n,
                            float angle_min, float angle_increment, float range_max, double u0) {
  ENTER();
  NEED(curr && prev && ranges, AMCLJ_ERR_INVALID, "filter_update: null argument");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "filter_update: no particles");
  int rc = stage_scan_impl(ctx, ranges, n, angle_min, angle_increment, range_max);
  if (rc) return rc;
  rc = upload_normals(ctx, normals);
  if (rc) return rc;
  if (ctx->p.s8_resampler == AMCLJ_RESAMPLE_SYSTEMATIC) {
    ctx->normals_staged = normals != nullptr;
    rc = amclj_filter_update_async(ctx, curr, prev, seed, u0);
    if (rc) return rc;
    rc = check_zero_total(ctx, "filter_update");
    collect_times(ctx);
    return rc;
  }
  /* roulette / tournament need the host in the loop (pf.clj:156-167 has an unbounded
   * attempt count) */
  stamp(ctx, 0);
  CK(reset_ctrl(ctx, true)); /* one clear per update: bounding box, max, total, scan tiles */
  ctx->ctrl_fresh = true;
  ctx->bbox_done = false;
  rc = enqueue_motion(ctx, curr, prev, normals != nullptr, seed);
  if (rc) return rc;
  stamp(ctx, 1);
  rc = enqueue_sensor(ctx, false);
  if (rc) return rc;
  stamp(ctx, 2);
  if (ctx->p.s8_resampler == AMCLJ_RESAMPLE_ROULETTE) {
    rc = enqueue_cdf(ctx, true, true);
    if (rc) return rc;
    rc = run_roulette(ctx, nullptr, nullptr, 0, seed ^ 0x9E3779B97F4A7C15ull);
  } else {
    rc = run_tournament(ctx, nullptr, 0, seed ^ 0x9E3779B97F4A7C15ull);
  }
  if (rc) return rc;
  stamp(ctx, 3);
  CK(cudaStreamSynchronize(ctx->stream));
  collect_times(ctx);
  return AMCLJ_OK;
}

int32_t amclj_filter_update_host(amclj_ctx *ctx, float *x, float *y, float *theta, float *w,
                                 int64_t n, const double curr[3], const double prev[3],
                                 const float *normals, uint64_t seed, const float *ranges,
                                 int32_t n_ranges, float angle_min, float angle_increment,
                                 float range_max, double u0) {
  ENTER();
  NEED(x && y && theta && n > 0 && n <= (1ll << 31) - 1, AMCLJ_ERR_INVALID, "filter_update_host: bad arguments");
  NEED(curr && prev && ranges, AMCLJ_ERR_INVALID, "filter_update_host: null argument");
  if (ctx->p.s8_resampler != AMCLJ_RESAMPLE_SYSTEMATIC) { /* host-in-the-loop resamplers: plain composition */
    int rc = amclj_set_particles(ctx, x, y, theta, nullptr, n);
    if (rc) return rc;
    rc = amclj_filter_update(ctx, curr, prev, normals, seed, ranges, n_ranges, angle_min, angle_increment,
                             range_max, u0);
    if (rc) return rc;
    return amclj_get_particles(ctx, x, y, theta, w);
  }
  /* upload -> update -> download enqueued back to back, one synchronisation at the end */
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "filter_update: u0 must be in [0,1)");
  int64_t keep = ctx->n;
  ctx->n = 0;
  int rc = ensure_particles(ctx, n);
  if (rc) { ctx->n = keep; return rc; }
  size_t b = sizeof(float) * (size_t)n;
  CK(cudaMemcpyAsync(ctx->x, x, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->y, y, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->th, theta, b, cudaMemcpyHostToDevice, ctx->stream));
  ctx->n = n;
  if (ctx->world == 1) { ctx->n_global = n; ctx->global_offset = 0; }
  ctx->have_raw = ctx->have_cdf = false;
  rc = stage_scan_impl(ctx, ranges, n_ranges, angle_min, angle_increment, range_max);
  if (rc) return rc;
  rc = upload_normals(ctx, normals);
  if (rc) return rc;
  ctx->normals_staged = normals != nullptr;
  rc = amclj_filter_update_async(ctx, curr, prev, seed, u0);
  if (rc) return rc;
  CK(cudaMemcpyAsync(x, ctx->x, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(y, ctx->y, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(theta, ctx->th, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (w) CK(cudaMemcpyAsync(w, ctx->w, b, cudaMemcpyDeviceToHost, ctx->stream));
  rc = check_zero_total(ctx, "filter_update_host");
  collect_times(ctx);
  return rc;
}

int32_t amclj_pose_estimate(amclj_ctx *ctx, double out[4]) {
  ENTER();
  NEED(out, AMCLJ_ERR_INVALID, "pose_estimate: null");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "pose_estimate: no particles");
  int nb = pose_blocks();
  CK(launch_pose(ctx->x, ctx->y, ctx->th, ctx->n, ctx->pose_part, ctx->stream));
  std::vector<double> h((size_t)nb * 4);
  CK(cudaMemcpyAsync(h.data(), ctx->pose_part, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double s[4] = {0, 0, 0, 0};
  for (int b = 0; b < nb; b++)
    for (int k = 0; k < 4; k++) s[k] += h[(size_t)b * 4 + k];
  /* sums only: a sharded caller adds the ranks' sums before dividing (out[k] / n_global) */
  double inv = 1.0 / (double)(ctx->world > 1 ? 1 : ctx->n);
  for (int k = 0; k < 4; k++) out[k] = s[k] * inv;
  return AMCLJ_OK;
}

int32_t amclj_get_stats(amclj_ctx *ctx, amclj_stats *out) {
  ENTER();
  NEED(out, AMCLJ_ERR_INVALID, "get_stats: null");
  CK(cudaStreamSynchronize(ctx->stream));
  collect_times(ctx);
  unsigned long long h[16] = {0};
  CK(cudaMemcpy(h, ctx->stats, sizeof(h), cudaMemcpyDeviceToHost));
  ctx->st.zero_weight_updates = h[12];
  ctx->st.ms_table_fill = ctx->pt.ms_fill;
  ctx->st.table_bytes = ctx->pt.tables_bytes;
  ctx->st.oob_probes = h[6];
  if (h[8] && ctx->last_diag) ctx->st.map_in_smem = (int)h[8] - 1 == DF_SMEM4; /* the placement that actually ran */
  ctx->st.rays = h[0];
  ctx->st.cells_visited = h[1];
  ctx->st.probes = h[2];
  ctx->st.hits = h[3];
  ctx->st.table_cells = ctx->last_diag ? h[9] : 0;
  ctx->st.table_particles = ctx->last_diag ? h[10] : 0;
  ctx->st.table_misses = ctx->last_diag ? h[11] : 0;
  ctx->st.n_particles = ctx->n;
  *out = ctx->st;
  return AMCLJ_OK;
}

/* ---- augmented MCL: pf.clj:243-279 ---------------------------------------------------- */
static int select_kth(amclj_ctx *ctx, const float *v, int64_t n, int64_t k, float *out) {
  /* k-th smallest (0-based): four histogram passes, most significant byte first */
  uint32_t prefix = 0, mask = 0;
  uint32_t h[256];
  for (int shift = 24; shift >= 0; shift -= 8) {
    CK(launch_select_hist(v, n, prefix, mask, shift, ctx->hist, ctx->sm_count, ctx->stream));
    CK(cudaMemcpyAsync(h, ctx->hist, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int b = 0;
    for (; b < 256; b++) {
      if (k < (int64_t)h[b]) break;
      k -= h[b];
    }
    NEED(b < 256, AMCLJ_ERR_STATE, "median: rank outside the histogram");
    prefix |= (uint32_t)b << shift;
    mask |= 255u << shift;
  }
  *out = ord_to_float_host(prefix);
  return AMCLJ_OK;
}

int32_t amclj_median_weight(amclj_ctx *ctx, double *out) {
  ENTER();
  NEED(out, AMCLJ_ERR_INVALID, "median: null");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "median: no particles");
  const float *v = ctx->have_raw ? ctx->raw : ctx->w;
  if (!ctx->hist) CK(dalloc(&ctx->hist, 256));
  float a = 0, b = 0;
  int rc = select_kth(ctx, v, ctx->n, ctx->n / 2, &a);
  if (rc) return rc;
  if (ctx->n % 2 == 0) { /* even count: mean of the two middle order statistics */
    rc = select_kth(ctx, v, ctx->n, ctx->n / 2 - 1, &b);
    if (rc) return rc;
    *out = ((double)a + (double)b) / 2.0;
  } else {
    *out = (double)a;
  }
  return AMCLJ_OK;
}

int32_t amclj_augmented_resample(amclj_ctx *ctx, double alpha_slow, double alpha_fast, uint64_t seed,
                                 int32_t *idx_out) {
  ENTER();
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "augmented: no particles");
  NEED(ctx->map.cells, AMCLJ_ERR_STATE, "augmented: no map set");
  double w_avg = 0;
  int rc = amclj_median_weight(ctx, &w_avg); /* pf.clj:271 */
  if (rc) return rc;
  ctx->w_slow += alpha_slow * (w_avg - ctx->w_slow); /* pf.clj:249-250 */
  ctx->w_fast += alpha_fast * (w_avg - ctx->w_fast);
  double thr = 1.0 - ctx->w_fast / ctx->w_slow;
  if (!(thr > 0.0)) thr = 0.0; /* (max 0.0 ...) */
  const float *wv = ctx->have_raw ? ctx->raw : ctx->w;
  stamp(ctx, 2);
  CK(launch_augmented(wv, ctx->x, ctx->y, ctx->th, ctx->n, (float)thr, ctx->map.cells, ctx->map.W,
                      ctx->map.H, ctx->map.res, seed, ctx->x2, ctx->y2, ctx->th2, ctx->w2, ctx->idx,
                      ctx->stream));
  swap_particles(ctx);
  ctx->have_raw = false;
  ctx->have_cdf = false;
  stamp(ctx, 3);
  if (idx_out)
    CK(cudaMemcpyAsync(idx_out, ctx->idx, sizeof(int32_t) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_augment_state(amclj_ctx *ctx, const double *set_slow_fast, double *get_slow_fast) {
  if (!ctx) return AMCLJ_ERR_INVALID;
  if (set_slow_fast) {
    ctx->w_slow = set_slow_fast[0];
    ctx->w_fast = set_slow_fast[1];
  }
  if (get_slow_fast) {
    get_slow_fast[0] = ctx->w_slow;
    get_slow_fast[1] = ctx->w_fast;
  }
  return AMCLJ_OK;
}

/* ---- sharded building blocks ------------------------------------------- */
int32_t amclj_shard_motion_sensor_async(amclj_ctx *ctx, const double curr[3], const double prev[3],
                                        uint64_t seed) {
  ENTER();
  NEED(curr && prev, AMCLJ_ERR_INVALID, "shard: null control");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "shard: no particles");
  stamp(ctx, 0);
  CK(reset_ctrl(ctx, true)); /* one clear per update: bounding box, max, total, scan tiles */
  ctx->ctrl_fresh = true;
  ctx->bbox_done = false;
  int rc = enqueue_motion(ctx, curr, prev, false, seed);
  if (rc) return rc;
  stamp(ctx, 1);
  rc = enqueue_sensor(ctx, true);
  if (rc) return rc;
  stamp(ctx, 2);
  return AMCLJ_OK;
}

int32_t amclj_shard_cdf_async(amclj_ctx *ctx) {
  ENTER();
  NEED(ctx->have_raw, AMCLJ_ERR_STATE, "shard_cdf: no sensor pass");
  return enqueue_cdf(ctx, true, true); /* AMCLJ_BUF_RAW_MAX already holds the global max */
}

int32_t amclj_shard_generate_async(amclj_ctx *ctx, uint64_t total_global, uint64_t cdf_offset,
                                   uint64_t r, int64_t m_lo, int64_t count) {
  ENTER();
  NEED(ctx->have_cdf, AMCLJ_ERR_STATE, "shard_generate: no CDF");
  NEED(count >= 0 && ctx->n_global > 0, AMCLJ_ERR_INVALID, "shard_generate: bad arguments");
  if (count == 0) return AMCLJ_OK;
  int rc = ensure_stage(ctx, count);
  if (rc) return rc;
  SysPlan p;
  p.total = total_global;
  p.n = (uint64_t)ctx->n_global;
  p.a = total_global / p.n;
  p.b = total_global % p.n;
  p.r = r;
  CK(launch_set_plan(p, ctx->plan, ctx->stream));
  GatherArgs g;
  memset(&g, 0, sizeof(g));
  g.cdf = ctx->cdf;
  g.n_local = ctx->n;
  g.cdf_offset = cdf_offset;
  g.plan = ctx->plan;
  g.m_lo = m_lo;
  g.count = count;
  g.x = ctx->x;
  g.y = ctx->y;
  g.th = ctx->th;
  g.orow = ctx->stage_out;
  g.oidx64 = ctx->stage_idx;
  g.gid0 = ctx->global_offset;
  g.wout = 1.0f / (float)ctx->n_global;
  CK(launch_systematic(g, ctx->stream));
  ctx->launches += 1;
  return AMCLJ_OK;
}

int32_t amclj_shard_get_stage_idx(amclj_ctx *ctx, int64_t *idx_out, int64_t count) {
  ENTER();
  NEED(idx_out && count >= 0 && count <= ctx->stage_cap, AMCLJ_ERR_INVALID, "shard_get_stage_idx: bad arguments");
  if (count == 0) return AMCLJ_OK;
  CK(cudaMemcpyAsync(idx_out, ctx->stage_idx, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_shard_adopt_async(amclj_ctx *ctx, int64_t count) {
  ENTER();
  NEED(count > 0 && count <= ctx->stage_cap, AMCLJ_ERR_INVALID, "shard_adopt: bad count");
  ctx->n = 0;
  int rc = ensure_particles(ctx, count);
  if (rc) return rc;
  CK(launch_adopt(ctx->stage_in, count, ctx->x, ctx->y, ctx->th, ctx->w, ctx->stream));
  ctx->launches += 1;
  ctx->n = count;
  ctx->have_raw = ctx->have_cdf = false;
  stamp(ctx, 3);
  return AMCLJ_OK;
}

/* ---- peer-memory resampling -------------------------------------------------------------- */
static void local_views(amclj_ctx *c, PeerView *a, PeerView *b) {
  /* A = the buffers that are current now (flip accounts for swaps since export) */
  float *cx = c->flip ? c->x2 : c->x, *cy = c->flip ? c->y2 : c->y, *ct = c->flip ? c->th2 : c->th;
  float *ax = c->flip ? c->x : c->x2, *ay = c->flip ? c->y : c->y2, *at = c->flip ? c->th : c->th2;
  a->x = cx; a->y = cy; a->th = ct; a->cdf = c->cdf; a->n = c->n; a->gid0 = c->global_offset;
  b->x = ax; b->y = ay; b->th = at; b->cdf = c->cdf; b->n = c->n; b->gid0 = c->global_offset;
}

int32_t amclj_shard_export(amclj_ctx *ctx, uint8_t *handles) {
  ENTER();
  NEED(handles, AMCLJ_ERR_INVALID, "shard_export: null");
  NEED(ctx->n > 0 && ctx->cdf, AMCLJ_ERR_STATE, "shard_export: no particles");
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->flip = 0;
  ctx->frozen = true;
  ctx->gen = 0;
  ctx->sync_gen = 0;
  CK(cudaMemsetAsync(ctx->mbox, 0, sizeof(unsigned long long) * MBOX_WORDS, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  void *bufs[AMCLJ_IPC_BUFFERS] = {ctx->x, ctx->y, ctx->th, ctx->x2, ctx->y2, ctx->th2, ctx->cdf, ctx->mbox};
  for (int k = 0; k < AMCLJ_IPC_BUFFERS; k++) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, bufs[k]));
    memcpy(handles + (size_t)k * AMCLJ_IPC_HANDLE_BYTES, &h, sizeof(h));
  }
  return AMCLJ_OK;
}

int32_t amclj_shard_import(amclj_ctx *ctx, int32_t rank, const uint8_t *handles, int64_t n_peer,
                           int64_t peer_global_offset) {
  ENTER();
  NEED(handles && rank >= 0 && rank < MAX_WORLD && n_peer > 0, AMCLJ_ERR_INVALID, "shard_import: bad arguments");
  void *p[AMCLJ_IPC_BUFFERS];
  for (int k = 0; k < AMCLJ_IPC_BUFFERS; k++) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)k * AMCLJ_IPC_HANDLE_BYTES, sizeof(h));
    if (ctx->ipc_opened[rank][k]) cudaIpcCloseMemHandle(ctx->ipc_opened[rank][k]);
    CK(cudaIpcOpenMemHandle(&p[k], h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_opened[rank][k] = p[k];
  }
  PeerView a = {(float *)p[0], (float *)p[1], (float *)p[2], (uint64_t *)p[6], n_peer, peer_global_offset};
  PeerView b = {(float *)p[3], (float *)p[4], (float *)p[5], (uint64_t *)p[6], n_peer, peer_global_offset};
  ctx->peers[0][rank] = a;
  ctx->peers[1][rank] = b;
  ctx->peer_mbox[rank] = reinterpret_cast<unsigned long long *>(p[7]);
  ctx->peer_set[rank] = true;
  return AMCLJ_OK;
}

int32_t amclj_shard_attach_local(amclj_ctx *ctx, int32_t rank, amclj_ctx *peer) {
  ENTER();
  NEED(peer && rank >= 0 && rank < MAX_WORLD && peer->n > 0, AMCLJ_ERR_INVALID, "shard_attach_local: bad arguments");
  if (peer == ctx) {
    ctx->flip = 0;
    ctx->frozen = true;
    ctx->gen = 0;
    ctx->sync_gen = 0;
    CK(cudaMemsetAsync(ctx->mbox, 0, sizeof(unsigned long long) * MBOX_WORDS, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  } else if (peer->device != ctx->device) { /* one process, several GPUs: plain peer access */
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    NEED(can, AMCLJ_ERR_INVALID, "shard_attach_local: device %d cannot access device %d", ctx->device, peer->device);
    cudaError_t pe = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (pe == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
    else CK(pe);
  }
  NEED(peer->flip == 0, AMCLJ_ERR_STATE, "shard_attach_local: attach before the first pull");
  if (peer != ctx && peer->device == ctx->device) ctx->peer_shares_device = true;
  peer->frozen = true;
  local_views(peer, &ctx->peers[0][rank], &ctx->peers[1][rank]);
  ctx->peer_mbox[rank] = peer->mbox;
  ctx->peer_set[rank] = true;
  return AMCLJ_OK;
}

/* ---- the whole sharded update behind ONE call per rank: exchanges through the mailboxes ------- */
static void mailbox_args(amclj_ctx *ctx, MailboxArgs &a) {
  memset(&a, 0, sizeof(a));
  a.mbox = ctx->mbox;
  for (int g = 0; g < ctx->world; g++) a.peer_mbox[g] = ctx->peer_mbox[g];
  a.rank = ctx->rank;
  a.world = ctx->world;
  a.gen = ctx->gen;
  a.rawmax_ord = ctx->rawmax_ord;
  a.rawmax = ctx->rawmax;
  a.total = ctx->total;
  a.totals_all = reinterpret_cast<unsigned long long *>(ctx->totals_all);
  a.error_flag = ctx->stats + 13; /* [13]: mailbox waits that timed out */
}

/* phase A: motion + sensor, publish the local max */
/* ranges != null: the per-message beam table of a scan whose geometry is staged, built and uploaded while the motion
 * kernel runs (amclj_shard_update_scan_async) */
static int shard_phase_a(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                         const float *ranges = nullptr, int n_ranges = 0, float range_max = 0.0f) {
  NEED(curr && prev, AMCLJ_ERR_INVALID, "shard_update: null control");
  NEED(ctx->n > 0, AMCLJ_ERR_STATE, "shard_update: no particles");
  NEED(ctx->beams.staged, AMCLJ_ERR_STATE, "shard_update: no scan staged");
  NEED(ctx->world >= 1 && ctx->world <= MAX_WORLD, AMCLJ_ERR_INVALID, "shard_update: world %d", ctx->world);
  for (int g = 0; g < ctx->world; g++)
    NEED(ctx->peer_set[g] && ctx->peer_mbox[g], AMCLJ_ERR_STATE, "shard_update: rank %d not connected", g);
  ctx->gen++;
  stamp(ctx, 0);
  CK(reset_ctrl(ctx, true));
  ctx->ctrl_fresh = true;
  ctx->bbox_done = false;
  ctx->time_main = ctx->time_main_env || (ctx->p.collect_stats & 2) != 0;
  ctx->stages_timed = ctx->time_main;
  int rc = enqueue_motion(ctx, curr, prev, ctx->normals_staged, seed);
  ctx->normals_staged = false;
  if (rc) return rc;
  if (ranges) {
    rc = restage_ranges(ctx, ranges, n_ranges, range_max);
    if (rc) return rc;
  }
  if (ctx->stages_timed) stamp(ctx, 1);
  rc = enqueue_sensor(ctx, false);
  if (rc) return rc;
  if (ctx->stages_timed) stamp(ctx, 2);
  MailboxArgs m;
  mailbox_args(ctx, m);
  CK(launch_mailbox(m, 0, ctx->stream));
  CK(cudaEventRecord(ctx->ev_pub[0], ctx->stream));
  ctx->launches += 1;
  return AMCLJ_OK;
}

/* phase B: wait for every rank's max, CDF against the global max, publish the local total */
static int shard_phase_b(amclj_ctx *ctx) {
  MailboxArgs m;
  mailbox_args(ctx, m);
  CK(launch_mailbox(m, 1, ctx->stream));
  ctx->max_in_ord = false; /* AMCLJ_BUF_RAW_MAX now holds the global max */
  int rc = enqueue_cdf(ctx, true, true);
  if (rc) return rc;
  CK(launch_mailbox(m, 2, ctx->stream));
  CK(cudaEventRecord(ctx->ev_pub[1], ctx->stream));
  ctx->launches += 2;
  return AMCLJ_OK;
}

static int enqueue_pull(amclj_ctx *ctx, double u0);

/* phase C: wait for every rank's total, then pull this rank's slots over peer memory */
static int shard_phase_c(amclj_ctx *ctx, double u0) {
  MailboxArgs m;
  mailbox_args(ctx, m);
  CK(launch_mailbox(m, 3, ctx->stream));
  ctx->launches += 1;
  return enqueue_pull(ctx, u0);
}

int32_t amclj_shard_update_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                 double u0) {
  ENTER();
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "shard_update: u0 must be in [0,1)");
  int rc = shard_phase_a(ctx, curr, prev, seed);
  if (rc) return rc;
  rc = shard_phase_b(ctx);
  if (rc) return rc;
  return shard_phase_c(ctx, u0);
}

int32_t amclj_host_register(amclj_ctx *ctx, void *ptr, int64_t bytes) {
  ENTER();
  NEED(ptr && bytes > 0, AMCLJ_ERR_INVALID, "host_register: null or empty buffer");
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    (void)cudaGetLastError();
    return AMCLJ_OK;
  }
  CK(e);
  return AMCLJ_OK;
}

int32_t amclj_host_unregister(amclj_ctx *ctx, void *ptr) {
  ENTER();
  NEED(ptr, AMCLJ_ERR_INVALID, "host_unregister: null buffer");
  CK(cudaStreamSynchronize(ctx->stream)); /* no copy from or into it may be in flight */
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    (void)cudaGetLastError();
    return AMCLJ_OK;
  }
  CK(e);
  return AMCLJ_OK;
}

int32_t amclj_shard_update_scan_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                      double u0, const float *ranges, int32_t n, float angle_min,
                                      float angle_increment, float range_max) {
  ENTER();
  NEED(ranges, AMCLJ_ERR_INVALID, "shard_update_scan: null ranges");
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "shard_update: u0 must be in [0,1)");
  int rc;
  if (scan_geometry_staged(ctx, n, angle_min, angle_increment, range_max)) {
    rc = shard_phase_a(ctx, curr, prev, seed, ranges, n, range_max);
  } else {
    rc = stage_scan_impl(ctx, ranges, n, angle_min, angle_increment, range_max);
    if (rc) return rc;
    rc = shard_phase_a(ctx, curr, prev, seed);
  }
  if (rc) return rc;
  rc = shard_phase_b(ctx);
  if (rc) return rc;
  return shard_phase_c(ctx, u0);
}

int32_t amclj_shard_rendezvous_async(amclj_ctx *ctx) {
  ENTER();
  NEED(ctx->world >= 1 && ctx->world <= MAX_WORLD, AMCLJ_ERR_INVALID, "shard_rendezvous: world %d", ctx->world);
  for (int g = 0; g < ctx->world; g++)
    NEED(ctx->peer_set[g] && ctx->peer_mbox[g], AMCLJ_ERR_STATE, "shard_rendezvous: rank %d not connected", g);
  NEED(!ctx->peer_shares_device, AMCLJ_ERR_STATE,
       "shard_rendezvous: a peer shares this GPU (kernels that wait on one another must not share a device)");
  MailboxArgs m;
  mailbox_args(ctx, m);
  m.gen = ++ctx->sync_gen;
  CK(launch_mailbox(m, 4, ctx->stream));
  return AMCLJ_OK;
}

int32_t amclj_shard_update_host(amclj_ctx *ctx, float *x, float *y, float *theta, float *w, int64_t n,
                                const double curr[3], const double prev[3], uint64_t seed, const float *ranges,
                                int32_t n_ranges, float angle_min, float angle_increment, float range_max,
                                double u0) {
  ENTER();
  NEED(x && y && theta && n > 0 && n == ctx->n, AMCLJ_ERR_INVALID,
       "shard_update_host: the shard holds %lld particles (buffers are fixed once connected)", (long long)ctx->n);
  NEED(curr && prev && ranges, AMCLJ_ERR_INVALID, "shard_update_host: null argument");
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "shard_update: u0 must be in [0,1)");
  const size_t b = sizeof(float) * (size_t)n;
  CK(cudaMemcpyAsync(ctx->x, x, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->y, y, b, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->th, theta, b, cudaMemcpyHostToDevice, ctx->stream));
  ctx->have_raw = ctx->have_cdf = false;
  int rc = stage_scan_impl(ctx, ranges, n_ranges, angle_min, angle_increment, range_max);
  if (rc) return rc;
  rc = amclj_shard_update_async(ctx, curr, prev, seed, u0);
  if (rc) return rc;
  CK(cudaMemcpyAsync(x, ctx->x, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(y, ctx->y, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(theta, ctx->th, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (w) CK(cudaMemcpyAsync(w, ctx->w, b, cudaMemcpyDeviceToHost, ctx->stream));
  rc = check_zero_total(ctx, "shard_update_host");
  collect_times(ctx);
  return rc;
}

int32_t amclj_shard_pull_async(amclj_ctx *ctx, double u0) {
  ENTER();
  return enqueue_pull(ctx, u0);
}

static int enqueue_pull(amclj_ctx *ctx, double u0) {
  NEED(ctx->have_cdf, AMCLJ_ERR_STATE, "shard_pull: no CDF");
  NEED(u0 >= 0.0 && u0 < 1.0, AMCLJ_ERR_INVALID, "shard_pull: u0 must be in [0,1)");
  NEED(ctx->world <= MAX_WORLD, AMCLJ_ERR_INVALID, "shard_pull: world > %d", MAX_WORLD);
  PullArgs a;
  memset(&a, 0, sizeof(a));
  for (int g = 0; g < ctx->world; g++) {
    NEED(ctx->peer_set[g], AMCLJ_ERR_STATE, "shard_pull: rank %d not imported", g);
    a.peer[g] = ctx->peers[ctx->flip][g];
  }
  a.world = ctx->world;
  a.totals = ctx->totals_all;
  a.u0_bits = (uint64_t)ldexp(u0, 64);
  a.n_global = ctx->n_global;
  a.slot0 = ctx->global_offset;
  a.count = ctx->n;
  a.ox = ctx->x2;
  a.oy = ctx->y2;
  a.oth = ctx->th2;
  a.ow = ctx->w2;
  a.oidx64 = ctx->stage_idx && ctx->stage_cap >= ctx->n ? ctx->stage_idx : nullptr;
  a.wout = 1.0f / (float)ctx->n_global;
  a.zero_flag = ctx->stats + 12; /* [12]: updates whose total was zero (never cleared) */
  CK(launch_pull(a, ctx->stream));
  ctx->launches += 1;
  swap_particles(ctx);
  ctx->have_raw = ctx->have_cdf = false;
  stamp(ctx, 3);
  return AMCLJ_OK;
}

int32_t amclj_plan_systematic(const uint64_t *totals, int32_t world, int64_t n_global, double u0,
                              uint64_t *r_out, uint64_t *cdf_offsets, int64_t *m_lo,
                              int64_t *m_cnt, int64_t *send_matrix) {
  if (!totals || world < 1 || n_global < 0 || !(u0 >= 0.0 && u0 < 1.0)) return AMCLJ_ERR_INVALID;
  typedef unsigned __int128 u128;
  uint64_t T = 0;
  std::vector<uint64_t> off((size_t)world);
  for (int g = 0; g < world; g++) {
    off[g] = T;
    T += totals[g];
  }
  uint64_t r = T ? (uint64_t)(((u128)(uint64_t)ldexp(u0, 64) * T) >> 64) : 0;
  if (r_out) *r_out = r;
  /* #{m : U_m < C} = ceil((C*N - r) / T), clamped to [0, N] */
  auto below = [&](uint64_t C) -> int64_t {
    u128 cn = (u128)C * (uint64_t)n_global;
    if (cn <= r) return 0;
    u128 v = (cn - r + T - 1) / T;
    return v > (u128)n_global ? n_global : (int64_t)v;
  };
  std::vector<int64_t> lo((size_t)world), cnt((size_t)world);
  for (int g = 0; g < world; g++) {
    if (cdf_offsets) cdf_offsets[g] = off[g];
    if (T == 0) {
      lo[g] = 0;
      cnt[g] = 0;
    } else {
      lo[g] = below(off[g]);
      cnt[g] = below(off[g] + totals[g]) - lo[g];
    }
    if (m_lo) m_lo[g] = lo[g];
    if (m_cnt) m_cnt[g] = cnt[g];
  }
  if (send_matrix) {
    /* slot m belongs to rank d with slots [d*q + min(d, rem), ...): n_global / world each,
     * the remainder to the low ranks */
    int64_t q = n_global / world, rem = n_global % world;
    auto start = [&](int d) { return (int64_t)d * q + (d < rem ? d : rem); };
    for (int g = 0; g < world; g++)
      for (int d = 0; d < world; d++) {
        int64_t a = lo[g] > start(d) ? lo[g] : start(d);
        int64_t b = lo[g] + cnt[g] < start(d + 1) ? lo[g] + cnt[g] : start(d + 1);
        send_matrix[(size_t)g * world + d] = b > a ? b - a : 0;
      }
  }
  return AMCLJ_OK;
}

int32_t amclj_device_ptr(amclj_ctx *ctx, int32_t which, void **ptr, int64_t *bytes) {
  ENTER();
  NEED(ptr, AMCLJ_ERR_INVALID, "device_ptr: null");
  void *p = nullptr;
  int64_t b = 0;
  switch (which) {
    case AMCLJ_BUF_X: p = ctx->x; b = 4 * ctx->cap; break;
    case AMCLJ_BUF_Y: p = ctx->y; b = 4 * ctx->cap; break;
    case AMCLJ_BUF_THETA: p = ctx->th; b = 4 * ctx->cap; break;
    case AMCLJ_BUF_W: p = ctx->w; b = 4 * ctx->cap; break;
    case AMCLJ_BUF_RAW: p = ctx->raw; b = 4 * ctx->cap; break;
    case AMCLJ_BUF_RAW_MAX: p = ctx->rawmax; b = 4; break;
    case AMCLJ_BUF_TOTAL: p = ctx->total; b = 8; break;
    case AMCLJ_BUF_STAGE_OUT: p = ctx->stage_out; b = 16 * ctx->stage_cap; break;
    case AMCLJ_BUF_STAGE_IN: p = ctx->stage_in; b = 16 * ctx->stage_cap; break;
    case AMCLJ_BUF_CDF: p = ctx->cdf; b = 8 * ctx->cap; break;
    case AMCLJ_BUF_TOTALS_ALL: p = ctx->totals_all; b = 8 * MAX_WORLD; break;
    default: return fail(ctx, AMCLJ_ERR_INVALID, "device_ptr: unknown buffer %d", which);
  }
  *ptr = p;
  if (bytes) *bytes = b;
  return AMCLJ_OK;
}

int32_t amclj_reserve_stage(amclj_ctx *ctx, int64_t rows) {
  ENTER();
  NEED(rows >= 0, AMCLJ_ERR_INVALID, "reserve_stage: bad rows");
  CK(cudaStreamSynchronize(ctx->stream));
  return ensure_stage(ctx, rows);
}


/* ---- one process, several GPUs: SURVEY 8b `amclj_create(const int* devices, int n_devices, ...)` ----
 * The reference has ONE caller thread making three calls per update (pf.clj:229-233, :297); a JVM host
 * cannot spawn a process per GPU.  An amclj_mctx owns one amclj_ctx per device, shards the particles in
 * contiguous blocks, replicates map and scan, and runs the sharded update of amclj_shard_update_async on
 * every device from the caller's thread: the kernels exchange {max, total} through peer-memory mailboxes
 * and pull their resampled slots over NVLink, so the host only enqueues.  The work is enqueued phase by
 * phase over all ranks (every publish before any wait), which keeps it deadlock-free even when several
 * ranks share one device (the emulated shards of the tests). */
struct amclj_mctx {
  int32_t n = 0;
  amclj_ctx *ctx[MAX_WORLD] = {};
  int64_t n_global = 0;
  bool connected = false;
  bool shared_device = false; /* two ranks on one GPU (emulated shards) */
  std::string err;
};

static int mfail(amclj_mctx *m, int rank, int rc) {
  if (rc && m) m->err = std::string("rank ") + std::to_string(rank) + ": " + amclj_last_error(m->ctx[rank]);
  return rc;
}
#define MEACH(call)                                  \
  do {                                               \
    for (int r_ = 0; r_ < m->n; r_++) {              \
      amclj_ctx *c = m->ctx[r_];                     \
      int rc_ = (call);                              \
      if (rc_) return mfail(m, r_, rc_);             \
    }                                                \
  } while (0)

static void mslots(const amclj_mctx *m, int rank, int64_t n_global, int64_t *lo, int64_t *cnt) {
  const int64_t q = n_global / m->n, rem = n_global % m->n;
  *lo = rank * q + (rank < rem ? rank : rem);
  *cnt = q + (rank < rem ? 1 : 0);
}

static int mconnect(amclj_mctx *m) {
  for (int r = 0; r < m->n; r++) { /* own rank first: it resets the mailbox and the buffer alternation */
    int rc = amclj_shard_attach_local(m->ctx[r], r, m->ctx[r]);
    if (rc) return mfail(m, r, rc);
  }
  for (int r = 0; r < m->n; r++)
    for (int g = 0; g < m->n; g++) {
      if (g == r) continue;
      int rc = amclj_shard_attach_local(m->ctx[r], g, m->ctx[g]);
      if (rc) return mfail(m, r, rc);
    }
  m->connected = true;
  return AMCLJ_OK;
}

int32_t amclj_create_multi(const int32_t *devices, int32_t n_devices, amclj_mctx **out) {
  if (!out) return AMCLJ_ERR_INVALID;
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > MAX_WORLD) return AMCLJ_ERR_INVALID;
  amclj_mctx *m = new amclj_mctx();
  for (int r = 0; r < n_devices; r++) {
    int rc = amclj_create(devices[r], &m->ctx[r]);
    if (rc) {
      for (int k = 0; k < r; k++) amclj_destroy(m->ctx[k]);
      delete m;
      return rc;
    }
    m->n = r + 1;
    for (int k = 0; k < r; k++) m->shared_device = m->shared_device || devices[k] == devices[r];
  }
  *out = m;
  return AMCLJ_OK;
}

int32_t amclj_destroy_multi(amclj_mctx *m) {
  if (!m) return AMCLJ_OK;
  for (int r = 0; r < m->n; r++) { /* nobody may still be reading a peer's buffers */
    cudaSetDevice(m->ctx[r]->device);
    cudaStreamSynchronize(m->ctx[r]->stream);
  }
  for (int r = 0; r < m->n; r++) amclj_destroy(m->ctx[r]);
  delete m;
  return AMCLJ_OK;
}

int32_t amclj_multi_size(const amclj_mctx *m) { return m ? m->n : 0; }
amclj_ctx *amclj_multi_ctx(amclj_mctx *m, int32_t rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }
const char *amclj_mlast_error(const amclj_mctx *m) { return m ? m->err.c_str() : "null mctx"; }

int32_t amclj_mset_params(amclj_mctx *m, const amclj_params *p) {
  if (!m) return AMCLJ_ERR_INVALID;
  MEACH(amclj_set_params(c, p));
  return AMCLJ_OK;
}

int32_t amclj_mset_map(amclj_mctx *m, const int8_t *cells, int32_t width, int32_t height, float resolution) {
  if (!m) return AMCLJ_ERR_INVALID;
  MEACH(amclj_set_map(c, cells, width, height, resolution));
  return AMCLJ_OK;
}

int32_t amclj_mset_particles(amclj_mctx *m, const float *x, const float *y, const float *theta, const float *w,
                             int64_t n_global) {
  if (!m || !x || !y || !theta || n_global < m->n) return AMCLJ_ERR_INVALID;
  const bool same_split = m->connected && m->n_global == n_global;
  for (int r = 0; r < m->n; r++) {
    int64_t lo, cnt;
    mslots(m, r, n_global, &lo, &cnt);
    amclj_ctx *c = m->ctx[r];
    if (!same_split) c->frozen = false; /* a new split: the buffers may be reallocated, peers are re-attached below */
    int rc = amclj_set_shard(c, r, m->n, lo, n_global);
    if (!rc) rc = amclj_set_particles(c, x + lo, y + lo, theta + lo, w ? w + lo : nullptr, cnt);
    if (rc) return mfail(m, r, rc);
  }
  m->n_global = n_global;
  return same_split ? AMCLJ_OK : mconnect(m);
}

int32_t amclj_minit_uniform(amclj_mctx *m, int64_t n_global, int32_t heading_mode, uint64_t seed) {
  if (!m || n_global < m->n) return AMCLJ_ERR_INVALID;
  const bool same_split = m->connected && m->n_global == n_global;
  for (int r = 0; r < m->n; r++) {
    int64_t lo, cnt;
    mslots(m, r, n_global, &lo, &cnt);
    amclj_ctx *c = m->ctx[r];
    if (!same_split) c->frozen = false;
    int rc = amclj_set_shard(c, r, m->n, lo, n_global);
    if (!rc) rc = amclj_init_uniform(c, cnt, heading_mode, seed); /* Philox keyed by the global particle id */
    if (rc) return mfail(m, r, rc);
  }
  m->n_global = n_global;
  return same_split ? AMCLJ_OK : mconnect(m);
}

int32_t amclj_mget_particles(amclj_mctx *m, float *x, float *y, float *theta, float *w) {
  if (!m || m->n_global <= 0) return AMCLJ_ERR_INVALID;
  for (int r = 0; r < m->n; r++) {
    int64_t lo, cnt;
    mslots(m, r, m->n_global, &lo, &cnt);
    int rc = amclj_get_particles(m->ctx[r], x ? x + lo : nullptr, y ? y + lo : nullptr, theta ? theta + lo : nullptr,
                                 w ? w + lo : nullptr);
    if (rc) return mfail(m, r, rc);
  }
  return AMCLJ_OK;
}

int64_t amclj_mnum_particles(const amclj_mctx *m) { return m ? m->n_global : 0; }

int32_t amclj_mstage_scan(amclj_mctx *m, const float *ranges, int32_t n, float angle_min, float angle_increment,
                          float range_max) {
  if (!m) return AMCLJ_ERR_INVALID;
  MEACH(amclj_stage_scan(c, ranges, n, angle_min, angle_increment, range_max));
  return AMCLJ_OK;
}

static int mphases(amclj_mctx *m, const double curr[3], const double prev[3], uint64_t seed, double u0) {
  if (!m->connected) {
    m->err = "mfilter_update: no particles yet (amclj_mset_particles / amclj_minit_uniform)";
    return AMCLJ_ERR_STATE;
  }
  if (!(u0 >= 0.0 && u0 < 1.0)) {
    m->err = "mfilter_update: u0 must be in [0,1)";
    return AMCLJ_ERR_INVALID;
  }
  for (int r = 0; r < m->n; r++) {
    if (cudaSetDevice(m->ctx[r]->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
    int rc = shard_phase_a(m->ctx[r], curr, prev, seed);
    if (rc) return mfail(m, r, rc);
  }
  /* Kernels that wait on one another must never share a GPU (nothing guarantees that they run at the same
   * time): when two ranks do share one — the emulated shards of the tests — every wait is ordered behind the
   * publishes it waits for by stream events, so that it finds its mailbox full and never spins.  On distinct
   * GPUs the mailbox alone carries the dependency and the host adds nothing. */
  auto after_publishes = [&](int r, int which) -> int {
    if (!m->shared_device) return AMCLJ_OK;
    for (int g = 0; g < m->n; g++)
      if (g != r && cudaStreamWaitEvent(m->ctx[r]->stream, m->ctx[g]->ev_pub[which], 0) != cudaSuccess) return AMCLJ_ERR_CUDA;
    return AMCLJ_OK;
  };
  for (int r = 0; r < m->n; r++) {
    if (cudaSetDevice(m->ctx[r]->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
    int rc = after_publishes(r, 0);
    if (!rc) rc = shard_phase_b(m->ctx[r]);
    if (rc) return mfail(m, r, rc);
  }
  for (int r = 0; r < m->n; r++) {
    if (cudaSetDevice(m->ctx[r]->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
    int rc = after_publishes(r, 1);
    if (!rc) rc = shard_phase_c(m->ctx[r], u0);
    if (rc) return mfail(m, r, rc);
  }
  return AMCLJ_OK;
}

int32_t amclj_mfilter_update_async(amclj_mctx *m, const double curr[3], const double prev[3], uint64_t seed,
                                   double u0) {
  if (!m) return AMCLJ_ERR_INVALID;
  return mphases(m, curr, prev, seed, u0);
}

int32_t amclj_msynchronize(amclj_mctx *m) {
  if (!m) return AMCLJ_ERR_INVALID;
  MEACH(amclj_synchronize(c));
  return AMCLJ_OK;
}

int32_t amclj_mfilter_update_host(amclj_mctx *m, float *x, float *y, float *theta, float *w, int64_t n_global,
                                  const double curr[3], const double prev[3], uint64_t seed, const float *ranges,
                                  int32_t n_ranges, float angle_min, float angle_increment, float range_max,
                                  double u0) {
  if (!m || !x || !y || !theta || !ranges || !curr || !prev) return AMCLJ_ERR_INVALID;
  if (!m->connected || n_global != m->n_global) { /* first call, or a different particle count: (re)shard */
    int rc = amclj_mset_particles(m, x, y, theta, nullptr, n_global);
    if (rc) return rc;
  } else {
    for (int r = 0; r < m->n; r++) {
      int64_t lo, cnt;
      mslots(m, r, n_global, &lo, &cnt);
      amclj_ctx *ctx = m->ctx[r];
      if (cudaSetDevice(ctx->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
      const size_t b = sizeof(float) * (size_t)cnt;
      cudaError_t e = cudaMemcpyAsync(ctx->x, x + lo, b, cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->y, y + lo, b, cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->th, theta + lo, b, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) return mfail(m, r, fail(ctx, AMCLJ_ERR_CUDA, "mfilter_update_host: %s", cudaGetErrorString(e)));
      ctx->have_raw = ctx->have_cdf = false;
    }
  }
  MEACH(amclj_stage_scan(c, ranges, n_ranges, angle_min, angle_increment, range_max));
  int rc = mphases(m, curr, prev, seed, u0);
  if (rc) return rc;
  for (int r = 0; r < m->n; r++) {
    int64_t lo, cnt;
    mslots(m, r, n_global, &lo, &cnt);
    amclj_ctx *ctx = m->ctx[r];
    if (cudaSetDevice(ctx->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
    const size_t b = sizeof(float) * (size_t)cnt;
    cudaError_t e = cudaMemcpyAsync(x + lo, ctx->x, b, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y + lo, ctx->y, b, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(theta + lo, ctx->th, b, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && w) e = cudaMemcpyAsync(w + lo, ctx->w, b, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) return mfail(m, r, fail(ctx, AMCLJ_ERR_CUDA, "mfilter_update_host: %s", cudaGetErrorString(e)));
  }
  for (int r = 0; r < m->n; r++) {
    amclj_ctx *ctx = m->ctx[r];
    if (cudaSetDevice(ctx->device) != cudaSuccess) return AMCLJ_ERR_CUDA;
    rc = check_zero_total(ctx, "mfilter_update_host");
    collect_times(ctx);
    if (rc) return mfail(m, r, rc);
  }
  return AMCLJ_OK;
}

int32_t amclj_mpose_estimate(amclj_mctx *m, double out[4]) {
  if (!m || !out || m->n_global <= 0) return AMCLJ_ERR_INVALID;
  double s[4] = {0, 0, 0, 0};
  for (int r = 0; r < m->n; r++) {
    double part[4];
    int rc = amclj_pose_estimate(m->ctx[r], part); /* local sums on a shard */
    if (rc) return mfail(m, r, rc);
    const double scale = m->n > 1 ? 1.0 : (double)m->ctx[r]->n;
    for (int k = 0; k < 4; k++) s[k] += part[k] * scale;
  }
  for (int k = 0; k < 4; k++) out[k] = s[k] / (double)m->n_global;
  return AMCLJ_OK;
}

}  // extern "C"
