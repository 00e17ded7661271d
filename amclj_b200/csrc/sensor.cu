// sensor.cu — K2: laser measurement model.  One lane per particle, walking that particle's
// beams one after the other: ray-cast through the occupancy grid, score with the beam mixture.
//
// Replaces (jvia/amclj): pf.clj:148-154 apply-sensor-model ->
//   sensor.clj:165-173 likelihood-field-range-finder-model -> :141-162 (beam loop, mixture)
//   -> :130-139 map-range -> :112-128 extract-line -> map.clj:48-57 uninhabitable?.
//
// The ray visits exactly the cells of the reference's integer Bresenham (extract-line), but
// does not test them one by one: the grid is staged as a Chebyshev distance field d(c) =
// "distance in cells to the nearest blocked cell (unknown, occupied or out of range)".
// A Bresenham step moves at most one cell along each axis, so from a cell with d(c) = d the
// next d-1 cells of the line are certainly free and the walk can jump d steps at once; the
// minor-axis position after any number of steps has the closed form
//   j(k) = floor((2*k*dy + dx) / (2*dx))
// (the reference's error accumulator, sensor.clj:123-127, unrolled), evaluated incrementally
// with a per-ray magic reciprocal.  First blocked cell, step count and range are therefore
// bit-identical to the cell-by-cell walk.
//
// Divergence: rays need 1..70 probes.  A warp task is 32 particles x a chunk of beams; every
// lane owns one particle and walks its beams in order, so particles that sit close together
// (pose tracking, resampled duplicates) run in lock step by themselves.  A lane whose ray has
// ended parks; when fewer than `refill` lanes are still walking, the parked lanes score their
// ray and set up their next beam together (while-while persistent lanes).  Each particle's
// weight is accumulated by one lane in beam order: deterministic, no cross-lane reduction.
#include "internal.h"
#include "sensor_common.cuh"

namespace amclj {

#ifndef PROBES_PER_VOTE
#define PROBES_PER_VOTE 8
#endif
#ifndef PROBES_PER_VOTE_WEDGE
#define PROBES_PER_VOTE_WEDGE 6
#endif

#ifndef WEDGE_THREADS
#define WEDGE_THREADS 1024
#define WEDGE_CTAS 1
#endif
/* the L1/L2 byte field is latency bound: 2 CTAs of 24 warps (40 registers) hide more of it than
 * one of 32 (C5: 107.5 -> 97.8 ms) */
#ifndef GLOBAL_THREADS
#define GLOBAL_THREADS 768
#define GLOBAL_CTAS 2
#endif
constexpr int mode_threads(int mode) {
  return mode == DF_WEDGE8 ? WEDGE_THREADS : (mode == DF_SMEM4 ? kMaxThreads : GLOBAL_THREADS);
}
constexpr int mode_ctas(int mode) { return mode == DF_WEDGE8 ? WEDGE_CTAS : (mode == DF_SMEM4 ? 1 : GLOBAL_CTAS); }
template <int MODE, bool DIAG, bool WIDE, bool TS, bool RS>
__global__ void __launch_bounds__(mode_threads(MODE), mode_ctas(MODE))
sensor_kernel(const __grid_constant__ SensorLaunch a) {
  extern __shared__ uint4 smem_dyn[];
  __shared__ uint64_t bar;
  if (a.skip_flag && *a.skip_flag == 1) return; /* a dense cloud: the ray tables took this update (raytable.cu) */
  /* spread-out clouds are walked in spatial (tile) order so that a CTA's probes share an L1-sized
   * neighbourhood; a compact cloud is local as it is */
  const int32_t *perm = a.perm;
  if (a.bbox) {
    const bool compact = cloud_is_compact(a.bbox, a.compact_m);
    if (compact) perm = nullptr;
    if (a.decision && blockIdx.x == 0 && threadIdx.x == 0) *a.decision = compact ? 1 : 0;
  }
  Field df;
  df.g = a.df;
  df.s = 0;
  if (MODE == DF_SMEM4) {
    /* stage the whole nibble-packed field with the bulk-copy engine (TMA, UBLKCP) */
    uint8_t *sdf = reinterpret_cast<uint8_t *>(smem_dyn);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t total = (uint32_t)a.df_bytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                   "r"(total)
                   : "memory");
      const uint32_t CH = 32768;
      for (uint32_t off = 0; off < total; off += CH) {
        uint32_t sz = total - off < CH ? total - off : CH;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(smem_u32(sdf + off)), "l"(a.df + off), "r"(sz), "r"(smem_u32(&bar))
            : "memory");
      }
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          "selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
    }
    df.s = smem_u32(sdf);
  }
  Tables tb;
  tb.mix = reinterpret_cast<const float4 *>(a.beam);
  tb.dir = reinterpret_cast<const float2 *>(a.beam + 4 * (size_t)a.nb);
  tb.div = a.divtab;
  tb.s_dir = tb.s_mix = tb.s_div = 0;
  uint32_t s_rows = 0, s_recs = 0;
  const uint32_t ring_last_row = (uint32_t)a.ring_nrows - 1u;
  if (TS) {
    /* [field | mix float4[nb] | dir float2[nb] | div uint2[ndiv]] */
    uint8_t *base = reinterpret_cast<uint8_t *>(smem_dyn) + (MODE == DF_SMEM4 ? a.df_bytes : 0);
    float4 *smix = reinterpret_cast<float4 *>(base);
    float2 *sdir = reinterpret_cast<float2 *>(smix + a.nb);
    uint2 *sdiv = reinterpret_cast<uint2 *>(sdir + a.nb);
    for (int i = threadIdx.x; i < a.nb; i += blockDim.x) { smix[i] = __ldg(tb.mix + i); sdir[i] = __ldg(tb.dir + i); }
    for (int i = threadIdx.x; i < a.ndiv; i += blockDim.x) sdiv[i] = __ldg(tb.div + i);
    tb.s_mix = smem_u32(smix); tb.s_dir = smem_u32(sdir); tb.s_div = smem_u32(sdiv);
    /* keep the three window addresses in registers: left alone, the compiler rebuilds each of
     * them from SR_CgaCtaId at every use */
    asm volatile("" : "+r"(tb.s_mix), "+r"(tb.s_dir), "+r"(tb.s_div));
    if (RS) { /* | rows uint4[ring_nrows] | set-up records uint4[2 * ring_entries] (16-byte aligned) */
      uint4 *srows = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(sdiv + a.ndiv) + 15) & ~(uintptr_t)15);
      uint4 *srecs = srows + a.ring_nrows;
      for (int i = threadIdx.x; i < a.ring_nrows; i += blockDim.x) srows[i] = __ldg(a.ring_rows + i);
      for (int i = threadIdx.x; i < 2 * a.ring_entries; i += blockDim.x) srecs[i] = __ldg(a.ring_recs + i);
      s_rows = smem_u32(srows); s_recs = smem_u32(srecs);
      asm volatile("" : "+r"(s_rows), "+r"(s_recs));
    }
    __syncthreads();
  }
  const ResDiv rdv = make_resdiv(a.res);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * wpb;
  const int nb = a.nb;
  constexpr bool BYTES = MODE == DF_GLOBAL8 || MODE == DF_WEDGE8;
  const int idx_scale = BYTES ? 1 : 4;
  const uint32_t DMASK = BYTES ? 255u : 15u;
  const int row = a.row_nib * idx_scale;
  const int64_t ngroups = (a.n + 31) >> 5;
  const int64_t ntasks = ngroups * a.chunks;
  float wmax = -INFINITY;
  unsigned long long n_cells = 0, n_probes = 0, n_hits = 0, n_rays = 0, n_oob = 0;
  const uint32_t idx_limit = (uint32_t)(a.df_bytes * (BYTES ? 1u : 8u)); /* DIAG bounds check */

  /* task = 32 consecutive particles (in walking order) x one chunk of beams; the warps of a CTA
   * take consecutive tasks, i.e. neighbouring particles */
  /* Distribution (a.walk_tail_pct >= 0): every CTA owns a contiguous run of the first tasks (its warps
   * draw from it through a shared counter, so neighbouring particles still run side by side and no
   * warp waits for another), and the last walk_tail_pct percent of the tasks are a common tail that
   * whoever finishes first draws from: with 6-7 tasks per warp a fixed share leaves warps idle for up
   * to one task in seven (ncu on C3: SM active cycles 5.9M .. 7.3M).  walk_tail_pct < 0: the fixed
   * interleaved share. */
  __shared__ int s_next;
  const bool dyn = a.walk_tail_pct >= 0 && a.walk_ctr != nullptr;
  const int64_t G = gridDim.x, bid = blockIdx.x;
  const int64_t nt_static = dyn ? ntasks - ntasks * a.walk_tail_pct / 100 : 0;
  const int64_t t_lo = nt_static * bid / G, n_own = nt_static * (bid + 1) / G - t_lo;
  if (dyn) {
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
  }
  for (int64_t it = 0;; it++) {
    int64_t task;
    if (dyn) {
      int kk = 0;
      if (lane == 0) kk = atomicAdd(&s_next, 1);
      kk = __shfl_sync(FULL, kk, 0);
      if (kk < n_own) {
        task = t_lo + kk;
      } else {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(a.walk_ctr, 1u);
        t = __shfl_sync(FULL, t, 0);
        if ((int64_t)t >= ntasks - nt_static) break;
        task = nt_static + t;
      }
    } else {
      task = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5) + it * nwarps;
      if (task >= ntasks) break;
    }
    const int chunk = (int)(task / ngroups);
    const int64_t slot = (task - (int64_t)chunk * ngroups) * 32 + lane;
    const bool valid = slot < a.n;
    const int64_t p = valid && perm ? (int64_t)__ldg(perm + slot) : slot;
    const int nseg = (nb + SEG - 1) / SEG;
    const int b_lo = (int)((int64_t)nseg * chunk / a.chunks) * SEG;
    const int b_hi = min(nb, (int)((int64_t)nseg * (chunk + 1) / a.chunks) * SEG);
    float ox = 0.f, oy = 0.f, th = 0.f;
    if (valid) { ox = __ldg(a.x + p); oy = __ldg(a.y + p); th = __ldg(a.th + p); }
    float st, ct;
    sincos_det(th, st, ct);
    if (a.s10) { /* S10: base->laser offset (sensor.clj:78-81), off by default */
      float nx = ox + fmaf(a.laser_x, ct, -(a.laser_y * st));
      float ny = oy + fmaf(a.laser_x, st, a.laser_y * ct);
      ox = nx; oy = ny;
    }
    /* once per task: the start cell by the reference's own double divide, and the offset inside it (contract.h) */
    const RayStart rst = make_ray_start(a, ox, oy, st, ct);
    const int x0 = rst.x0, y0 = rst.y0;
    const bool start_in = (uint32_t)x0 < (uint32_t)a.W && (uint32_t)y0 < (uint32_t)a.H;
    /* padded field: cell (x, y) lives at (y+1)*row + (x+1); entry 0 is a ring cell (d = 0),
     * which makes an out-of-range start an immediate hit (the start cell is tested first) */
    const int idx0 = start_in ? (y0 + 1) * row + (x0 + 1) * idx_scale : 0;
    const int xoff = x0, yoff = y0 - a.ring_E; /* RS: end cell relative to the start, see below */
    const bool dbg = DIAG && a.dbg_count > 0 && valid && p >= a.dbg_first && p < a.dbg_first + a.dbg_count;

    /* the ray in flight; a parked lane (am == 0) turns every update below into a no-op */
    uint32_t idx = 0; /* unsigned: 32 byte planes of a large map reach past 2^31 */
    int k = 0, j = 0, err = 0, dy = 0, ndx = 0, kmax = 0, stepA = 0, stepB = 0;
    uint32_t M = 0, sh = 0, am = 0;
    int dg_a0 = 0, dg_b0 = 0, dg_xs = 0, dg_ys = 0, dg_steep = 0, dg_probes = 0; /* DIAG only */
    bool have_ray = false;
    float acc = 0.0f, prod = 1.0f;
    int pexp = 0;
    long long qacc = 0;
    int jb = valid ? b_lo : b_hi; /* this lane's current beam */
    int seg_end = min(b_lo + SEG, b_hi);

    for (;;) {
      if (am == 0) {
        /* ---- (1) score the ray that has ended: sensor.clj:149-162 beam mixture ---------- */
        if (have_ray) {
          const bool hit = k <= kmax; /* a miss leaves k beyond the last cell to test */
          /* sensor.clj:118-120 hypot(x - x0, y - y0) * resolution on a hit, :121 max-range on a miss
           * (branch-free: the root of a miss's leftover counters is computed and dropped) */
          const uint32_t d2 = (uint32_t)k * (uint32_t)k + (uint32_t)j * (uint32_t)j;
          const float rng = hit ? __fmul_rn(sqrt_int(d2), a.res) : a.miss_value;
          const float4 mx = ld_mix<TS>(tb, jb); /* obs, pz2's exponential, pz3 + pz4: beam-only terms */
          float z = mx.x - rng;
          float pz1 = hit_gauss(z, a.k2, a.z_hit);
          float pz2 = z < 0.0f ? mx.y : 0.0f;
          float sum = (pz1 + pz2) + mx.z;
          if (a.s7) { /* sum of logs as the log of a running product, flushed before underflow */
            logprod_mul(acc, prod, pexp, sum);
          } else {
            acc += (sum * sum) * sum; /* sensor.clj:162 (pow . 3) */
          }
          if (DIAG) {
            int kk = hit ? k : kmax, jj = j;
            n_rays++; n_cells += (unsigned)kk + 1u; n_probes += (unsigned)dg_probes; n_hits += hit ? 1u : 0u;
            if (dbg) {
              if (!hit) jj = ndx < 0 ? (int)((2ll * kk * dy - ndx) / (-2ll * ndx)) : 0;
              int ca = dg_a0 + dg_xs * kk, cb2 = dg_b0 + dg_ys * jj;
              int64_t o = (p - a.dbg_first) * nb + jb;
              if (a.dbg_hit_x) a.dbg_hit_x[o] = dg_steep ? cb2 : ca;
              if (a.dbg_hit_y) a.dbg_hit_y[o] = dg_steep ? ca : cb2;
              if (a.dbg_steps) a.dbg_steps[o] = kk;
              if (a.dbg_hit) a.dbg_hit[o] = (uint8_t)hit;
              if (a.dbg_range) a.dbg_range[o] = rng;
            }
          }
          have_ray = false;
          jb++;
          if (jb == seg_end) { /* close the segment: one rounding to fixed point */
            float v = a.s7 ? logprod_close(acc, prod, pexp) : acc;
            qacc += __float2ll_rn(fmaxf(v, FIX_FLOOR) * FIX_SCALE);
            acc = 0.0f; prod = 1.0f; pexp = 0;
            seg_end = min(seg_end + SEG, b_hi);
          }
        }
        /* ---- (2) set up the next beam: sensor.clj:130-139 map-range -------------------- */
        if (jb < b_hi) {
          const float2 bd = ld_dir<TS>(tb, jb);
          float cb = bd.x, sb = bd.y;
          float c = cb, s = sb;
          if (a.s1) { /* S1: theta + obs-bearing by the angle-addition identity */
            c = fmaf(ct, cb, -(st * sb));
            s = fmaf(st, cb, ct * sb);
          }
          /* sensor.clj:99-103 project-pose, map.clj:36-46 world->map: the cell of the end point (contract.h: one
           * float FMA per coordinate relative to the start cell, the reference's double computation next to a .5
           * boundary) — once, for the record look-up and for the computed set-up alike */
          int x1, y1;
          end_cell(a, rdv, rst, c, s, ox, oy, th, jb, x1, y1);
          bool from_table = false;
          if (RS) {
            /* Everything sensor.clj:130-139 derives after the two roundings (steep?, the swaps, calc-step,
             * dx, dy, the miss rule, the reciprocal, the strides and the direction class) depends only on
             * the end cell RELATIVE to the start cell, and that offset lies on a thin ring around the
             * circle of the ray's length: one record per ring cell, filled once per scan geometry
             * (ring_records_kernel below, from this very code), read here with two LDS.128.
             * floor(q + 0.5) as an FP32 add with round-down against 1.5 * 2^23 (exact below 2^22, which
             * the host checks for every start inside the map; a start outside hits at step 0 whatever
             * record it gets, so its garbage offset only has to stay inside the table). */
            const int ex = x1 - xoff;
            const uint32_t ry = (uint32_t)(y1 - yoff);
            const uint32_t ryc = min(ry, ring_last_row);
            uint4 rw;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rw.x), "=r"(rw.y), "=r"(rw.z), "=r"(rw.w) : "r"(s_rows + ryc * 16u));
            const uint32_t tt = (uint32_t)abs(ex) - rw.y;
            const bool ok = tt < rw.z && ry <= ring_last_row;
            from_table = ok || !start_in;
            if (from_table) {
              const uint32_t ent = ok ? tt + (ex < 0 ? rw.w : rw.x) : 0u;
              uint4 r0, r1;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(s_recs + ent * 32u));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(s_recs + ent * 32u + 16u));
              dy = (int)r0.x; ndx = (int)r0.y; kmax = (int)r0.z; M = r0.w;
              sh = WIDE ? (r1.x & 31u) : 0u; stepA = (int)r1.y; stepB = (int)r1.z;
              idx = (uint32_t)idx0 + (start_in ? r1.w : 0u);
              k = 0; j = 0; err = 0;
              if (DIAG) {
                const bool steep = (r1.x >> 8) & 1u;
                dg_a0 = steep ? y0 : x0; dg_b0 = steep ? x0 : y0;
                dg_xs = (r1.x >> 9) & 1u ? -1 : 1; dg_ys = (r1.x >> 10) & 1u ? -1 : 1; dg_steep = steep; dg_probes = 0;
              }
            }
          }
          if (!from_table) {
          int adx = abs(x1 - x0), ady = abs(y1 - y0);
          bool steep = ady > adx; /* sensor.clj:87-91 */
          int a0 = steep ? y0 : x0, b0 = steep ? x0 : y0;
          int a1, b1;
          if (steep) {
            a1 = y1; b1 = x1;
          } else if (a.s3) {
            a1 = x1; b1 = y1;
          } else { /* sensor.clj:138 `(if steep [y1 x1] [x0 y0])`: end = start */
            a1 = a0; b1 = b0;
          }
          int xs = a0 < a1 ? 1 : -1, ys = b0 < b1 ? 1 : -1; /* sensor.clj:108-110, ties -> -1 */
          int dx = abs(a0 - a1);
          dy = abs(b0 - b1);
          /* the miss test runs AFTER the cell test at step k: S4=1 `x == max-x`;
           * S4=0 (sensor.clj:121) `x > max-x` */
          if (a.s4) kmax = dx;
          else if (xs > 0) kmax = dx + 1;
          else if (dx > 0) kmax = 0;
          else { /* dx = dy = 0: `2*err >= 0` always holds -> diagonal walk to a blocked cell */
            kmax = 0x3fffffff; dx = 1; dy = 1;
          }
          uint2 e = ld_div<TS>(tb, min(dx, a.ndiv - 1)); /* dx < ndiv unless the start is out of range */
          M = e.x; sh = e.y;
          stepA = xs * (steep ? row : idx_scale);
          stepB = ys * (steep ? idx_scale : row);
          idx = idx0; k = 0; j = 0; err = 0; ndx = -dx;
          if (MODE == DF_WEDGE8 && start_in) { /* the plane of this ray's direction class */
            /* plane = ((steep*4 + (xs<0)*2 + (ys<0)) * BINS + bin), bin = min(BINS-1, floor(dy*BINS/dx))
             * without a division; accumulated directly as a byte offset */
            const uint32_t P = (uint32_t)a.wedge_plane;
            uint32_t off = steep ? 4 * WEDGE_BINS * P : 0;
            off += xs < 0 ? 2 * WEDGE_BINS * P : 0;
            off += ys < 0 ? WEDGE_BINS * P : 0;
            const int dyb = dy * WEDGE_BINS;
#pragma unroll
            for (int t = 1; t < WEDGE_BINS; t++) off += (dyb >= t * dx) ? P : 0;
            idx += off;
          }
          if (DIAG) { dg_a0 = a0; dg_b0 = b0; dg_xs = xs; dg_ys = ys; dg_steep = steep; dg_probes = 0; }
          }
          have_ray = true;
          am = DMASK;
        }
      }
      uint32_t act = __ballot_sync(FULL, am != 0);
      if (act == 0) break;
      /* lanes that will need another set-up later; when none does, walk to the end */
      const uint32_t more = __ballot_sync(FULL, jb + 1 < b_hi);
      const int thresh = more ? a.refill : 1;
      /* ---- (3) walk: sensor.clj:112-128 extract-line, d cells per probe ------------------ */
      do {
#pragma unroll
        for (int u = 0; u < (MODE == DF_WEDGE8 ? PROBES_PER_VOTE_WEDGE : PROBES_PER_VOTE); u++) {
          /* branch-free: a parked lane reads d = 0, which makes every update a no-op */
          if (DIAG && idx >= idx_limit) { n_oob++; idx = 0; } /* never taken: see tests */
          const uint32_t d = df_probe_parked<MODE>(df, idx, am);
          if (DIAG) dg_probes += am ? 1 : 0;
          k += (int)d;
          const bool fin = d == 0 || k > kmax; /* blocked cell, or every cell to the end is free */
          const int e2 = err + (int)d * dy;
          const uint32_t t = (uint32_t)(2 * e2 - ndx);
          const uint32_t jd = WIDE ? (__umulhi(t, M) >> sh) : __umulhi(t, M);
          err = e2 + (int)jd * ndx;
          j += (int)jd;
          idx += (uint32_t)((int)d * stepA + (int)jd * stepB);
          if (BYTES && PARK_ON_ZERO) idx = fin ? 0u : idx; /* entry 0: d = 0 (df_probe_parked) */
          am = fin ? 0u : am;
        }
        act = __ballot_sync(FULL, am != 0);
      } while (__popc(act) >= thresh);
    }
    if (valid) {
      if (a.chunks > 1) {
        a.partials[(int64_t)chunk * a.n + p] = qacc; /* summed (as integers) by combine_kernel */
      } else {
        float v = fixed_to_raw(qacc);
        if (a.raw) a.raw[p] = v;
        wmax = fmaxf(wmax, v);
      }
    }
  }
  if (dyn) { /* the last CTA out leaves the tail counter at zero for the next launch */
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(a.walk_ctr + 1, 1u) == gridDim.x - 1) {
      a.walk_ctr[0] = 0;
      a.walk_ctr[1] = 0;
    }
  }
  if (a.rawmax_ord && a.chunks == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, o));
    if (lane == 0 && wmax > -INFINITY) atomicMax(a.rawmax_ord, ord_of_float(wmax));
  }
  if (DIAG && a.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_rays += __shfl_xor_sync(FULL, n_rays, o);
      n_cells += __shfl_xor_sync(FULL, n_cells, o);
      n_probes += __shfl_xor_sync(FULL, n_probes, o);
      n_hits += __shfl_xor_sync(FULL, n_hits, o);
      n_oob += __shfl_xor_sync(FULL, n_oob, o);
    }
    if (lane == 0) {
      if (blockIdx.x == 0 && threadIdx.x == 0) a.stats[8] = (unsigned long long)MODE + 1ull;
      atomicAdd(a.stats + 0, n_rays);
      atomicAdd(a.stats + 1, n_cells);
      atomicAdd(a.stats + 2, n_probes);
      atomicAdd(a.stats + 3, n_hits);
      atomicAdd(a.stats + 6, n_oob);
    }
  }
}

// raw[p] = sum over chunks of the fixed-point partial weights, plus the max for the normaliser
__global__ void __launch_bounds__(256)
combine_kernel(const long long *__restrict__ part, int chunks, int64_t n, float *__restrict__ raw,
               uint32_t *rawmax_ord, const int32_t *skip_flag) {
  if (skip_flag && *skip_flag == 1) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float v = -INFINITY;
  if (i < n) {
    long long q = __ldg(part + i);
    for (int c = 1; c < chunks; c++) q += __ldg(part + (int64_t)c * n + i);
    v = fixed_to_raw(q);
    raw[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  if ((threadIdx.x & 31) == 0 && rawmax_ord && v > -INFINITY) atomicMax(rawmax_ord, ord_of_float(v));
}

// One set-up record per ring cell (ex, ey) = end cell - start cell: the line parameters the walk
// starts from, computed by the same expressions as set-up (2) above.
//   rec[2e]   = {dy, -dx, kmax, M}    rec[2e+1] = {sh | steep << 8 | (xs<0) << 9 | (ys<0) << 10, stepA, stepB, class offset}
__global__ void __launch_bounds__(256)
ring_records_kernel(const SensorLaunch a, int mode, const short2 *__restrict__ ents, int n_entries, uint4 *recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const short2 en = ents[i];
  if (en.x <= -32767) { /* unused slot / sentinel: never indexed by a valid offset */
    recs[2 * i] = make_uint4(0u, (uint32_t)-1, 0u, 0u);
    recs[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const bool bytes = mode == DF_GLOBAL8 || mode == DF_WEDGE8;
  const int idx_scale = bytes ? 1 : 4, row = a.row_nib * idx_scale;
  const int ex = en.x, ey = en.y;
  const bool steep = abs(ey) > abs(ex);                       /* sensor.clj:87-91 */
  const int da = steep ? ey : (a.s3 ? ex : 0), db = steep ? ex : (a.s3 ? ey : 0); /* :137-138 */
  const int xs = da > 0 ? 1 : -1, ys = db > 0 ? 1 : -1;       /* :108-110, ties -> -1 */
  int dx = abs(da), dy = abs(db), kmax;
  if (a.s4) kmax = dx;
  else if (xs > 0) kmax = dx + 1;
  else if (dx > 0) kmax = 0;
  else { kmax = 0x3fffffff; dx = 1; dy = 1; }
  const uint2 e = a.divtab[min(dx, a.ndiv - 1)];
  const int stepA = xs * (steep ? row : idx_scale), stepB = ys * (steep ? idx_scale : row);
  uint32_t off = 0;
  if (mode == DF_WEDGE8) {
    const uint32_t P = (uint32_t)a.wedge_plane;
    off = steep ? 4 * WEDGE_BINS * P : 0;
    off += xs < 0 ? 2 * WEDGE_BINS * P : 0;
    off += ys < 0 ? WEDGE_BINS * P : 0;
    const int dyb = dy * WEDGE_BINS;
    for (int t = 1; t < WEDGE_BINS; t++) off += (dyb >= t * dx) ? P : 0;
  }
  const uint32_t flags = (steep ? 1u : 0u) | (xs < 0 ? 2u : 0u) | (ys < 0 ? 4u : 0u);
  recs[2 * i] = make_uint4((uint32_t)dy, (uint32_t)(-dx), (uint32_t)kmax, e.x);
  recs[2 * i + 1] = make_uint4((e.y & 31u) | (flags << 8), (uint32_t)stepA, (uint32_t)stepB, off);
}

cudaError_t launch_ring_records(const SensorLaunch &a, int mode, const short2 *ents, int n_entries, uint4 *recs,
                                cudaStream_t s) {
  if (n_entries <= 0) return cudaSuccess;
  ring_records_kernel<<<(n_entries + 255) / 256, 256, 0, s>>>(a, mode, ents, n_entries, recs);
  return cudaGetLastError();
}

static size_t table_bytes(const SensorLaunch &a) { return (size_t)a.nb * 24 + (size_t)a.ndiv * 8; }
static size_t ring_bytes(const SensorLaunch &a) { return 16 + (size_t)a.ring_nrows * 16 + (size_t)a.ring_entries * 32; }

template <int MODE, bool DIAG, bool WIDE, bool TS, bool RS>
static cudaError_t launch_t(const amclj_ctx *c, const SensorLaunch &a, cudaStream_t s) {
  auto kern = sensor_kernel<MODE, DIAG, WIDE, TS, RS>;
  size_t smem = (MODE == DF_SMEM4 ? a.df_bytes : 0) + (TS ? table_bytes(a) : 0) + (RS ? ring_bytes(a) : 0);
  if (smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (MODE != DF_SMEM4 && c->walk_carveout >= 0) { /* the rest of the SM's 256 KB serves the field as L1 */
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, c->walk_carveout > 100 ? 100 : c->walk_carveout);
    if (e != cudaSuccess) return e;
  }
  /* one CTA per SM (the staged field owns the SM's shared memory); warps sized to the work */
  int64_t ntasks = ((a.n + 31) / 32) * a.chunks;
  int64_t want_warps = (ntasks + c->sm_count - 1) / c->sm_count;
  const int max_wpb = mode_threads(MODE) / 32;
  int wpb = (int)(want_warps < 4 ? 4 : (want_warps > max_wpb ? max_wpb : want_warps));
  int64_t blocks = (ntasks + wpb - 1) / wpb;
  int per_sm = MODE == DF_SMEM4 ? 1 : (MODE == DF_WEDGE8 ? WEDGE_CTAS : (GLOBAL_CTAS > 2 ? GLOBAL_CTAS : 2));
  if (blocks > (int64_t)c->sm_count * per_sm) blocks = (int64_t)c->sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  if (c->time_main) cudaEventRecord(c->ev[5], s);
  kern<<<(unsigned)blocks, wpb * 32, smem, s>>>(a);
  if (c->time_main) cudaEventRecord(c->ev[6], s);
  return cudaGetLastError();
}

template <int MODE, bool TS, bool RS>
static cudaError_t launch_w(const amclj_ctx *c, const SensorLaunch &a, bool diag, bool wide,
                            cudaStream_t s) {
  if (diag) return wide ? launch_t<MODE, true, true, TS, RS>(c, a, s) : launch_t<MODE, true, false, TS, RS>(c, a, s);
  return wide ? launch_t<MODE, false, true, TS, RS>(c, a, s) : launch_t<MODE, false, false, TS, RS>(c, a, s);
}
template <int MODE>
static cudaError_t launch_m(const amclj_ctx *c, const SensorLaunch &a, bool diag, bool wide,
                            cudaStream_t s) {
  /* tables ride in shared memory behind the field when the SM has room for them; so do the ring
   * rows and set-up records (CTAs per SM share the budget) */
  const size_t field = MODE == DF_SMEM4 ? a.df_bytes : 0;
  const size_t per_sm = MODE == DF_SMEM4 ? 1 : (MODE == DF_WEDGE8 ? WEDGE_CTAS : (GLOBAL_CTAS > 2 ? GLOBAL_CTAS : 2));
  const size_t budget = (size_t)c->max_smem_optin / per_sm;
  bool ts = field + table_bytes(a) + 64 <= (size_t)c->max_smem_optin;
  bool rs = ts && a.ring_ok && a.ring_recs && field + table_bytes(a) + ring_bytes(a) + 1024 <= budget;
  if (rs) return launch_w<MODE, true, true>(c, a, diag, wide, s);
  return ts ? launch_w<MODE, true, false>(c, a, diag, wide, s) : launch_w<MODE, false, false>(c, a, diag, wide, s);
}

/* beams of one particle group are split into `chunks` warp tasks when there are too few
 * particles to give every resident warp a few tasks */
int sensor_chunks(const amclj_ctx *c, int64_t n, int nb) {
  int64_t groups = (n + 31) / 32;
  int64_t slots = (int64_t)c->sm_count * 32;
  int64_t want = (4 * slots + groups - 1) / groups;
  int64_t cap = (nb + SEG - 1) / SEG; /* a task is at least one segment of beams */
  if (want > cap) want = cap;
  if (want > 16) want = 16;
  return want < 1 ? 1 : (int)want;
}

cudaError_t launch_sensor(const amclj_ctx *c, const SensorLaunch &a_in, int mode, bool diag,
                          cudaStream_t s) {
  if (a_in.n <= 0 || a_in.nb <= 0) return cudaSuccess;
  SensorLaunch a = a_in;
  bool wide = a.div_wide != 0;
  float *final_raw = a.raw;
  int nseg = (a.nb + SEG - 1) / SEG;
  if (a.chunks > nseg) a.chunks = nseg;
  if (a.chunks < 1 || !a.partials) a.chunks = 1;
  cudaError_t e;
  switch (mode) {
    case DF_SMEM4: e = launch_m<DF_SMEM4>(c, a, diag, wide, s); break;
    case DF_GLOBAL8: e = launch_m<DF_GLOBAL8>(c, a, diag, wide, s); break;
    case DF_WEDGE8: e = launch_m<DF_WEDGE8>(c, a, diag, wide, s); break;
    default: e = launch_m<DF_GLOBAL4>(c, a, diag, wide, s); break;
  }
  if (e != cudaSuccess) return e;
  if (a.chunks > 1 && final_raw) {
    combine_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a.partials, a.chunks, a.n, final_raw,
                                                                 a.rawmax_ord, a.skip_flag);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace amclj
