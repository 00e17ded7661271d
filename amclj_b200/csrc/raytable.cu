// raytable.cu — K2b: the laser measurement model for DENSE particle clouds (pose tracking).
//
// Replaces the same reference code as sensor.cu (pf.clj:148-154 -> sensor.clj:165-173 -> :141-162
// -> :130-139 map-range -> :112-128 extract-line), with the same results bit for bit.
//
// map-range turns both ends of a ray into integer cells first (sensor.clj:134-135, map.clj:36-46)
// and everything after that — steep?, the swaps, calc-step, the whole extract-line loop — sees
// only those four integers.  So the range of a ray is a pure function of (start cell, end cell).
// A tracking cloud puts thousands of particles on each start cell, and every end cell of a ray of
// r cells lies in a thin ring around it (|(ex, ey)| within sqrt(2) of r: both ends are rounded
// independently).  Per update:
//   1. rt_hist_plan_kernel: histogram of the particles over the start cells of the cloud's bounding
//      box; its last CTA plans the update — cells holding at least `thr` particles are "dense" and
//      get a table, and the walking order is [dense cells, cell by cell | the rest];
//   2. rt_scatter_fill_kernel: the first CTAs finish the counting sort, the others walk ONE ray per
//      (dense cell, ring entry) with the same distance-field walk as sensor.cu and store its range:
//      a table per dense cell (lgfloor, 5.6 m rays: 1,922 entries = 7.7 KB);
//   3. rt_lookup_kernel scores every (particle, beam): rotate the beam by the particle's heading,
//      round the end point exactly as sensor.cu does, look the range up, beam mixture.  Particles of
//      sparse cells, and any end cell outside the ring (never seen; tests assert it), walk their ray
//      directly — the table is an accelerator, never an approximation;
//   4. rt_finish_kernel: fixed-point accumulators -> raw weights and their maximum.
// Nothing is kept from one update to the next: tables are rebuilt from the map every time.
// Weights are accumulated per 30-beam segment in 32.32 fixed point exactly as in sensor.cu, so the
// two paths agree bit for bit and the policy that picks one never changes a result.
#include "internal.h"
#include "sensor_common.cuh"
#include "raycast.cuh"

namespace amclj {

constexpr int RT_THREADS = 256;
static_assert(RT_PI == 32, "a work item is one warp of particles");
#ifndef RT_CTAS
#define RT_CTAS 4
#endif

__device__ __forceinline__ float ord_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// The cloud's bounding box in cells (the box itself comes from the motion kernel / bbox_kernel as
// four ordered-uint maxima).  round_cell is monotone, so every particle's start cell lies inside.
__device__ __forceinline__ bool cell_box(const uint32_t *bbox, float res, int &cx0, int &cy0, int &bw, int &bh) {
  float xmax = ord_to_float(bbox[0]), xmin = -ord_to_float(bbox[1]);
  float ymax = ord_to_float(bbox[2]), ymin = -ord_to_float(bbox[3]);
  if (!(fabsf(xmax) < 1.0e30f && fabsf(xmin) < 1.0e30f && fabsf(ymax) < 1.0e30f && fabsf(ymin) < 1.0e30f))
    return false; /* empty box, NaN or infinite coordinates */
  long long xa = round_cell_start(xmin, res), xb = round_cell_start(xmax, res);
  long long ya = round_cell_start(ymin, res), yb = round_cell_start(ymax, res);
  long long w = xb - xa + 1, h = yb - ya + 1;
  if (w <= 0 || h <= 0 || w > RT_BIN_CAP || h > RT_BIN_CAP || w * h > RT_BIN_CAP) return false;
  cx0 = (int)xa; cy0 = (int)ya; bw = (int)w; bh = (int)h;
  return true;
}

__device__ __forceinline__ int bin_of(float px, float py, float res, int cx0, int cy0, int bw, int bh) {
  int cx = min(max(round_cell_start(px, res) - cx0, 0), bw - 1), cy = min(max(round_cell_start(py, res) - cy0, 0), bh - 1);
  return cy * bw + cx;
}

// Counting sort of the particles by start cell.  A tracking cloud piles thousands of particles on
// a handful of cells, so counts (and later ranks) are taken in shared memory per CTA and only one
// atomic per (CTA, occupied cell) reaches L2.
constexpr int RT_SORT_THREADS = 512, RT_SORT_PER_THREAD = 4;
constexpr int RT_SORT_TILE = RT_SORT_THREADS * RT_SORT_PER_THREAD;

struct PlanArgs {
  int64_t n;
  int32_t thr, max_slots, max_items, force;
  RtSlot *slots;
  int32_t *item0;
  RtCtl *ctl;
};

// The plan (run by the last CTA of the histogram): classify the cells, decide, and lay the walking
// order out as [particles of dense cells, cell by cell | particles of sparse cells].  Dense cell k
// gets table slot k and the work items [item0[k], item0[k+1]) of 32 particles each.
__device__ void rt_plan(const PlanArgs &pa, uint32_t *bins, int cx0, int cy0, int bw, int bh, int4 *s_warp, int4 *s_total) {
  const int nthreads = blockDim.x, nbins = bw * bh, per = (nbins + nthreads - 1) / nthreads;
  const int lo = min((int)threadIdx.x * per, nbins), hi = min(lo + per, nbins);
  int4 v = make_int4(0, 0, 0, 0); /* dense particles, dense items, dense cells, sparse particles */
  for (int b = lo; b < hi; b++) {
    int c = (int)__ldcg(bins + b); /* the other CTAs' counts, straight from L2 */
    if (c >= pa.thr) { v.x += c; v.y += (c + RT_PI - 1) / RT_PI; v.z += 1; }
    else v.w += c;
  }
  /* block-wide exclusive scan of the four counters */
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = nthreads >> 5;
  int4 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int tx = __shfl_up_sync(0xffffffffu, inc.x, o), ty = __shfl_up_sync(0xffffffffu, inc.y, o);
    int tz = __shfl_up_sync(0xffffffffu, inc.z, o), tw = __shfl_up_sync(0xffffffffu, inc.w, o);
    if (lane >= o) { inc.x += tx; inc.y += ty; inc.z += tz; inc.w += tw; }
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int4 w = lane < nwarps ? s_warp[lane] : make_int4(0, 0, 0, 0), wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int tx = __shfl_up_sync(0xffffffffu, wi.x, o), ty = __shfl_up_sync(0xffffffffu, wi.y, o);
      int tz = __shfl_up_sync(0xffffffffu, wi.z, o), tw = __shfl_up_sync(0xffffffffu, wi.w, o);
      if (lane >= o) { wi.x += tx; wi.y += ty; wi.z += tz; wi.w += tw; }
    }
    if (lane == 31) *s_total = wi;
    s_warp[lane] = make_int4(wi.x - w.x, wi.y - w.y, wi.z - w.z, wi.w - w.w);
  }
  __syncthreads();
  const int4 base = s_warp[wid], tot = *s_total;
  int run_dn = base.x + inc.x - v.x, run_di = base.y + inc.y - v.y, run_dc = base.z + inc.z - v.z,
      run_sn = base.w + inc.w - v.w;
  /* tables pay when most particles sit in dense cells (or when the caller forces the path) */
  const bool dense = tot.z > 0 && tot.z <= pa.max_slots && tot.y <= pa.max_items &&
                     (pa.force || 2ll * tot.x >= (long long)pa.n);
  RtCtl *ctl = pa.ctl;
  if (threadIdx.x == 0) {
    ctl->dense = dense ? 1 : 0;
    ctl->n_dense = dense ? tot.x : 0;
    ctl->n_items = dense ? tot.y : 0;
    ctl->n_cells = dense ? tot.z : 0;
    ctl->cx0 = cx0; ctl->cy0 = cy0; ctl->bw = bw; ctl->bh = bh;
    ctl->hist_done = 0; /* ready for the next update */
    ctl->next_tail = 0;
    if (dense) pa.item0[tot.z] = tot.y;
  }
  if (!dense) return;
  for (int b = lo; b < hi; b++) {
    int c = (int)__ldcg(bins + b);
    if (c >= pa.thr) {
      RtSlot sl;
      sl.cx = cx0 + b % bw; sl.cy = cy0 + b / bw; sl.first = run_dn; sl.cnt = c;
      pa.slots[run_dc] = sl;
      pa.item0[run_dc] = run_di;
      bins[b] = (uint32_t)run_dn;
      run_dn += c;
      run_di += (c + RT_PI - 1) / RT_PI;
      run_dc++;
    } else {
      bins[b] = (uint32_t)(tot.x + run_sn);
      run_sn += c;
    }
  }
}

__global__ void __launch_bounds__(RT_SORT_THREADS)
rt_hist_plan_kernel(const float *__restrict__ x, const float *__restrict__ y, float res,
                    const uint32_t *__restrict__ bbox, uint32_t *bins, long long *qacc, const PlanArgs pa) {
  extern __shared__ uint32_t sh_bins[];
  __shared__ int4 s_warp[32];
  __shared__ int4 s_total;
  __shared__ uint32_t s_ticket;
  const int64_t n = pa.n, base = (int64_t)blockIdx.x * RT_SORT_TILE;
#pragma unroll
  for (int k = 0; k < RT_SORT_PER_THREAD; k++) { /* the weight accumulators of this update */
    int64_t i = base + k * RT_SORT_THREADS + threadIdx.x;
    if (i < n) qacc[i] = 0;
  }
  int cx0, cy0, bw, bh;
  if (!cell_box(bbox, res, cx0, cy0, bw, bh)) { /* every CTA sees the same box */
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      RtCtl *ctl = pa.ctl;
      ctl->dense = 0; ctl->n_dense = 0; ctl->n_items = 0; ctl->n_cells = 0; ctl->next_tail = 0;
    }
    return;
  }
  const int nbins = bw * bh;
  for (int b = threadIdx.x; b < nbins; b += RT_SORT_THREADS) sh_bins[b] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RT_SORT_PER_THREAD; k++) {
    int64_t i = base + k * RT_SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh_bins[bin_of(__ldg(x + i), __ldg(y + i), res, cx0, cy0, bw, bh)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += RT_SORT_THREADS) {
    uint32_t c = sh_bins[b];
    if (c) atomicAdd(&bins[b], c);
  }
  /* the last CTA to get here has every count in L2 and plans the update */
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&pa.ctl->hist_done, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  rt_plan(pa, bins, cx0, cy0, bw, bh, s_warp, &s_total);
}

// Second half of the sort (the first CTAs) and the table fill (the others) in one launch: both
// need only the plan.  Scatter: rank inside this CTA's share of each cell in shared memory, one
// reservation per occupied cell.  Fill: one ray per (dense cell, ring entry); neighbouring threads
// take neighbouring end cells of the same start cell: near-identical rays, converged warps, probes
// that share cache lines.
template <int MODE, bool WIDE, bool DIAG>
__global__ void __launch_bounds__(RT_SORT_THREADS, 2)
rt_scatter_fill_kernel(const __grid_constant__ RtLaunch L, int scatter_blocks) {
  const RtCtl *ctl = L.ctl;
  if (ctl->dense != 1) return;
  const SensorLaunch &a = L.s;
  if ((int)blockIdx.x < scatter_blocks) {
    extern __shared__ uint32_t sh_bins[];
    const int cx0 = ctl->cx0, cy0 = ctl->cy0, bw = ctl->bw, bh = ctl->bh, nbins = bw * bh;
    for (int b = threadIdx.x; b < nbins; b += RT_SORT_THREADS) sh_bins[b] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RT_SORT_TILE;
    int bin[RT_SORT_PER_THREAD];
    uint32_t rank[RT_SORT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < RT_SORT_PER_THREAD; k++) {
      int64_t i = base + k * RT_SORT_THREADS + threadIdx.x;
      bin[k] = -1;
      if (i < a.n) {
        bin[k] = bin_of(__ldg(a.x + i), __ldg(a.y + i), a.res, cx0, cy0, bw, bh);
        rank[k] = atomicAdd(&sh_bins[bin[k]], 1u);
      }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += RT_SORT_THREADS) {
      uint32_t c = sh_bins[b];
      if (c) sh_bins[b] = atomicAdd(&L.bins[b], c);
    }
    __syncthreads();
    /* walking order, and next to it what every look-up of the particle starts from: the ray origin
     * (S10: the laser, sensor.clj:78-81) and the heading's sine and cosine, once instead of per segment */
#pragma unroll
    for (int k = 0; k < RT_SORT_PER_THREAD; k++) {
      if (bin[k] < 0) continue;
      const int64_t i = base + k * RT_SORT_THREADS + threadIdx.x;
      const uint32_t pos = sh_bins[bin[k]] + rank[k];
      float ox = __ldg(a.x + i), oy = __ldg(a.y + i), st, ct;
      sincos_det(__ldg(a.th + i), st, ct);
      if (a.s10) {
        float nx = ox + fmaf(a.laser_x, ct, -(a.laser_y * st));
        float ny = oy + fmaf(a.laser_x, st, a.laser_y * ct);
        ox = nx; oy = ny;
      }
      L.perm[pos] = (int32_t)i;
      L.pstate[pos] = make_float4(ox, oy, st, ct);
    }
    return;
  }
  const int64_t total = (int64_t)ctl->n_cells * L.n_entries;
  const int64_t stride = (int64_t)(gridDim.x - scatter_blocks) * blockDim.x;
  unsigned long long n_probes = 0, n_oob = 0;
  for (int64_t w = (int64_t)(blockIdx.x - scatter_blocks) * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int slot = (int)(w / L.n_entries), e = (int)(w - (int64_t)slot * L.n_entries);
    const short2 en = __ldg(L.ents + e);
    if (en.x == -32768) continue; /* the unused slot of a merged row */
    if (en.x == -32767) { /* the sentinel behind the ring: what the fast path reads for an end cell off it */
      const float nanv = __int_as_float(0x7fc00000);
      if (DIAG) reinterpret_cast<float4 *>(L.tables)[w] = make_float4(nanv, nanv, nanv, nanv);
      else reinterpret_cast<float *>(L.tables)[w] = nanv;
      continue;
    }
    const int cx = L.slots[slot].cx, cy = L.slots[slot].cy;
    RayOut o;
    const float rng = cast_ray<MODE, WIDE, DIAG>(a, a.df, cx, cy, cx + en.x, cy + en.y, o);
    if (DIAG) {
      reinterpret_cast<float4 *>(L.tables)[w] =
          make_float4(rng, __int_as_float(o.kk | (o.hit << 31)), __int_as_float(o.hx), __int_as_float(o.hy));
      n_probes += (unsigned)o.probes; n_oob += (unsigned)o.oob;
    } else {
      reinterpret_cast<float *>(L.tables)[w] = rng;
    }
  }
  if (DIAG && a.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_probes += __shfl_xor_sync(FULL, n_probes, o);
      n_oob += __shfl_xor_sync(FULL, n_oob, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(a.stats + 2, n_probes);
      atomicAdd(a.stats + 6, n_oob);
    }
  }
}

// One 30-beam segment of one particle, every range from the table of the particle's start cell.
// Straight-line per beam (the compiler interleaves two beams): an end cell outside the ring reads
// the table's NaN entry, the segment comes back NaN and the caller redoes it on the general path
// (a NaN scan range takes the same detour and gets the same answer as before, only slower).
template <bool S7, bool TS>
__device__ __forceinline__ float table_segment(const SensorLaunch &a, const RtLaunch &L, const ResDiv &rdv,
                                               const float *__restrict__ table, uint32_t s_mix, uint32_t s_dir,
                                               uint32_t s_rows, const float4 *mixt, const float2 *dirt, float ox, float oy,
                                               float st, float ct, float th, const RayStart &rst, int jb0, int jb1) {
  float acc = 0.0f, prod = 1.0f;
  int pexp = 0;
  const uint32_t last_row = (uint32_t)L.nrows - 1u, nan_entry = (uint32_t)L.n_entries - 1u;
  /* S1 off = rotation by zero: fma(1, cb, -(0 * sb)) and fma(0, cb, 1 * sb) give back (cb, sb) */
  const float rc = a.s1 ? ct : 1.0f, rs = a.s1 ? st : 0.0f;
#pragma unroll 2
  for (int jb = jb0; jb < jb1; jb++) {
    const float2 bd = TS ? lds_f2(s_dir + (uint32_t)jb * 8u) : __ldg(dirt + jb);
    const float c = fmaf(rc, bd.x, -(rs * bd.y)), s = fmaf(rs, bd.x, rc * bd.y); /* S1: angle addition */
    /* sensor.clj:130-135: the end cell relative to the start cell — one float FMA per coordinate, the reference's
     * double computation next to a .5 boundary (contract.h) */
    int ex, ey;
    if (a.cr.K != 0.0f) {
      if (end_offsets_fast(c, s, rst.fKx, rst.fKy, a.cr, ex, ey)) {
        const int2 e = end_cells_exact(a, ox, oy, th, jb);
        ex = e.x - rst.x0; ey = e.y - rst.y0;
      }
    } else { /* very long rays: plain float rounding of the end point, floor(q + 0.5) as an add with round-down
              * against 1.5 * 2^23 (a start inside a table cell keeps the quotient far from the int32 range) */
      const float ex_w = fmaf(a.R, c, ox), ey_w = fmaf(a.R, s, oy);
      const float qx0 = ex_w * rdv.r, qy0 = ey_w * rdv.r;
      const float qx = fmaf(fmaf(-qx0, rdv.res, ex_w), rdv.r, qx0);
      const float qy = fmaf(fmaf(-qy0, rdv.res, ey_w), rdv.r, qy0);
      ex = __float_as_int(__fadd_rd(qx + 0.5f, 12582912.0f)) - 0x4B400000 - rst.x0;
      ey = __float_as_int(__fadd_rd(qy + 0.5f, 12582912.0f)) - 0x4B400000 - rst.y0;
    }
    const uint32_t ry = (uint32_t)(ey + L.E);
    /* ring entry of the end cell; off the ring: the NaN entry, which poisons the segment */
    const uint32_t ryc = min(ry, last_row);
    const uint4 rw = TS ? lds_u4(s_rows + ryc * 16u) : __ldg(L.rows + ryc);
    const uint32_t tt = (uint32_t)abs(ex) - rw.y; /* row = {first entry, lo, cnt, first entry of the negative arc} */
    const bool ok = tt < rw.z && ry <= last_row;
    const uint32_t ent = ok ? tt + (ex < 0 ? rw.w : rw.x) : nan_entry;
    const float rng = __ldg(table + ent);
    /* sensor.clj:149-162 beam mixture, exactly as in sensor.cu */
    const float4 mx = TS ? lds_f4(s_mix + (uint32_t)jb * 16u) : __ldg(mixt + jb);
    const float z = mx.x - rng;
    const float pz1 = hit_gauss(z, a.k2, a.z_hit);
    const float pz2 = z < 0.0f ? mx.y : 0.0f;
    const float sum = (pz1 + pz2) + mx.z;
    if (S7) logprod_mul(acc, prod, pexp, sum);
    else acc += (sum * sum) * sum;
  }
  return S7 ? logprod_close(acc, prod, pexp) : acc;
}

// Warp task ("unit") = 32 particles x one 30-beam segment; lane = particle.  Units over dense cells
// take their ranges from that cell's table; units over the particles of sparse cells walk.  Every
// CTA owns a contiguous run of the dense units (so at any time an SM works on one or two cells and
// their tables stay in L1) plus every G-th sparse unit, longest first; its warps draw from that list
// through a shared counter.  Segment sums are added into a 64-bit fixed-point accumulator per
// particle (integer atomics: order-free), which rt_finish_kernel turns into the raw weight.
template <int MODE, bool WIDE, bool DIAG, bool TS>
__global__ void __launch_bounds__(RT_THREADS, RT_CTAS)
rt_lookup_kernel(const __grid_constant__ RtLaunch L) {
  const SensorLaunch &a = L.s;
  const RtCtl *ctl = L.ctl;
  const bool dense = ctl->dense == 1;
  if (!dense && !L.must_run) return;
  extern __shared__ uint4 smem_dyn[];
  __shared__ int s_next;
  const int nb = a.nb, nseg = (nb + SEG - 1) / SEG;
  const float4 *mixt = reinterpret_cast<const float4 *>(a.beam);
  const float2 *dirt = reinterpret_cast<const float2 *>(a.beam + 4 * (size_t)nb);
  uint32_t s_mix = 0, s_dir = 0, s_rows = 0;
  if (threadIdx.x == 0) s_next = 0;
  if (TS) { /* [mix float4[nb] | rows uint4[nrows] | dir float2[nb]] */
    float4 *smix = reinterpret_cast<float4 *>(smem_dyn);
    uint4 *srows = reinterpret_cast<uint4 *>(smix + nb);
    float2 *sdir = reinterpret_cast<float2 *>(srows + L.nrows);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) { smix[i] = __ldg(mixt + i); sdir[i] = __ldg(dirt + i); }
    for (int i = threadIdx.x; i < L.nrows; i += blockDim.x) srows[i] = __ldg(L.rows + i);
    s_mix = smem_u32(smix); s_dir = smem_u32(sdir); s_rows = smem_u32(srows);
    asm volatile("" : "+r"(s_mix), "+r"(s_dir), "+r"(s_rows));
  }
  __syncthreads();
  if (a.decision && blockIdx.x == 0 && threadIdx.x == 0)
    *a.decision = dense ? 2 : ((a.bbox && cloud_is_compact(a.bbox, a.compact_m)) ? 1 : 0);
  const int64_t n_items_dense = dense ? ctl->n_items : 0;
  const int64_t n_dense = dense ? ctl->n_dense : 0;
  const int n_tabs = dense ? ctl->n_cells : 0;
  const int64_t n_sparse_items = (a.n - n_dense + RT_PI - 1) / RT_PI;
  const int64_t ud = n_items_dense * nseg, us = n_sparse_items * nseg;
  const int64_t G = gridDim.x, bid = blockIdx.x;
  /* the last few percent of the dense units are a common tail: whoever finishes first takes it */
  const int64_t ud_static = ud - ud * L.tail_pct / 100;
  const int64_t d_lo = ud_static * bid / G, d_hi = ud_static * (bid + 1) / G;
  const int64_t ns_mine = us > bid ? (us - bid + G - 1) / G : 0; /* sparse units bid, bid + G, ... */
  const int64_t n_mine = ns_mine + (d_hi - d_lo);
  const ResDiv rdv = make_resdiv(a.res);
  const bool fast_round = rdv.lim > 0.0f;
  const uint32_t nrows = (uint32_t)L.nrows;
  const int lane = threadIdx.x & 31;
  unsigned long long n_cells = 0, n_probes = 0, n_hits = 0, n_rays = 0, n_oob = 0, n_miss = 0;

  int c_slot = 0, c_i0 = 0, c_i1 = 0; /* the dense cell of the previous unit: [c_i0, c_i1) are its items */
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(&s_next, 1);
    k = __shfl_sync(FULL, k, 0);
    int64_t u;
    if (k < n_mine) {
      u = k < ns_mine ? ud + bid + (int64_t)k * G : d_lo + (k - ns_mine);
    } else { /* this CTA's share is done: help with the common tail */
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(&L.ctl->next_tail, 1u);
      t = __shfl_sync(FULL, t, 0);
      if ((int64_t)t >= ud - ud_static) break;
      u = ud_static + t;
    }
    const int64_t item = u / nseg;
    const int sgm = (int)(u - item * nseg);
    int64_t first;
    int cnt, cx = 0, cy = 0;
    const float *table = nullptr;
    if (item < n_items_dense) {
      if (item < c_i0 || item >= c_i1) { /* last slot whose first item is <= item */
        int lo_s = 0, hi_s = n_tabs;
        while (hi_s - lo_s > 1) {
          const int mid = (lo_s + hi_s) >> 1;
          if ((int64_t)__ldg(L.item0 + mid) <= item) lo_s = mid; else hi_s = mid;
        }
        c_slot = lo_s; c_i0 = __ldg(L.item0 + lo_s); c_i1 = __ldg(L.item0 + lo_s + 1);
      }
      const RtSlot sl = L.slots[c_slot];
      const int off = (int)(item - c_i0) * RT_PI;
      first = sl.first + off; cnt = min(RT_PI, sl.cnt - off); cx = sl.cx; cy = sl.cy;
      table = reinterpret_cast<const float *>(L.tables) + (size_t)c_slot * L.n_entries * (DIAG ? 4 : 1);
    } else {
      first = n_dense + (item - n_items_dense) * RT_PI;
      cnt = (int)min((int64_t)RT_PI, a.n - first);
    }
    if (lane >= cnt) continue;
    int64_t p;
    float ox, oy, st, ct;
    if (dense) { /* the sort left the ray origin and the heading's sine / cosine in walking order */
      p = (int64_t)__ldg(L.perm + first + lane);
      const float4 ps = __ldg(L.pstate + first + lane);
      ox = ps.x; oy = ps.y; st = ps.z; ct = ps.w;
    } else {
      p = first + lane;
      ox = __ldg(a.x + p); oy = __ldg(a.y + p);
      sincos_det(__ldg(a.th + p), st, ct);
      if (a.s10) { /* S10: base->laser offset (sensor.clj:78-81), off by default */
        float nx = ox + fmaf(a.laser_x, ct, -(a.laser_y * st));
        float ny = oy + fmaf(a.laser_x, st, a.laser_y * ct);
        ox = nx; oy = ny;
      }
    }
    const RayStart rst = make_ray_start(a, ox, oy, st, ct);
    const int x0 = rst.x0, y0 = rst.y0;
    /* the table belongs to cell (cx, cy): a particle that is not there walks (the sort and this
     * kernel round the same coordinates the same way, so this fires only with S10) */
    const bool use_table = table != nullptr && x0 == cx && y0 == cy && fast_round && abs(x0) < (1 << 21) && abs(y0) < (1 << 21) &&
                           (rst.fast || a.cr.K == 0.0f); /* a NaN heading takes the general path */
    const int jb0 = sgm * SEG, jb1 = min(jb0 + SEG, nb);
    float v = 0.0f;
    bool general = DIAG || !use_table;
    if (!general) {
      v = a.s7 ? table_segment<true, TS>(a, L, rdv, table, s_mix, s_dir, s_rows, mixt, dirt, ox, oy, st, ct, __ldg(a.th + p), rst, jb0, jb1)
               : table_segment<false, TS>(a, L, rdv, table, s_mix, s_dir, s_rows, mixt, dirt, ox, oy, st, ct, __ldg(a.th + p), rst, jb0, jb1);
      general = v != v;
    }
    if (general) { /* sparse cells, diagnostics, and the segments the fast path gave back */
      const bool dbg = DIAG && a.dbg_count > 0 && p >= a.dbg_first && p < a.dbg_first + a.dbg_count;
      float acc = 0.0f, prod = 1.0f;
      int pexp = 0;
      for (int jb = jb0; jb < jb1; jb++) {
        const float2 bd = TS ? lds_f2(s_dir + (uint32_t)jb * 8u) : __ldg(dirt + jb);
        float c = bd.x, s = bd.y;
        if (a.s1) {
          c = fmaf(ct, bd.x, -(st * bd.y));
          s = fmaf(st, bd.x, ct * bd.y);
        }
        bool ok = false;
        uint32_t ent = 0;
        if (use_table) {
          int x1, y1;
          end_cell(a, rdv, rst, c, s, ox, oy, __ldg(a.th + p), jb, x1, y1);
          const int ex = x1 - x0, ey = y1 - y0;
          const uint32_t ry = (uint32_t)(ey + L.E);
          if (ry < nrows) {
            const uint4 rw = TS ? lds_u4(s_rows + ry * 16u) : __ldg(L.rows + ry);
            const uint32_t tt = (uint32_t)abs(ex) - rw.y;
            ok = tt < rw.z;
            ent = tt + (ex < 0 ? rw.w : rw.x);
          }
        }
        float rng;
        RayOut o;
        if (ok) {
          if (DIAG) {
            const float4 te = __ldg(reinterpret_cast<const float4 *>(table) + ent);
            rng = te.x;
            const int kh = __float_as_int(te.y);
            o.kk = kh & 0x7fffffff; o.hit = (kh >> 31) & 1;
            o.hx = __float_as_int(te.z); o.hy = __float_as_int(te.w);
            o.probes = 0; o.oob = 0;
          } else {
            rng = __ldg(table + ent);
          }
        } else {
          rng = cast_ray_slow<MODE, WIDE, DIAG>(a, ox, oy, __ldg(a.th + p), jb, o);
          if (DIAG && use_table) n_miss++; /* an end cell outside the ring: never seen */
        }
        const float4 mx = TS ? lds_f4(s_mix + (uint32_t)jb * 16u) : __ldg(mixt + jb);
        const float z = mx.x - rng;
        const float pz1 = hit_gauss(z, a.k2, a.z_hit);
        const float pz2 = z < 0.0f ? mx.y : 0.0f;
        const float sum = (pz1 + pz2) + mx.z;
        if (a.s7) logprod_mul(acc, prod, pexp, sum);
        else acc += (sum * sum) * sum;
        if (DIAG) {
          n_rays++; n_cells += (unsigned)o.kk + 1u; n_probes += (unsigned)o.probes; n_hits += (unsigned)o.hit;
          n_oob += (unsigned)o.oob;
          if (dbg) {
            int64_t q = (p - a.dbg_first) * nb + jb;
            if (a.dbg_hit_x) a.dbg_hit_x[q] = o.hx;
            if (a.dbg_hit_y) a.dbg_hit_y[q] = o.hy;
            if (a.dbg_steps) a.dbg_steps[q] = o.kk;
            if (a.dbg_hit) a.dbg_hit[q] = (uint8_t)o.hit;
            if (a.dbg_range) a.dbg_range[q] = rng;
          }
        }
      }
      v = a.s7 ? logprod_close(acc, prod, pexp) : acc;
    }
    /* close the segment: one rounding to fixed point; segments add as integers */
    atomicAdd(reinterpret_cast<unsigned long long *>(L.qacc) + p,
              (unsigned long long)__float2ll_rn(fmaxf(v, FIX_FLOOR) * FIX_SCALE));
  }
  if (DIAG && a.stats) {
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_rays += __shfl_xor_sync(FULL, n_rays, o);
      n_cells += __shfl_xor_sync(FULL, n_cells, o);
      n_probes += __shfl_xor_sync(FULL, n_probes, o);
      n_hits += __shfl_xor_sync(FULL, n_hits, o);
      n_oob += __shfl_xor_sync(FULL, n_oob, o);
      n_miss += __shfl_xor_sync(FULL, n_miss, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 11, n_miss);
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.stats[8] = dense ? 16ull + (unsigned long long)MODE : (unsigned long long)MODE + 1ull;
        a.stats[9] = (unsigned long long)(dense ? ctl->n_cells : 0);
        a.stats[10] = (unsigned long long)n_dense;
      }
      atomicAdd(a.stats + 0, n_rays);
      atomicAdd(a.stats + 1, n_cells);
      atomicAdd(a.stats + 2, n_probes);
      atomicAdd(a.stats + 3, n_hits);
      atomicAdd(a.stats + 6, n_oob);
    }
  }
}

// raw[p] = the fixed-point sum of the particle's segments as a float, plus the max for the normaliser
__global__ void __launch_bounds__(256)
rt_finish_kernel(const RtLaunch L) {
  const SensorLaunch &a = L.s;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  /* the histogram of the next update starts from zero (the scatter left write cursors behind) */
  for (int64_t b = i; b < RT_BIN_CAP; b += (int64_t)gridDim.x * blockDim.x) L.bins[b] = 0;
  if (L.ctl->dense != 1 && !L.must_run) return;
  float v = -INFINITY;
  if (i < a.n) {
    v = fixed_to_raw(L.qacc[i]);
    if (a.raw) a.raw[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  if ((threadIdx.x & 31) == 0 && a.rawmax_ord && v > -INFINITY) atomicMax(a.rawmax_ord, ord_of_float(v));
}

template <int MODE, bool WIDE, bool DIAG>
static cudaError_t launch_rt_t(const amclj_ctx *c, const RtLaunch &L, cudaStream_t s) {
  /* persistent grids: the amount of work is decided on the device */
  const int64_t grid = (int64_t)c->sm_count * RT_CTAS;
  const int64_t grid_lookup = grid * (c->rt.waves > 0 ? c->rt.waves : 1);
  const int scatter_blocks = (int)((L.s.n + RT_SORT_TILE - 1) / RT_SORT_TILE);
  const size_t sort_smem = sizeof(uint32_t) * RT_BIN_CAP;
  auto sf = rt_scatter_fill_kernel<MODE, WIDE, DIAG>;
  cudaError_t e = cudaFuncSetAttribute(sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem);
  if (e != cudaSuccess) return e;
  sf<<<(unsigned)(scatter_blocks + c->sm_count * 2), RT_SORT_THREADS, sort_smem, s>>>(L, scatter_blocks);
  /* beam tables and ring rows ride in shared memory when RT_CTAS CTAs of them fit the SM */
  const size_t smem = (size_t)L.s.nb * 24 + (size_t)L.nrows * 16;
  if (c->time_main) cudaEventRecord(c->ev[5], s);
  if (smem * RT_CTAS + 4096 <= (size_t)c->max_smem_optin) {
    auto kern = rt_lookup_kernel<MODE, WIDE, DIAG, true>;
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid_lookup, RT_THREADS, smem, s>>>(L);
  } else {
    rt_lookup_kernel<MODE, WIDE, DIAG, false><<<(unsigned)grid_lookup, RT_THREADS, 0, s>>>(L);
  }
  if (c->time_main) cudaEventRecord(c->ev[6], s);
  rt_finish_kernel<<<(unsigned)((L.s.n + 255) / 256), 256, 0, s>>>(L);
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_rt_m(const amclj_ctx *c, const RtLaunch &L, bool diag, bool wide, cudaStream_t s) {
  if (diag) return wide ? launch_rt_t<MODE, true, true>(c, L, s) : launch_rt_t<MODE, false, true>(c, L, s);
  return wide ? launch_rt_t<MODE, true, false>(c, L, s) : launch_rt_t<MODE, false, false>(c, L, s);
}

/* once per device: the histogram takes 64 KB of dynamic shared memory */
cudaError_t rt_prepare_device() {
  const int sort_smem = (int)(sizeof(uint32_t) * RT_BIN_CAP);
  return cudaFuncSetAttribute(rt_hist_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sort_smem);
}

cudaError_t launch_raytable(const amclj_ctx *c, const RtLaunch &L, int mode, bool diag, cudaStream_t s) {
  const SensorLaunch &a = L.s;
  if (a.n <= 0 || a.nb <= 0) return cudaSuccess;
  const size_t sort_smem = sizeof(uint32_t) * RT_BIN_CAP; /* a counter per cell of the largest box */
  const unsigned blocks = (unsigned)((a.n + RT_SORT_TILE - 1) / RT_SORT_TILE);
  PlanArgs pa;
  pa.n = a.n; pa.thr = L.thr; pa.max_slots = L.max_slots; pa.max_items = L.max_items; pa.force = L.must_run == 2 ? 1 : 0;
  pa.slots = L.slots; pa.item0 = L.item0; pa.ctl = L.ctl;
  rt_hist_plan_kernel<<<blocks, RT_SORT_THREADS, sort_smem, s>>>(a.x, a.y, a.res, a.bbox, L.bins, L.qacc, pa);
  const bool wide = a.div_wide != 0;
  switch (mode) {
    case DF_WEDGE8: return launch_rt_m<DF_WEDGE8>(c, L, diag, wide, s);
    case DF_GLOBAL8: return launch_rt_m<DF_GLOBAL8>(c, L, diag, wide, s);
    default: return launch_rt_m<DF_GLOBAL4>(c, L, diag, wide, s);
  }
}

}  // namespace amclj
