// ptable.cu — K2c: the laser measurement model on PERSISTENT, map-wide ray tables.
//
// Replaces the same reference code as sensor.cu (pf.clj:148-154 -> sensor.clj:165-173 -> :141-162
// -> :130-139 map-range -> :112-128 extract-line), with the same results bit for bit.
//
// map-range rounds both ends of a ray to integer cells before anything else looks at them
// (sensor.clj:134-135, map.clj:36-46): the range of a ray is a pure function of (start cell, end
// cell), and the map does not change between updates.  So the tables of raytable.cu need not be
// rebuilt every update for the cells a dense cloud happens to stand on: ONE table per free cell
// of the map (lgfloor: 78,513 cells x 1,923 end cells of a 5.6 m ray = 604 MB of float ranges,
// filled once in a few milliseconds by walking 151 M rays with sensor.cu's own distance-field
// walk) serves every later update of every cloud.  Per update:
//   1. pt_hist_kernel / pt_scan_kernel / pt_scatter_kernel: counting sort of the particles by the
//      table slot of their start cell (warp-aggregated atomics, so a tracking cloud that piles
//      thousands of particles on one cell does not serialise), leaving next to the permutation what
//      every look-up of a particle starts from (ray origin, sine and cosine of the heading);
//   2. pt_lookup_kernel: every warp owns a contiguous run of the sorted particles and walks through
//      the cells of that run: the cell's table comes into the warp's own shared-memory buffer by one
//      bulk copy (cp.async.bulk + mbarrier, the next cell's copy in flight while this one is used),
//      and the (particle, 30-beam segment) pairs of the cell are spread over the lanes: rotate the
//      beam by the heading, round the end point exactly as sensor.cu does, ring index, one LDS,
//      beam mixture.  Each table is read from HBM once per update;
//   3. pt_finish_kernel: fixed-point accumulators -> raw weights and their maximum.
// Segment sums are rounded once to 32.32 fixed point and added as integers exactly as in
// sensor.cu / raytable.cu, so the three paths agree bit for bit and which one ran can never be
// seen in a result.  An end cell off the ring reads the table's NaN entry and the segment is
// redone by walking (never seen; the table is an accelerator, never an approximation).
#include "internal.h"
#include "raycast.cuh"
#include "sensor_common.cuh"

namespace amclj {

constexpr int PT_SORT_THREADS = 256;

/* ray origin (S10: the laser, sensor.clj:78-81) and heading sine / cosine of particle i */
__device__ __forceinline__ float4 pt_origin(const SensorLaunch &a, int64_t i) {
  float ox = __ldg(a.x + i), oy = __ldg(a.y + i), st, ct;
  sincos_det(__ldg(a.th + i), st, ct);
  if (a.s10) {
    float nx = ox + fmaf(a.laser_x, ct, -(a.laser_y * st));
    float ny = oy + fmaf(a.laser_x, st, a.laser_y * ct);
    ox = nx; oy = ny;
  }
  return make_float4(ox, oy, st, ct);
}

// ---- one-time fill: one ray per (free cell, ring entry) ----------------------------------------
template <int MODE, bool WIDE>
__global__ void __launch_bounds__(256)
pt_fill_kernel(const __grid_constant__ PtLaunch L) {
  const SensorLaunch &a = L.s;
  const int64_t total = (int64_t)L.n_slots * L.stride;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float nanv = __int_as_float(0x7fc00000);
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int slot = (int)(w / L.stride), e = (int)(w - (int64_t)slot * L.stride);
    float v = nanv; /* padding, the unused slot of a merged row, and the sentinel behind the ring */
    if (e < L.n_entries) {
      const short2 en = __ldg(L.ents + e);
      if (en.x > -32767) {
        const int2 c = __ldg(L.cell_of_slot + slot);
        RayOut o;
        v = cast_ray<MODE, WIDE, false>(a, a.df, c.x, c.y, c.x + en.x, c.y + en.y, o);
      }
    }
    L.tables[w] = v;
  }
}

// ---- per update: counting sort by table slot ---------------------------------------------------
__global__ void __launch_bounds__(PT_SORT_THREADS)
pt_hist_kernel(const __grid_constant__ PtLaunch L) {
  const SensorLaunch &a = L.s;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int key = -1 - lane; /* lanes past the end: a key nobody shares */
  if (i < a.n) {
    L.qacc[i] = 0; /* the weight accumulators of this update */
    const float4 o = pt_origin(a, i);
    int x0, y0;
    x0 = round_cell_start(o.x, a.res); y0 = round_cell_start(o.y, a.res);
    const bool in = (uint32_t)x0 < (uint32_t)a.W && (uint32_t)y0 < (uint32_t)a.H;
    key = in ? __ldg(L.slot_of_cell + (size_t)y0 * a.W + x0) : L.n_slots;
    L.pkey[i] = key;
  }
  const uint32_t peers = __match_any_sync(FULL, key);
  if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(L.bin_count + key, (uint32_t)__popc(peers));
  if (blockIdx.x == 0) /* the bin scan's ticket and tile descriptors start from zero */
    for (int t = threadIdx.x; t < L.scan_words; t += blockDim.x) L.scan_state[t] = 0ull;
}

// Exclusive scan of the bin counts (n_slots + 1 bins; lgfloor: 78,514), which also re-zeroes them: single
// pass, decoupled look-back over tiles of 2048 bins (the same scheme as filter.cu's weight scan; one CTA
// walking all bins took 150 us).  scan_state = {ticket, pad, descriptor per tile}, zeroed by the histogram.
constexpr int PT_SCAN_THREADS = 256, PT_SCAN_ITEMS = 8, PT_SCAN_TILE = PT_SCAN_THREADS * PT_SCAN_ITEMS;
constexpr unsigned long long PT_AGG = 1ull << 62, PT_PREFIX = 2ull << 62, PT_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long pt_ld_volatile(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(PT_SCAN_THREADS)
pt_scan_kernel(uint32_t *bin_count, uint32_t *bin_start, uint32_t *cursor, int nbins, unsigned long long *scan_state) {
  __shared__ uint32_t warp_tot[PT_SCAN_THREADS / 32];
  __shared__ uint32_t s_excl, s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(reinterpret_cast<uint32_t *>(scan_state), 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  unsigned long long *desc = scan_state + 1;
  const int base = (int)tile * PT_SCAN_TILE + (int)threadIdx.x * PT_SCAN_ITEMS;
  uint32_t c[PT_SCAN_ITEMS];
  if (base + PT_SCAN_ITEMS <= nbins) { /* base is a multiple of 8: two 128-bit loads */
    const uint4 u = *reinterpret_cast<const uint4 *>(bin_count + base), v = *reinterpret_cast<const uint4 *>(bin_count + base + 4);
    c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w; c[4] = v.x; c[5] = v.y; c[6] = v.z; c[7] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < PT_SCAN_ITEMS; k++) c[k] = base + k < nbins ? bin_count[base + k] : 0u;
  }
  uint32_t run = 0;
#pragma unroll
  for (int k = 0; k < PT_SCAN_ITEMS; k++) { const uint32_t t = c[k]; c[k] = run; run += t; } /* exclusive inside the thread */
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  uint32_t woff = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < PT_SCAN_THREADS / 32; k++) {
    const uint32_t ws = warp_tot[k];
    if (k < wid) woff += ws;
    tile_total += ws;
  }
  if (wid == 0) {
    unsigned long long excl = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&desc[0], PT_PREFIX | tile_total);
    } else {
      if (lane == 0) atomicExch(&desc[tile], PT_AGG | tile_total);
      int look = (int)tile - 1;
      for (;;) {
        const int t = look - lane;
        unsigned long long sv;
        do {
          sv = t >= 0 ? pt_ld_volatile(&desc[t]) : PT_PREFIX;
        } while (__any_sync(FULL, (sv >> 62) == 0));
        const uint32_t is_prefix = __ballot_sync(FULL, (sv >> 62) == 2);
        const int first = is_prefix ? __ffs(is_prefix) - 1 : 32;
        unsigned long long contrib = lane <= first ? (sv & PT_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(FULL, contrib, o);
        excl += contrib;
        if (is_prefix) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(&desc[tile], PT_PREFIX | (excl + tile_total));
    }
    if (lane == 0) {
      s_excl = (uint32_t)excl;
      if ((int)(tile + 1) * PT_SCAN_TILE >= nbins) bin_start[nbins] = (uint32_t)excl + tile_total; /* = n */
    }
  }
  __syncthreads();
  const uint32_t off = s_excl + woff + inc - run;
#pragma unroll
  for (int k = 0; k < PT_SCAN_ITEMS; k++)
    if (base + k < nbins) {
      bin_start[base + k] = off + c[k];
      cursor[base + k] = off + c[k];
      bin_count[base + k] = 0;
    }
}

__global__ void __launch_bounds__(PT_SORT_THREADS)
pt_scatter_kernel(const __grid_constant__ PtLaunch L) {
  const SensorLaunch &a = L.s;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int key = i < a.n ? L.pkey[i] : -1 - lane;
  const uint32_t peers = __match_any_sync(FULL, key);
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (key >= 0 && lane == leader) base = atomicAdd(L.cursor + key, (uint32_t)__popc(peers));
  base = __shfl_sync(FULL, base, leader);
  if (key < 0) return;
  const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
  L.perm[pos] = (int32_t)i;
  /* what every look-up of this particle starts from (contract.h): the offset inside its start cell (+ K) per axis,
   * heading sine / cosine — computed once per particle instead of once per segment */
  const float4 o = pt_origin(a, i);
  float fKx, fKy;
  start_cell_frac(o.x, a.res, a.cr.K, fKx);
  start_cell_frac(o.y, a.res, a.cr.K, fKy);
  L.pstate[pos] = make_float4(fKx, fKy, o.z, o.w);
  /* for the near-tie end cells: the heading's sine and cosine in double, and the ray origin */
  double st_d = 0.0, ct_d = 1.0; /* S1 off: rotation by zero */
  if (a.s1) sincos_det_d((double)__ldg(a.th + i), st_d, ct_d);
  L.pstate_d[2 * (size_t)pos] = make_double2(st_d, ct_d);
  L.pstate_d[2 * (size_t)pos + 1] = make_double2((double)o.x, (double)o.y);
}

// ---- the look-up -------------------------------------------------------------------------------
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// One 30-beam segment of one particle, every range from the staged table of its start cell:
// the arithmetic of raytable.cu's table_segment (and of sensor.cu's set-up) with the table in
// shared memory.  DBG: also leave every range in the parity probe's buffer.
__device__ __forceinline__ uint32_t lds_u1(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

/* exact end cells (contract.h) of those of the beams jb, jb + 1 whose float offset lies next to a .5 boundary
 * (which & 1, which & 2), as (offset x, ring row) like the fast half; out of line and once per PAIR of beams, so
 * that the hot loop below stays two straight-line halves.  psd: {(sin, cos) of the heading, ray origin}. */
#ifdef AMCLJ_PT_EXACT_INLINE
#define PT_EXACT_ATTR __forceinline__
#else
#define PT_EXACT_ATTR __noinline__
#endif
static __device__ PT_EXACT_ATTR int4 pt_exact_pair(const SensorLaunch &a, const PtLaunch &L, const double2 *psd, int slot, int jb,
                                                  int which, int4 cells) {
  const double rinv_d = L.rinv_d;
  const int E = L.E;
  const double2 hd = __ldg(psd); /* S1 off: (0, 1), which gives back (cb, sb) */
  const double2 od = __ldg(psd + 1);
  const double Rd = (double)a.R, res = (double)a.res;
  /* the start cell: the table's own cell (the sort put the particle there by round_cell_start of the same origin);
   * a particle on the zero table — off the map or on a blocked cell — reads 0 whatever entry it asks for */
  const int2 c0 = slot < L.n_slots ? __ldg(L.cell_of_slot + slot) : make_int2(0, 0);
  const int x0 = c0.x, y0 = c0.y;
#pragma unroll
  for (int u = 0; u < 2; u++) {
    if (!(which & (1 << u))) continue;
    const double cbd = __ldg(a.beam_d + 2 * (jb + u)), sbd = __ldg(a.beam_d + 2 * (jb + u) + 1);
    const double cd = fma(hd.y, cbd, -(hd.x * sbd)), sd = fma(hd.x, cbd, hd.y * sbd);
    const double vx = fma(Rd, cd, od.x), vy = fma(Rd, sd, od.y);
    const int cx = cell_exact_d(vx, res, rinv_d), cy = cell_exact_d(vy, res, rinv_d); /* NaN -> -1e9: off the ring */
    if (u == 0) { cells.x = cx - x0; cells.y = cy - y0 + E; } else { cells.z = cx - x0; cells.w = cy - y0 + E; }
  }
  return cells;
}

// ---- packed fp32 (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE round-to-nearest operations per issue slot) ----------
// The look-up kernel is bound by instruction issue, and its beams go in pairs anyway: the float arithmetic of the
// two beams of a pair rides in one instruction each.  Every lane of an .f32x2 operation rounds exactly like the
// scalar operation (no fusing, no flush-to-zero), so the results stay bit-identical to sensor.cu / raytable.cu.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void upk2i(f32x2 v, int &lo, int &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, f32x2 &a, f32x2 &b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ f32x2 lds_64(uint32_t addr) {
  f32x2 a;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr));
  return a;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); /* the instruction of hit_gauss */
  return e;
}

// One 30-beam segment of one particle, every range from the staged table of its start cell.  Beams go in
// pairs: [end cells of both, packed float arithmetic] -> [one test: is either next to a .5 boundary? then both are
// re-taken in double, out of line] -> [ring index and table entry of each, packed beam mixture of both].  Each half is
// straight-line.  The beam constants are staged per PAIR of beams (pt_lookup_kernel): s_dir float4 (cos a, cos b,
// sin a, sin b), s_mixa float4 (obs a, obs b, e2 a, e2 b), s_mixb float2 (p34 a, p34 b); jb0 is even (SEG is).
static_assert(SEG % 2 == 0, "pt_segment takes its beams in pairs that start on an even beam");
template <bool S7, bool DBG, bool R32>
__device__ __forceinline__ float pt_segment(const SensorLaunch &a, const PtLaunch &L, uint32_t s_tab, uint32_t s_mixa,
                                            uint32_t s_mixb, uint32_t s_dir, uint32_t s_rows, float fKx, float fKy, float st,
                                            float ct, const double2 *psd, int slot, int jb0, int jb1, float *dbg_row) {
  float acc = 0.0f, prod = 1.0f;
  int pexp = 0;
  const uint32_t last_row = (uint32_t)L.nrows - 1u, nan_entry = (uint32_t)L.n_entries - 1u;
  /* S1 off = rotation by zero: fma(1, cb, -(0 * sb)) and fma(0, cb, 1 * sb) give back (cb, sb) */
  const float rc = a.s1 ? ct : 1.0f, rs = a.s1 ? st : 0.0f;
  f32x2 rc2 = pk2(rc, rc), rs2 = pk2(rs, rs), nrs2 = pk2(-rs, -rs), fKx2 = pk2(fKx, fKx), fKy2 = pk2(fKy, fKy);
  asm volatile("" : "+l"(rc2), "+l"(rs2), "+l"(nrs2), "+l"(fKx2), "+l"(fKy2)); /* kept as pairs: not rebuilt per pair of beams */
  const CellRefine cr = a.cr;
  const f32x2 rcell2 = pk2(cr.rc, cr.rc), nk2 = pk2(-a.k2, -a.k2);
  const int coff_row = cr.coff - L.E; /* ring rows start at offset -E */

  /* second half, per beam: ring entry of the end cell (offset ex, ring row ryi) -> range from the table */
  auto range_of = [&](int ex, int ryi) -> float {
    const uint32_t ry = (uint32_t)ryi;
    const uint32_t ryc = min(ry, R32 ? last_row + 1u : last_row);
    uint32_t r_base, r_lo, r_cnt; /* the row's first entry, first |ex|, entries per arc (the negative arc follows the positive) */
    if (R32) { /* one word per row: a random 4-byte read meets ~3.5-way bank conflicts, a 16-byte one ~8 wavefronts */
      const uint32_t rw = lds_u1(s_rows + ryc * 4u); /* base : 16 | lo : 8 | cnt : 8; the row behind the last is empty */
      r_base = rw & 0xffffu; r_lo = __byte_perm(rw, 0u, 0x4442); r_cnt = rw >> 24;
    } else {
      const uint4 rw = lds_u4(s_rows + ryc * 16u);
      r_base = rw.x; r_lo = rw.y; r_cnt = rw.z;
    }
    const uint32_t tt = (uint32_t)abs(ex) - r_lo;
    const bool ok = R32 ? tt < r_cnt : (tt < r_cnt && ry <= last_row);
    const uint32_t ent = ok ? r_base + tt + (ex < 0 ? r_cnt : 0u) : nan_entry;
    return lds_f1(s_tab + ent * 4u);
  };

  int jb = jb0;
  for (; jb + 1 < jb1; jb += 2) {
    const uint32_t k = (uint32_t)jb >> 1;
    /* first half: sensor.clj:130-135 / map.clj:36-46, the end cells relative to the start cell — one FMA per
     * coordinate (contract.h) — as (offset x, ring row) */
    f32x2 C, S;
    lds_2x64(s_dir + k * 16u, C, S);
    const f32x2 cd = fma2(rc2, C, mul2(nrs2, S)); /* S1: angle addition, c = fma(rc, cb, -(rs * sb)) */
    const f32x2 sd = fma2(rs2, C, mul2(rc2, S));  /*                     s = fma(rs, cb, rc * sb)    */
    int uxa, uxb, uya, uyb;
    upk2i(fma2(rcell2, cd, fKx2), uxa, uxb);
    upk2i(fma2(rcell2, sd, fKy2), uya, uyb);
    int xa = (uxa >> cr.fb) - cr.coff, ya = (uya >> cr.fb) - coff_row;
    int xb = (uxb >> cr.fb) - cr.coff, yb = (uyb >> cr.fb) - coff_row;
    const uint32_t na = min((uint32_t)uxa & cr.mask, (uint32_t)uya & cr.mask); /* 0: next to a .5 boundary */
    const uint32_t nb = min((uint32_t)uxb & cr.mask, (uint32_t)uyb & cr.mask);
    if (min(na, nb) == 0u) { /* 0.1 % of the pairs at fb = 14 */
      const int4 e = pt_exact_pair(a, L, psd, slot, jb, (na == 0u ? 1 : 0) | (nb == 0u ? 2 : 0), make_int4(xa, ya, xb, yb));
      xa = e.x; ya = e.y; xb = e.z; yb = e.w;
    }
    /* second half: ranges from the table, sensor.clj:149-162 beam mixture of both */
    const float rga = range_of(xa, ya), rgb = range_of(xb, yb);
    if (DBG && dbg_row) { dbg_row[jb] = rga; dbg_row[jb + 1] = rgb; }
    f32x2 OBS, E2;
    lds_2x64(s_mixa + k * 16u, OBS, E2);
    const f32x2 P34 = lds_64(s_mixb + k * 8u);
    const f32x2 Z = sub2(OBS, pk2(rga, rgb));
    float za, zb, ta, tb, e2a, e2b;
    upk2(Z, za, zb);
    upk2(mul2(mul2(Z, Z), nk2), ta, tb); /* -(z * z) * k2 */
    upk2(E2, e2a, e2b);
    /* z_hit * e stays scalar: ptxas (12.9) contracts a packed multiply that feeds a packed add into one FFMA2 —
     * whatever --fmad and the .rn modifiers say — and a fused (z_hit * e + pz2) would round differently from the
     * other sensor paths; no other packed product here feeds a sum */
    const f32x2 PZ1 = pk2(a.z_hit * ex2_approx(ta), a.z_hit * ex2_approx(tb));
    const f32x2 PZ2 = pk2(za < 0.0f ? e2a : 0.0f, zb < 0.0f ? e2b : 0.0f);
    const f32x2 SUM = add2(add2(PZ1, PZ2), P34);
    float sa, sb;
    if (S7) {
      upk2(SUM, sa, sb);
      logprod_mul(acc, prod, pexp, sa);
      logprod_mul(acc, prod, pexp, sb);
    } else {
      upk2(mul2(mul2(SUM, SUM), SUM), sa, sb);
      acc += sa;
      acc += sb;
    }
  }
  if (jb < jb1) { /* a last segment with an odd number of beams: the first half of its pair */
    const uint32_t k = (uint32_t)jb >> 1;
    const float4 bd = lds_f4(s_dir + k * 16u);
    const float c = fmaf(rc, bd.x, -(rs * bd.z)), s = fmaf(rs, bd.x, rc * bd.z);
    const int ux = __float_as_int(fmaf(cr.rc, c, fKx)), uy = __float_as_int(fmaf(cr.rc, s, fKy));
    int xa = (ux >> cr.fb) - cr.coff, ya = (uy >> cr.fb) - coff_row;
    if (min((uint32_t)ux & cr.mask, (uint32_t)uy & cr.mask) == 0u) {
      const int4 e = pt_exact_pair(a, L, psd, slot, jb, 1, make_int4(xa, ya, 0, 0));
      xa = e.x; ya = e.y;
    }
    const float rng = range_of(xa, ya);
    if (DBG && dbg_row) dbg_row[jb] = rng;
    const float4 ma = lds_f4(s_mixa + k * 16u);
    const float2 mb = lds_f2(s_mixb + k * 8u);
    const float z = ma.x - rng;
    const float pz1 = hit_gauss(z, a.k2, a.z_hit);
    const float pz2 = z < 0.0f ? ma.z : 0.0f;
    const float sum = (pz1 + pz2) + mb.x;
    if (S7) logprod_mul(acc, prod, pexp, sum);
    else acc += (sum * sum) * sum;
  }
  return S7 ? logprod_close(acc, prod, pexp) : acc;
}

/* a segment whose table look-up came back NaN (an end cell off the ring, a NaN pose): walk its rays */
template <int MODE, bool WIDE>
__device__ __noinline__ float pt_segment_walk(const SensorLaunch &a, float ox, float oy, float th, int jb0, int jb1,
                                              float *dbg_row) {
  const float4 *mixt = reinterpret_cast<const float4 *>(a.beam);
  float acc = 0.0f, prod = 1.0f;
  int pexp = 0;
  for (int jb = jb0; jb < jb1; jb++) {
    RayOut o;
    const float rng = cast_ray_slow<MODE, WIDE, false>(a, ox, oy, th, jb, o);
    if (dbg_row) dbg_row[jb] = rng;
    const float4 mx = __ldg(mixt + jb);
    const float z = mx.x - rng;
    const float pz1 = hit_gauss(z, a.k2, a.z_hit);
    const float pz2 = z < 0.0f ? mx.y : 0.0f;
    const float sum = (pz1 + pz2) + mx.z;
    if (a.s7) logprod_mul(acc, prod, pexp, sum);
    else acc += (sum * sum) * sum;
  }
  return a.s7 ? logprod_close(acc, prod, pexp) : acc;
}

constexpr int PT_MAX_WARPS = 24;
constexpr int PT_TAIL_CHUNK = 64; /* sorted particles per unit of the common tail */

template <int MODE, bool WIDE, bool DBG>
__global__ void __launch_bounds__(PT_MAX_WARPS * 32, 1)
pt_lookup_kernel(const __grid_constant__ PtLaunch L) {
  const SensorLaunch &a = L.s;
  extern __shared__ uint4 smem_dyn[];
  __shared__ uint64_t bars[PT_MAX_WARPS][2];
  const int nb = a.nb, nseg = (nb + SEG - 1) / SEG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  /* beam constants per PAIR of beams (pt_segment's packed arithmetic):
   * [mixa float4[np] (obs a, obs b, e2 a, e2 b) | dir float4[np] (cos a, cos b, sin a, sin b) | rows uint4[nrows] |
   *  mixb float2[np] (p34 a, p34 b) | pad to 16 | per warp: nbuf tables of stride floats] */
  const int np = (nb + 1) >> 1;
  float4 *smixa = reinterpret_cast<float4 *>(smem_dyn);
  float4 *sdir = smixa + np;
  uint4 *srows = reinterpret_cast<uint4 *>(sdir + np);
  float2 *smixb = reinterpret_cast<float2 *>(srows + L.nrows);
  float *stab = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(smixb + np) + 15) & ~(uintptr_t)15) +
                (size_t)wid * L.nbuf * L.stride;
  {
    const float4 *mixt = reinterpret_cast<const float4 *>(a.beam);
    const float2 *dirt = reinterpret_cast<const float2 *>(a.beam + 4 * (size_t)nb);
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
      const int ja = 2 * i, jc = min(2 * i + 1, nb - 1); /* an odd beam count: the last pair repeats its first half (never read) */
      const float4 ma = __ldg(mixt + ja), mb = __ldg(mixt + jc);
      const float2 da = __ldg(dirt + ja), db = __ldg(dirt + jc);
      smixa[i] = make_float4(ma.x, mb.x, ma.y, mb.y);
      smixb[i] = make_float2(ma.z, mb.z);
      sdir[i] = make_float4(da.x, db.x, da.y, db.y);
    }
    if (L.rows32) {
      uint32_t *srows32 = reinterpret_cast<uint32_t *>(srows);
      for (int i = threadIdx.x; i < L.nrows; i += blockDim.x) {
        const uint4 rw = __ldg(L.rows + i);
        srows32[i] = rw.x | (rw.y << 16) | (rw.z << 24);
      }
      if (threadIdx.x == 0) srows32[L.nrows] = 0u; /* the sentinel row: no end cell, i.e. the NaN entry */
    } else {
      for (int i = threadIdx.x; i < L.nrows; i += blockDim.x) srows[i] = __ldg(L.rows + i);
    }
  }
  uint32_t s_mixa = smem_u32(smixa), s_mixb = smem_u32(smixb), s_dir = smem_u32(sdir), s_rows = smem_u32(srows);
  asm volatile("" : "+r"(s_mixa), "+r"(s_mixb), "+r"(s_dir), "+r"(s_rows));
  const uint32_t s_tab0 = smem_u32(stab), tab_bytes = (uint32_t)L.stride * 4u;
  const uint32_t bar0 = smem_u32(&bars[wid][0]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  /* this warp's runs of the sorted particles: a contiguous share of the first (100 - tail_pct) percent, then
   * chunks of the rest drawn from a common counter by whoever finishes first */
  const int64_t n = a.n, Wt = (int64_t)gridDim.x * wpb, g = (int64_t)blockIdx.x * wpb + wid;
  const int64_t n_static = n - n * L.tail_pct / 100;
  /* a unit of the tail: a quarter of a warp's own share, 8 .. 64 sorted particles (a dense cloud of 100k
   * particles gives every warp ~28: units of 64 would be the longest thing in the kernel) */
  const int64_t chunk = max((int64_t)8, min((int64_t)PT_TAIL_CHUNK, n / (Wt * 4)));
  const int nbins = L.n_slots + 1;
  const uint32_t *start = L.bin_start;
  float wmax = -INFINITY;
  int it = 0; /* tables staged so far by this warp (mbarrier phase) */
  uint32_t q_lo = (uint32_t)(n_static * g / Wt), q_hi = (uint32_t)(n_static * (g + 1) / Wt);
  for (bool own = true;; own = false) {
  if (!own) { /* next unit of the common tail: units shrink as the tail runs out (a quarter of what is left per warp,
               * 8 .. chunk particles), so that the SMs finish within a few hundred particle-segments of one another
               * (fixed units of 64 left them 1.79 M .. 1.98 M active cycles apart) */
    uint32_t c = 0, want = 0;
    if (lane == 0) {
      const int64_t left = (n - n_static) - (int64_t)*reinterpret_cast<volatile uint32_t *>(L.tail_ctr);
      want = (uint32_t)max((int64_t)L.tail_min, min(chunk, left / (L.tail_div * Wt)));
      c = atomicAdd(L.tail_ctr, want);
    }
    c = __shfl_sync(FULL, c, 0);
    want = __shfl_sync(FULL, want, 0);
    const int64_t lo64 = n_static + (int64_t)c;
    if (lo64 >= n) break;
    q_lo = (uint32_t)lo64;
    q_hi = (uint32_t)min(lo64 + (int64_t)want, n);
  }
  if (q_lo >= q_hi) continue;
  int s; /* the slot holding position q_lo: the last one with start[s] <= q_lo */
  {
    int lo_s = 0, hi_s = nbins;
    while (hi_s - lo_s > 1) {
      const int mid = (lo_s + hi_s) >> 1;
      if (__ldg(start + mid) <= q_lo) lo_s = mid; else hi_s = mid;
    }
    s = lo_s;
  }
  uint32_t q = q_lo;

  /* the next cell of the run: slot, first position, particles */
  auto next_cell = [&](int &slot, uint32_t &p0, int &cnt) -> bool {
    if (q >= q_hi) return false;
    for (;;) { /* advance s to the slot holding q: the first s' >= s with start[s' + 1] > q */
      const int t = s + 1 + lane;
      const uint32_t v = t <= nbins ? __ldg(start + t) : 0xffffffffu;
      const uint32_t m = __ballot_sync(FULL, v > q);
      if (m) { s += __ffs(m) - 1; break; }
      s += 32;
    }
    const uint32_t end = min(__ldg(start + s + 1), q_hi);
    slot = s; p0 = q; cnt = (int)(end - q);
    q = end;
    return true;
  };
  auto issue = [&](int slot, int buf) { /* lane 0: bulk copy of the slot's table into buffer buf */
    if (lane == 0) {
      const uint32_t bar = bar0 + 8u * (uint32_t)buf, dst = s_tab0 + (uint32_t)buf * tab_bytes;
      const float *src = L.tables + (size_t)slot * L.stride;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the buffer's last readers are done (syncwarp) */
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tab_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(tab_bytes), "r"(bar)
                   : "memory");
    }
  };

  int c_slot = 0, n_slot = 0, c_cnt = 0, n_cnt = 0;
  uint32_t c_p0 = 0, n_p0 = 0;
  /* nbuf = 2: the next cell's table is in flight while this one is used; nbuf = 1: one buffer per warp (more
   * warps fit the SM and cover the copy's latency for one another) */
  const bool dbl = L.nbuf == 2;
  bool have = next_cell(c_slot, c_p0, c_cnt);
  if (have) issue(c_slot, dbl ? (it & 1) : 0);
  for (; have; it++) {
    bool have_n = false;
    if (dbl) {
      have_n = next_cell(n_slot, n_p0, n_cnt);
      if (have_n) issue(n_slot, (it + 1) & 1);
    }
    const int buf = dbl ? (it & 1) : 0;
    { /* wait for this cell's table */
      const uint32_t bar = bar0 + 8u * (uint32_t)buf, parity = (uint32_t)(dbl ? (it >> 1) : it) & 1u;
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
      }
    }
    const uint32_t s_tab = s_tab0 + (uint32_t)buf * tab_bytes;
    const int ntask = c_cnt * nseg; /* (segment, particle) pairs, particle fastest: a warp shares its beams */
    for (int t = lane; t < ntask; t += 32) {
      const int sgm = t / c_cnt, j = t - sgm * c_cnt;
      const uint32_t pos = c_p0 + (uint32_t)j;
      const int64_t p = (int64_t)__ldg(L.perm + pos);
      const float4 ps = __ldg(L.pstate + pos);
      const int jb0 = sgm * SEG, jb1 = min(jb0 + SEG, nb);
      float *dbg_row = nullptr;
      if (DBG && a.dbg_range && p >= a.dbg_first && p < a.dbg_first + a.dbg_count)
        dbg_row = a.dbg_range + (p - a.dbg_first) * nb;
      float v;
      if (L.rows32)
        v = a.s7 ? pt_segment<true, DBG, true>(a, L, s_tab, s_mixa, s_mixb, s_dir, s_rows, ps.x, ps.y, ps.z, ps.w, L.pstate_d + 2 * (size_t)pos, c_slot, jb0, jb1, dbg_row)
                 : pt_segment<false, DBG, true>(a, L, s_tab, s_mixa, s_mixb, s_dir, s_rows, ps.x, ps.y, ps.z, ps.w, L.pstate_d + 2 * (size_t)pos, c_slot, jb0, jb1, dbg_row);
      else
        v = a.s7 ? pt_segment<true, DBG, false>(a, L, s_tab, s_mixa, s_mixb, s_dir, s_rows, ps.x, ps.y, ps.z, ps.w, L.pstate_d + 2 * (size_t)pos, c_slot, jb0, jb1, dbg_row)
                 : pt_segment<false, DBG, false>(a, L, s_tab, s_mixa, s_mixb, s_dir, s_rows, ps.x, ps.y, ps.z, ps.w, L.pstate_d + 2 * (size_t)pos, c_slot, jb0, jb1, dbg_row);
      if (v != v) {
        const double2 od = __ldg(L.pstate_d + 2 * (size_t)pos + 1); /* the ray origin (float values, kept as doubles) */
        v = pt_segment_walk<MODE, WIDE>(a, (float)od.x, (float)od.y, __ldg(a.th + p), jb0, jb1, dbg_row);
        if (a.stats) atomicAdd(a.stats + 11, 1ull); /* table_misses: expected 0 */
      }
      /* close the segment: one rounding to fixed point; segments add as integers */
      const long long fx = __float2ll_rn(fmaxf(v, FIX_FLOOR) * FIX_SCALE);
      if (!L.fuse_finish) {
        atomicAdd(reinterpret_cast<unsigned long long *>(L.qacc) + p, (unsigned long long)fx);
      } else {
        /* the accumulator's top 4 bits count the segments added so far; the low 60 bits hold the sum of the
         * segments, each biased by 2^53 so that it is positive (|segment| < 2^52 by the floor) and the field
         * can never carry into the count (15 x 2^54 < 2^60).  Whoever adds the particle's last segment has its
         * complete sum and finishes it here — raw weight and running maximum — instead of a kernel of its own */
        constexpr long long BIAS = 1ll << 53;
        const unsigned long long add = (unsigned long long)(fx + BIAS) + (1ull << 60);
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(L.qacc) + p, add);
        if ((int)(old >> 60) == nseg - 1) {
          const long long q = (long long)((old + add) & 0x0fffffffffffffffull) - (long long)nseg * BIAS;
          const float r = fixed_to_raw(q);
          if (a.raw) a.raw[p] = r;
          wmax = fmaxf(wmax, r);
        }
      }
    }
    __syncwarp();
    if (!dbl) {
      have_n = next_cell(n_slot, n_p0, n_cnt);
      if (have_n) issue(n_slot, 0);
    }
    c_slot = n_slot; c_p0 = n_p0; c_cnt = n_cnt;
    have = have_n;
  }
  } /* ranges */
  if (L.fuse_finish && a.rawmax_ord) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, o));
    if (lane == 0 && wmax > -INFINITY) atomicMax(a.rawmax_ord, ord_of_float(wmax));
  }
}

// raw[p] = the fixed-point sum of the particle's segments as a float, plus the max for the normaliser.  Four
// particles per thread (two 128-bit loads, one 128-bit store) and ONE atomic per CTA: one atomicMax per warp on
// the same word (39,000 of them for 1.25 M particles) was most of this kernel's 28 us.
constexpr int PT_FIN_THREADS = 256, PT_FIN_ITEMS = 4;
__global__ void __launch_bounds__(PT_FIN_THREADS)
pt_finish_kernel(const long long *__restrict__ qacc, int64_t n, float *__restrict__ raw, uint32_t *rawmax_ord) {
  __shared__ float warp_max[PT_FIN_THREADS / 32];
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PT_FIN_ITEMS;
  float v[PT_FIN_ITEMS];
  if (i0 + PT_FIN_ITEMS <= n) {
    const longlong2 q0 = __ldg(reinterpret_cast<const longlong2 *>(qacc + i0));
    const longlong2 q1 = __ldg(reinterpret_cast<const longlong2 *>(qacc + i0) + 1);
    v[0] = fixed_to_raw(q0.x); v[1] = fixed_to_raw(q0.y); v[2] = fixed_to_raw(q1.x); v[3] = fixed_to_raw(q1.y);
    if (raw) *reinterpret_cast<float4 *>(raw + i0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < PT_FIN_ITEMS; k++) {
      v[k] = -INFINITY;
      if (i0 + k < n) {
        v[k] = fixed_to_raw(__ldg(qacc + i0 + k));
        if (raw) raw[i0 + k] = v[k];
      }
    }
  }
  float m = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0 && rawmax_ord) {
#pragma unroll
    for (int k = 1; k < PT_FIN_THREADS / 32; k++) m = fmaxf(m, warp_max[k]);
    if (m > -INFINITY) atomicMax(rawmax_ord, ord_of_float(m));
  }
}

size_t pt_lookup_smem(const PtLaunch &L, int warps) {
  return (size_t)((L.s.nb + 1) / 2) * 40 + (size_t)L.nrows * 16 + 32 + (size_t)warps * L.nbuf * L.stride * 4;
}
int pt_max_warps() { return PT_MAX_WARPS; }

template <int MODE, bool WIDE>
static cudaError_t launch_pt_fill_t(const amclj_ctx *c, const PtLaunch &L, cudaStream_t s) {
  pt_fill_kernel<MODE, WIDE><<<(unsigned)(c->sm_count * 16), 256, 0, s>>>(L);
  return cudaGetLastError();
}

cudaError_t launch_pt_fill(const amclj_ctx *c, const PtLaunch &L, int mode, cudaStream_t s) {
  const bool wide = L.s.div_wide != 0;
  /* the zero table: every beam of a particle on a non-free cell hits at step 0 (range 0) */
  cudaError_t e = cudaMemsetAsync(L.tables + (size_t)L.n_slots * L.stride, 0, sizeof(float) * (size_t)L.stride, s);
  if (e != cudaSuccess || L.n_slots == 0) return e;
  switch (mode) {
    case DF_WEDGE8: return wide ? launch_pt_fill_t<DF_WEDGE8, true>(c, L, s) : launch_pt_fill_t<DF_WEDGE8, false>(c, L, s);
    case DF_GLOBAL8: return wide ? launch_pt_fill_t<DF_GLOBAL8, true>(c, L, s) : launch_pt_fill_t<DF_GLOBAL8, false>(c, L, s);
    default: return wide ? launch_pt_fill_t<DF_GLOBAL4, true>(c, L, s) : launch_pt_fill_t<DF_GLOBAL4, false>(c, L, s);
  }
}

template <int MODE, bool WIDE, bool DBG>
static cudaError_t launch_pt_lookup_t(const amclj_ctx *c, const PtLaunch &L, cudaStream_t s) {
  auto kern = pt_lookup_kernel<MODE, WIDE, DBG>;
  const size_t smem = pt_lookup_smem(L, L.warps);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  /* persistent: one CTA per SM; fewer when there are not enough particles to give every warp a few */
  int64_t ctas = c->sm_count;
  const int64_t want = (L.s.n + (int64_t)L.warps * 4 - 1) / ((int64_t)L.warps * 4);
  if (ctas > want) ctas = want < 1 ? 1 : want;
  if (c->time_main) cudaEventRecord(c->ev[5], s);
  kern<<<(unsigned)ctas, L.warps * 32, smem, s>>>(L);
  if (c->time_main) cudaEventRecord(c->ev[6], s);
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_pt_lookup_m(const amclj_ctx *c, const PtLaunch &L, bool wide, bool dbg, cudaStream_t s) {
  if (dbg) return wide ? launch_pt_lookup_t<MODE, true, true>(c, L, s) : launch_pt_lookup_t<MODE, false, true>(c, L, s);
  return wide ? launch_pt_lookup_t<MODE, true, false>(c, L, s) : launch_pt_lookup_t<MODE, false, false>(c, L, s);
}

cudaError_t launch_pt_update(const amclj_ctx *c, const PtLaunch &L, int mode, bool dbg, bool hist_done, cudaStream_t s) {
  const SensorLaunch &a = L.s;
  if (a.n <= 0 || a.nb <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((a.n + PT_SORT_THREADS - 1) / PT_SORT_THREADS);
  if (!hist_done) pt_hist_kernel<<<blocks, PT_SORT_THREADS, 0, s>>>(L); /* else: the motion kernel took this pass along */
  pt_scan_kernel<<<(unsigned)((L.n_slots + 1 + PT_SCAN_TILE - 1) / PT_SCAN_TILE), PT_SCAN_THREADS, 0, s>>>(L.bin_count, L.bin_start, L.cursor,
                                                                                                                L.n_slots + 1, L.scan_state);
  pt_scatter_kernel<<<blocks, PT_SORT_THREADS, 0, s>>>(L);
  const bool wide = a.div_wide != 0;
  cudaError_t e;
  switch (mode) {
    case DF_WEDGE8: e = launch_pt_lookup_m<DF_WEDGE8>(c, L, wide, dbg, s); break;
    case DF_GLOBAL8: e = launch_pt_lookup_m<DF_GLOBAL8>(c, L, wide, dbg, s); break;
    default: e = launch_pt_lookup_m<DF_GLOBAL4>(c, L, wide, dbg, s); break;
  }
  if (e != cudaSuccess) return e;
  if (!L.fuse_finish) {
    const int64_t per_cta = (int64_t)PT_FIN_THREADS * PT_FIN_ITEMS;
    pt_finish_kernel<<<(unsigned)((a.n + per_cta - 1) / per_cta), PT_FIN_THREADS, 0, s>>>(L.qacc, a.n, a.raw, a.rawmax_ord);
  }
  return cudaGetLastError();
}

}  // namespace amclj
