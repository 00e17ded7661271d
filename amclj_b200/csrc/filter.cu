// filter.cu — every hot-path kernel except the ray caster: K1 motion sample, map staging
// (distance transform), K3 weight normalisation -> fixed point -> decoupled-look-back
// prefix sum, K4 systematic binary-search gather, K5 roulette / tournament selection,
// K6 pose estimate, and on-device particle initialisation.
//
// Reference functions replaced (jvia/amclj): pf.clj:139-145 apply-motion-model ->
// motion.clj:24-45; pf.clj:156-167 roulette-selection; pf.clj:182-195 tournament-selection;
// pf.clj:67-84 pose-estimate; map.clj:59-84 uniform-initialization, :87-111
// gaussian-initialization.
#include <string.h>

#include "internal.h"
#include "sensor_common.cuh"

namespace amclj {

// ======================================================================================
// K1 motion.clj:24-45 sample-motion-model-odometry*.  rot1/trans/rot2 and the three sigmas
// are launch-uniform (computed once on the host, motion.clj:68-79 did it per particle);
// per particle: three normals, one sincos, three FMAs.  24 B/particle of HBM traffic.
// ======================================================================================
__device__ __forceinline__ uint32_t ord_of_float_m(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ void philox_normals3(uint64_t seed, uint64_t gid, uint32_t tag,
                                                float &n1, float &n2, float &n3) {
  uint32_t r[4];
  philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0u, tag, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  float ra = sqrtf(-2.0f * logf(u24(r[0]))), rb = sqrtf(-2.0f * logf(u24(r[2])));
  float a = 6.283185307179586f * u24(r[1]), b = 6.283185307179586f * u24(r[3]);
  n1 = ra * cosf(a);
  n2 = ra * sinf(a);
  n3 = rb * cosf(b);
}

/* one particle: motion.clj:24-45 with the launch-uniform control; returns its table slot when the sort's first
 * pass rides along (else -1) */
__device__ __forceinline__ int motion_one(const MotionLaunch &a, int64_t i, float &x, float &y, float &th, float n1,
                                          float n2, float n3) {
  float r1 = fmaf(-n1, a.sd1, a.rot1);                          /* motion.clj:35 */
  float tr = a.trans_zero ? 0.0f : fmaf(-n2, a.sdt, a.trans);   /* motion.clj:36-37 */
  float r2 = fmaf(-n3, a.sd2, a.rot2);                          /* motion.clj:38 */
  float sn, cs;
  sincos_det(th + r1, sn, cs);
  const float nx = fmaf(tr, cs, x), ny = fmaf(tr, sn, y);     /* motion.clj:39-44 */
  const float nth = wrap_heading(th + (r1 + r2)); /* quat/rotate by rot1* + rot2*; the next to-heading is in (-pi, pi] */
  x = nx; y = ny; th = nth;
  int hkey = -1;
  if (a.hist_slot_of_cell) { /* first pass of the persistent-table sort (ptable.cu pt_hist_kernel), same arithmetic */
    float ox = nx, oy = ny;
    if (a.hist_s10) {
      float st, ct;
      sincos_det(nth, st, ct);
      ox = nx + fmaf(a.hist_laser_x, ct, -(a.hist_laser_y * st));
      oy = ny + fmaf(a.hist_laser_x, st, a.hist_laser_y * ct);
    }
    const int x0 = round_cell_start(ox, a.hist_res), y0 = round_cell_start(oy, a.hist_res);
    const bool in = (uint32_t)x0 < (uint32_t)a.hist_W && (uint32_t)y0 < (uint32_t)a.hist_H;
    hkey = in ? __ldg(a.hist_slot_of_cell + (size_t)y0 * a.hist_W + x0) : a.hist_n_slots;
    a.hist_pkey[i] = hkey;
    a.hist_qacc[i] = 0;
  }
  return hkey;
}

// Every thread moves FOUR consecutive particles: x, y, theta (and the injected normals) stream through
// 128-bit coalesced loads and stores; the last, partial group of a set whose size is not a multiple of four
// goes element by element.
__global__ void __launch_bounds__(256) motion_kernel(const MotionLaunch a) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i0 = g * 4;
  const int lane = threadIdx.x & 31;
  uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  float px[4], py[4], pt[4], nz[12];
  const int cnt = i0 < a.n ? (int)min((int64_t)4, a.n - i0) : 0;
  const bool vec = cnt == 4 && a.vec_ok;
  if (vec) {
    const float4 vx = *reinterpret_cast<const float4 *>(a.x + i0), vy = *reinterpret_cast<const float4 *>(a.y + i0),
                 vt = *reinterpret_cast<const float4 *>(a.th + i0);
    px[0] = vx.x; px[1] = vx.y; px[2] = vx.z; px[3] = vx.w;
    py[0] = vy.x; py[1] = vy.y; py[2] = vy.z; py[3] = vy.w;
    pt[0] = vt.x; pt[1] = vt.y; pt[2] = vt.z; pt[3] = vt.w;
    if (a.normals) {
      const float4 *nn = reinterpret_cast<const float4 *>(a.normals + 3 * i0); /* 48 bytes per group: 16-byte aligned */
      const float4 na = __ldg(nn), nb = __ldg(nn + 1), nc = __ldg(nn + 2);
      nz[0] = na.x; nz[1] = na.y; nz[2] = na.z; nz[3] = na.w; nz[4] = nb.x; nz[5] = nb.y;
      nz[6] = nb.z; nz[7] = nb.w; nz[8] = nc.x; nz[9] = nc.y; nz[10] = nc.z; nz[11] = nc.w;
    }
  } else {
    for (int u = 0; u < cnt; u++) {
      px[u] = a.x[i0 + u]; py[u] = a.y[i0 + u]; pt[u] = a.th[i0 + u];
      if (a.normals)
        for (int k = 0; k < 3; k++) nz[3 * u + k] = __ldg(a.normals + 3 * (i0 + u) + k);
    }
  }
  int hkey[4] = {-1, -1, -1, -1};
#pragma unroll
  for (int u = 0; u < 4; u++) {
    if (u < cnt) {
      float n1, n2, n3;
      if (a.normals) { n1 = nz[3 * u]; n2 = nz[3 * u + 1]; n3 = nz[3 * u + 2]; }
      else philox_normals3(a.seed, (uint64_t)(a.gid0 + i0 + u), TAG_MOTION, n1, n2, n3);
      hkey[u] = motion_one(a, i0 + u, px[u], py[u], pt[u], n1, n2, n3);
      m0 = max(m0, ord_of_float_m(px[u])); m1 = max(m1, ord_of_float_m(-px[u]));
      m2 = max(m2, ord_of_float_m(py[u])); m3 = max(m3, ord_of_float_m(-py[u]));
    }
  }
  if (vec) {
    *reinterpret_cast<float4 *>(a.x + i0) = make_float4(px[0], px[1], px[2], px[3]);
    *reinterpret_cast<float4 *>(a.y + i0) = make_float4(py[0], py[1], py[2], py[3]);
    *reinterpret_cast<float4 *>(a.th + i0) = make_float4(pt[0], pt[1], pt[2], pt[3]);
  } else {
    for (int u = 0; u < cnt; u++) { a.x[i0 + u] = px[u]; a.y[i0 + u] = py[u]; a.th[i0 + u] = pt[u]; }
  }
  if (a.hist_slot_of_cell) { /* histogram with one atomic per distinct cell per warp and round */
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int key = u < cnt ? hkey[u] : -1 - lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(a.hist_bin_count + key, (uint32_t)__popc(peers));
    }
    if (blockIdx.x == 0)
      for (int t = threadIdx.x; t < a.hist_scan_words; t += blockDim.x) a.hist_scan_state[t] = 0ull;
  }
  if (a.bbox) { /* bounding box of the moved cloud, for the sensor stage's walking-order decision */
    m0 = __reduce_max_sync(0xffffffffu, m0);
    m1 = __reduce_max_sync(0xffffffffu, m1);
    m2 = __reduce_max_sync(0xffffffffu, m2);
    m3 = __reduce_max_sync(0xffffffffu, m3);
    __shared__ uint32_t sm[4];
    if (threadIdx.x < 4) sm[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { /* block-level first: the four global words are shared by everyone */
      atomicMax(&sm[0], m0);
      atomicMax(&sm[1], m1);
      atomicMax(&sm[2], m2);
      atomicMax(&sm[3], m3);
    }
    __syncthreads();
    if (threadIdx.x < 4) atomicMax(a.bbox + threadIdx.x, sm[threadIdx.x]);
  }
}

cudaError_t launch_motion(const MotionLaunch &a_in, cudaStream_t s) {
  if (a_in.n <= 0 || a_in.is_static) return cudaSuccess; /* motion.clj:26-27 */
  MotionLaunch a = a_in;
  auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  a.vec_ok = al16(a.x) && al16(a.y) && al16(a.th) && (!a.normals || al16(a.normals));
  const int64_t groups = (a.n + 3) / 4;
  motion_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ======================================================================================
// Map staging: Chebyshev (L-infinity) distance to the nearest blocked cell, separable:
// d(x,y) = min over rows y' of max(|y'-y|, rowdist(x,y')).  The field is padded by a ring
// of blocked cells, which also stands for everything beyond it (map.clj:19-24: out of
// range reads -1, which is uninhabitable).
// ======================================================================================
__global__ void rowdist_kernel(const int8_t *__restrict__ cells, int W, int H, int cap,
                               uint8_t *__restrict__ rd) {
  int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
  int PW = W + 2;
  if (px >= PW) return;
  auto blocked = [&](int qx) -> bool {
    if (qx <= 0 || qx >= W + 1 || py == 0 || py == H + 1) return true;
    return cells[(size_t)(py - 1) * W + (qx - 1)] != 0; /* map.clj:48-57 */
  };
  int d = 0;
  if (!blocked(px)) {
    for (d = 1; d < cap; d++)
      if (blocked(px - d) || blocked(px + d)) break;
  }
  rd[(size_t)py * PW + px] = (uint8_t)d;
}

__global__ void coldist_kernel(const uint8_t *__restrict__ rd, int W, int H, int cap,
                               uint8_t *__restrict__ df4, int row_bytes4,
                               uint8_t *__restrict__ df8, int pitch8) {
  /* one thread per PAIR of cells so that a nibble-packed byte has a single writer */
  int pp = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
  int PW = W + 2, PH = H + 2;
  if (2 * pp >= PW) return;
  int dd[2];
  for (int h = 0; h < 2; h++) {
    int px = 2 * pp + h;
    int best = 0;
    if (px < PW) {
      best = rd[(size_t)py * PW + px];
      for (int dy = 1; dy < best; dy++) {
        int up = py - dy >= 0 ? rd[(size_t)(py - dy) * PW + px] : 0;
        int dn = py + dy < PH ? rd[(size_t)(py + dy) * PW + px] : 0;
        int m = min(up, dn);
        best = min(best, max(dy, m));
      }
    }
    dd[h] = min(best, cap);
  }
  if (df4) df4[(size_t)py * row_bytes4 + pp] = (uint8_t)(min(dd[0], 15) | (min(dd[1], 15) << 4));
  if (df8) {
    df8[(size_t)py * pitch8 + 2 * pp] = (uint8_t)min(dd[0], 255);
    if (2 * pp + 1 < pitch8) df8[(size_t)py * pitch8 + 2 * pp + 1] = (uint8_t)min(dd[1], 255);
  }
}

cudaError_t build_distance_fields(amclj_ctx *c, bool want8, cudaStream_t s) {
  DeviceMap &m = c->map;
  int PW = m.W + 2, PH = m.H + 2;
  uint8_t *rd = nullptr;
  cudaError_t e = cudaMalloc(&rd, (size_t)PW * PH);
  if (e != cudaSuccess) return e;
  int cap = want8 ? 255 : 15;
  dim3 g1((PW + 127) / 128, PH), g2(((PW + 1) / 2 + 127) / 128, PH);
  rowdist_kernel<<<g1, 128, 0, s>>>(m.cells, m.W, m.H, cap, rd);
  cudaMemsetAsync(m.df4, 0, m.bytes4, s);
  if (want8) cudaMemsetAsync(m.df8, 0, m.bytes8, s);
  coldist_kernel<<<g2, 128, 0, s>>>(rd, m.W, m.H, cap, m.df4, m.row_bytes4, want8 ? m.df8 : nullptr,
                                    m.pitch8);
  e = cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize(s);
  cudaFree(rd);
  return e != cudaSuccess ? e : e2;
}

// ======================================================================================
// K3 normalise -> fixed point -> prefix sum.
// pf.clj:159-160 divides every weight by a float sum; a float prefix sum depends on the
// summation order, so the contract here is integer: q_i = floor(w_i / w_max * 2^32)
// elementwise, then an exact u64 inclusive scan (single pass, decoupled look-back).
// ======================================================================================
__device__ __forceinline__ float float_of_ord(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void max_kernel(const float *__restrict__ v, int64_t n, uint32_t *out_ord) {
  float m = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, __ldg(v + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomicMax(out_ord, ord_of_float(m));
}

__global__ void decode_max_kernel(const uint32_t *ord, float *out) {
  uint32_t o = *ord;
  *out = o == 0 ? -INFINITY : float_of_ord(o);
}

cudaError_t launch_max(const float *v, int64_t n, uint32_t *ord, float *out, int sm_count,
                       cudaStream_t s) {
  cudaMemsetAsync(ord, 0, sizeof(uint32_t), s);
  if (n > 0) {
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    max_kernel<<<(unsigned)blocks, 256, 0, s>>>(v, n, ord);
  }
  decode_max_kernel<<<1, 1, 0, s>>>(ord, out);
  return cudaGetLastError();
}

cudaError_t launch_decode_max(const uint32_t *ord, float *out, cudaStream_t s) {
  decode_max_kernel<<<1, 1, 0, s>>>(ord, out);
  return cudaGetLastError();
}

// bounding box of the particle set as four maxima (x, -x, y, -y) in ordered-uint form, so that
// a zeroed control block is the identity and a NaN coordinate poisons the box (-> "not compact")
__global__ void __launch_bounds__(256)
bbox_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t n, uint32_t *bbox) {
  uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float px = __ldg(x + i), py = __ldg(y + i);
    m0 = max(m0, ord_of_float(px));
    m1 = max(m1, ord_of_float(-px));
    m2 = max(m2, ord_of_float(py));
    m3 = max(m3, ord_of_float(-py));
  }
  m0 = __reduce_max_sync(0xffffffffu, m0);
  m1 = __reduce_max_sync(0xffffffffu, m1);
  m2 = __reduce_max_sync(0xffffffffu, m2);
  m3 = __reduce_max_sync(0xffffffffu, m3);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(bbox + 0, m0);
    atomicMax(bbox + 1, m1);
    atomicMax(bbox + 2, m2);
    atomicMax(bbox + 3, m3);
  }
}

cudaError_t launch_bbox(const float *x, const float *y, int64_t n, uint32_t *bbox, int sm_count,
                        cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > (int64_t)sm_count * 4) blocks = (int64_t)sm_count * 4;
  bbox_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, y, n, bbox);
  return cudaGetLastError();
}

// Walking order for spread-out clouds: a counting sort of the particle indices by square map tile
// (row-major tiles) and heading sector.  The order inside a tile is whatever the atomics give — it only decides
// which lanes walk together, never a result.  Skipped (the kernels return) for a compact cloud.
// Tile side: 8 cells (shift 3) on a densely covered map; on a sparsely covered one (C5: one particle per
// 16 cells) the side doubles until a tile holds ~128 particles, so that the 32 lanes of a warp still share
// their heading sector and the tile counters stay few (tile_order_shift).
constexpr int TILE_SHIFT_MIN = 3, TILE_SHIFT_MAX = 8;
int tile_order_shift(int64_t n, int64_t free_cells) {
  int s = TILE_SHIFT_MIN;
  while (s < TILE_SHIFT_MAX && (n << (2 * s)) < 128ll * free_cells) s++;
  return s;
}
#ifndef AMCLJ_HEADING_BINS
#define AMCLJ_HEADING_BINS 8
#endif
constexpr int HEADING_BINS = AMCLJ_HEADING_BINS; /* power of two: secondary key, heading sector */
static int tile_count_at(int W, int H, int shift) { return ((W >> shift) + 1) * ((H >> shift) + 1) * HEADING_BINS; }
int tile_count(int W, int H) { return tile_count_at(W, H, TILE_SHIFT_MIN); } /* the most there can be */

__device__ __forceinline__ int tile_of(float px, float py, float pth, float res, int W, int H, int TILE_SHIFT) {
  int cx = min(max(round_cell(px, res), 0), W - 1), cy = min(max(round_cell(py, res), 0), H - 1);
  int hb = __float2int_rd(pth * (HEADING_BINS / 6.283185307179586f)) & (HEADING_BINS - 1);
  return ((cy >> TILE_SHIFT) * ((W >> TILE_SHIFT) + 1) + (cx >> TILE_SHIFT)) * HEADING_BINS + hb;
}

__global__ void __launch_bounds__(256)
tile_hist_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th,
                 int64_t n, float res, int W, int H, int shift, uint32_t *counters, const uint32_t *bbox, float compact_m) {
  if (bbox && cloud_is_compact(bbox, compact_m)) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&counters[tile_of(__ldg(x + i), __ldg(y + i), __ldg(th + i), res, W, H, shift)], 1u);
}

__global__ void __launch_bounds__(1024)
tile_scan_kernel(uint32_t *counters, int ntiles, const uint32_t *bbox, float compact_m) {
  if (bbox && cloud_is_compact(bbox, compact_m)) return;
  __shared__ uint32_t part[1024];
  const int per = (ntiles + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, ntiles);
  uint32_t sum = 0;
  for (int t = lo; t < hi; t++) sum += counters[t];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { /* 1024 partial sums: a serial pass is fine */
    uint32_t run = 0;
    for (int t = 0; t < 1024; t++) { uint32_t v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int t = lo; t < hi; t++) { uint32_t v = counters[t]; counters[t] = run; run += v; }
}

__global__ void __launch_bounds__(256)
tile_scatter_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th,
                    int64_t n, float res, int W, int H, int shift, uint32_t *cursors, int32_t *perm, const uint32_t *bbox,
                    float compact_m) {
  if (bbox && cloud_is_compact(bbox, compact_m)) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm[atomicAdd(&cursors[tile_of(__ldg(x + i), __ldg(y + i), __ldg(th + i), res, W, H, shift)], 1u)] = (int32_t)i;
}

cudaError_t launch_tile_order(const float *x, const float *y, const float *th, int64_t n, float res, int W, int H,
                              uint32_t *counters, int32_t *perm, const uint32_t *bbox, float compact_m,
                              int shift, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (shift < TILE_SHIFT_MIN || shift > TILE_SHIFT_MAX) shift = TILE_SHIFT_MIN;
  int ntiles = tile_count_at(W, H, shift);
  cudaMemsetAsync(counters, 0, sizeof(uint32_t) * (size_t)ntiles, s);
  unsigned blocks = (unsigned)((n + 255) / 256);
  tile_hist_kernel<<<blocks, 256, 0, s>>>(x, y, th, n, res, W, H, shift, counters, bbox, compact_m);
  tile_scan_kernel<<<1, 1024, 0, s>>>(counters, ntiles, bbox, compact_m);
  tile_scatter_kernel<<<blocks, 256, 0, s>>>(x, y, th, n, res, W, H, shift, counters, perm, bbox, compact_m);
  return cudaGetLastError();
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr uint64_t ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t weight_to_fixed(float v, float vmax, int is_log) {
  /* log-weights: w = exp(v - max) in (0, 1], max maps to exactly 1; linear: w / w_max */
  if (is_log) return fixed_weight(exp_det(v - vmax), 1.0f);
  return fixed_weight(v, vmax);
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const float *__restrict__ v, const float *__restrict__ vmaxp,
            const uint32_t *__restrict__ vmax_ord, int is_log, int64_t n,
            uint64_t *__restrict__ cdf, uint64_t *tile_state, uint32_t *tile_counter,
            uint64_t *total_out) {
  __shared__ uint64_t warp_sums[SCAN_THREADS / 32];
  __shared__ uint64_t s_excl;
  __shared__ uint32_t s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  /* the max arrives as a float (all-reduced across ranks) or as the ordered-uint the sensor
   * kernel accumulated with atomicMax */
  float vmax;
  if (vmax_ord) {
    uint32_t o = *vmax_ord;
    vmax = o == 0 ? -INFINITY : float_of_ord(o);
  } else {
    vmax = *vmaxp;
  }
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t q[SCAN_ITEMS];
  if (base + SCAN_ITEMS <= n) { /* 128-bit loads: base is a multiple of 8 */
    const float4 *v4 = reinterpret_cast<const float4 *>(v + base);
    float4 a = __ldg(v4), b = __ldg(v4 + 1);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) q[k] = weight_to_fixed(f[k], vmax, is_log);
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
      q[k] = base + k < n ? weight_to_fixed(__ldg(v + base + k), vmax, is_log) : 0;
  }
  uint64_t run = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { run += q[k]; q[k] = run; }
  /* block-wide exclusive scan of the per-thread totals */
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  uint64_t woff = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < SCAN_THREADS / 32; k++) {
    uint64_t ws = warp_sums[k];
    if (k < wid) woff += ws;
    tile_total += ws;
  }
  uint64_t thread_excl = woff + inc - run;
  /* decoupled look-back (Merrill & Garland): publish the aggregate, then walk back */
  if (wid == 0) {
    uint64_t excl = 0;
    if (tile == 0) {
      if (lane == 0) {
        __threadfence();
        atomicExch((unsigned long long *)&tile_state[0], ST_PREFIX | tile_total);
      }
    } else {
      if (lane == 0) atomicExch((unsigned long long *)&tile_state[tile], ST_AGG | tile_total);
      int64_t look = (int64_t)tile - 1;
      for (;;) {
        int64_t t = look - lane;
        uint64_t sv;
        do { /* spin until every descriptor of the window is published */
          sv = t >= 0 ? ld_volatile_u64(&tile_state[t]) : ST_PREFIX;
        } while (__any_sync(0xffffffffu, (sv >> 62) == 0));
        uint32_t is_prefix = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
        int first = is_prefix ? __ffs(is_prefix) - 1 : 32; /* nearest tile with a prefix */
        uint64_t contrib = lane <= first ? (sv & ST_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (is_prefix) break;
        look -= 32;
      }
      if (lane == 0)
        atomicExch((unsigned long long *)&tile_state[tile], ST_PREFIX | (excl + tile_total));
    }
    if (lane == 0) {
      s_excl = excl;
      if ((int64_t)(tile + 1) * SCAN_TILE >= n) *total_out = excl + tile_total;
    }
  }
  __syncthreads();
  const uint64_t off = s_excl + thread_excl;
  if (base + SCAN_ITEMS <= n) {
    ulonglong2 *o2 = reinterpret_cast<ulonglong2 *>(cdf + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k += 2) o2[k / 2] = make_ulonglong2(off + q[k], off + q[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
      if (base + k < n) cdf[base + k] = off + q[k];
  }
}

int64_t scan_tiles(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

/* total_out, tile_counter and tile_state[0..tiles) must be zero (the ctx control block is
 * cleared once per update) */
cudaError_t launch_scan(const float *v, const float *vmaxp, const uint32_t *vmax_ord, int is_log,
                        int64_t n, uint64_t *cdf, uint64_t *tile_state, uint32_t *tile_counter,
                        uint64_t *total_out, cudaStream_t s) {
  int64_t tiles = scan_tiles(n);
  if (tiles == 0) return cudaSuccess;
  scan_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(v, vmaxp, vmax_ord, is_log, n, cdf, tile_state,
                                                        tile_counter, total_out);
  return cudaGetLastError();
}

// ======================================================================================
// K4 low-variance / systematic resampling (Probabilistic Robotics table 4.4) as a gather:
// output slot m looks up the particle whose CDF interval contains U_m.
// ======================================================================================
// r = floor(u0 * T) with u0 given as a 0.64 fixed-point fraction; T = a*N + b.
__global__ void plan_kernel(const uint64_t *total, uint64_t u0_bits, int64_t n_global, SysPlan *out) {
  SysPlan p;
  p.total = *total;
  p.n = (uint64_t)n_global;
  p.a = p.total / p.n;
  p.b = p.total % p.n;
  p.r = __umul64hi(u0_bits, p.total);
  *out = p;
}
cudaError_t launch_plan(const uint64_t *total, uint64_t u0_bits, int64_t n_global, SysPlan *out,
                        cudaStream_t s) {
  plan_kernel<<<1, 1, 0, s>>>(total, u0_bits, n_global, out);
  return cudaGetLastError();
}
__global__ void set_plan_kernel(const SysPlan p, SysPlan *out) { *out = p; }
cudaError_t launch_set_plan(const SysPlan &p, SysPlan *out, cudaStream_t s) {
  set_plan_kernel<<<1, 1, 0, s>>>(p, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) systematic_gather_kernel(const GatherArgs a) {
  __shared__ SysPlan s_plan;
  if (threadIdx.x == 0) {
    if (a.plan) {
      s_plan = *a.plan;
    } else { /* r = floor(u0 * T) with u0 as a 0.64 fraction; T = a*N + b */
      SysPlan p;
      p.total = *a.total_ptr;
      p.n = (uint64_t)a.n_global;
      p.a = p.total / p.n;
      p.b = p.total % p.n;
      p.r = __umul64hi(a.u0_bits, p.total);
      s_plan = p;
    }
  }
  __syncthreads();
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const SysPlan plan = s_plan;
  /* The thresholds of a CTA's slots are monotone, so their sources lie between the source of its first slot
   * and the source of its last: two threads search the whole CDF (~21 dependent loads for a million entries),
   * the others only that window (a few hundred entries the CTA shares in L1). */
  __shared__ int64_t s_win[2];
  if (plan.total != 0 && (threadIdx.x == 0 || threadIdx.x == blockDim.x - 1)) {
    const int64_t kk = min(k, a.count - 1);
    const uint64_t key = sys_threshold(plan, (uint64_t)(a.m_lo + kk)) - a.cdf_offset;
    int64_t lo = 0, hi = a.n_local;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (__ldg(a.cdf + mid) > key) hi = mid; else lo = mid + 1;
    }
    s_win[threadIdx.x == 0 ? 0 : 1] = lo;
  }
  __syncthreads();
  if (k >= a.count) return;
  int64_t i;
  if (plan.total == 0) {
    /* every weight is zero (e.g. a zero-probability beam under log-weights): there is no CDF to draw
     * from.  Keep the particle that sits in the slot, exactly as pull_kernel does, and count the
     * event so that the synchronising callers can report it (AMCLJ_ERR_STATE). */
    if (k == 0 && a.zero_flag) atomicAdd(a.zero_flag, 1ull);
    i = a.m_lo + k < a.n_local ? a.m_lo + k : a.n_local - 1;
  } else {
    uint64_t U = sys_threshold(plan, (uint64_t)(a.m_lo + k));
    uint64_t key = U - a.cdf_offset; /* caller guarantees U >= cdf_offset */
    int64_t lo = s_win[0], hi = min(s_win[1] + 1, a.n_local);
    while (lo < hi) { /* first i with cdf[i] > key */
      int64_t mid = (lo + hi) >> 1;
      if (__ldg(a.cdf + mid) > key) hi = mid; else lo = mid + 1;
    }
    i = lo < a.n_local ? lo : a.n_local - 1;
  }
  float px = __ldg(a.x + i), py = __ldg(a.y + i), pt = __ldg(a.th + i);
  if (a.ox) { a.ox[k] = px; a.oy[k] = py; a.oth[k] = pt; a.ow[k] = a.wout; }
  if (a.orow) a.orow[k] = make_float4(px, py, pt, a.wout);
  if (a.oidx) a.oidx[k] = (int32_t)i;
  if (a.oidx64) a.oidx64[k] = a.gid0 + i;
}

cudaError_t launch_systematic(const GatherArgs &a, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  systematic_gather_kernel<<<(unsigned)((a.count + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ======================================================================================
// K5 pf.clj:156-167 roulette-selection: attempt t draws idx ~ U{0..N-1}, u ~ U[0,1) and is
// accepted when wn[idx] > u; the output keeps acceptance order.  Attempts are independent,
// so a super-chunk of them is tested in parallel and compacted in attempt order (count ->
// block scan -> look-back over blocks -> regenerate and write).
// ======================================================================================
__device__ __forceinline__ void roulette_attempt(uint64_t seed, uint64_t t, int64_t n,
                                                 int32_t &idx, uint32_t &u) {
  uint64_t b = t >> 1;
  uint32_t r[4];
  philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), 0u, TAG_ROULETTE, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  int k = (int)(t & 1) * 2;
  idx = (int32_t)(((uint64_t)r[k] * (uint64_t)n) >> 32);
  u = r[k + 1];
}

__device__ __forceinline__ bool roulette_test(const RouletteArgs &a, int64_t t, int32_t &idx) {
  uint32_t u;
  if (a.sidx) {
    idx = __ldg(a.sidx + t);
    u = __ldg(a.su + t);
    if (idx < 0 || idx >= a.n) return false;
  } else {
    roulette_attempt(a.seed, (uint64_t)t, a.n, idx, u);
  }
  uint64_t hiC = __ldg(a.cdf + idx), loC = idx > 0 ? __ldg(a.cdf + idx - 1) : 0;
  return roulette_accept(hiC - loC, u, a.total);
}

__global__ void __launch_bounds__(256) roulette_kernel(const RouletteArgs a) {
  __shared__ uint32_t warp_sums[8];
  __shared__ uint64_t s_excl;
  __shared__ uint32_t s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t tb = a.t0 + ((int64_t)tile * 256 + threadIdx.x) * a.per_thread;
  uint32_t cnt = 0;
  for (int k = 0; k < a.per_thread; k++) {
    int64_t t = tb + k;
    int32_t idx;
    if (t < a.t_end && roulette_test(a, t, idx)) cnt++;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  uint32_t woff = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t ws = warp_sums[k];
    if (k < wid) woff += ws;
    tile_total += ws;
  }
  if (wid == 0) {
    uint64_t excl = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch((unsigned long long *)&a.tile_state[0], ST_PREFIX | tile_total);
    } else {
      if (lane == 0) atomicExch((unsigned long long *)&a.tile_state[tile], ST_AGG | tile_total);
      int64_t look = (int64_t)tile - 1;
      for (;;) {
        int64_t t = look - lane;
        uint64_t sv;
        do {
          sv = t >= 0 ? ld_volatile_u64(&a.tile_state[t]) : ST_PREFIX;
        } while (__any_sync(0xffffffffu, (sv >> 62) == 0));
        uint32_t is_prefix = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
        int first = is_prefix ? __ffs(is_prefix) - 1 : 32;
        uint64_t contrib = lane <= first ? (sv & ST_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (is_prefix) break;
        look -= 32;
      }
      if (lane == 0)
        atomicExch((unsigned long long *)&a.tile_state[tile], ST_PREFIX | (excl + tile_total));
    }
    if (lane == 0) {
      s_excl = excl;
      if (tile == gridDim.x - 1) *a.accepted_out = excl + tile_total;
    }
  }
  __syncthreads();
  if (cnt == 0) return;
  int64_t slot = a.got_before + (int64_t)s_excl + woff + (inc - cnt);
  for (int k = 0; k < a.per_thread && slot < a.n; k++) {
    int64_t t = tb + k;
    int32_t idx;
    if (t < a.t_end && roulette_test(a, t, idx)) {
      a.idx_out[slot] = idx;
      if (slot == a.n - 1) *a.last_attempt = (unsigned long long)(t + 1);
      slot++;
    }
  }
}

cudaError_t launch_roulette(const RouletteArgs &a, int64_t tiles, cudaStream_t s) {
  cudaMemsetAsync(a.tile_state, 0, sizeof(uint64_t) * (size_t)tiles, s);
  cudaMemsetAsync(a.tile_counter, 0, sizeof(uint32_t), s);
  cudaMemsetAsync(a.accepted_out, 0, sizeof(unsigned long long), s);
  roulette_kernel<<<(unsigned)tiles, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// pf.clj:182-195 tournament-selection: N rounds of two uniform draws, keep the heavier
// (`(> w1 w2)` ? p1 : p2 — ties keep the second).
__global__ void __launch_bounds__(256)
tournament_kernel(const float *__restrict__ w, int64_t n, const int32_t *__restrict__ sidx,
                  uint64_t seed, int32_t *__restrict__ idx_out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int32_t a, b;
  if (sidx) {
    a = __ldg(sidx + 2 * k);
    b = __ldg(sidx + 2 * k + 1);
  } else {
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)((uint64_t)k >> 32), 0u, TAG_TOURNAMENT, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    a = (int32_t)(((uint64_t)r[0] * (uint64_t)n) >> 32);
    b = (int32_t)(((uint64_t)r[1] * (uint64_t)n) >> 32);
  }
  idx_out[k] = (__ldg(w + a) > __ldg(w + b)) ? a : b;
}

cudaError_t launch_tournament(const float *w, int64_t n, const int32_t *sidx, uint64_t seed,
                              int32_t *idx_out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  tournament_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, n, sidx, seed, idx_out);
  return cudaGetLastError();
}

// gather by an index list; the carried weight is the normalised weight q_i / T
// (pf.clj:160 `(/ w normalizer)` under the integer contract) or a constant.
__global__ void __launch_bounds__(256)
gather_idx_kernel(const int32_t *__restrict__ idx, int64_t n, const float *__restrict__ x,
                  const float *__restrict__ y, const float *__restrict__ th,
                  int wmode, const uint64_t *__restrict__ cdf, uint64_t total,
                  const float *__restrict__ wsrc, float wconst, float *ox, float *oy, float *oth,
                  float *ow) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int32_t i = __ldg(idx + k);
  ox[k] = __ldg(x + i);
  oy[k] = __ldg(y + i);
  oth[k] = __ldg(th + i);
  float wv = wconst;
  if (wmode == 1) {
    uint64_t q = __ldg(cdf + i) - (i > 0 ? __ldg(cdf + i - 1) : 0);
    wv = total ? (float)((double)q / (double)total) : 0.0f;
  } else if (wmode == 2) {
    wv = __ldg(wsrc + i);
  }
  ow[k] = wv;
}

cudaError_t launch_gather_idx(const int32_t *idx, int64_t n, const float *x, const float *y,
                              const float *th, int wmode, const uint64_t *cdf, uint64_t total,
                              const float *wsrc, float wconst, float *ox, float *oy, float *oth,
                              float *ow, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  gather_idx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, x, y, th, wmode, cdf, total,
                                                               wsrc, wconst, ox, oy, oth, ow);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
adopt_rows_kernel(const float4 *__restrict__ rows, int64_t n, float *x, float *y, float *th,
                  float *w) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 r = __ldg(rows + k);
  x[k] = r.x; y[k] = r.y; th[k] = r.z; w[k] = r.w;
}

cudaError_t launch_adopt(const float4 *rows, int64_t n, float *x, float *y, float *th, float *w,
                         cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  adopt_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rows, n, x, y, th, w);
  return cudaGetLastError();
}

__global__ void fill_kernel(float *p, int64_t n, float v) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) p[k] = v;
}
cudaError_t launch_fill(float *p, int64_t n, float v, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}

// ======================================================================================
// K6 pf.clj:67-84 pose-estimate: unweighted mean of x, y and of the planar quaternion
// components (0, 0, sin(theta/2), cos(theta/2)).
// ======================================================================================
constexpr int POSE_BLOCKS = 296;
__global__ void __launch_bounds__(256)
pose_kernel(const float *__restrict__ x, const float *__restrict__ y,
            const float *__restrict__ th, int64_t n, double *__restrict__ part) {
  double sx = 0, sy = 0, sz = 0, sw = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s, c;
    sincosf(0.5f * __ldg(th + i), &s, &c);
    sx += (double)__ldg(x + i);
    sy += (double)__ldg(y + i);
    sz += (double)s;
    sw += (double)c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  __shared__ double sm[8][4];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[wid][0] = sx; sm[wid][1] = sy; sm[wid][2] = sz; sm[wid][3] = sw; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int k = 0; k < 8; k++) t += sm[k][threadIdx.x];
    part[blockIdx.x * 4 + threadIdx.x] = t;
  }
}
int pose_blocks() { return POSE_BLOCKS; }
cudaError_t launch_pose(const float *x, const float *y, const float *th, int64_t n, double *part,
                        cudaStream_t s) {
  pose_kernel<<<POSE_BLOCKS, 256, 0, s>>>(x, y, th, n, part);
  return cudaGetLastError();
}

// ======================================================================================
// map.clj:59-84 uniform-initialization: proposals (x, y) ~ U[0,W) x U[0,H) in cell units,
// kept when cell (int x, int y) is inhabitable (map.clj:52-55 truncates), then scaled by
// the resolution (map.clj:29-34).  Each particle owns a Philox stream, so the result does
// not depend on how particles are spread over GPUs.
// ======================================================================================
__global__ void __launch_bounds__(256)
init_uniform_kernel(const int8_t *__restrict__ cells, int W, int H, float res, int64_t gid0,
                    int64_t n, int heading_mode, uint64_t seed, float *x, float *y, float *th,
                    float *w, float w0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t gid = (uint64_t)(gid0 + i);
  for (uint32_t at = 0;; at++) {
    uint32_t r[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), at, TAG_INIT, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float px = u24(r[0]) * (float)W, py = u24(r[1]) * (float)H;
    int cx = (int)px, cy = (int)py;
    bool bad = cx >= W || cy >= H || cx < 0 || cy < 0 || cells[(size_t)cy * W + cx] != 0;
    if (at < 100000u && bad) continue;
    x[i] = px * res;
    y[i] = py * res;
    if (heading_mode == 0) /* map.clj:67-68: N(0,1) rad */
      th[i] = sqrtf(-2.0f * logf(u24(r[2]))) * cosf(6.283185307179586f * u24(r[3]));
    else
      th[i] = (u24(r[2]) * 2.0f - 1.0f) * 3.14159265358979f;
    w[i] = w0;
    break;
  }
}

cudaError_t launch_init_uniform(const int8_t *cells, int W, int H, float res, int64_t gid0,
                                int64_t n, int heading_mode, uint64_t seed, float *x, float *y,
                                float *th, float *w, float w0, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  init_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cells, W, H, res, gid0, n,
                                                                 heading_mode, seed, x, y, th, w, w0);
  return cudaGetLastError();
}

// map.clj:87-111 gaussian-initialization: x, y ~ N(pose, sd_xy), heading rotated by
// N(0, sd_theta) (the reference passes sd = rot-mean = 0, map.clj:106).
__global__ void __launch_bounds__(256)
init_gaussian_kernel(float mx, float my, float mth, float sd_xy, float sd_th, int64_t gid0,
                     int64_t n, uint64_t seed, float *x, float *y, float *th, float *w, float w0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float n1, n2, n3;
  philox_normals3(seed, (uint64_t)(gid0 + i), TAG_GAUSS, n1, n2, n3);
  x[i] = fmaf(sd_xy, n1, mx);
  y[i] = fmaf(sd_xy, n2, my);
  th[i] = fmaf(sd_th, n3, mth);
  w[i] = w0;
}

cudaError_t launch_init_gaussian(float mx, float my, float mth, float sd_xy, float sd_th,
                                 int64_t gid0, int64_t n, uint64_t seed, float *x, float *y,
                                 float *th, float *w, float w0, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  init_gaussian_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mx, my, mth, sd_xy, sd_th, gid0,
                                                                  n, seed, x, y, th, w, w0);
  return cudaGetLastError();
}


// ======================================================================================
// pf.clj:243-279 augmented MCL (amcl*): w_avg = median of the weights; w_slow / w_fast are
// exponential averages of it; every output slot either injects a fresh uniform particle
// (map.clj:71-84) when Math/random > max(0, 1 - fast/slow) — the reference's inequality,
// inverted relative to Probabilistic Robotics table 8.3 — or runs roulette-step
// (pf.clj:169-180: rejection against the UN-normalised weights).
// ======================================================================================
constexpr uint32_t TAG_AUGMENT = 0x4155474du;

// k-th smallest by radix select on the order-preserving uint image of the floats: one
// 256-bin histogram pass per byte, most significant first.
__global__ void __launch_bounds__(256)
select_hist_kernel(const float *__restrict__ v, int64_t n, uint32_t prefix, uint32_t prefix_mask,
                   int shift, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t o = ord_of_float(__ldg(v + i));
    if ((o & prefix_mask) == prefix) atomicAdd(&sh[(o >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

cudaError_t launch_select_hist(const float *v, int64_t n, uint32_t prefix, uint32_t prefix_mask,
                               int shift, uint32_t *hist, int sm_count, cudaStream_t s) {
  cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), s);
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
  select_hist_kernel<<<(unsigned)blocks, 256, 0, s>>>(v, n, prefix, prefix_mask, shift, hist);
  return cudaGetLastError();
}

float ord_to_float_host(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

__global__ void __launch_bounds__(256)
augmented_kernel(const float *__restrict__ wgt, const float *__restrict__ x,
                 const float *__restrict__ y, const float *__restrict__ th, int64_t n,
                 float inject_threshold, const int8_t *__restrict__ cells, int W, int H, float res,
                 uint64_t seed, float *ox, float *oy, float *oth, float *ow, int32_t *oidx) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  uint32_t r[4];
  auto draw = [&](uint32_t attempt) {
    philox4x32_10((uint32_t)m, (uint32_t)((uint64_t)m >> 32), attempt, TAG_AUGMENT, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
  };
  draw(0u);
  if (u24(r[0]) > inject_threshold) { /* pf.clj:254-256: a fresh particle from uniform-initialization */
    for (uint32_t at = 1;; at++) {
      draw(at);
      float px = u24(r[0]) * (float)W, py = u24(r[1]) * (float)H;
      int cx = (int)px, cy = (int)py;
      bool bad = cx >= W || cy >= H || cx < 0 || cy < 0 || cells[(size_t)cy * W + cx] != 0;
      if (at < 100000u && bad) continue;
      ox[m] = px * res;
      oy[m] = py * res;
      oth[m] = sqrtf(-2.0f * logf(u24(r[2]))) * cosf(6.283185307179586f * u24(r[3])); /* map.clj:67-68 */
      ow[m] = 0.0f;
      oidx[m] = -1;
      return;
    }
  }
  for (uint32_t at = 1;; at++) { /* pf.clj:169-180 roulette-step */
    draw(at);
    int32_t i = (int32_t)(((uint64_t)r[0] * (uint64_t)n) >> 32);
    float wi = __ldg(wgt + i);
    if (wi > u24(r[1]) || at >= 100000u) {
      ox[m] = __ldg(x + i);
      oy[m] = __ldg(y + i);
      oth[m] = __ldg(th + i);
      ow[m] = wi;
      oidx[m] = i;
      return;
    }
  }
}

cudaError_t launch_augmented(const float *wgt, const float *x, const float *y, const float *th,
                             int64_t n, float inject_threshold, const int8_t *cells, int W, int H,
                             float res, uint64_t seed, float *ox, float *oy, float *oth, float *ow,
                             int32_t *oidx, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  augmented_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(wgt, x, y, th, n, inject_threshold, cells, W,
                                                              H, res, seed, ox, oy, oth, ow, oidx);
  return cudaGetLastError();
}


// ======================================================================================
// Sharded systematic resampling as a PULL over peer memory: every rank fills its own output
// slots [slot0, slot0 + count).  Slot m's threshold U_m falls into the CDF interval of exactly
// one rank g (exclusive prefix of the all-gathered totals); the thread binary-searches rank g's
// CDF and copies that particle — plain loads on rank g's memory, local or across NVLink/NVSwitch
// (IPC-mapped peer pointers).  No staging rows, no all-to-all, no host round trip; with evenly
// spread weights almost every slot resolves locally.  Same integer arithmetic as the single-GPU
// gather, so the global list is identical.
// ======================================================================================
__global__ void __launch_bounds__(256) pull_kernel(const PullArgs a) {
  __shared__ SysPlan s_plan;
  __shared__ uint64_t s_off[MAX_WORLD + 1];
  if (threadIdx.x == 0) {
    uint64_t T = 0;
    for (int g = 0; g < a.world; g++) {
      s_off[g] = T;
      T += a.totals[g];
    }
    s_off[a.world] = T;
    SysPlan p;
    p.total = T;
    p.n = (uint64_t)a.n_global;
    p.a = T / p.n;
    p.b = T % p.n;
    p.r = __umul64hi(a.u0_bits, T);
    s_plan = p;
    if (T == 0 && blockIdx.x == 0) atomicAdd(a.zero_flag, 1ull);
  }
  __syncthreads();
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const SysPlan plan = s_plan;
  /* as in the single-GPU gather: when the CTA's first and last slot draw from the same rank, the sources of
   * all its slots lie between theirs, and only two threads search that rank's whole CDF */
  __shared__ int64_t s_win[2];
  __shared__ int s_wing[2];
  if (plan.total != 0 && (threadIdx.x == 0 || threadIdx.x == blockDim.x - 1)) {
    const int64_t kk = min(k, a.count - 1);
    const uint64_t U = sys_threshold(plan, (uint64_t)(a.slot0 + kk));
    int gg = 0;
    for (; gg + 1 < a.world && U >= s_off[gg + 1]; gg++) {}
    const uint64_t key = U - s_off[gg];
    const uint64_t *cdf = a.peer[gg].cdf;
    int64_t lo = 0, hi = a.peer[gg].n;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (cdf[mid] > key) hi = mid; else lo = mid + 1;
    }
    s_win[threadIdx.x == 0 ? 0 : 1] = lo;
    s_wing[threadIdx.x == 0 ? 0 : 1] = gg;
  }
  __syncthreads();
  if (k >= a.count) return;
  const int64_t m = a.slot0 + k;
  int g = 0;
  int64_t i = 0;
  if (plan.total == 0) { /* degenerate: all weights zero -> keep the particle that sits in the slot */
    for (g = 0; g + 1 < a.world && m >= a.peer[g + 1].gid0; g++) {}
    i = m - a.peer[g].gid0;
  } else {
    const uint64_t U = sys_threshold(plan, (uint64_t)m);
    for (g = 0; g + 1 < a.world && U >= s_off[g + 1]; g++) {}
    const uint64_t key = U - s_off[g];
    const uint64_t *cdf = a.peer[g].cdf;
    int64_t lo = 0, hi = a.peer[g].n;
    if (s_wing[0] == s_wing[1] && s_wing[0] == g) { lo = s_win[0]; hi = min(s_win[1] + 1, a.peer[g].n); }
    while (lo < hi) { /* first i with cdf[i] > key */
      int64_t mid = (lo + hi) >> 1;
      if (cdf[mid] > key) hi = mid; else lo = mid + 1;
    }
    i = lo < a.peer[g].n ? lo : a.peer[g].n - 1;
  }
  a.ox[k] = a.peer[g].x[i];
  a.oy[k] = a.peer[g].y[i];
  a.oth[k] = a.peer[g].th[i];
  a.ow[k] = a.wout;
  if (a.oidx64) a.oidx64[k] = a.peer[g].gid0 + i;
}

// ======================================================================================
// The two exchanges of a sharded update, done by the kernels themselves over peer
// memory instead of by a collective library: every rank owns a mailbox with one slot per rank;
// a rank PUBLISHES by storing into its slot of every rank's mailbox (plain stores over NVLink, or
// into its own memory), tagged with the update's generation, and WAITS by polling its own
// mailbox — local memory — until every slot carries this generation.  12 bytes per rank and
// update; no host round trip, nothing for the host framework to schedule.
//   max:    one 64-bit word {generation : 32 | ordered-uint of the rank's largest raw weight : 32}
//   totals: the 64-bit fixed-point total, then (after a system-scope fence) its generation word;
//           the reader takes them in the opposite order.  The publishing kernel runs after the
//           scan in stream order, so a peer that has seen the generation may read that rank's CDF.
// ======================================================================================
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void publish_max_kernel(const MailboxArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  const unsigned long long word = ((unsigned long long)a.gen << 32) | (unsigned long long)(*a.rawmax_ord);
  st_sys_u64(a.peer_mbox[g] + MBOX_MAX + a.rank, word);
}

__global__ void wait_max_kernel(const MailboxArgs a) {
  const int g = threadIdx.x;
  uint32_t ord = 0;
  if (g < a.world) {
    unsigned long long w;
    long long spins = 0;
    do { w = ld_sys_u64(a.mbox + MBOX_MAX + g); } while ((uint32_t)(w >> 32) != a.gen && ++spins < MBOX_SPIN_LIMIT);
    if ((uint32_t)(w >> 32) != a.gen && a.error_flag) atomicAdd(a.error_flag, 1ull); /* a rank never published: fail loudly */
    ord = (uint32_t)w;
  }
  ord = __reduce_max_sync(0xffffffffu, ord);
  if (g == 0) *a.rawmax = ord == 0 ? -INFINITY : float_of_ord(ord); /* the global max, as decode_max_kernel leaves the local one */
}

__global__ void publish_total_kernel(const MailboxArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  st_sys_u64(a.peer_mbox[g] + MBOX_TOTAL + a.rank, (unsigned long long)*a.total);
  __threadfence_system();
  st_sys_u64(a.peer_mbox[g] + MBOX_TOTAL_GEN + a.rank, (unsigned long long)a.gen);
}

__global__ void wait_total_kernel(const MailboxArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  long long spins = 0;
  while ((uint32_t)ld_sys_u64(a.mbox + MBOX_TOTAL_GEN + g) != a.gen && ++spins < MBOX_SPIN_LIMIT) {}
  if (spins >= MBOX_SPIN_LIMIT && a.error_flag) atomicAdd(a.error_flag, 1ull);
  __threadfence_system();
  a.totals_all[g] = ld_sys_u64(a.mbox + MBOX_TOTAL + g);
}

// A barrier among the shards on the DEVICE time line: every rank stores the call's count into its slot of every
// rank's mailbox and polls its own until all slots carry it.  What follows it in the stream starts at the same
// instant on every GPU (to within a store over NVLink), whatever the skew between the hosts that enqueued it.
__global__ void rendezvous_kernel(const MailboxArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  st_sys_u64(a.peer_mbox[g] + MBOX_SYNC + a.rank, (unsigned long long)a.gen);
  long long spins = 0;
  while ((uint32_t)ld_sys_u64(a.mbox + MBOX_SYNC + g) != a.gen && ++spins < MBOX_SPIN_LIMIT) {}
  if (spins >= MBOX_SPIN_LIMIT && a.error_flag) atomicAdd(a.error_flag, 1ull);
}

cudaError_t launch_mailbox(const MailboxArgs &a, int what, cudaStream_t s) {
  switch (what) {
    case 4: rendezvous_kernel<<<1, 32, 0, s>>>(a); break;
    case 0: publish_max_kernel<<<1, 32, 0, s>>>(a); break;
    case 1: wait_max_kernel<<<1, 32, 0, s>>>(a); break;
    case 2: publish_total_kernel<<<1, 32, 0, s>>>(a); break;
    default: wait_total_kernel<<<1, 32, 0, s>>>(a); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_pull(const PullArgs &a, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  pull_kernel<<<(unsigned)((a.count + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}


// ======================================================================================
// Directional ("wedge") distance fields.  For a direction class (steep, x step, y step, slope
// bin [lo, hi]) and a free cell c, w(c) is the smallest k >= 1 such that some cell a Bresenham
// line of that class could occupy k steps after c is blocked: major offset k, minor offset in
// [floor(k*lo), floor(k*hi) + 1]  (the line's own offset is floor(k*s) or floor(k*s) + 1 for its
// slope s in [lo, hi], whatever its error state at c).  All cells of the line at steps 1..w-1 are
// therefore free, so the walk may jump w steps — the same exactness argument as for the isotropic
// field, but obstacles beside or behind the ray no longer shorten the jump.
// ======================================================================================
__global__ void __launch_bounds__(128)
wedge_build_kernel(const int8_t *__restrict__ cells, int W, int H, int pitch, size_t plane,
                   uint8_t *__restrict__ out) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y, cls = blockIdx.z;
  if (px >= W + 2) return;
  const int bin = cls % WEDGE_BINS, ysn = (cls / WEDGE_BINS) & 1, xsn = (cls / (2 * WEDGE_BINS)) & 1,
            steep = (cls / (4 * WEDGE_BINS)) & 1;
  const int xs = xsn ? -1 : 1, ys = ysn ? -1 : 1;
  auto blocked = [&](int qx, int qy) -> bool { /* padded coordinates */
    if (qx <= 0 || qx >= W + 1 || qy <= 0 || qy >= H + 1) return true;
    return cells[(size_t)(qy - 1) * W + (qx - 1)] != 0;
  };
  int d = 0;
  if (!blocked(px, py)) {
    d = 255;
    for (int k = 1; k < 255 && d == 255; k++) {
      int jlo = (k * bin) / WEDGE_BINS, jhi = (k * (bin + 1)) / WEDGE_BINS + 1;
      for (int j = jlo; j <= jhi; j++) {
        int da = xs * k, db = ys * j;
        int qx = steep ? px + db : px + da, qy = steep ? py + da : py + db;
        if (blocked(qx, qy)) { d = k; break; }
      }
    }
  }
  out[(size_t)cls * plane + (size_t)py * pitch + px] = (uint8_t)d;
}

cudaError_t build_wedge_fields(amclj_ctx *c, cudaStream_t s) {
  DeviceMap &m = c->map;
  m.wedge_plane = (size_t)m.pitch8 * (m.H + 2);
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&m.wedge), m.wedge_plane * WEDGE_CLASSES);
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(m.wedge, 0, m.wedge_plane * WEDGE_CLASSES, s);
  dim3 g((m.W + 2 + 127) / 128, m.H + 2, WEDGE_CLASSES);
  wedge_build_kernel<<<g, 128, 0, s>>>(m.cells, m.W, m.H, m.pitch8, m.wedge_plane, m.wedge);
  e = cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

}  // namespace amclj
