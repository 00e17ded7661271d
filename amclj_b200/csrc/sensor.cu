// sensor.cu — K2: laser measurement model.  One warp per particle, one lane per beam:
// ray-cast through the occupancy grid and score with the beam mixture.
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
// (the reference's error accumulator, sensor.clj:123-127, unrolled).  First blocked cell,
// step count and range are therefore bit-identical to the cell-by-cell walk.
#include "internal.h"

namespace amclj {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint32_t ord_of_float(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct RayOut {
  int k, j, hit;
  int probes;
};

template <int MODE>
__device__ __forceinline__ uint32_t df_probe(const uint8_t *__restrict__ df, int idx) {
  if (MODE == DF_GLOBAL8) return __ldg(df + idx);
  uint32_t byte = (MODE == DF_SMEM4) ? df[idx >> 1] : __ldg(df + (idx >> 1));
  return (byte >> ((idx & 1) << 2)) & 15u;
}

// sensor.clj:130-139 map-range + :112-128 extract-line for one beam.
template <int MODE, bool DIAG>
__device__ __forceinline__ float cast_ray(const SensorLaunch &a, const uint8_t *__restrict__ df,
                                          float ox, float oy, float ct, float st, int x0, int y0,
                                          bool start_in, float cb, float sb, RayOut &out,
                                          int &hx, int &hy, int &steps) {
  float c = cb, s = sb;
  if (a.s1) { /* S1: theta + obs-bearing by the angle-addition identity */
    c = fmaf(ct, cb, -(st * sb));
    s = fmaf(st, cb, ct * sb);
  }
  /* sensor.clj:99-103 project-pose, map.clj:36-46 world->map */
  int x1 = round_cell(fmaf(a.R, c, ox), a.res);
  int y1 = round_cell(fmaf(a.R, s, oy), a.res);
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
  int dx = abs(a0 - a1), dy = abs(b0 - b1);
  /* the miss test runs AFTER the cell test at step k: S4=1 `x == max-x`;
   * S4=0 (sensor.clj:121) `x > max-x` */
  int kmax;
  if (a.s4) kmax = dx;
  else if (xs > 0) kmax = dx + 1;
  else if (dx > 0) kmax = 0;
  else { /* dx = dy = 0: `2*err >= 0` always holds -> diagonal walk until a blocked cell */
    kmax = 0x3fffffff; dx = 1; dy = 1;
  }
  uint32_t M; int sh;
  if (dx < a.ndiv) {
    uint2 e = __ldg(a.divtab + dx);
    M = e.x; sh = (int)e.y;
  } else if (dx > 0) { /* never taken for validated ranges; kept exact */
    uint32_t D = 2u * (uint32_t)dx;
    int L = 31 - __clz(D);
    M = (uint32_t)(((1ull << (31 + L)) + D - 1) / D);
    sh = L - 1;
  } else { M = 0; sh = 0; }
  int stepA = xs * (steep ? a.row_nib : 1), stepB = ys * (steep ? 1 : a.row_nib);
  /* padded field: cell (x, y) lives at (y+1)*row + (x+1); entry 0 is a ring cell (d = 0),
   * which makes an out-of-range start an immediate hit (the start cell is tested first) */
  int idx = start_in ? (y0 + 1) * a.row_nib + (x0 + 1) : 0;
  int k = 0, j = 0, err = 0, hit, probes = 0;
  for (;;) {
    uint32_t d = df_probe<MODE>(df, idx);
    if (DIAG) probes++;
    if (d == 0) { hit = 1; break; }
    k += (int)d;
    if (k > kmax) { hit = 0; break; }
    err += (int)d * dy;
    uint32_t t = (uint32_t)(2 * err + dx);
    int jd = (int)(__umulhi(t, M) >> sh);
    err -= jd * dx;
    j += jd;
    idx += (int)d * stepA + jd * stepB;
  }
  float rng;
  if (hit) {
    uint32_t d2 = (uint32_t)k * (uint32_t)k + (uint32_t)j * (uint32_t)j;
    rng = __fmul_rn(__fsqrt_rn(__uint2float_rn(d2)), a.res);
  } else {
    rng = a.miss_value;
    k = kmax;
  }
  out.k = k; out.hit = hit; out.probes = probes;
  if (DIAG) {
    if (!hit) j = dx > 0 ? (int)((2ll * k * dy + dx) / (2ll * dx)) : 0;
    int ca = a0 + xs * k, cbb = b0 + ys * j;
    hx = steep ? cbb : ca;
    hy = steep ? ca : cbb;
    steps = k;
  }
  out.j = j;
  return rng;
}

constexpr int kMaxThreads = 1024;

template <int MODE, bool DIAG>
__global__ void __launch_bounds__(kMaxThreads, 1) sensor_kernel(const SensorLaunch a) {
  extern __shared__ uint4 smem_dyn[];
  __shared__ uint64_t bar;
  const uint8_t *df = a.df;
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
    df = sdf;
  }
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * wpb;
  const float *cbp = a.beam, *sbp = a.beam + a.nb, *obsp = a.beam + 2 * a.nb,
              *e2p = a.beam + 3 * a.nb, *p34p = a.beam + 4 * a.nb;
  float wmax = -INFINITY;
  unsigned long long n_cells = 0, n_probes = 0, n_hits = 0, n_rays = 0;
  for (int64_t p = (int64_t)(threadIdx.x >> 5) * gridDim.x + blockIdx.x; p < a.n; p += nwarps) {
    float ox = __ldg(a.x + p), oy = __ldg(a.y + p), th = __ldg(a.th + p);
    float st, ct;
    sincos_det(th, st, ct);
    if (a.s10) { /* S10: base->laser offset (sensor.clj:78-81), off by default */
      float nx = ox + fmaf(a.laser_x, ct, -(a.laser_y * st));
      float ny = oy + fmaf(a.laser_x, st, a.laser_y * ct);
      ox = nx; oy = ny;
    }
    int x0 = round_cell(ox, a.res), y0 = round_cell(oy, a.res);
    bool start_in = (uint32_t)x0 < (uint32_t)a.W && (uint32_t)y0 < (uint32_t)a.H;
    bool dbg = DIAG && a.dbg_count > 0 && p >= a.dbg_first && p < a.dbg_first + a.dbg_count;
    float acc = 0.0f;
    for (int jb = lane; jb < a.nb; jb += 32) {
      float cb = __ldg(cbp + jb), sb = __ldg(sbp + jb), obs = __ldg(obsp + jb);
      RayOut ro;
      int hx = 0, hy = 0, steps = 0;
      float rng = cast_ray<MODE, DIAG>(a, df, ox, oy, ct, st, x0, y0, start_in, cb, sb, ro, hx,
                                       hy, steps);
      /* sensor.clj:149-162 beam mixture; pz2's exponential and pz3/pz4 depend on the beam
       * only and come precomputed (e2, p34) */
      float z = obs - rng;
      float pz1 = a.z_hit * expf(-(z * z) * a.inv2s2);
      float pz2 = z < 0.0f ? __ldg(e2p + jb) : 0.0f;
      float sum = (pz1 + pz2) + __ldg(p34p + jb);
      acc += a.s7 ? logf(sum) : (sum * sum) * sum;
      if (DIAG) {
        n_rays++; n_cells += (unsigned)ro.k + 1u; n_probes += (unsigned)ro.probes;
        n_hits += (unsigned)ro.hit;
        if (dbg) {
          int64_t o = (p - a.dbg_first) * a.nb + jb;
          if (a.dbg_hit_x) a.dbg_hit_x[o] = hx;
          if (a.dbg_hit_y) a.dbg_hit_y[o] = hy;
          if (a.dbg_steps) a.dbg_steps[o] = steps;
          if (a.dbg_hit) a.dbg_hit[o] = (uint8_t)ro.hit;
          if (a.dbg_range) a.dbg_range[o] = rng;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && a.raw) a.raw[p] = acc;
    wmax = fmaxf(wmax, acc);
  }
  if (lane == 0 && a.rawmax_ord && wmax > -INFINITY) atomicMax(a.rawmax_ord, ord_of_float(wmax));
  if (DIAG && a.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_rays += __shfl_xor_sync(0xffffffffu, n_rays, o);
      n_cells += __shfl_xor_sync(0xffffffffu, n_cells, o);
      n_probes += __shfl_xor_sync(0xffffffffu, n_probes, o);
      n_hits += __shfl_xor_sync(0xffffffffu, n_hits, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, n_rays);
      atomicAdd(a.stats + 1, n_cells);
      atomicAdd(a.stats + 2, n_probes);
      atomicAdd(a.stats + 3, n_hits);
    }
  }
}

template <int MODE, bool DIAG>
static cudaError_t launch_t(const amclj_ctx *c, const SensorLaunch &a, cudaStream_t s) {
  auto kern = sensor_kernel<MODE, DIAG>;
  size_t smem = MODE == DF_SMEM4 ? a.df_bytes : 0;
  if (smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  /* one CTA per SM (the staged field owns the SM's shared memory); warps sized to the work */
  int64_t want_warps = (a.n + c->sm_count - 1) / c->sm_count;
  int wpb = (int)(want_warps < 4 ? 4 : (want_warps > 32 ? 32 : want_warps));
  int64_t blocks = (a.n + wpb - 1) / wpb;
  int per_sm = MODE == DF_SMEM4 ? 1 : 2;
  if (blocks > (int64_t)c->sm_count * per_sm) blocks = (int64_t)c->sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, wpb * 32, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_sensor(const amclj_ctx *c, const SensorLaunch &a, int mode, bool diag,
                          cudaStream_t s) {
  if (a.n <= 0 || a.nb <= 0) return cudaSuccess;
  switch (mode) {
    case DF_SMEM4: return diag ? launch_t<DF_SMEM4, true>(c, a, s) : launch_t<DF_SMEM4, false>(c, a, s);
    case DF_GLOBAL8: return diag ? launch_t<DF_GLOBAL8, true>(c, a, s) : launch_t<DF_GLOBAL8, false>(c, a, s);
    default: return diag ? launch_t<DF_GLOBAL4, true>(c, a, s) : launch_t<DF_GLOBAL4, false>(c, a, s);
  }
}

}  // namespace amclj
