// raycast.cuh — one ray of sensor.clj:130-139 map-range / :112-128 extract-line from integer start and
// end cells, with the distance-field jumps of sensor.cu; shared by the per-update ray tables
// (raytable.cu) and the persistent map-wide ray tables (ptable.cu).
#pragma once
#include "internal.h"
#include "sensor_common.cuh"

namespace amclj {

struct RayOut { /* diagnostic view of one ray */
  int kk;       /* steps along the major axis (a miss: up to the last cell tested) */
  int hit;
  int hx, hy;   /* blocked cell (a miss: the last cell tested) */
  int probes, oob;
};

// One ray from cell (x0, y0) towards cell (x1, y1): sensor.clj:130-139 map-range after the two
// roundings, :112-128 extract-line with distance-field jumps.  The same arithmetic as the walk
// in sensor.cu (set-up (2), probe step (3), range (1)), one probe per loop trip.
template <int MODE, bool WIDE, bool DIAG>
__device__ __forceinline__ float cast_ray(const SensorLaunch &a, const uint8_t *field, int x0, int y0, int x1, int y1,
                                          RayOut &out) {
  constexpr bool BYTES = MODE == DF_GLOBAL8 || MODE == DF_WEDGE8;
  constexpr int idx_scale = BYTES ? 1 : 4;
  constexpr uint32_t DMASK = BYTES ? 255u : 15u;
  Field df;
  df.g = field;
  df.s = 0;
  const int row = a.row_nib * idx_scale;
  const bool start_in = (uint32_t)x0 < (uint32_t)a.W && (uint32_t)y0 < (uint32_t)a.H;
  uint32_t idx = start_in ? (uint32_t)((y0 + 1) * row + (x0 + 1) * idx_scale) : 0u;
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
  int dx = abs(a0 - a1), dy = abs(b0 - b1), kmax;
  if (a.s4) kmax = dx;
  else if (xs > 0) kmax = dx + 1;
  else if (dx > 0) kmax = 0;
  else { kmax = 0x3fffffff; dx = 1; dy = 1; } /* diagonal walk to a blocked cell */
  const uint2 e = __ldg(a.divtab + min(dx, a.ndiv - 1));
  const uint32_t M = e.x, sh = e.y;
  const int stepA = xs * (steep ? row : idx_scale), stepB = ys * (steep ? idx_scale : row);
  const int ndx = -dx;
  int k = 0, j = 0, err = 0;
  if (MODE == DF_WEDGE8 && start_in) { /* the plane of this ray's direction class */
    const uint32_t P = (uint32_t)a.wedge_plane;
    uint32_t off = steep ? 4 * WEDGE_BINS * P : 0;
    off += xs < 0 ? 2 * WEDGE_BINS * P : 0;
    off += ys < 0 ? WEDGE_BINS * P : 0;
    const int dyb = dy * WEDGE_BINS;
#pragma unroll
    for (int t = 1; t < WEDGE_BINS; t++) off += (dyb >= t * dx) ? P : 0;
    idx += off;
  }
  const uint32_t idx_limit = (uint32_t)(a.df_bytes * (BYTES ? 1u : 8u));
  int probes = 0, oob = 0;
  uint32_t am = DMASK;
  while (am) {
    if (DIAG && idx >= idx_limit) { oob++; idx = 0; }
    const uint32_t d = df_probe<MODE>(df, idx, am);
    if (DIAG) probes++;
    k += (int)d;
    const bool fin = d == 0 || k > kmax;
    const int e2 = err + (int)d * dy;
    const uint32_t t = (uint32_t)(2 * e2 - ndx);
    const uint32_t jd = WIDE ? (__umulhi(t, M) >> sh) : __umulhi(t, M);
    err = e2 + (int)jd * ndx;
    j += (int)jd;
    idx += (uint32_t)((int)d * stepA + (int)jd * stepB);
    am = fin ? 0u : am;
  }
  const bool hit = k <= kmax;
  const uint32_t d2 = (uint32_t)k * (uint32_t)k + (uint32_t)j * (uint32_t)j;
  const float rng = hit ? __fmul_rn(sqrt_int(d2), a.res) : a.miss_value; /* sensor.clj:118-121 */
  if (DIAG) {
    int kk = hit ? k : kmax, jj = j;
    if (!hit) jj = ndx < 0 ? (int)((2ll * kk * dy - ndx) / (-2ll * ndx)) : 0;
    int ca = a0 + xs * kk, cb2 = b0 + ys * jj;
    out.kk = kk; out.hit = hit ? 1 : 0;
    out.hx = steep ? cb2 : ca; out.hy = steep ? ca : cb2;
    out.probes = probes; out.oob = oob;
  }
  return rng;
}

/* the rare paths of the look-up kernels, kept out of line so that their beam loops stay small: beam jb of the
 * particle at (ox, oy) with heading th, everything recomputed with the full handling of sensor.cu (a start off
 * the map, huge / NaN coordinates) and the refinement */
template <int MODE, bool WIDE, bool DIAG>
__device__ __noinline__ float cast_ray_slow(const SensorLaunch &a, float ox, float oy, float th, int jb, RayOut &out) {
  const ResDiv rdv = make_resdiv(a.res);
  float st, ct;
  sincos_det(th, st, ct);
  const RayStart rst = make_ray_start(a, ox, oy, st, ct);
  const float2 bd = __ldg(reinterpret_cast<const float2 *>(a.beam + 4 * (size_t)a.nb) + jb); /* (cos, sin) of the beam */
  float c = bd.x, s = bd.y;
  if (a.s1) {
    c = fmaf(ct, bd.x, -(st * bd.y));
    s = fmaf(st, bd.x, ct * bd.y);
  }
  int x1, y1;
  end_cell(a, rdv, rst, c, s, ox, oy, th, jb, x1, y1);
  return cast_ray<MODE, WIDE, DIAG>(a, a.df, rst.x0, rst.y0, x1, y1, out);
}

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

}  // namespace amclj
