/* A plain-C caller of libamclj_b200.so: the ABI of include/amclj_b200.h used the way a non-Python
 * host (the JNA shim, a C++ node) would.  8x8 map with a wall at x = 6; one filter update.
 * Built and run by tests/test_gpu_c_harness.py:  gcc abi_harness.c -I include -L amclj_b200 -lamclj_b200 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "amclj_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int32_t rc_ = (call);                                                        \
    if (rc_ != AMCLJ_OK) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? amclj_last_error(ctx) : "(no ctx)"); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(void) {
  amclj_ctx *ctx = NULL;
  if (amclj_abi_version() != AMCLJ_ABI_VERSION) return 2;
  CHECK(amclj_create(0, &ctx));
  amclj_params p;
  CHECK(amclj_params_default(&p, AMCLJ_PRESET_NS));
  CHECK(amclj_set_params(ctx, &p));
  int8_t cells[64];
  memset(cells, 0, sizeof(cells));
  for (int y = 0; y < 8; y++) cells[y * 8 + 6] = 100;
  CHECK(amclj_set_map(ctx, cells, 8, 8, 1.0f));
  enum { N = 64 };
  float x[N], y[N], th[N], w[N];
  for (int i = 0; i < N; i++) { x[i] = 1.0f + 0.05f * i; y[i] = 3.0f; th[i] = 0.0f; }
  CHECK(amclj_set_particles(ctx, x, y, th, NULL, N));
  int32_t hx[4], hy[4], steps[4];
  uint8_t hit[4];
  float rng[4];
  CHECK(amclj_raycast_debug(ctx, 4, 0.0f, 1.5707964f, 10.0f, 20, 1, hx, hy, steps, hit, rng));
  /* particle 20 sits at (2.0, 3.0): beam 0 (+x) meets the wall at x = 6 after 4 steps */
  if (!(hit[0] == 1 && hx[0] == 6 && hy[0] == 3 && steps[0] == 4 && rng[0] == 4.0f)) {
    fprintf(stderr, "unexpected ray: hit %d cell (%d,%d) steps %d range %g\n", hit[0], hx[0], hy[0], steps[0], rng[0]);
    return 3;
  }
  float ranges[4] = {4.0f, 5.0f, 2.0f, 3.0f};
  double curr[3] = {0.1, 0.0, 0.0}, prev[3] = {0.0, 0.0, 0.0};
  CHECK(amclj_filter_update(ctx, curr, prev, NULL, 42u, ranges, 4, 0.0f, 1.5707964f, 10.0f, 0.25));
  CHECK(amclj_get_particles(ctx, x, y, th, w));
  double est[4];
  CHECK(amclj_pose_estimate(ctx, est));
  float sum = 0;
  for (int i = 0; i < N; i++) sum += w[i];
  if (fabsf(sum - 1.0f) > 1e-5f) { fprintf(stderr, "weights sum to %g\n", sum); return 4; }
  /* the scan says the wall is 4 m ahead: the resampled cloud sits near x = 2 */
  if (!(est[0] > 1.5 && est[0] < 2.6)) { fprintf(stderr, "estimate x = %g\n", est[0]); return 5; }
  amclj_stats st;
  CHECK(amclj_get_stats(ctx, &st));
  printf("C_HARNESS_OK estimate x %.3f y %.3f, update %.3f ms\n", est[0], est[1], st.ms_total);
  CHECK(amclj_destroy(ctx));
  return 0;
}
