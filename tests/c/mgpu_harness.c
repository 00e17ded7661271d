/* One process, one caller thread, several GPUs — the way a JVM host would drive them (the reference's
 * filter loop is one `go` block, pf.clj:297): amclj_create_multi over the devices on the command line (a
 * device may repeat: emulated shards), a synthetic walled map, uniform initialisation, three filter updates
 * through amclj_mfilter_update_async and, once, through the host-buffer call; the same updates on ONE
 * amclj_ctx holding all particles must give identical particles, bit for bit.
 *   gcc mgpu_harness.c -I include -L amclj_b200 -lamclj_b200;  ./mgpu_harness 200000 0 0     (2 shards on GPU 0)
 *                                                              ./mgpu_harness 1000000 0 1 2 3 4 5 6 7 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "amclj_b200.h"

#define MCHECK(call)                                                                              \
  do {                                                                                            \
    int32_t rc_ = (call);                                                                         \
    if (rc_ != AMCLJ_OK) {                                                                        \
      fprintf(stderr, "%s -> %d: %s | %s\n", #call, rc_, m ? amclj_mlast_error(m) : "", one ? amclj_last_error(one) : ""); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

enum { W = 400, H = 300, NB = 90 };

int main(int argc, char **argv) {
  amclj_mctx *m = NULL;
  amclj_ctx *one = NULL;
  if (argc < 3) { fprintf(stderr, "usage: %s n_particles device [device ...]\n", argv[0]); return 2; }
  const long long n = atoll(argv[1]);
  int32_t devs[16];
  const int nd = argc - 2;
  if (nd > 16 || n < nd) return 2;
  for (int i = 0; i < nd; i++) devs[i] = atoi(argv[2 + i]);
  /* a hall with a border wall, a few inner walls and a patch of unknown cells */
  int8_t *cells = calloc((size_t)W * H, 1);
  for (int x = 0; x < W; x++) { cells[x] = 100; cells[(H - 1) * W + x] = 100; }
  for (int y = 0; y < H; y++) { cells[y * W] = 100; cells[y * W + W - 1] = 100; }
  for (int y = 40; y < 220; y++) cells[y * W + 130] = 100;
  for (int x = 200; x < 360; x++) cells[150 * W + x] = 100;
  for (int y = 230; y < 260; y++) for (int x = 30; x < 80; x++) cells[y * W + x] = -1;
  float ranges[NB];
  for (int i = 0; i < NB; i++) ranges[i] = 1.0f + 0.04f * (float)((i * 37) % 97);
  const float amin = -3.1415927f, ainc = 6.2831855f / NB, rmax = 5.0f;
  amclj_params p;
  amclj_params_default(&p, AMCLJ_PRESET_NS);
  MCHECK(amclj_create_multi(devs, nd, &m));
  MCHECK(amclj_create(devs[0], &one));
  MCHECK(amclj_mset_params(m, &p));
  MCHECK(amclj_set_params(one, &p));
  MCHECK(amclj_mset_map(m, cells, W, H, 0.05f));
  MCHECK(amclj_set_map(one, cells, W, H, 0.05f));
  MCHECK(amclj_minit_uniform(m, n, 1, 77u));
  MCHECK(amclj_init_uniform(one, n, 1, 77u));
  float *a[4], *b[4];
  for (int k = 0; k < 4; k++) { a[k] = malloc(sizeof(float) * (size_t)n); b[k] = malloc(sizeof(float) * (size_t)n); }
  MCHECK(amclj_mstage_scan(m, ranges, NB, amin, ainc, rmax));
  long long bad = 0;
  for (int it = 0; it < 3; it++) {
    double curr[3] = {0.2 * (it + 1), 0.0, 0.1 * (it + 1)}, prev[3] = {0.2 * it, 0.0, 0.1 * it};
    const double u0 = 0.123 + 0.2 * it;
    MCHECK(amclj_mfilter_update_async(m, curr, prev, 1000u + (unsigned)it, u0));
    MCHECK(amclj_msynchronize(m));
    MCHECK(amclj_filter_update(one, curr, prev, NULL, 1000u + (unsigned)it, ranges, NB, amin, ainc, rmax, u0));
    MCHECK(amclj_mget_particles(m, a[0], a[1], a[2], a[3]));
    MCHECK(amclj_get_particles(one, b[0], b[1], b[2], b[3]));
    for (int k = 0; k < 3; k++) bad += memcmp(a[k], b[k], sizeof(float) * (size_t)n) != 0;
  }
  /* host-buffer entry point: PoseArray in, resampled PoseArray out, over all devices */
  {
    double curr[3] = {0.8, 0.0, 0.4}, prev[3] = {0.6, 0.0, 0.3};
    MCHECK(amclj_mfilter_update_host(m, a[0], a[1], a[2], a[3], n, curr, prev, 2000u, ranges, NB, amin, ainc, rmax, 0.61));
    MCHECK(amclj_filter_update_host(one, b[0], b[1], b[2], b[3], n, curr, prev, NULL, 2000u, ranges, NB, amin, ainc, rmax, 0.61));
    for (int k = 0; k < 3; k++) bad += memcmp(a[k], b[k], sizeof(float) * (size_t)n) != 0;
  }
  double em[4], eo[4];
  MCHECK(amclj_mpose_estimate(m, em));
  MCHECK(amclj_pose_estimate(one, eo));
  if (fabs(em[0] - eo[0]) > 1e-9 * fabs(eo[0]) + 1e-12 || fabs(em[1] - eo[1]) > 1e-9 * fabs(eo[1]) + 1e-12) bad += 100;
  amclj_stats st;
  MCHECK(amclj_get_stats(amclj_multi_ctx(m, 0), &st));
  printf("%s shards=%d n=%lld mismatching_arrays=%lld rank0_update_ms=%.3f sensor_path=%d\n", bad ? "MGPU_MISMATCH" : "MGPU_OK",
         nd, n, bad, st.ms_total, st.sensor_path);
  amclj_destroy_multi(m);
  amclj_destroy(one);
  return bad ? 3 : 0;
}
