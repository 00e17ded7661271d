/*
 * amclj_b200.h — C ABI of libamclj_b200.so: the B200-native implementation of the
 * per-update particle-filter hot path of jvia/amclj.
 *
 * The reference has no FFI/plugin interface of its own: the hot path is plain Clojure
 * called from mcl* (reference src/amclj/pf.clj:229-233).  The entry points below are the
 * smallest seam that keeps those call signatures, so a Clojure/JNA shim
 * can keep `apply-motion-model`, `apply-sensor-model`, `roulette-selection` and the
 * per-particle model functions unchanged.  Each entry point cites the reference function
 * it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns int32 status: AMCLJ_OK (0) or a negative AMCLJ_ERR_* code;
 *     amclj_last_error(ctx) gives a human-readable message owned by the ctx (valid until
 *     the next call on that ctx).  No exception crosses the boundary.
 *   - the caller owns every host buffer; the library owns all device memory behind the
 *     opaque amclj_ctx.  One ctx drives ONE GPU (one rank of a sharded filter).
 *   - a ctx is not thread-safe; distinct ctxs are independent (the reference has exactly
 *     one caller thread: the single `go` block at pf.clj:297).
 *   - calls are synchronous on return unless the name ends in `_async`; `_async` calls
 *     only enqueue work on the ctx stream (amclj_set_stream).
 *   - there is NO CPU fallback: without a usable CUDA device amclj_create fails with
 *     AMCLJ_ERR_NODEVICE.
 */
#ifndef AMCLJ_B200_H
#define AMCLJ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMCLJ_ABI_VERSION 2

#define AMCLJ_OK 0
#define AMCLJ_ERR_INVALID (-1)  /* bad argument */
#define AMCLJ_ERR_CUDA (-2)     /* CUDA runtime error, see amclj_last_error */
#define AMCLJ_ERR_STATE (-3)    /* call order: no map / particles / scan yet */
#define AMCLJ_ERR_NOMEM (-4)
#define AMCLJ_ERR_NODEVICE (-5) /* no usable CUDA device: there is no CPU fallback */

#define AMCLJ_PRESET_REF 0 /* the reference as written, bug for bug */
#define AMCLJ_PRESET_NS 1  /* north-star semantics (Probabilistic Robotics) */

#define AMCLJ_RESAMPLE_ROULETTE 0   /* pf.clj:156-167 */
#define AMCLJ_RESAMPLE_SYSTEMATIC 1 /* PR table 4.4, integer CDF */
#define AMCLJ_RESAMPLE_TOURNAMENT 2 /* pf.clj:182-195 */

/*
 * Model constants and the ten semantic switches.  POD, passed by
 * pointer.  Defaults = the Clojure `def` constants: sensor.clj:11-21, motion.clj:12-16.
 */
typedef struct amclj_params {
  /* S1 sensor.clj:131,133 — 0: ray direction = obs-bearing only (REF); 1: theta+bearing */
  int32_t s1_use_heading;
  /* S2 sensor.clj:12,133,121 — 0: ray length and miss value = global max_range (-1.0);
   * 1: the scan's rangeMax */
  int32_t s2_use_scan_range_max;
  /* S3 sensor.clj:137-138 — 0: non-steep rays get end = start (diagonal walk); 1: proper */
  int32_t s3_proper_endpoint;
  /* S4 sensor.clj:121 — 0: `(> x max-x)` tested after the cell test; 1: stop after
   * testing the endpoint cell */
  int32_t s4_endpoint_termination;
  /* S5 sensor.clj:151-153 — 0: z_hit doubles as sigma; 1: sigma_hit */
  int32_t s5_use_sigma_hit;
  /* S6 sensor.clj:13,144-145 — beams = range(0, n, int((n-1)/(max_beams-1)));
   * <= 0 means every beam */
  int32_t s6_max_beams;
  /* S7 sensor.clj:162,173 — 0: weight = sum (p)^3; 1: log-weight = sum log p */
  int32_t s7_log_weights;
  /* S8 pf.clj:156-167 — resampler used by amclj_filter_update (AMCLJ_RESAMPLE_*) */
  int32_t s8_resampler;
  /* S9 motion.clj:35 — 0: rot1 variance a1*rot1^2 + a2*rot2^2 (REF); 1: a2*trans^2 (PR) */
  int32_t s9_pr_rot1_variance;
  /* S10 sensor.clj:78-81,142 — 0: base->laser offset ignored (REF and NS default) */
  int32_t s10_apply_laser_offset;
  int32_t collect_stats; /* bit 0: count cells visited / probes (diagnostic kernels); bit 1: time the dominant
                          * sensor kernel alone (amclj_stats.ms_sensor_main; two more event records per update) */
  /* distance-field placement: 0 auto (maps whose 32 directional planes fit 64 MB: the planes,
   * walked in place for a compact particle cloud and in map-tile order for a spread-out one;
   * larger maps: one byte per cell in L2), 1 isotropic nibbles in shared memory, 2 isotropic bytes
   * in global memory, 3 isotropic nibbles in global memory,
   * 4 directional byte fields (32 direction classes) through L1/L2, 5 the same with the particles
   * walked in map-tile order (spread-out clouds), 6 per-start-cell ray tables for every occupied cell
   * (mode 8 takes them by itself when most particles share their start cell with many others),
   * 7 persistent map-wide ray tables: one table per FREE CELL OF THE MAP, filled once per (map, ray length,
   * S3/S4) and looked up by every later update of any cloud; mode 0 uses them whenever the map is small
   * enough (lgfloor: 604 MB) and falls back to the policy of mode 8 otherwise,
   * 8 the auto policy without the persistent tables (walk kernel / per-update tables picked per update) */
  int32_t df_mode;
  float max_range;    /* sensor.clj:12  (-1.0) */
  float z_hit;        /* sensor.clj:14 */
  float z_short;      /* sensor.clj:15 */
  float z_max;        /* sensor.clj:16 */
  float z_rand;       /* sensor.clj:17 */
  float sigma_hit;    /* sensor.clj:18 */
  float lambda_short; /* sensor.clj:19 */
  float laser_x;      /* sensor.clj:80 (0.135), used only when s10 != 0 */
  float laser_y;
  float alpha1; /* motion.clj:12 */
  float alpha2; /* motion.clj:13 */
  float alpha3; /* motion.clj:14 */
  float alpha4; /* motion.clj:15 */
  float reserved1[3];
} amclj_params;

typedef struct amclj_stats {
  double ms_motion;   /* last update, device time from CUDA events on the ctx stream */
  double ms_sensor;
  double ms_resample;
  double ms_total;
  uint64_t rays;          /* rays cast in the last sensor pass */
  uint64_t cells_visited; /* Bresenham cells the reference would have tested */
  uint64_t probes;        /* distance-field probes actually issued */
  uint64_t hits;
  int64_t n_particles;
  int32_t beams_used;
  int32_t map_in_smem; /* 1 when the last sensor pass walked the shared-memory field */
  uint64_t oob_probes; /* diagnostic passes only: probes outside the staged field (must be 0) */
  uint64_t table_cells;     /* diagnostic passes only: start cells that got a ray table (0: the walk kernel ran) */
  uint64_t table_particles; /* ... and the particles standing on them */
  uint64_t table_misses;    /* ... and their rays whose end cell was not in the table (walked instead; expected 0) */
  double ms_sensor_main;    /* the dominant kernel of the last sensor pass alone (sensor_kernel, or rt_lookup_kernel) */
  int32_t sensor_path;      /* what the last sensor pass ran on: 0 walk in map-tile order, 1 walk in place, 2 per-update ray tables,
                             * 3 persistent map-wide ray tables, -1 not known yet */
  int32_t kernels_launched; /* kernels the last filter update enqueued */
  uint64_t zero_weight_updates; /* updates so far whose weights were all zero: the resampler kept the particles
                                 * as they were (synchronous calls return AMCLJ_ERR_STATE for such an update) */
  double ms_table_fill;     /* what the last one-time fill of the persistent ray tables took (0: none yet) */
  uint64_t table_bytes;     /* device memory held by the persistent ray tables */
} amclj_stats;

typedef struct amclj_ctx amclj_ctx;

/* ---- lifecycle ------------------------------------------------------------------- */
int32_t amclj_abi_version(void);
/* Fill *p with a preset (AMCLJ_PRESET_REF / AMCLJ_PRESET_NS). */
int32_t amclj_params_default(amclj_params *p, int32_t preset);
/* One ctx per GPU.  `device` is a CUDA ordinal. */
int32_t amclj_create(int32_t device, amclj_ctx **out);
int32_t amclj_destroy(amclj_ctx *ctx);
const char *amclj_last_error(const amclj_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. the host framework's current stream).
 * NULL restores the ctx's own stream. */
int32_t amclj_set_stream(amclj_ctx *ctx, void *cuda_stream);
int32_t amclj_synchronize(amclj_ctx *ctx);
int32_t amclj_set_params(amclj_ctx *ctx, const amclj_params *p);
int32_t amclj_get_params(const amclj_ctx *ctx, amclj_params *p);
/* Shard identity for a particle set split over `world` GPUs: this ctx holds
 * global particles [global_offset, global_offset + n_local) of n_global. */
int32_t amclj_set_shard(amclj_ctx *ctx, int32_t rank, int32_t world, int64_t global_offset,
                        int64_t n_global);

/* ---- map: nav_msgs/OccupancyGrid as amclj sees it (nav_msgs.clj:27-38, map.clj:19-57) */
/* cells: int8 [height][width] row-major, -1 unknown / 0 free / 100 occupied. */
int32_t amclj_set_map(amclj_ctx *ctx, const int8_t *cells, int32_t width, int32_t height,
                      float resolution);

/* ---- particles: PoseArray <-> SoA (geometry_msgs.clj:54-72,224-242) ----------------- */
/* theta = quat/to-heading of the orientation (quaternions.clj:82-97), done by the shim. */
int32_t amclj_set_particles(amclj_ctx *ctx, const float *x, const float *y,
                            const float *theta, const float *w, int64_t n);
int32_t amclj_get_particles(amclj_ctx *ctx, float *x, float *y, float *theta, float *w);
int64_t amclj_num_particles(const amclj_ctx *ctx);
/* The same two calls for the reference's own record shape: poses7 = [n][7] doubles, one
 * geometry_msgs/Pose per row as (position.x, position.y, position.z, orientation.x, .y, .z, .w)
 * (geometry_msgs.clj:9,30,54).  set: theta = quat/to-heading(orientation), quaternions.clj:82-97 — pole
 * guards at 0.4999 / -0.499 included, no normalisation of q, position.z dropped as the filter never
 * reads it.  get: orientation = quat/from-heading(theta), quaternions.clj:57-62 = (0, 0, sin(theta/2),
 * cos(theta/2)), position.z = 0.  Both in double on the host, as on the JVM.  w (float[n]) may be NULL. */
int32_t amclj_set_poses_quat(amclj_ctx *ctx, const double *poses7, const float *w, int64_t n);
int32_t amclj_get_poses_quat(amclj_ctx *ctx, double *poses7, float *w);
/* map.clj:71-84 uniform-initialization / map.clj:108-111 gaussian-initialization, on device
 * (in-kernel Philox).  heading_mode 0: N(0,1) rad (REF, map.clj:67-68); 1: U(-pi,pi). */
int32_t amclj_init_uniform(amclj_ctx *ctx, int64_t n, int32_t heading_mode, uint64_t seed);
int32_t amclj_init_gaussian(amclj_ctx *ctx, int64_t n, double x, double y, double theta,
                            double sd_xy, double sd_theta, uint64_t seed);

/* ---- motion: apply-motion-model pf.clj:139-145 -> motion.clj:24-45,68-79 ----------- */
/* curr/prev = (x, y, yaw) of the odom->base transforms.  normals: host float [n][3]
 * standard normals (rot1, trans, rot2 draw per particle, motion.clj:35-38) or NULL to
 * draw them in-kernel from Philox4x32-10 keyed by (seed, global particle id). */
int32_t amclj_motion_update(amclj_ctx *ctx, const double curr[3], const double prev[3],
                            const float *normals, uint64_t seed);

/* ---- sensor: apply-sensor-model pf.clj:148-154 -> sensor.clj:141-173 ---------------- */
/* LaserScan fields angleMin, angleIncrement, rangeMax, ranges (sensor_msgs.clj:8-12). */
int32_t amclj_sensor_update(amclj_ctx *ctx, const float *ranges, int32_t n, float angle_min,
                            float angle_increment, float range_max);
/* Raw per-particle weight of the last sensor pass: sum p^3 or sum log p (S7). */
int32_t amclj_get_raw_weights(amclj_ctx *ctx, float *raw);
/* Parity probe for map-range / extract-line (sensor.clj:112-139): ray-casts particles
 * [first, first+count) for every used beam; outputs are [count][beams_used] (any may be
 * NULL).  steps = major-axis steps taken to the hit (or last tested step on a miss). */
int32_t amclj_raycast_debug(amclj_ctx *ctx, int32_t n, float angle_min, float angle_increment,
                            float range_max, int64_t first, int64_t count, int32_t *hit_x,
                            int32_t *hit_y, int32_t *steps, uint8_t *hit, float *range_m);
/* Number of beams selected by S6 for an n-beam scan. */
int32_t amclj_beams_used(const amclj_ctx *ctx, int32_t n);
/* Entries of one per-start-cell ray table for the staged scan geometry (the end cells a ray of that
 * length can reach, sensor.clj:130-135), or 0 when the ray-table path is unavailable. */
int32_t amclj_ray_table_entries(const amclj_ctx *ctx);
/* The ring itself, host only (no ctx, no device): for a ray of r_cells cells, row i (end-cell offset
 * ey = i - *half_rows) admits |ex| in [lo[i], lo[i] + cnt[i]).  Returns the entries per table (0: the
 * geometry is out of range); fills at most max_rows rows.  The CPU tests check it against the
 * rounding of sensor.clj:134-135 by brute force. */
int32_t amclj_ray_ring(double r_cells, int32_t max_rows, int32_t *lo, int32_t *cnt, int32_t *half_rows);

/* ---- normalise + resample: roulette-selection pf.clj:156-167 ----------------------- */
/* Linear weights (w = raw, or exp(raw - max raw) under S7) -> fixed point
 * q_i = floor(w_i / w_max * 2^32), inclusive integer CDF.  total_out = sum q_i. */
int32_t amclj_normalize(amclj_ctx *ctx, int32_t from_raw, uint64_t *total_out,
                        float *wmax_out);
/* mode AMCLJ_RESAMPLE_*.  u0 in [0,1): the single systematic draw.  stream_idx/stream_u:
 * injected roulette/tournament attempt stream of stream_len entries (idx in [0,n),
 * u = numerator of u/2^32) or NULL for the Philox stream keyed by `seed`.
 * idx_out (host int32 [n]) receives the selected source indices, or NULL. */
int32_t amclj_resample(amclj_ctx *ctx, int32_t mode, double u0, const int32_t *stream_idx,
                       const uint32_t *stream_u, int64_t stream_len, uint64_t seed,
                       int32_t *idx_out);

/* Source indices chosen by the last resample (host int32 [n]). */
int32_t amclj_get_resample_indices(amclj_ctx *ctx, int32_t *idx_out);

/* ---- fused update: one mcl* iteration's hot path (pf.clj:229-233) ------------------- */
/* motion -> sensor -> normalise -> resample, device-resident in and out.  The scan is
 * copied to the device inside the call. */
int32_t amclj_filter_update(amclj_ctx *ctx, const double curr[3], const double prev[3],
                            const float *normals, uint64_t seed, const float *ranges,
                            int32_t n, float angle_min, float angle_increment,
                            float range_max, double u0);
/* Enqueue-only variant: scan must have been staged with amclj_stage_scan. */
int32_t amclj_stage_scan(amclj_ctx *ctx, const float *ranges, int32_t n, float angle_min,
                         float angle_increment, float range_max);
int32_t amclj_filter_update_async(amclj_ctx *ctx, const double curr[3], const double prev[3],
                                  uint64_t seed, double u0);
/* Page-lock a caller-owned host buffer (a JNA Memory, a direct ByteBuffer, a malloc'ed array) so that the
 * copies of the *_host calls run at the bus's speed instead of through the driver's staging buffers (measured
 * on B200, 1.25 M particles: 3.99 ms per update with plain pageable buffers, the pinned figure once they are
 * registered).  Once per buffer, for as long as the host reuses it; unregister before freeing it.  No reference
 * counterpart (amclj keeps its particles in JVM vectors); the buffers are those of INTEGRATION.md's shim. */
int32_t amclj_host_register(amclj_ctx *ctx, void *ptr, int64_t bytes);
int32_t amclj_host_unregister(amclj_ctx *ctx, void *ptr);
/* amclj_stage_scan + amclj_filter_update_async in one call (one mcl* iteration, pf.clj:229-233, for a device-resident
 * cloud).  When the scan has the geometry of the staged one — every message of a laser after its first — the motion
 * kernel, which needs nothing of the scan, is enqueued first and the message's beam table is built and uploaded
 * while it runs; otherwise exactly the two calls.  Results are identical either way. */
int32_t amclj_filter_update_scan_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                       double u0, const float *ranges, int32_t n, float angle_min,
                                       float angle_increment, float range_max);
/* The same step for a host-resident PoseArray: uploads x,y,theta, runs the fused update,
 * downloads the resampled particles and weights (in place). */
int32_t amclj_filter_update_host(amclj_ctx *ctx, float *x, float *y, float *theta, float *w,
                                 int64_t n, const double curr[3], const double prev[3],
                                 const float *normals, uint64_t seed, const float *ranges,
                                 int32_t n_ranges, float angle_min, float angle_increment,
                                 float range_max, double u0);

/* ---- augmented MCL pf.clj:243-279 (amcl*; not wired into monte-carlo-localization) ------ */
/* incanter median of the current weights (raw weights of the last sensor pass, else w). */
int32_t amclj_median_weight(amclj_ctx *ctx, double *out);
/* augmented-selection: w_avg = median weight, w_slow/w_fast updated with alpha_slow/alpha_fast
 * (pf.clj:16-17: 0.01 / 0.1), then every slot injects a uniform particle when
 * random > max(0, 1 - fast/slow) (pf.clj:254, the reference's inequality) or runs roulette-step
 * (pf.clj:169-180).  Per-slot Philox streams keyed by `seed`.  idx_out: source index, -1 for an
 * injected particle (host int32 [n], may be NULL). */
int32_t amclj_augmented_resample(amclj_ctx *ctx, double alpha_slow, double alpha_fast, uint64_t seed,
                                 int32_t *idx_out);
/* The w-slow / w-fast atoms (pf.clj:243-244, both start at 1): either pointer may be NULL. */
int32_t amclj_augment_state(amclj_ctx *ctx, const double *set_slow_fast, double *get_slow_fast);

/* ---- pose-estimate pf.clj:67-84: unweighted mean of x, y, and of (qz, qw) ----------- */
/* out = {mean x, mean y, mean qz, mean qw}.  On a shard (world > 1) out holds the local
 * SUMS; the caller adds the ranks' sums and divides by n_global. */
int32_t amclj_pose_estimate(amclj_ctx *ctx, double out[4]);

int32_t amclj_get_stats(amclj_ctx *ctx, amclj_stats *out);
/* Measured ceiling of the ray caster's inner operation on this GPU: dependent distance-field
 * lookups per second from shared memory (which = 0, 128 KiB nibble table), from an L2-resident
 * 64 MiB byte table (which = 1) or from an L1-resident 64 KiB byte table (which = 2), all SMs,
 * 32 warps each. */
int32_t amclj_microbench(amclj_ctx *ctx, int32_t which, double *lookups_per_s);

/* ---- sharded (multi-GPU) building blocks, SURVEY 8e --------------------------------- */
/* Stage 1: motion + sensor, enqueue only; the local max raw weight lands in a device
 * float (amclj_device_ptr(AMCLJ_BUF_RAW_MAX)) for the all-reduce(max). */
int32_t amclj_shard_motion_sensor_async(amclj_ctx *ctx, const double curr[3],
                                        const double prev[3], uint64_t seed);
/* Stage 2: fixed-point CDF of the local weights against the GLOBAL max (read from
 * AMCLJ_BUF_RAW_MAX after the all-reduce); local total lands in AMCLJ_BUF_TOTAL (u64). */
int32_t amclj_shard_cdf_async(amclj_ctx *ctx);
/* Stage 3: systematic offspring for global slots [m_lo, m_lo+count) whose thresholds fall
 * in this rank's CDF interval [cdf_offset, cdf_offset + local total).  Output: float4
 * (x, y, theta, w) rows in AMCLJ_BUF_STAGE_OUT, global source index in idx (optional
 * host int32/int64 readback through amclj_shard_get_stage_idx). */
int32_t amclj_shard_generate_async(amclj_ctx *ctx, uint64_t total_global, uint64_t cdf_offset,
                                   uint64_t r, int64_t m_lo, int64_t count);
/* Global source index of each generated row (host int64 [count]); parity probe. */
int32_t amclj_shard_get_stage_idx(amclj_ctx *ctx, int64_t *idx_out, int64_t count);
/* Stage 4: adopt `count` float4 rows from AMCLJ_BUF_STAGE_IN as the new local particles. */
int32_t amclj_shard_adopt_async(amclj_ctx *ctx, int64_t count);
/* Peer-memory variant of stages 3-4 (no staging rows, no all-to-all, no host round trip): every
 * rank PULLS its own output slots from whichever rank's CDF interval holds the slot's threshold,
 * reading that rank's CDF and particles through NVLink peer pointers.
 *   export: CUDA IPC handles (AMCLJ_IPC_BUFFERS x AMCLJ_IPC_HANDLE_BYTES) of this ctx's particle
 *           and CDF buffers; freezes their size.  import: map rank `rank`'s buffers (n_peer
 *           particles starting at global id peer_global_offset).  attach_local: the same for a ctx
 *           of this process (also used for the ctx's own rank).
 *   pull:   needs the all-gathered totals in AMCLJ_BUF_TOTALS_ALL (uint64[world]) and the CDF of
 *           stage 2; fills this rank's slots [global_offset, global_offset + n) and swaps them in.
 * Every rank must call pull the same number of times (the buffer sets alternate in lock step). */
#define AMCLJ_IPC_HANDLE_BYTES 64
#define AMCLJ_IPC_BUFFERS 8
int32_t amclj_shard_export(amclj_ctx *ctx, uint8_t *handles);
int32_t amclj_shard_import(amclj_ctx *ctx, int32_t rank, const uint8_t *handles, int64_t n_peer,
                           int64_t peer_global_offset);
int32_t amclj_shard_attach_local(amclj_ctx *ctx, int32_t rank, amclj_ctx *peer);
int32_t amclj_shard_pull_async(amclj_ctx *ctx, double u0);
/* The whole sharded update behind ONE call per rank, without a collective library and without a host
 * round trip: motion + sensor, then the ranks exchange {largest raw weight, fixed-point total} through
 * peer-memory mailboxes written and polled by small kernels (12 bytes per rank and update), then pull.
 * Needs export / import (or attach_local) of every rank, the scan staged, and the same call made the same
 * number of times on every rank (the mailbox entries carry the update's generation).  Replaces
 * shard_motion_sensor_async + all-reduce + shard_cdf_async + all-gather + shard_pull_async. */
int32_t amclj_shard_update_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                 double u0);
/* amclj_stage_scan + amclj_shard_update_async in one call, as amclj_filter_update_scan_async. */
int32_t amclj_shard_update_scan_async(amclj_ctx *ctx, const double curr[3], const double prev[3], uint64_t seed,
                                      double u0, const float *ranges, int32_t n, float angle_min,
                                      float angle_increment, float range_max);
/* A barrier among the connected shards on the DEVICE time line (no reference counterpart: amclj runs one JVM):
 * enqueues a one-warp kernel that stores this call's count into every rank's mailbox and returns when every rank's
 * count has arrived, so whatever follows it in the stream starts at the same instant on every GPU, however far
 * apart the hosts enqueued it.  Same number of calls on every rank; every rank on its own GPU (AMCLJ_ERR_STATE
 * when a peer shares this device: kernels that wait on one another must not share a GPU). */
int32_t amclj_shard_rendezvous_async(amclj_ctx *ctx);
/* The same for a host-resident shard: upload x, y, theta of this rank's n particles, stage the scan, update,
 * download the rank's resampled slots (in place).  n must equal the connected shard size. */
int32_t amclj_shard_update_host(amclj_ctx *ctx, float *x, float *y, float *theta, float *w, int64_t n,
                                const double curr[3], const double prev[3], uint64_t seed, const float *ranges,
                                int32_t n_ranges, float angle_min, float angle_increment, float range_max,
                                double u0);
/* Host-only planning (no GPU needed): for per-rank totals T_g, systematic offset
 * r = floor(u0 * T), and slots split evenly (n_global / world per rank, remainder to the
 * low ranks) computes m_lo[g], m_cnt[g] and the world x world send matrix (row = source). */
int32_t amclj_plan_systematic(const uint64_t *totals, int32_t world, int64_t n_global,
                              double u0, uint64_t *r_out, uint64_t *cdf_offsets,
                              int64_t *m_lo, int64_t *m_cnt, int64_t *send_matrix);

#define AMCLJ_BUF_X 0
#define AMCLJ_BUF_Y 1
#define AMCLJ_BUF_THETA 2
#define AMCLJ_BUF_W 3
#define AMCLJ_BUF_RAW 4
#define AMCLJ_BUF_RAW_MAX 5   /* float[1] */
#define AMCLJ_BUF_TOTAL 6     /* uint64[1] */
#define AMCLJ_BUF_STAGE_OUT 7 /* float4[capacity] */
#define AMCLJ_BUF_STAGE_IN 8  /* float4[capacity] */
#define AMCLJ_BUF_CDF 9       /* uint64[n] */
#define AMCLJ_BUF_TOTALS_ALL 10 /* uint64[16]: all-gather target for the per-rank totals */
/* Raw device pointer + size in bytes of a library-owned buffer, for zero-copy interop
 * with the host framework's collectives (the pointer stays valid until the particle count
 * changes or the ctx is destroyed). */
int32_t amclj_device_ptr(amclj_ctx *ctx, int32_t which, void **ptr, int64_t *bytes);
/* Make sure the staging buffers hold at least `rows` float4 rows. */
int32_t amclj_reserve_stage(amclj_ctx *ctx, int64_t rows);

/* ---- one process, several GPUs (`amclj_create(const int* devices, int n_devices, ...)`) ------
 * One caller thread, as in the reference (pf.clj:297): an amclj_mctx owns one amclj_ctx per listed device
 * (a device may be listed more than once: emulated shards), splits the particles into contiguous blocks
 * (n_global / n each, the remainder to the low ranks), replicates map and scan, and runs
 * amclj_shard_update_async on every device.  Results are identical to one amclj_ctx holding all particles.
 * Per-rank calls that have no m-variant (amclj_get_stats, amclj_set_stream, ...) go through amclj_multi_ctx. */
typedef struct amclj_mctx amclj_mctx;
int32_t amclj_create_multi(const int32_t *devices, int32_t n_devices, amclj_mctx **out);
int32_t amclj_destroy_multi(amclj_mctx *m);
int32_t amclj_multi_size(const amclj_mctx *m);
amclj_ctx *amclj_multi_ctx(amclj_mctx *m, int32_t rank);
const char *amclj_mlast_error(const amclj_mctx *m);
int32_t amclj_mset_params(amclj_mctx *m, const amclj_params *p);
int32_t amclj_mset_map(amclj_mctx *m, const int8_t *cells, int32_t width, int32_t height, float resolution);
/* x, y, theta, w: host arrays of ALL n_global particles (w may be NULL). */
int32_t amclj_mset_particles(amclj_mctx *m, const float *x, const float *y, const float *theta, const float *w,
                             int64_t n_global);
int32_t amclj_mget_particles(amclj_mctx *m, float *x, float *y, float *theta, float *w);
int64_t amclj_mnum_particles(const amclj_mctx *m);
int32_t amclj_minit_uniform(amclj_mctx *m, int64_t n_global, int32_t heading_mode, uint64_t seed);
int32_t amclj_mstage_scan(amclj_mctx *m, const float *ranges, int32_t n, float angle_min, float angle_increment,
                          float range_max);
/* One mcl* iteration's hot path (pf.clj:229-233) over all devices, enqueue only / with host buffers. */
int32_t amclj_mfilter_update_async(amclj_mctx *m, const double curr[3], const double prev[3], uint64_t seed,
                                   double u0);
int32_t amclj_mfilter_update_host(amclj_mctx *m, float *x, float *y, float *theta, float *w, int64_t n_global,
                                  const double curr[3], const double prev[3], uint64_t seed, const float *ranges,
                                  int32_t n_ranges, float angle_min, float angle_increment, float range_max,
                                  double u0);
int32_t amclj_msynchronize(amclj_mctx *m);
/* pose-estimate pf.clj:67-84 over all shards: {mean x, mean y, mean qz, mean qw}. */
int32_t amclj_mpose_estimate(amclj_mctx *m, double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* AMCLJ_B200_H */
