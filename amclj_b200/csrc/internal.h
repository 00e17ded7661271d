// internal.h — context layout and kernel-launcher prototypes of libamclj_b200 (not part of
// the C ABI; include/amclj_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/amclj_b200.h"
#include "contract.h"

namespace amclj {

// ---------------------------------------------------------------------------------------
// occupancy grid on the device: a Chebyshev distance field ("how many Bresenham steps are
// certainly free from here"), padded by a one-cell ring of zeros (out of range = blocked,
// map.clj:19-24), nibble-packed when it has to live in shared memory.
// ---------------------------------------------------------------------------------------
struct DeviceMap {
  int32_t W = 0, H = 0;
  float res = 0.f;
  int8_t *cells = nullptr;   // [H][W] as uploaded (init kernels, debug)
  uint8_t *df4 = nullptr;    // nibble-packed, (H+2) rows of row_bytes4, distance capped at 15
  uint8_t *df8 = nullptr;    // one byte per cell, (H+2) x pitch8, capped at 255
  int32_t row_bytes4 = 0;    // multiple of 16
  int32_t pitch8 = 0;
  size_t bytes4 = 0, bytes8 = 0;
  int64_t free_cells = 0;    // cells with value 0 (counted at set_map)
  uint8_t *wedge = nullptr;  // WEDGE_CLASSES planes of (H+2) x pitch8 bytes: directional distance fields
  size_t wedge_plane = 0;
  bool fits_smem = false;
};

struct BeamTable {       // float4 mix[nb] = (obs, e2, p34, 0) then float2 dir[nb] = (cb, sb)
  float *dev = nullptr;
  double *dev_d = nullptr;   // double2 (cos, sin) per beam (near-tie refinement)
  amclj::CellRefine cr = {0.0f, 0, 0, 0u, 0.0f};
  int32_t nb = 0, cap = 0;
  int32_t n_scan = 0;
  float angle_min = 0, angle_inc = 0, range_max = 0;
  bool staged = false;
  /* what the staged directions were computed from (stage_scan recomputes them only when this changes) */
  bool dirs_valid = false;
  const void *dirs_pin = nullptr;
  int32_t dirs_n = 0, dirs_step = 0, dirs_nb = 0;
  float dirs_amin = 0, dirs_ainc = 0;
};

struct DivTable {        // magic reciprocals for the Bresenham minor-axis division
  uint2 *dev = nullptr;
  int32_t n = 0, cap = 0;
  int32_t wide = 0, dcap = 0;
};

struct SensorLaunch {
  const float *x, *y, *th;
  float *raw;
  int64_t n;
  const uint8_t *df;
  int32_t row_nib;   // nibbles (df4) or bytes (df8) per padded row
  int32_t wedge_plane;  // bytes per direction class (DF_WEDGE8)
  int32_t W, H;
  size_t df_bytes;
  float res;
  const float *beam;
  int32_t nb;
  const uint2 *divtab;   // (M, shift) per dx
  int32_t tab_smem;      // 1: beam and reciprocal tables are staged into shared memory
  int32_t ndiv;
  int32_t div_wide;  // 1: j-step = umulhi(t, M) >> sh; 0: shift-free (short rays)
  int32_t refill;    // set up new beams when fewer lanes than this (of 32) are still walking
  int32_t chunks;    // beam chunks per particle group (warp tasks = groups x chunks)
  long long *partials;  // [chunks][n] fixed-point partial weights when chunks > 1
  const uint32_t *bbox;  // particle bounding box (auto policy), see launch_bbox
  float compact_m;       // bounding-box side (metres) up to which the particle set counts as compact
  const int32_t *perm;   // walking order (particle indices sorted by map tile), or null
  int32_t *decision;     // out: 1 when the cloud was compact (host hint for the next update)
  uint32_t *walk_ctr;        // {next task of the common tail, CTAs finished}: zero between launches
  int32_t walk_tail_pct;     // share of the tasks drawn dynamically at the end (< 0: fixed interleaved shares)
  const int32_t *skip_flag;  // when set and *skip_flag == 1 the ray-table path took this update: return at once
  float R;            // ray length (S2)
  float miss_value;
  int32_t s1, s3, s4, s7, s10;
  float z_hit, inv2s2, laser_x, laser_y;
  float k2;  // inv2s2 * log2(e): the hit Gaussian is one ex2
  CellRefine cr;          // near-tie refinement of end cells (contract.h); cr.K == 0: off
  const double *beam_d;   // double2 (cos, sin) per beam, for the refinement's double path
  double rinv_d;          // 1 / resolution in double (host)
  // ring of end cells around the start cell and one line set-up record per ring cell (walk kernel)
  const uint4 *ring_rows;
  const uint4 *ring_recs;
  int32_t ring_nrows, ring_E, ring_entries;
  int32_t ring_ok;   // records match this launch (mode, switches, map, geometry) and every end cell quotient stays below 2^22
  uint32_t *rawmax_ord;          // ordered-uint max of raw (may be null)
  unsigned long long *stats;     // rays, cells, probes, hits (may be null)
  // parity probe
  int64_t dbg_first, dbg_count;
  int32_t *dbg_hit_x, *dbg_hit_y, *dbg_steps;
  uint8_t *dbg_hit;
  float *dbg_range;
};

struct MotionLaunch {
  float *x, *y, *th;
  int64_t n;
  int64_t gid0;
  const float *normals;  // [n][3] or null -> Philox
  uint64_t seed;
  int32_t is_static, trans_zero;
  int32_t vec_ok;  // set by launch_motion: every array is 16-byte aligned (128-bit loads and stores)
  uint32_t *bbox;  // when set: accumulate the bounding box of the moved cloud (4 ordered-uint maxima)
  float rot1, trans, rot2, sd1, sdt, sd2;
  /* optional: the first pass of the persistent-table sort fused in (ptable.cu's pt_hist_kernel): table slot of
   * the moved particle's start cell, histogram, cleared accumulator.  hist_slot_of_cell == null: off */
  const int32_t *hist_slot_of_cell;
  int32_t hist_W, hist_H, hist_n_slots, hist_s10;
  float hist_res, hist_laser_x, hist_laser_y;
  uint32_t *hist_bin_count;
  int32_t *hist_pkey;
  long long *hist_qacc;
  unsigned long long *hist_scan_state;
  int32_t hist_scan_words;
};

// ---------------------------------------------------------------------------------------
// ray tables (raytable.cu): for a dense cloud many particles share a start cell, and a ray's
// result is a function of its integer start and end cells only (sensor.clj:130-139).
// ---------------------------------------------------------------------------------------
constexpr int RT_PI = 32;           // particles per work item: one warp, lane = particle
constexpr int RT_BIN_CAP = 16384;   // start cells the cloud's bounding box may cover (128 x 128)
struct RtSlot { int32_t cx, cy, first, cnt; };  // a dense start cell: its particles are perm[first, first + cnt)
struct RtCtl {
  int32_t dense;    // 1: this update runs on the ray tables
  int32_t n_dense;  // particles in dense cells (they come first in the walking order)
  int32_t n_items;  // work items over dense cells
  int32_t n_cells;  // dense cells = tables to fill
  int32_t cx0, cy0, bw, bh;
  uint32_t hist_done;  // CTAs of the histogram that have finished (the last one plans)
  uint32_t next_tail;  // next unit of the look-up kernel's common tail
  uint32_t pad[6];
};
struct RayTables {
  // end-cell ring of a ray of r cells, built on the host per scan geometry: row ey + E ->
  // {base, lo, cnt, base + cnt}: |ex| in [lo, lo + cnt) is entry base + (|ex| - lo) (+ cnt when ex < 0)
  uint4 *rows = nullptr;
  short2 *ents = nullptr;   // entry -> (ex, ey); x = -32768 marks the unused slot of a merged row
  int32_t nrows = 0, E = 0, n_entries = 0, rows_cap = 0, ents_cap = 0;
  double r_cells = -1.0;
  bool have_ring = false;     // rows / ents describe the staged scan geometry
  bool available = false;     // ... and the ray-table path may use them
  uint4 *recs = nullptr;      // [2 * n_entries] line set-up records of the walk kernel (sensor.cu)
  int32_t recs_cap = 0, recs_mode = -1;
  long long recs_key[6] = {0, 0, 0, 0, 0, 0};
  double recs_r = -1.0;
  bool floor_trick_ok = false; // every end-cell quotient of a start inside the map stays below 2^22   // the field placement the records were built for (-1: none)
  uint32_t *bins = nullptr;   // [RT_BIN_CAP] particles per start cell, then write cursors
  RtSlot *slots = nullptr;    // [RT_BIN_CAP] the dense cells, one table each
  int32_t *item0 = nullptr;   // [RT_BIN_CAP + 1] first work item of each dense cell
  void *tables = nullptr;     // [max_slots][n_entries] float range (diagnostic passes: float4)
  size_t tables_bytes = 0;
  RtCtl *ctl = nullptr;
  float4 *pstate = nullptr;   // [pstate_cap] per-particle look-up state in walking order
  int64_t pstate_cap = 0;
  int32_t thr = 0;            // particles per cell from which a table pays (0: derived)
  size_t budget = 256ull << 20;
  int32_t tail_pct = 30;
  int32_t waves = 1;          // look-up grid = resident CTAs x waves (smaller CTAs balance better)
};
struct RtLaunch {
  SensorLaunch s;
  const uint4 *rows;
  const short2 *ents;
  int32_t nrows, E, n_entries;
  uint32_t *bins;
  RtSlot *slots;
  int32_t *item0;
  void *tables;
  RtCtl *ctl;
  int32_t *perm;
  float4 *pstate;    // [n] (origin x, origin y, sin, cos) of the particles in walking order
  long long *qacc;   // [n] 32.32 fixed-point weight accumulators
  int32_t thr, max_slots, max_items;
  int32_t tail_pct;  // share of the dense units handed out dynamically at the end of the look-up kernel
  int32_t must_run;  // 1: the walk kernel is not launched for this update, so this path takes every particle
};

// ---------------------------------------------------------------------------------------
// persistent map-wide ray tables (ptable.cu).  The map is static and a ray's range is a pure
// function of (start cell, end cell) (sensor.clj:130-139 rounds both ends before anything else
// looks at them), so ONE table per free cell of the map, filled once per (map, ray length,
// S3/S4, resolution), serves every later update of every cloud, spread out or dense: per update
// the particles are counting-sorted by start cell and every (particle, 30-beam segment) looks its
// ranges up in the cell's table, which the warp stages into shared memory with the bulk-copy
// engine.  A particle standing on a non-free or out-of-range cell sees range 0 on every beam
// (extract-line tests the start cell first): those share one all-zero table.
// ---------------------------------------------------------------------------------------
struct PersistentTables {
  int32_t n_slots = 0;             // free cells of the map = tables
  int32_t *slot_of_cell = nullptr; // [H][W] -> table slot, or n_slots (the zero table)
  int2 *cell_of_slot = nullptr;    // [n_slots]
  float *tables = nullptr;         // [n_slots + 1][stride]; the last one is all zeros
  size_t tables_bytes = 0;
  int32_t stride = 0;              // floats per table: n_entries rounded up to 4 (16-byte bulk copies)
  uint32_t *bin_count = nullptr;   // [n_slots + 2] particles per slot (zeroed by the scan)
  uint32_t *bin_start = nullptr;   // [n_slots + 2] exclusive prefix; [n_slots + 1] = n
  uint32_t *cursor = nullptr;      // [n_slots + 2] write cursors of the scatter
  int32_t *pkey = nullptr;         // [pkey_cap] slot of every particle
  double2 *pstate_d = nullptr;     // [2 * pkey_cap] per particle {heading (sin, cos), ray origin (x, y)} in double, sorted order
  int64_t pkey_cap = 0;
  bool enable = true;              // AMCLJ_PTABLE=0 turns the path off
  bool built = false;
  size_t budget = 4ull << 30;      // AMCLJ_PT_BUDGET_MB
  int32_t warps = 0;               // look-up warps per CTA (0: as many as the shared memory takes, at most 16)
  bool fuse_finish = false;        // AMCLJ_PT_FUSE=1: the look-up kernel finishes the particles itself (measured slower:
                                   // an atomic that returns its old value stalls the warp, a plain reduction does not)
  int32_t nbuf = 0;                // table buffers per warp requested (0: one, so that more warps fit; AMCLJ_PT_NBUF)
  int32_t nbuf_used = 1, warps_req = 0;
  int32_t tail_div = 4, tail_min = 8; // AMCLJ_PT_TAIL_DIV, AMCLJ_PT_TAIL_MIN
  int32_t tail_pct = 15;           // AMCLJ_PT_TAIL_PCT (measured: 0 -> 1.046, 15 -> 0.994, 30 -> 1.014 ms on the C4 share)
  bool fuse_motion = true;         // the motion kernel takes the sort's histogram pass along (AMCLJ_PT_FUSE_MOTION=0: pt_hist_kernel)
  bool hist_done = false;          // ... and did so for the update in progress
  int32_t rows32 = 1;              // AMCLJ_PT_ROWS32
  unsigned long long *scan_state = nullptr;  // bin scan ticket + tile descriptors
  int32_t scan_words = 0;
  // what the tables were built for
  int64_t key_map_gen = -1;
  double key_r = -1.0;
  int32_t key_s3 = -1, key_s4 = -1, key_entries = -1;
  uint32_t key_R_bits = 0, key_res_bits = 0;
  int32_t fill_mode = -1;
  double ms_fill = 0.0;            // what the one-time fill took
};
struct PtLaunch {
  SensorLaunch s;
  const uint4 *rows;
  const short2 *ents;
  int32_t nrows, E, n_entries, stride;
  int32_t n_slots;
  const int32_t *slot_of_cell;
  const int2 *cell_of_slot;
  float *tables;
  uint32_t *bin_count, *bin_start, *cursor;
  int32_t *pkey;
  int32_t *perm;
  float4 *pstate;
  double2 *pstate_d;   // [2n] {(sin, cos) of the heading, ray origin} in double, sorted order: read only on the near-tie path
  double rinv_d;       // 1 / resolution in double
  long long *qacc;
  int32_t warps;   // look-up warps per CTA
  int32_t fuse_finish;  // 1: the look-up kernel finishes a particle when its last segment lands (<= 15 segments)
  int32_t nbuf;         // table buffers per look-up warp: 2 (next table in flight) or 1
  unsigned long long *scan_state;  // bin scan: {ticket | tail counter, descriptor per tile}
  int32_t scan_words;
  int32_t tail_pct;     // share of the sorted particles handed out in chunks at the end of the look-up
  int32_t tail_div, tail_min;  // a unit of that tail: what is left / (tail_div x warps), at least tail_min particles
  uint32_t *tail_ctr;   // next chunk of that tail (zeroed by the histogram)
  int32_t rows32;       // 1: ring rows packed into one word in shared memory (base < 8192, lo < 256, cnt < 2048)
};

constexpr int MAX_WORLD = 16;
struct PeerView {   /* one rank's shard as seen from this GPU (own memory or NVLink peer memory) */
  const float *x, *y, *th;
  const uint64_t *cdf;
  int64_t n, gid0;
};
}  // namespace amclj

struct amclj_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [5], [6]: around the dominant sensor kernel
  bool time_main = false, time_main_env = false;  // collect_stats bit 1 (or AMCLJ_TIME_MAIN=1): events around that kernel
  int32_t launches = 0;   // kernels enqueued by the update in progress
  bool stages_timed = true;  // the last update recorded its inner stage events (ev[1], ev[2])
  std::string err;
  amclj_params p;
  amclj::DeviceMap map;
  amclj::BeamTable beams;
  amclj::DivTable divs;
  // particles (SoA), double-buffered for the resampling gather
  int64_t n = 0, cap = 0;
  float *x = nullptr, *y = nullptr, *th = nullptr, *w = nullptr, *raw = nullptr;
  float *x2 = nullptr, *y2 = nullptr, *th2 = nullptr, *w2 = nullptr;
  uint64_t *cdf = nullptr;
  int32_t *idx = nullptr;       // last resample's source indices
  float *normals = nullptr;     // staged injected normals [n][3]
  long long *partials = nullptr;  // sensor chunk partials (32.32 fixed point)
  int32_t *perm = nullptr;        // walking order of the particles (sorted by map tile)
  uint32_t *tile_counters = nullptr;
  int64_t partials_cap = 0;
  // scalars on the device: one control block [ord | total | counter | tile_state...] so that a
  // single memset per update clears everything the sensor max and the scan accumulate into
  uint64_t *ctrl = nullptr;
  bool scan_clean = false;
  bool ctrl_fresh = false;   // control block cleared for this update already
  bool bbox_done = false;    // the motion kernel of this update left the bounding box
  int32_t *decision = nullptr;      // device flag: the last cloud was compact
  volatile int32_t *pin_decision = nullptr;  // its (asynchronously refreshed) host copy: a hint only
  bool last_diag = false;   // the last sensor pass ran the diagnostic kernel (stats are fresh)
  bool max_in_ord = false;  // the weight max of the last pass lives in rawmax_ord only
  uint32_t *rawmax_ord = nullptr;  // ordered-uint of max raw
  float *rawmax = nullptr;         // float[1] (AMCLJ_BUF_RAW_MAX)
  uint64_t *total = nullptr;       // u64[1]  (AMCLJ_BUF_TOTAL)
  unsigned long long *stats = nullptr;  // rays, cells, probes, hits | roulette: accepted, last attempt
  uint64_t *scan_state = nullptr;  // decoupled look-back tile descriptors
  int64_t scan_state_cap = 0;
  uint32_t *scan_counter = nullptr;
  double *pose_part = nullptr;     // pose-estimate partials
  amclj::SysPlan *plan = nullptr;  // device copy of the systematic plan
  // roulette scratch
  uint32_t *r_counts = nullptr;
  int64_t r_counts_cap = 0;
  // sharding
  int32_t rank = 0, world = 1;
  int64_t global_offset = 0, n_global = 0;
  // peer-memory resampling: both particle buffer sets of every rank (A = current at export)
  amclj::PeerView peers[2][amclj::MAX_WORLD];
  bool peer_set[amclj::MAX_WORLD] = {};
  void *ipc_opened[amclj::MAX_WORLD][8] = {};
  unsigned long long *mbox = nullptr;                        /* this rank's mailbox (device memory, zeroed) */
  unsigned long long *peer_mbox[amclj::MAX_WORLD] = {};      /* every rank's mailbox as seen from here */
  uint32_t gen = 0;                                          /* sharded updates so far (the mailbox tag) */
  uint32_t sync_gen = 0;                                     /* rendezvous calls so far (amclj_shard_rendezvous_async) */
  bool peer_shares_device = false;                           /* a peer runs on this GPU (emulated shards) */
  cudaEvent_t ev_pub[2] = {nullptr, nullptr};                /* recorded after this rank published its max / its total */
  int flip = 0;            // swaps since export, mod 2
  bool frozen = false;     // buffers exported: no reallocation
  uint64_t *totals_all = nullptr;  // [MAX_WORLD] all-gathered totals
  float4 *stage_out = nullptr, *stage_in = nullptr;
  int64_t stage_cap = 0;
  int64_t *stage_idx = nullptr;
  uint64_t zero_seen = 0;  /* zero-total updates already reported (device counter: stats[12]) */
  uint64_t mbox_errors_seen = 0;  /* mailbox time-outs already reported (stats[13]) */
  bool have_raw = false, have_cdf = false, cdf_from_raw = false, normals_staged = false;
  float last_wmax = 0.f;
  uint64_t last_total = 0;
  amclj_stats st;
  int32_t refill = 6, refill_wedge = 12, refill_global = 16;
  bool order_large_maps = true;       // auto policy: tile-order walk on the L2 byte field as well
  size_t wedge_budget = 64ull << 20;  // auto policy: build the directional planes when they fit this
  float compact_cells = 320.f;  // auto policy: directional fields when the particle bounding box is this small
  uint32_t *bbox = nullptr;    // 4 ordered-uint maxima: x, -x, y, -y (inside ctrl)
  double w_slow = 1.0, w_fast = 1.0;  // pf.clj:243-244
  uint32_t *hist = nullptr;           // 256-bin histogram of the radix select
  int32_t force_chunks = 0;
  amclj::RayTables rt;
  amclj::PersistentTables pt;
  uint32_t *walk_ctr = nullptr;  // device {tail counter, CTAs finished} of the walk kernel
  int32_t tile_shift = 0;      // walking-order tile side (2^shift cells); 0: from the particle density (AMCLJ_TILE_SHIFT)
  int32_t walk_carveout = 20;  // shared-memory carve-out (percent) asked for the walk kernel on the global fields, the rest is L1
                               // (C5: 75.9 -> 72.1 ms against the driver's own choice); -1: driver's choice
  int32_t walk_tail_pct = 0;   // spread-out clouds: CTA-contiguous task runs, this share of the tasks as a common tail
  int32_t walk_tail_pct_global = 20; // the same on the L2-resident byte field of a large map, whose regions differ in wall
                               // density (C5: SM active cycles 149 M .. 184 M without a tail; 82.5 -> 75.5 ms)
  int64_t map_gen = 0;     // bumped by every set_map
  bool ring_setup = true;  // walk kernel reads its line set-up from per-ring-cell records (AMCLJ_RING_SETUP=0: computes it)
  bool rt_enable = true;   // auto policy may take the ray-table path (AMCLJ_RAYTABLE=0 turns it off)
  bool exact_cells = true; // near-tie refinement of end cells (AMCLJ_EXACT_CELLS=0: plain float rounding)
  void *pinned = nullptr;  // small pinned scratch for scalar readbacks
  void *pin_tab = nullptr; // pinned staging of the beam table
  size_t pin_tab_bytes = 0;
};

namespace amclj {
// sensor.cu
cudaError_t launch_sensor(const amclj_ctx *c, const SensorLaunch &a, int mode, bool diag,
                          cudaStream_t s);
enum { DF_SMEM4 = 0, DF_GLOBAL8 = 1, DF_GLOBAL4 = 2, DF_WEDGE8 = 3 };
// ptable.cu
cudaError_t launch_pt_fill(const amclj_ctx *c, const PtLaunch &L, int mode, cudaStream_t s);
cudaError_t launch_pt_update(const amclj_ctx *c, const PtLaunch &L, int mode, bool dbg, bool hist_done, cudaStream_t s);
size_t pt_lookup_smem(const PtLaunch &L, int warps);
int pt_max_warps();
// raytable.cu
cudaError_t launch_raytable(const amclj_ctx *c, const RtLaunch &L, int mode, bool diag, cudaStream_t s);
cudaError_t rt_prepare_device();
constexpr int WEDGE_BINS = 4;                   /* slope bins per octant */
constexpr int WEDGE_CLASSES = 8 * WEDGE_BINS;   /* steep x xs x ys x bin */
cudaError_t build_wedge_fields(amclj_ctx *c, cudaStream_t s);
cudaError_t launch_ring_records(const SensorLaunch &a, int mode, const short2 *ents, int n_entries, uint4 *recs,
                                cudaStream_t s);
int sensor_chunks(const amclj_ctx *c, int64_t n, int nb);

// filter.cu
struct GatherArgs {
  const uint64_t *cdf;
  int64_t n_local;
  uint64_t cdf_offset;
  const SysPlan *plan;  /* device pointer, or null: derive it from total_ptr / u0_bits / n_global */
  const uint64_t *total_ptr;
  uint64_t u0_bits;
  int64_t n_global;
  int64_t m_lo, count;
  const float *x, *y, *th;
  float *ox, *oy, *oth, *ow;  /* SoA outputs (single GPU) or null */
  float4 *orow;               /* row outputs (sharded) or null */
  int32_t *oidx;
  int64_t *oidx64;
  int64_t gid0;
  float wout;
  unsigned long long *zero_flag; /* counts updates whose fixed-point total was zero (may be null) */
};
struct PullArgs {
  PeerView peer[MAX_WORLD];
  int32_t world;
  const uint64_t *totals;  /* device [world]: every rank's fixed-point weight total */
  uint64_t u0_bits;
  int64_t n_global, slot0, count;
  float *ox, *oy, *oth, *ow;
  int64_t *oidx64;
  float wout;
  unsigned long long *zero_flag;
};
cudaError_t launch_pull(const PullArgs &a, cudaStream_t s);
/* mailbox exchange of a sharded update (filter.cu): word offsets inside a rank's mailbox */
constexpr int MBOX_MAX = 0, MBOX_TOTAL = MAX_WORLD, MBOX_TOTAL_GEN = 2 * MAX_WORLD, MBOX_SYNC = 3 * MAX_WORLD, MBOX_WORDS = 4 * MAX_WORLD;
struct MailboxArgs {
  unsigned long long *mbox;                 /* this rank's mailbox (its own memory) */
  unsigned long long *peer_mbox[MAX_WORLD]; /* every rank's mailbox as seen from this GPU */
  int32_t rank, world;
  uint32_t gen;
  const uint32_t *rawmax_ord;   /* local max (ordered uint) */
  float *rawmax;                /* out: the global max */
  const uint64_t *total;        /* local fixed-point total */
  unsigned long long *totals_all; /* out: every rank's total */
  unsigned long long *error_flag; /* counts waits that gave up (a rank that never published) */
};
/* polls before a wait gives up (a few seconds): a mis-wired exchange must not hang the GPU */
constexpr long long MBOX_SPIN_LIMIT = 1ll << 24;
/* what: 0 publish max, 1 wait for every max, 2 publish total, 3 wait for every total, 4 rendezvous (gen = its own count) */
cudaError_t launch_mailbox(const MailboxArgs &a, int what, cudaStream_t s);

struct RouletteArgs {
  const uint64_t *cdf;
  int64_t n;
  uint64_t total;
  uint64_t seed;
  const int32_t *sidx;  /* injected stream or null */
  const uint32_t *su;
  int64_t t0, t_end;    /* attempts [t0, t_end) */
  int per_thread;
  int64_t got_before;
  int32_t *idx_out;     /* [n] */
  uint64_t *tile_state;
  uint32_t *tile_counter;
  unsigned long long *accepted_out; /* accepted in this launch */
  unsigned long long *last_attempt; /* attempt index (+1) of the acceptance filling slot n-1 */
};
cudaError_t launch_motion(const MotionLaunch &a, cudaStream_t s);
cudaError_t launch_tile_order(const float *x, const float *y, const float *th, int64_t n, float res, int W,
                              int H,
                              uint32_t *counters, int32_t *perm, const uint32_t *bbox, float compact_m,
                              int shift /* tile side 2^shift cells, 3..8 (tile_order_shift) */,
                              cudaStream_t s);
int tile_count(int W, int H);
int tile_order_shift(int64_t n, int64_t free_cells);
cudaError_t launch_bbox(const float *x, const float *y, int64_t n, uint32_t *bbox, int sm_count,
                        cudaStream_t s);
cudaError_t build_distance_fields(amclj_ctx *c, bool want8, cudaStream_t s);
cudaError_t launch_max(const float *v, int64_t n, uint32_t *ord, float *out, int sm_count,
                       cudaStream_t s);
cudaError_t launch_decode_max(const uint32_t *ord, float *out, cudaStream_t s);
int64_t scan_tiles(int64_t n);
cudaError_t launch_scan(const float *v, const float *vmaxp, const uint32_t *vmax_ord, int is_log,
                        int64_t n, uint64_t *cdf, uint64_t *tile_state, uint32_t *tile_counter,
                        uint64_t *total_out, cudaStream_t s);
cudaError_t launch_plan(const uint64_t *total, uint64_t u0_bits, int64_t n_global, SysPlan *out,
                        cudaStream_t s);
cudaError_t launch_set_plan(const SysPlan &p, SysPlan *out, cudaStream_t s);
cudaError_t launch_systematic(const GatherArgs &a, cudaStream_t s);
cudaError_t launch_roulette(const RouletteArgs &a, int64_t tiles, cudaStream_t s);
cudaError_t launch_tournament(const float *w, int64_t n, const int32_t *sidx, uint64_t seed,
                              int32_t *idx_out, cudaStream_t s);
/* wmode 0: constant wconst, 1: q_i / total from the cdf, 2: copy wsrc[i] */
cudaError_t launch_gather_idx(const int32_t *idx, int64_t n, const float *x, const float *y,
                              const float *th, int wmode, const uint64_t *cdf, uint64_t total,
                              const float *wsrc, float wconst, float *ox, float *oy, float *oth,
                              float *ow, cudaStream_t s);
cudaError_t launch_adopt(const float4 *rows, int64_t n, float *x, float *y, float *th, float *w,
                         cudaStream_t s);
cudaError_t launch_fill(float *p, int64_t n, float v, cudaStream_t s);
cudaError_t launch_select_hist(const float *v, int64_t n, uint32_t prefix, uint32_t prefix_mask,
                               int shift, uint32_t *hist, int sm_count, cudaStream_t s);
float ord_to_float_host(uint32_t o);
cudaError_t launch_augmented(const float *wgt, const float *x, const float *y, const float *th,
                             int64_t n, float inject_threshold, const int8_t *cells, int W, int H,
                             float res, uint64_t seed, float *ox, float *oy, float *oth, float *ow,
                             int32_t *oidx, cudaStream_t s);
int pose_blocks();
cudaError_t launch_pose(const float *x, const float *y, const float *th, int64_t n, double *part,
                        cudaStream_t s);
cudaError_t launch_init_uniform(const int8_t *cells, int W, int H, float res, int64_t gid0,
                                int64_t n, int heading_mode, uint64_t seed, float *x, float *y,
                                float *th, float *w, float w0, cudaStream_t s);
cudaError_t launch_init_gaussian(float mx, float my, float mth, float sd_xy, float sd_th,
                                 int64_t gid0, int64_t n, uint64_t seed, float *x, float *y,
                                 float *th, float *w, float w0, cudaStream_t s);
}  // namespace amclj
