/*
 * pfslam.h -- C ABI of the B200-native particle-filter SLAM engine (libpfslam.so).
 *
 * Drop-in boundary for the per-frame hot path of michaelwillett/GPU-ICP-SLAM.  Each entry point
 * names the reference interface it replaces (paths relative to the reference repo).  The
 * reference's interface is five C++ free functions over file-static state (src/kernel.h:14-24);
 * this ABI makes the engine an explicit handle, takes N / map / device at run time instead of
 * #defines (kernel.cu:29-52), returns error codes instead of exit() (kernel.h:42-59), and never
 * synchronises the device except where a result is handed to the host.  The C++ wrappers with the
 * reference's exact kernel.h signatures live in gpu-icp-slam_b200/csrc/kernel_h_compat.cpp.
 *
 * Plain pointers and sizes only: no torch, glm or thrust types cross this boundary.
 */
#ifndef PFSLAM_H
#define PFSLAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFSLAM_ABI_VERSION 1

/* error codes (0 = ok); pfslam_last_error() gives the message for the calling thread */
#define PFSLAM_OK              0
#define PFSLAM_ERR_ARG         1
#define PFSLAM_ERR_CUDA        2
#define PFSLAM_ERR_STATE       3
#define PFSLAM_ERR_UNSUPPORTED 4

/* map representation */
#define PFSLAM_PATH_GRID2D 0   /* occupancy grid: PFMeasurementUpdate/PFUpdateMap, kernel.cu:307,551 */
#define PFSLAM_PATH_KD     1   /* kd-tree point cloud: PFMeasurementUpdateKD/PFUpdateMapKD, kernel.cu:1311,1406 */

/* scoring kernels; both produce the same integers as kernEvaluateParticles (kernel.cu:277) */
#define PFSLAM_SCORE_EXACT    0 /* one cosf/sinf per (particle, beam), the reference's expression */
#define PFSLAM_SCORE_FILTERED 1 /* hoisted trig + rounding guard band, exact fallback for near-ties */
#define PFSLAM_SCORE_TILED    2 /* FILTERED + TMA-staged 128x128 grid windows in shared memory (default) */

/* reference quirks reproduced by default */
#define PFSLAM_QUIRK_Q1_HALF_WEIGHT_SYNC 1u /* kernel.cu:337: only ceil(N/2) weights persist   */
#define PFSLAM_QUIRKS_REFERENCE (PFSLAM_QUIRK_Q1_HALF_WEIGHT_SYNC)

typedef struct pfslam_engine pfslam_engine;

/* Run-time replacement for PARTICLE_COUNT / LIDAR_SIZE / scene->maps[0] (kernel.cu:30,43,119-120). */
typedef struct pfslam_config {
    int32_t  abi_version;        /* PFSLAM_ABI_VERSION */
    int32_t  n_particles;        /* particles held by THIS engine (this GPU's shard) */
    int32_t  n_particles_global; /* particles of the whole filter; == n_particles on one GPU */
    int32_t  particle_offset;    /* global index of local particle 0 (multiple of 1024 if sharded) */
    int32_t  n_ranks;            /* GPUs sharing the filter (1 = single GPU) */
    int32_t  n_beams;            /* LIDAR_SIZE, 1081 */
    float    map_scale_x, map_scale_y; /* Patch.scale, metres (data/map_settings.txt: 40 40) */
    float    map_res_x, map_res_y;     /* Patch.resolution, metres per cell (0.025) */
    int32_t  device;             /* CUDA device ordinal */
    int32_t  path;               /* PFSLAM_PATH_* */
    int32_t  score_mode;         /* PFSLAM_SCORE_* */
    uint32_t quirks;             /* PFSLAM_QUIRK_* bit set */
    int32_t  kd_capacity;        /* kd path: node capacity (KD_MAX_SIZE, kernel.cu:77); 0 = 2 097 152 */
} pfslam_config;

/* Per-frame results, the engine-side equivalent of robotPos + the timers' inputs
 * (kernel.cu:64, :472-474, :323-326). */
typedef struct pfslam_frame_result {
    float   pose[3];             /* robotPos = best particle (x, y, theta) */
    int32_t fit_min, fit_max;    /* minmax_element of the scores */
    int32_t best_index;          /* global index of the arg-max (lowest index on ties) */
    float   sum_w, sum_w2, neff; /* PFResample's r, r2, Neff */
    int32_t resampled;           /* 1 if Neff < 0.7 N triggered the resample */
    int32_t n_free_cells;        /* cells cleared / reinforced by the map update this frame */
    int32_t n_wall_cells;
    int32_t n_slow_evals;        /* (particle, beam) pairs the filtered scorer re-did exactly */
    int32_t kd_size;             /* kd path: nodes in the tree after this frame (kdSize, kernel.cu:80) */
    int32_t kd_inserted;         /* kd path: nodes inserted this frame */
    int32_t exchange_timeout;    /* sharded engines: 1 once a peer-exchange wait hit its time limit (results invalid) */
    int32_t resample_count;      /* steps that resampled so far (reset by the explicit-pose test entry points) */
    int32_t wait_extrema_ns;     /* sharded engines: time this rank's step kernels have spent waiting for the peers' */
    int32_t wait_tiles_ns;       /*   extrema / tile sums so far (nanoseconds, accumulated; wraps after ~2 s of waiting) */
    int32_t n_windows;           /* tiled scorer: shared-memory windows placed for this frame's scan ... */
    int32_t n_wide_beams;        /* ... and beams whose hit box over the cloud fits none (scored through global memory) */
} pfslam_frame_result;

/* ---- life cycle: particleFilterInit(Scene*) / particleFilterFree(), kernel.cu:107-178 ---- */
void pfslam_default_config(pfslam_config *cfg);
int  pfslam_create(const pfslam_config *cfg, pfslam_engine **out);
int  pfslam_destroy(pfslam_engine *e);
const char *pfslam_last_error(void);
/* Run the engine's kernels on a caller-owned stream (e.g. torch's current stream) */
int  pfslam_set_stream(pfslam_engine *e, void *cuda_stream);

/* ---- the per-frame step: particleFilter(uchar4*, int frame, Lidar*), kernel.cu:1702; 2D order of
 * README.md:41-50: PFMotionUpdate, PFMeasurementUpdate, PFUpdateMap, PFResample ----
 * scan: n_beams float32 ranges in HOST memory (lidar->scans[frame]).  Blocks until the pose is
 * on the host.  particleFilterStep is the name BASELINE.json uses for the same call. */
int  pfslam_step(pfslam_engine *e, const float *scan, int32_t frame, pfslam_frame_result *out);
int  particleFilterStep(pfslam_engine *e, const float *scan, int32_t frame, float pose_out[3]);
/* Streaming variant of pfslam_step (the input side of src/main.cpp:199-206 + src/lidar.cpp without the blocking): the
 * scan is copied into a slot of a pinned ring and the step is enqueued -- one graph launch that pulls the scan from the
 * slot and publishes the frame result next to it; pfslam_wait blocks until that step's result is on the host.  Up to
 * PFSLAM_RING_DEPTH steps may be in flight (PFSLAM_ERR_STATE when the ring is full); tickets complete in order.
 * Grid path on a non-default stream only (PFSLAM_ERR_UNSUPPORTED otherwise). */
#define PFSLAM_RING_DEPTH 8
int  pfslam_submit(pfslam_engine *e, const float *scan, int32_t frame, int32_t *ticket);
int  pfslam_wait(pfslam_engine *e, int32_t ticket, pfslam_frame_result *out);

/* Same step without any host synchronisation: the scan is taken from DEVICE memory and the result
 * stays on the device until pfslam_fetch_result(). */
int  pfslam_step_async(pfslam_engine *e, const float *scan_dev, int32_t frame);
int  pfslam_fetch_result(pfslam_engine *e, pfslam_frame_result *out);

/* ---- the step's phases, for multi-GPU hosts that run a collective between them (see "multi-GPU" below) and for
 * function-level parity tests.  All asynchronous on the engine stream. ---- */
int  pfslam_upload_scan(pfslam_engine *e, const float *scan_host);                 /* kernel.cu:315 */
int  pfslam_phase_motion(pfslam_engine *e, int32_t frame);                         /* kernel.cu:400 */
int  pfslam_phase_score(pfslam_engine *e, const float *scan_dev);                  /* kernel.cu:319-326 */
/* ... all-gather PFSLAM_BUF_EXTREMA_LOCAL -> PFSLAM_BUF_EXTREMA_ALL ... */
int  pfslam_phase_weights(pfslam_engine *e);                                       /* kernel.cu:329-338, :456-463 */
int  pfslam_phase_map(pfslam_engine *e, const float *scan_dev);                    /* kernel.cu:551-577 */
/* ... all-gather PFSLAM_BUF_TILES_LOCAL -> _ALL, PFSLAM_BUF_POSE_LOCAL -> _ALL ... */
int  pfslam_phase_resample(pfslam_engine *e, int32_t frame);                       /* kernel.cu:472-486 */

/* Hosts that capture the phase calls into their own CUDA graph (dist.py does, together with the
 * all-gathers) switch the engine to external parameters: the phases then stop pushing {scan, frame}
 * themselves and read whatever the last pfslam_set_params put in the device StepParams -- a
 * stream-ordered 16-byte copy issued before each graph replay. */
int  pfslam_set_external_params(pfslam_engine *e, int32_t on);
int  pfslam_set_params(pfslam_engine *e, const float *scan_dev, int32_t frame);

/* occupancy-grid update alone (PFUpdateMap, kernel.cu:551) for an explicit pose */
int  pfslam_update_grid(pfslam_engine *e, const float *scan_host, const float pose[3]);
/* scoring alone (kernEvaluateParticles, kernel.cu:277): fit_out[n_particles] on the host */
int  pfslam_score_particles(pfslam_engine *e, const float *scan_host, int32_t *fit_out);

/* ---- state access: getPCData(), kernel.cu:803-813 (copies; the reference lends static arrays) ---- */
int  pfslam_get_particles(pfslam_engine *e, float *x, float *y, float *theta, float *w); /* each n_particles */
int  pfslam_set_particles(pfslam_engine *e, const float *x, const float *y, const float *theta, const float *w);
int  pfslam_get_grid(pfslam_engine *e, int8_t *grid_out);      /* map_w*map_h, idx = x*map_w + y */
int  pfslam_set_grid(pfslam_engine *e, const int8_t *grid_in);
int  pfslam_get_map_dim(pfslam_engine *e, int32_t *map_w, int32_t *map_h);
int  pfslam_get_pose(pfslam_engine *e, float pose[3]);
int  pfslam_synchronize(pfslam_engine *e);

/* ---- kd-tree point-cloud path (PFSLAM_PATH_KD) ----
 * nodes use the reference's KDTree::Node layout (kdtree.hpp:16-27): 8 x 4 B
 * {int axis, left, right, parent; float x, y, z, w}. */
/* kd-tree NN lookup alone (findCorrespondenceIndexKD, kernel.cu:924): n queries (x,y,z) -> node index */
int  pfslam_kd_nn(pfslam_engine *e, const float *q_xyz, int32_t n, int32_t *idx_out);
/* getPCData's kd part (kernel.cu:810-811): copies up to cap nodes, returns the tree size */
int  pfslam_get_kd(pfslam_engine *e, void *nodes_out, int32_t cap, int32_t *n_nodes);
int  pfslam_set_kd(pfslam_engine *e, const void *nodes_in, int32_t n_nodes);
/* measurement hook (bench.py's kd roofline line): mean number of tree nodes one NN walk of the scorer loads,
 * over the first n_sample particles x all in-range beams of the current scan */
int  pfslam_kd_mean_visits(pfslam_engine *e, int32_t n_sample, double *mean_visits);
/* ICP refinement alone (transformPointICP, kernel.cu:993-1093): one point-to-point step of `scan` against
 * the tree.  `robot_prev` is the pose the targets are built from (the global robotPos the reference reads
 * in kernGetWallsKD, kernel.cu:1011), `start` the pose that is corrected (the best particle's). */
int  pfslam_kd_icp(pfslam_engine *e, const float *scan_host, const float robot_prev[3], const float start[3],
                   float pose_out[3]);
/* On a PFSLAM_PATH_KD engine pfslam_update_grid() is PFUpdateMapKD alone (kernel.cu:1406-1540) for an
 * explicit robotPos: masks, point lists, NN, weight updates, inserts (first call: the first-scan build). */

/* ---- multi-GPU: particles sharded N/R per engine, map replicated (new; the reference is single-GPU) ----
 * Default transport: PEER MEMORY.  Every sharded engine (n_ranks > 1) owns one exchange region;
 * once each engine knows the regions of all ranks, pfslam_step / pfslam_step_async run the whole
 * sharded frame as one CUDA graph: the kernels store score extrema and weight-tile sums straight
 * into the peers' regions over NVLink, raise per-rank sequence flags there, spin (bounded by
 * PFSLAM_PEER_TIMEOUT_MS, default 10 000) on their own flags, and the resampler loads the drawn
 * particle's pose from its owner.  No collective library call and no host round trip per frame.
 * All ranks must step in lock step with the same frame numbers.  Sequence:
 *   one process per GPU:  pfslam_ipc_export -> exchange the 64-byte handles (any transport) ->
 *                         pfslam_ipc_connect for every other rank -> pfslam_exchange_ready -> barrier
 *   engines in one process: pfslam_exchange_region + pfslam_connect_peer, then pfslam_exchange_ready */
#define PFSLAM_IPC_HANDLE_BYTES 64
int  pfslam_exchange_region(pfslam_engine *e, void **dev_ptr, int64_t *bytes);
int  pfslam_ipc_export(pfslam_engine *e, void *handle_out /* PFSLAM_IPC_HANDLE_BYTES */);
int  pfslam_ipc_connect(pfslam_engine *e, int32_t rank, const void *handle);
int  pfslam_connect_peer(pfslam_engine *e, int32_t rank, void *peer_region);
int  pfslam_exchange_ready(pfslam_engine *e);

/* ---- alternative transport: device buffers a multi-GPU host all-gathers between the phase calls above
 * (dist.py with exchange="collective": torch.distributed all_gather_into_tensor) ---- */
#define PFSLAM_BUF_EXTREMA_LOCAL 0  /* 8 x 4 B: {min, max, argmax global idx, x, y, theta, 0, 0}   */
#define PFSLAM_BUF_EXTREMA_ALL   1  /* n_ranks x 8 x 4 B                                            */
#define PFSLAM_BUF_TILES_LOCAL   2  /* {tile sums of w, tile sums of w^2, lm[n]} float32, see bytes  */
#define PFSLAM_BUF_TILES_ALL     3  /* n_ranks x the above                                           */
#define PFSLAM_BUF_POSE_LOCAL    4  /* x[n], y[n], theta[n] float32                                  */
#define PFSLAM_BUF_POSE_ALL      5  /* n_ranks x the above                                           */
#define PFSLAM_BUF_SCAN          6  /* n_beams float32: the engine's device copy of the scan         */
int  pfslam_device_buffer(pfslam_engine *e, int32_t which, void **dev_ptr, int64_t *bytes);

/* number of kernels this engine has launched (bench.py's gpu_launches) */
int64_t pfslam_launch_count(pfslam_engine *e);

/* measurement hook for bench.py's roofline: runs the scoring phase once on the engine stream with
 * CUDA events around the dominant kernel alone (k_score_fast or k_score_exact) and around the whole
 * phase; returns both durations in milliseconds after synchronising */
int  pfslam_profile_score(pfslam_engine *e, float *ms_kernel, float *ms_phase);

/* in-step variant: while enabled, every scoring phase (pfslam_step, pfslam_step_async,
 * pfslam_phase_score) brackets its dominant kernel with a CUDA event pair (ring of 4096); pfslam_profile_read
 * synchronises, returns the mean duration in ms over the recorded launches and clears the ring */
int  pfslam_profile_enable(pfslam_engine *e, int32_t on);
int  pfslam_profile_read(pfslam_engine *e, float *ms_kernel_mean, int32_t *n_launches);

/* per-kernel breakdown: while on, single-GPU grid steps run as plain, serialised launches (no graph, no
 * side-by-side branches) with a CUDA event after every kernel; _read synchronises and returns, per lap id
 * 0..PFSLAM_LAP_COUNT-1 (pfslam_lap_name gives the kernel), the mean duration in ms and the launch count */
#define PFSLAM_LAP_COUNT 10
int  pfslam_profile_laps(pfslam_engine *e, int32_t on);
int  pfslam_profile_laps_read(pfslam_engine *e, float ms_mean[PFSLAM_LAP_COUNT], int32_t count[PFSLAM_LAP_COUNT]);
const char *pfslam_lap_name(int32_t id);

/* tuning hook: in-graph timeline.  While on, every kernel of the 2D step folds %globaltimer into a [first entry, last
 * exit] pair; (on, out): copy the pairs recorded so far (2 * PFSLAM_TRACE_COUNT words, nanoseconds), then reset.
 * The caller synchronises the engine's stream around the call. */
#define PFSLAM_TRACE_COUNT 16
int  pfslam_debug_trace(int32_t on, uint64_t *out);
const char *pfslam_trace_name(int32_t id);

/* tuning hook: with PFSLAM_STAGED_DEBUG & 16 the scoring kernel's blocks stamp their phase boundaries
 * (12 words per block, see pf_score_staged.cuh); copies the first n_words of that table */
int  pfslam_debug_staged_timing(uint64_t *out, int32_t n_words);

/* test hook: libdevice cosf/sinf of n host floats evaluated on the device (the functions the
 * reference's kernels call, kernel.cu:185-186); test hook: lets the test suite validate its CPU emulation of them */
int  pfslam_debug_trig(int32_t device, const float *x_host, int64_t n, float *cos_out, float *sin_out);

#ifdef __cplusplus
}
#endif
#endif
