// pfslam.cu -- host side of libpfslam.so: engine state, the per-frame launch sequence and the
// extern "C" ABI declared in include/pfslam.h.  B200 (sm_100a) only; there is no CPU fallback:
// every entry point either runs the CUDA kernels or returns an error.
//
// Reference boundary being replaced: src/kernel.h:14-24 / src/kernel.cu:107-178, :307-621,
// :1702-1768 of michaelwillett/GPU-ICP-SLAM.
#include "../../include/pfslam.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "pf_kernels2d.cuh"
#include "pf_score_filtered.cuh"
#include "pf_score_tiled.cuh"
#include "pf_score_staged.cuh"
#include "pf_kernels_kd.cuh"

#include <algorithm>

using namespace pf;

static thread_local std::string g_last_error;

static int set_error(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return set_error(PFSLAM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    } while (0)

struct pfslam_engine {
    pfslam_config cfg{};
    MapGeom geom{};
    int n = 0, n_global = 0, gidx0 = 0, n_ranks = 1;
    int n_tiles = 0;              // local tiles
    long long tiles_block = 0;    // floats per rank in the tiles buffer
    int n_score_blocks = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // device state
    float *x = nullptr, *y = nullptr, *th = nullptr;   // POSE_LOCAL: one allocation x|y|th
    float *w = nullptr;
    int *fit = nullptr;
    int8_t *grid = nullptr;
    unsigned *free_bits = nullptr, *wall_bits = nullptr;
    unsigned *free_stamp = nullptr; // grid path: per-cell epoch stamps of k_map_free's claims (never cleared between frames)
    size_t bits_bytes = 0;
    float *scan = nullptr, *angle = nullptr;
    int *blk_min = nullptr; long long *blk_maxkey = nullptr;
    Extrema *ext_local = nullptr, *ext_all = nullptr;
    float *tiles_local = nullptr, *tiles_all = nullptr;
    float *pose_all = nullptr;
    float *prefix = nullptr;
    FrameResult *res = nullptr;
    int *counters = nullptr;
    ScoreFilteredWork *fwork = nullptr;
    int *score_partial = nullptr;
    TiledWork *twork = nullptr;
    double2 *angle_cs = nullptr;
    float4 *pcs = nullptr;         // {x, y, cos(theta), sin(theta)} per particle, for the staged scorer
    bool prefix_fused = false;
    bool resample_follows_weights = false;   // k_resample is launched right after k_weights_scan on the same stream
    bool bounds_valid = false;     // cloud bounds in twork were produced by k_motion for the current poses
    CUtensorMap tmap;
    int score_mode = 0;            // effective mode (TILED falls back to FILTERED when unsupported)
    bool snap_push = true;         // peer shards: snapshots pushed to every peer (k_snapshot_push); PFSLAM_SNAPSHOT=pull: gathered remotely
    bool tail_fused = false;       // PFSLAM_TAIL=fused: weights + prefix + resample as one launch (k_weights_resample, measured slower)
    int n_sms = 0;
    bool staged = false;           // scorer generation: k_score_tiled (default) or k_score_staged (PFSLAM_TILED_KERNEL=staged)
    int tiled_grid = 0;            // k_score_tiled grid: SMs x resident blocks per SM
    // pinned host staging
    // pinned + device-mapped staging: slot 0 serves the blocking calls (upload_scan, fetch_result, ...), slots
    // 1..kRing the streaming ring of pfslam_submit / pfslam_wait (host pointer, device alias)
    float *h_scan = nullptr, *h_scan_dev = nullptr;
    FrameResult *h_res = nullptr, *h_res_dev = nullptr;
    bool io_capture = false;               // capturing the host-API flavour of the step graph
    cudaEvent_t ring_ev[PFSLAM_RING_DEPTH] = {};
    int ring_ticket[PFSLAM_RING_DEPTH] = {};   // ticket occupying the slot (0 = free)
    int next_ticket = 1;
    long long launches = 0;
    // kd-tree point-cloud path
    KdNode *kd = nullptr; int kd_cap = 0;
    KdSearch *kds = nullptr;               // 16-byte search shadow of kd (planar trees)
    int kd_size_ub = 0;                    // host-side upper bound of the device tree size
    bool kd_flat = true;                   // every node has z == 0 and a valid axis: the scorer walks the shadow
    int kd_walk = 2;                       // shadow format / visit body: 2 = branch-free visit (default), 1 = round 1's (PFSLAM_KD_WALK=1)
    KdState *ks = nullptr;
    int *bits_blk = nullptr; int n_bits_blk = 0;
    int *free_cells = nullptr, *wall_cells = nullptr; int pc_cap = 8192;
    float2 *kd_pts = nullptr; int *kd_nn_idx = nullptr, *kd_ins_index = nullptr;
    int *kd_claim = nullptr; int kd_stamp = 0;   // once-per-node-per-pass weight updates (k_kd_weights)
    bool kd_empty = true;                  // no tree yet (kdSize == 0, kernel.cu:1714)
    std::vector<KdNode> h_kd;
    float *kd_q = nullptr; int *kd_qi = nullptr; int kd_q_cap = 0;    // pfslam_kd_nn's query / result buffers (grown on demand)
    unsigned long long *kd_cnt = nullptr;
    // per-step parameters (device copy + pinned ring) and the captured step graph
    StepParams *sp = nullptr;
    StepParams *h_sp = nullptr;            // kParamSlots pinned slots
    StepParams cur{};                      // what the device copy will hold once the stream drains
    unsigned long long n_param_pushes = 0;
    bool cur_valid = false;        // `cur` is what the device StepParams hold (or will, in stream order)
    cudaEvent_t lap_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphNode_t graph_param_node = nullptr;      // the k_motion kernel node: the step's parameters are its last argument
    // the same step with the host API's copies inside: scan H2D from the pinned staging buffer at the head,
    // frame result D2H into pinned memory at the tail (pfslam_step = one launch + one synchronisation)
    cudaGraph_t graph_io = nullptr;
    cudaGraphExec_t graph_io_exec = nullptr;
    cudaGraphNode_t graph_io_param_node = nullptr;
    StepParams capture_params = {};                   // k_motion's by-value argument while a step graph is being captured
    bool graph_failed = false;
    bool use_graph = true;
    int graph_kernels = 0, graph_io_kernels = 0;
    bool in_capture = false;
    bool external_params = false;
    // shard exchange (pf_xchg.cuh): xc_host = single GPU / host-run collectives, xc_p2p = peer memory
    unsigned char *xreg = nullptr; size_t xreg_bytes = 0;
    size_t xoff_flags = 0, xoff_ext = 0, xoff_tiles = 0, xoff_snap = 0;
    Xchg xc_host{}, xc_p2p{};
    const Xchg *cur_xc = nullptr;
    void *peer_base[kMaxRanks] = {};
    bool peer_ipc[kMaxRanks] = {};
    bool p2p_ready = false;
    int seq = 0;                   // step sequence number (StepParams.seq)
    // independent kernels of a step run side by side: an auxiliary stream forked / joined with events
    // (branches of the step graph once captured)
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork[3] = {nullptr, nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    bool overlap = true;
    // serialised per-kernel timing (pfslam_profile_laps)
    bool laps_on = false;
    LapRec laps;
    // in-step kernel timing
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    int prof_n = 0;
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int kParamSlots = 256;

// next pinned parameter slot; slots are recycled after a full lap, guarded by one event per quarter lap
static int next_param_slot(pfslam_engine *e, StepParams **slot)
{
    const int s = (int)(e->n_param_pushes % kParamSlots);
    if (s % (kParamSlots / 4) == 0 && e->n_param_pushes >= (unsigned long long)kParamSlots)
        CUDA_TRY(cudaEventSynchronize(e->lap_ev[s / (kParamSlots / 4)]));
    *slot = &e->h_sp[s];
    return PFSLAM_OK;
}
static int param_slot_used(pfslam_engine *e)
{
    const int s = (int)(e->n_param_pushes % kParamSlots);
    e->n_param_pushes++;
    if (s % (kParamSlots / 4) == kParamSlots / 4 - 1) CUDA_TRY(cudaEventRecord(e->lap_ev[s / (kParamSlots / 4)], e->stream));
    return PFSLAM_OK;
}

// make the device StepParams equal (scan, frame) for the kernels enqueued after this call
static int push_params(pfslam_engine *e, const float *scan, int frame)
{
    if (e->in_capture || e->external_params) return PFSLAM_OK;   // the graph's k_motion argument / the host does it
    if (e->cur.scan == scan && e->cur.frame == frame && e->cur.seq == e->seq && e->cur_valid) return PFSLAM_OK;
    StepParams *slot = nullptr;
    int rc = next_param_slot(e, &slot);
    if (rc) return rc;
    slot->scan = scan; slot->frame = frame; slot->seq = e->seq; slot->scan_src = nullptr; slot->res_host = nullptr;
    CUDA_TRY(cudaMemcpyAsync(e->sp, slot, sizeof(StepParams), cudaMemcpyHostToDevice, e->stream));
    e->cur = *slot; e->cur_valid = true;
    return param_slot_used(e);
}

constexpr int kMotionArgs = 17;       // k_motion's parameter count; the by-value StepParams is the last one
template <class... A> constexpr int kernel_arity(void (*)(A...)) { return (int)sizeof...(A); }
template <class R, class... A> struct last_arg { using type = typename last_arg<A...>::type; };
template <class R> struct last_arg<R> { using type = R; };
template <class... A> constexpr bool ends_in_step_params(void (*)(A...)) { return std::is_same<typename last_arg<A...>::type, const StepParams>::value || std::is_same<typename last_arg<A...>::type, StepParams>::value; }
static_assert(kernel_arity(k_motion) == kMotionArgs && ends_in_step_params(k_motion), "launch_graph patches k_motion's last argument");

extern "C" {

const char *pfslam_last_error(void) { return g_last_error.c_str(); }

void pfslam_default_config(pfslam_config *c)
{
    memset(c, 0, sizeof *c);
    c->abi_version = PFSLAM_ABI_VERSION;
    c->n_particles = 1000;                 // PARTICLE_COUNT, kernel.cu:30
    c->n_particles_global = 1000;
    c->particle_offset = 0;
    c->n_ranks = 1;
    c->n_beams = 1081;                     // LIDAR_SIZE, kernel.cu:43
    c->map_scale_x = 40.0f; c->map_scale_y = 40.0f;   // data/map_settings.txt
    c->map_res_x = 0.025f; c->map_res_y = 0.025f;
    c->device = 0;
    c->path = PFSLAM_PATH_GRID2D;
    c->score_mode = PFSLAM_SCORE_TILED;
    c->quirks = PFSLAM_QUIRKS_REFERENCE;
}

int pfslam_destroy(pfslam_engine *e)
{
    if (!e) return PFSLAM_OK;               // particleFilterFree() is called before Init (main.cpp:194)
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->x); cudaFree(e->w); cudaFree(e->fit); cudaFree(e->grid);
    cudaFree(e->free_bits); cudaFree(e->wall_bits); cudaFree(e->free_stamp); cudaFree(e->scan); cudaFree(e->angle);
    cudaFree(e->blk_min); cudaFree(e->blk_maxkey); cudaFree(e->ext_local);
    if (e->ext_all != e->ext_local) cudaFree(e->ext_all);
    cudaFree(e->tiles_local);
    if (e->tiles_all != e->tiles_local) cudaFree(e->tiles_all);
    cudaFree(e->pose_all); cudaFree(e->prefix); cudaFree(e->res); cudaFree(e->counters);
    cudaFree(e->fwork); cudaFree(e->score_partial); cudaFree(e->twork); cudaFree(e->pcs); cudaFree(e->angle_cs); cudaFree(e->sp);
    cudaFree(e->kd); cudaFree(e->kds); cudaFree(e->ks); cudaFree(e->bits_blk); cudaFree(e->free_cells); cudaFree(e->wall_cells);
    cudaFree(e->kd_pts); cudaFree(e->kd_nn_idx); cudaFree(e->kd_ins_index); cudaFree(e->kd_claim);
    cudaFree(e->kd_q); cudaFree(e->kd_qi); cudaFree(e->kd_cnt);
    for (int r = 0; r < kMaxRanks; r++) if (e->peer_ipc[r] && e->peer_base[r]) cudaIpcCloseMemHandle(e->peer_base[r]);
    cudaFree(e->xreg);
    cudaFreeHost(e->h_sp);
    for (auto ev : e->lap_ev) if (ev) cudaEventDestroy(ev);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->graph) cudaGraphDestroy(e->graph);
    if (e->graph_io_exec) cudaGraphExecDestroy(e->graph_io_exec);
    if (e->graph_io) cudaGraphDestroy(e->graph_io);
    cudaFreeHost(e->h_scan); cudaFreeHost(e->h_res);
    for (auto ev : e->ring_ev) if (ev) cudaEventDestroy(ev);
    for (auto ev : e->prof_ev) cudaEventDestroy(ev);
    for (auto ev : e->laps.ev) cudaEventDestroy(ev);
    for (int i = 0; i < 3; i++) if (e->ev_fork[i]) cudaEventDestroy(e->ev_fork[i]);
    for (int i = 0; i < 2; i++) if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    if (e->aux) cudaStreamDestroy(e->aux);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return PFSLAM_OK;
}

static int engine_alloc(pfslam_engine *e)
{
    const int n = e->n;
    const size_t ncell = (size_t)e->geom.w * e->geom.h;
    CUDA_TRY(cudaMalloc(&e->x, sizeof(float) * 3 * n));
    e->y = e->x + n; e->th = e->x + 2 * n;
    CUDA_TRY(cudaMalloc(&e->w, sizeof(float) * n));
    CUDA_TRY(cudaMalloc(&e->fit, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&e->grid, ncell));
    e->bits_bytes = ((ncell + 31) / 32) * 4;
    CUDA_TRY(cudaMalloc(&e->free_bits, e->bits_bytes));
    CUDA_TRY(cudaMalloc(&e->wall_bits, e->bits_bytes));
    CUDA_TRY(cudaMalloc(&e->free_stamp, sizeof(unsigned) * (size_t)ncell));
    CUDA_TRY(cudaMalloc(&e->scan, sizeof(float) * (e->cfg.n_beams + 32)));
    CUDA_TRY(cudaMalloc(&e->angle, sizeof(float) * (e->cfg.n_beams + 32)));
    e->n_score_blocks = score_partial_count(n);
    CUDA_TRY(cudaMalloc(&e->blk_min, sizeof(int) * e->n_score_blocks));
    CUDA_TRY(cudaMalloc(&e->blk_maxkey, sizeof(long long) * e->n_score_blocks));
    CUDA_TRY(cudaMalloc(&e->ext_local, sizeof(Extrema)));
    e->n_tiles = ceil_div(n, kTile);
    e->tiles_block = 2ll * e->n_tiles + n;
    CUDA_TRY(cudaMalloc(&e->tiles_local, sizeof(float) * e->tiles_block));
    if (e->n_ranks > 1) {
        CUDA_TRY(cudaMalloc(&e->ext_all, sizeof(Extrema) * e->n_ranks));
        CUDA_TRY(cudaMalloc(&e->tiles_all, sizeof(float) * e->tiles_block * e->n_ranks));
    } else {
        e->ext_all = e->ext_local;
        e->tiles_all = e->tiles_local;
    }
    CUDA_TRY(cudaMalloc(&e->pose_all, sizeof(float) * (e->n_ranks == 1 ? 4 : 3) * (size_t)n * e->n_ranks));
    CUDA_TRY(cudaMalloc(&e->prefix, sizeof(float) * ((size_t)e->n_tiles * e->n_ranks + 1)));
    {   // single GPU, or a host that runs its own collectives between the phases
        Xchg &h = e->xc_host;
        h.n_ranks = e->n_ranks; h.rank = e->n_ranks > 1 ? e->gidx0 / n : 0; h.parity_mask = 0; h.timeout_ms = 0;
        h.tiles_block = e->tiles_block; h.sum_off = 0; h.lm_off = 2ll * e->n_tiles; h.snap_stride = 0;
        h.ext_all = e->ext_all; h.tiles_all = e->tiles_all;
        h.snap = e->n_ranks == 1 ? e->pose_all : nullptr;
        h.snap_aos = e->n_ranks == 1 ? 1 : 0;
        for (int r = 0; r < e->n_ranks && r < kMaxRanks; r++) h.pose_src[r] = e->pose_all + (size_t)r * 3 * n;
        e->cur_xc = &e->xc_host;
    }
    if (e->n_ranks > 1) {
        // the exchange region peers store into / load from (pf_xchg.cuh); same layout on every rank
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t tb = (size_t)n + (((size_t)2 * e->n_tiles + 3) & ~(size_t)3);
        e->xoff_flags = 0;
        e->xoff_ext = up(sizeof(int) * 2 * kMaxRanks);
        e->xoff_tiles = up(e->xoff_ext + sizeof(Extrema) * 2 * kMaxRanks);
        e->xoff_snap = up(e->xoff_tiles + sizeof(float) * 2 * e->n_ranks * tb);
        e->xreg_bytes = up(e->xoff_snap + sizeof(float) * 2 * 4 * (size_t)n * e->n_ranks);     // snapshots [parity][rank][n] float4
        CUDA_TRY(cudaMalloc(&e->xreg, e->xreg_bytes));
        CUDA_TRY(cudaMemsetAsync(e->xreg, 0, e->xreg_bytes, e->stream));
        Xchg &x = e->xc_p2p;
        x.n_ranks = e->n_ranks; x.rank = e->gidx0 / n; x.parity_mask = 1;
        const char *to = getenv("PFSLAM_PEER_TIMEOUT_MS");
        x.timeout_ms = to ? (unsigned)atoi(to) : 10000u;
        x.tiles_block = (long long)tb; x.sum_off = n; x.lm_off = 0; x.snap_stride = 4ll * n * e->n_ranks; x.snap_aos = 1;
        x.ext_all = reinterpret_cast<Extrema *>(e->xreg + e->xoff_ext);
        x.tiles_all = reinterpret_cast<float *>(e->xreg + e->xoff_tiles);
        x.snap = reinterpret_cast<float *>(e->xreg + e->xoff_snap) + 4ll * n * x.rank;          // own slot of parity 0
        x.flags = reinterpret_cast<int *>(e->xreg + e->xoff_flags);
        x.off_ext = (long long)e->xoff_ext; x.off_tiles = (long long)e->xoff_tiles; x.off_flags = (long long)e->xoff_flags; x.off_snap = (long long)e->xoff_snap;
        e->peer_base[x.rank] = e->xreg;
    }
    CUDA_TRY(cudaMalloc(&e->res, sizeof(FrameResult)));
    CUDA_TRY(cudaMalloc(&e->counters, sizeof(int) * 16));
    CUDA_TRY(cudaMalloc(&e->fwork, sizeof(ScoreFilteredWork)));
    CUDA_TRY(cudaMalloc(&e->score_partial, sizeof(int) * (size_t)score_tiled_rows() * n));
    CUDA_TRY(cudaMalloc(&e->twork, sizeof(TiledWork)));
    CUDA_TRY(cudaMalloc(&e->pcs, sizeof(float4) * (size_t)n));
    CUDA_TRY(cudaMemsetAsync(e->pcs, 0, sizeof(float4) * (size_t)n, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->twork, 0, sizeof(TiledWork), e->stream));
    CUDA_TRY(cudaMalloc(&e->sp, sizeof(StepParams)));
    CUDA_TRY(cudaMemsetAsync(e->sp, 0, sizeof(StepParams), e->stream));
    CUDA_TRY(cudaMallocHost(&e->h_sp, sizeof(StepParams) * kParamSlots));
    for (auto &ev : e->lap_ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (e->cfg.path == PFSLAM_PATH_KD) {
        e->kd_cap = e->cfg.kd_capacity > 0 ? e->cfg.kd_capacity : (1 << 21);
        CUDA_TRY(cudaMalloc(&e->kd, sizeof(KdNode) * (size_t)e->kd_cap));
        CUDA_TRY(cudaMalloc(&e->kds, sizeof(KdSearch) * (size_t)e->kd_cap));
        { const char *kw = getenv("PFSLAM_KD_WALK"); e->kd_walk = (kw && atoi(kw) == 1) ? 1 : 2; }
        CUDA_TRY(cudaMalloc(&e->ks, sizeof(KdState)));
        CUDA_TRY(cudaMemsetAsync(e->ks, 0, sizeof(KdState), e->stream));
        e->n_bits_blk = ceil_div((int)(e->bits_bytes / 4), kBitsBlockWords);
        CUDA_TRY(cudaMalloc(&e->bits_blk, sizeof(int) * 2 * e->n_bits_blk));
        CUDA_TRY(cudaMalloc(&e->free_cells, sizeof(int) * e->pc_cap));
        CUDA_TRY(cudaMalloc(&e->wall_cells, sizeof(int) * e->pc_cap));
        CUDA_TRY(cudaMalloc(&e->kd_pts, sizeof(float2) * 2 * e->pc_cap));
        CUDA_TRY(cudaMalloc(&e->kd_nn_idx, sizeof(int) * 2 * e->pc_cap));
        CUDA_TRY(cudaMalloc(&e->kd_ins_index, sizeof(int) * e->pc_cap));
        CUDA_TRY(cudaMalloc(&e->kd_claim, sizeof(int) * (size_t)e->kd_cap));
        CUDA_TRY(cudaMemsetAsync(e->kd_claim, 0, sizeof(int) * (size_t)e->kd_cap, e->stream));
    }
    CUDA_TRY(cudaHostAlloc(&e->h_scan, sizeof(float) * e->cfg.n_beams * (PFSLAM_RING_DEPTH + 1), cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(&e->h_res, sizeof(FrameResult) * (PFSLAM_RING_DEPTH + 1), cudaHostAllocMapped));
    memset(e->h_res, 0, sizeof(FrameResult) * (PFSLAM_RING_DEPTH + 1));
    for (auto &ev : e->ring_ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaHostGetDevicePointer(&e->h_scan_dev, e->h_scan, 0));
    CUDA_TRY(cudaHostGetDevicePointer(&e->h_res_dev, e->h_res, 0));
    // initial state: kernel.cu:122-132
    CUDA_TRY(cudaMemsetAsync(e->x, 0, sizeof(float) * 3 * n, e->stream));
    std::vector<float> ones(n, 1.0f);
    CUDA_TRY(cudaMemcpyAsync(e->w, ones.data(), sizeof(float) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->grid, -100 & 0xff, ncell, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->free_bits, 0, e->bits_bytes, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->wall_bits, 0, e->bits_bytes, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->free_stamp, 0, sizeof(unsigned) * (size_t)e->geom.w * e->geom.h, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->scan, 0, sizeof(float) * (e->cfg.n_beams + 32), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->fit, 0, sizeof(int) * n, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->res, 0, sizeof(FrameResult), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->counters, 0, sizeof(int) * 16, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->fwork, 0, sizeof(ScoreFilteredWork), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->tiles_local, 0, sizeof(float) * e->tiles_block, e->stream));
    k_init_beams<<<ceil_div(e->cfg.n_beams + 32, 128), 128, 0, e->stream>>>(e->angle, e->cfg.n_beams + 32);
    CUDA_TRY(cudaMalloc(&e->angle_cs, sizeof(double2) * (e->cfg.n_beams + 32)));
    k_init_beam_trig<<<ceil_div(e->cfg.n_beams + 32, 128), 128, 0, e->stream>>>(e->angle, e->cfg.n_beams + 32, e->angle_cs);
    k_bounds_reset<<<1, 32, 0, e->stream>>>(e->twork);
    e->launches += 3;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

// CUDA loads kernels lazily, and a first-use load can wait for the device to drain.  A sharded engine's
// kernels spin on flags that another engine's (another stream's) kernels raise, so every kernel must be
// resident before the first step: a load blocking behind a spinning kernel would stall its own peer.
static int preload_kernels()
{
    cudaFuncAttributes a;
#define PF_PRELOAD(k) CUDA_TRY(cudaFuncGetAttributes(&a, k))
    PF_PRELOAD(k_motion); PF_PRELOAD(k_cloud_bounds); PF_PRELOAD(k_bounds_reset); PF_PRELOAD(k_tile_prep);
    PF_PRELOAD(k_beam_prep); CUDA_TRY(cudaFuncGetAttributes(&a, k_score_tiled<256, 4>)); CUDA_TRY(cudaFuncGetAttributes(&a, k_score_tiled<512, 2>)); PF_PRELOAD(k_score_fast); PF_PRELOAD(k_score_exact);
    CUDA_TRY(cudaFuncGetAttributes(&a, staged_kernel()));
    PF_PRELOAD(k_score_combine); PF_PRELOAD(k_score_combine_rows); PF_PRELOAD(k_extrema);
    PF_PRELOAD(k_weights_scan); PF_PRELOAD(k_weights_resample); PF_PRELOAD(k_prefix); PF_PRELOAD(k_resample); PF_PRELOAD(k_map_free); PF_PRELOAD(k_map_wall);
    PF_PRELOAD(k_score_kd<0>); PF_PRELOAD(k_score_kd<1>); PF_PRELOAD(k_score_kd<2>); PF_PRELOAD(k_kd_shadow<1>); PF_PRELOAD(k_kd_shadow<2>); PF_PRELOAD(k_icp); PF_PRELOAD(k_kd_mark); PF_PRELOAD(k_bits_count); PF_PRELOAD(k_bits_offsets);
    PF_PRELOAD(k_bits_scatter); PF_PRELOAD(k_kd_points_nn); PF_PRELOAD(k_kd_weights); PF_PRELOAD(k_kd_insert);
    PF_PRELOAD(k_kd_finish); PF_PRELOAD(k_kd_nn); PF_PRELOAD(k_xc_wait); PF_PRELOAD(k_publish_result); PF_PRELOAD(k_snapshot_push);
#undef PF_PRELOAD
    return PFSLAM_OK;
}

int pfslam_create(const pfslam_config *cfg, pfslam_engine **out)
{
    if (!cfg || !out) return set_error(PFSLAM_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != PFSLAM_ABI_VERSION)
        return set_error(PFSLAM_ERR_ARG, "abi_version %d != %d", cfg->abi_version, PFSLAM_ABI_VERSION);
    if (cfg->n_particles <= 0 || cfg->n_beams <= 0 || cfg->n_beams > 4096)
        return set_error(PFSLAM_ERR_ARG, "bad n_particles/n_beams");
    if (cfg->map_res_x <= 0.f || cfg->map_res_y <= 0.f || cfg->map_scale_x <= 0.f || cfg->map_scale_y <= 0.f)
        return set_error(PFSLAM_ERR_ARG, "bad map scale/resolution");
    if (cfg->n_ranks < 1 || cfg->n_particles_global < cfg->n_particles)
        return set_error(PFSLAM_ERR_ARG, "bad sharding");
    if (cfg->n_ranks > 1 && (cfg->n_particles % kTile != 0 || cfg->particle_offset % kTile != 0 ||
                             (long long)cfg->n_particles * cfg->n_ranks != cfg->n_particles_global))
        return set_error(PFSLAM_ERR_ARG, "sharded engines need n_particles %% 1024 == 0 and equal shards");
    if (cfg->path != PFSLAM_PATH_GRID2D && cfg->path != PFSLAM_PATH_KD)
        return set_error(PFSLAM_ERR_ARG, "unknown path %d", cfg->path);
    if (cfg->n_ranks > kMaxRanks)
        return set_error(PFSLAM_ERR_UNSUPPORTED, "at most %d ranks", kMaxRanks);
    if (cfg->score_mode < PFSLAM_SCORE_EXACT || cfg->score_mode > PFSLAM_SCORE_TILED)
        return set_error(PFSLAM_ERR_ARG, "bad score_mode");
    // Every kernel indexes the grid as x * map_w + y, like the reference (kernel.cu:251, :1436), which is only
    // consistent for square maps; the reference's scene file has one RES and its maps are square.
    if ((int)(cfg->map_scale_x / cfg->map_res_x) != (int)(cfg->map_scale_y / cfg->map_res_y))
        return set_error(PFSLAM_ERR_UNSUPPORTED, "non-square maps (%d x %d cells) are not supported",
                         (int)(cfg->map_scale_x / cfg->map_res_x), (int)(cfg->map_scale_y / cfg->map_res_y));
    // the global tile prefix (k_prefix / the fused last block) holds 2 floats per 1024-particle tile in shared memory
    if (((long long)cfg->n_particles_global + kTile - 1) / kTile > kMaxPrefixTiles)
        return set_error(PFSLAM_ERR_UNSUPPORTED, "at most %d particles in total", kMaxPrefixTiles * kTile);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev)
        return set_error(PFSLAM_ERR_ARG, "device %d of %d", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    pfslam_engine *e = new pfslam_engine();
    e->cfg = *cfg;
    e->n = cfg->n_particles; e->n_global = cfg->n_particles_global;
    e->gidx0 = cfg->particle_offset; e->n_ranks = cfg->n_ranks;
    // map_dim = ivec2(scale / resolution), kernel.cu:120 (float division, truncation)
    e->geom.w = (int)(cfg->map_scale_x / cfg->map_res_x);
    e->geom.h = (int)(cfg->map_scale_y / cfg->map_res_y);
    e->geom.scale_x = cfg->map_scale_x; e->geom.scale_y = cfg->map_scale_y;
    e->geom.res_x = cfg->map_res_x; e->geom.res_y = cfg->map_res_y;
    if (e->geom.w <= 0 || e->geom.h <= 0 || (long long)e->geom.w * e->geom.h > (1ll << 30)) {
        delete e; return set_error(PFSLAM_ERR_ARG, "bad map dimensions");
    }
    cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete e; return set_error(PFSLAM_ERR_CUDA, "stream: %s", cudaGetErrorString(ce)); }
    e->own_stream = true;
    {   // the side branches yield to the critical path
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        ce = cudaStreamCreateWithPriority(&e->aux, cudaStreamNonBlocking, lo);
    }
    for (int i = 0; i < 3 && ce == cudaSuccess; i++) ce = cudaEventCreateWithFlags(&e->ev_fork[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && ce == cudaSuccess; i++) ce = cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming);
    if (ce != cudaSuccess) { pfslam_destroy(e); return set_error(PFSLAM_ERR_CUDA, "aux stream: %s", cudaGetErrorString(ce)); }
    { const char *no = getenv("PFSLAM_NO_OVERLAP"); e->overlap = !(no && atoi(no) != 0); }
    // scorer generation: k_score_tiled (3 blocks per SM, one window per block at a time) is the default -- inside the step
    // graph it still finishes ~10 us ahead of k_score_staged (1 block per SM, 5 windows resident, all beams in one kernel),
    // whose block-wide stage loads nothing else on the SM can hide; PFSLAM_TILED_KERNEL=staged selects the latter
    { const char *tk = getenv("PFSLAM_TILED_KERNEL"); e->staged = tk && strcmp(tk, "staged") == 0; }
    { const char *sm = getenv("PFSLAM_SNAPSHOT"); e->snap_push = !(sm && strcmp(sm, "pull") == 0); }
    { const char *tl = getenv("PFSLAM_TAIL"); e->tail_fused = tl && strcmp(tl, "fused") == 0; }
    if (cudaDeviceGetAttribute(&e->n_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess) e->n_sms = 0;
    { const char *dg = getenv("PFSLAM_STAGED_DEBUG"); const int v = dg ? atoi(dg) : 0; cudaMemcpyToSymbol(g_staged_dbg, &v, sizeof v); }
    int rc = engine_alloc(e);
    if (rc != PFSLAM_OK) { std::string keep = g_last_error; pfslam_destroy(e); g_last_error = keep; return rc; }
    if ((rc = preload_kernels()) != PFSLAM_OK) { std::string keep = g_last_error; pfslam_destroy(e); g_last_error = keep; return rc; }
    rc = score_filtered_setup(e->cfg.device);
    if (rc != 0) { pfslam_destroy(e); return set_error(PFSLAM_ERR_CUDA, "scoring kernel setup failed"); }
    // the fixed-point scorers need an integral map centre c0 = 0.5*scale/res and maps below 2048 cells;
    // the tiled scorer additionally needs a TMA-legal row stride and at most 2048 beams
    {
        const float c0x = (0.5f * cfg->map_scale_x) / cfg->map_res_x, c0y = (0.5f * cfg->map_scale_y) / cfg->map_res_y;
        const bool fixed_ok = c0x == (float)(int)c0x && c0y == (float)(int)c0y && e->geom.w <= 2000 && e->geom.h <= 2000;
        e->score_mode = cfg->score_mode;
        if (!fixed_ok) e->score_mode = PFSLAM_SCORE_EXACT;
        if (e->score_mode == PFSLAM_SCORE_TILED) {
            if (cfg->n_beams > kMaxGroups * kChunkBeams || make_grid_tensor_map(&e->tmap, e->grid, e->geom.w, e->geom.h) != 0 ||
                (e->tiled_grid = (e->staged ? score_staged_setup(cfg->device) : score_tiled_setup(cfg->device))) <= 0)
                e->score_mode = PFSLAM_SCORE_FILTERED;
        }
    }
    *out = e;
    return PFSLAM_OK;
}

int pfslam_set_stream(pfslam_engine *e, void *cuda_stream)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->own_stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
    e->stream = (cudaStream_t)cuda_stream;
    // the legacy default stream cannot be captured into a graph: plain launches there
    e->use_graph = cuda_stream != nullptr && cuda_stream != (void *)cudaStreamLegacy;
    return PFSLAM_OK;
}

int pfslam_set_external_params(pfslam_engine *e, int32_t on)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->external_params = on != 0;
    return PFSLAM_OK;
}

int pfslam_set_params(pfslam_engine *e, const float *scan_dev, int32_t frame)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    StepParams *slot = nullptr;
    int rc = next_param_slot(e, &slot);
    if (rc) return rc;
    slot->scan = scan_dev ? scan_dev : e->scan; slot->frame = frame; slot->seq = e->seq; slot->scan_src = nullptr; slot->res_host = nullptr;
    CUDA_TRY(cudaMemcpyAsync(e->sp, slot, sizeof(StepParams), cudaMemcpyHostToDevice, e->stream));
    e->cur = *slot; e->cur_valid = true;
    return param_slot_used(e);
}

// ---- shard exchange over peer memory (pf_xchg.cuh) ----------------------------------------------
int pfslam_exchange_region(pfslam_engine *e, void **dev_ptr, int64_t *bytes)
{
    if (!e || !dev_ptr || !bytes) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (!e->xreg) return set_error(PFSLAM_ERR_STATE, "single-rank engines have no exchange region");
    *dev_ptr = e->xreg; *bytes = (int64_t)e->xreg_bytes;
    return PFSLAM_OK;
}

int pfslam_ipc_export(pfslam_engine *e, void *handle_out)
{
    if (!e || !handle_out) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (!e->xreg) return set_error(PFSLAM_ERR_STATE, "single-rank engines have no exchange region");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == PFSLAM_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, e->xreg));
    memcpy(handle_out, &h, sizeof h);
    return PFSLAM_OK;
}

static int set_peer(pfslam_engine *e, int rank, void *base, bool ipc)
{
    if (rank < 0 || rank >= e->n_ranks) return set_error(PFSLAM_ERR_ARG, "rank %d of %d", rank, e->n_ranks);
    if (rank == e->xc_p2p.rank) return PFSLAM_OK;          // own region is already in place
    if (e->peer_ipc[rank] && e->peer_base[rank]) cudaIpcCloseMemHandle(e->peer_base[rank]);
    e->peer_base[rank] = base; e->peer_ipc[rank] = ipc;
    e->p2p_ready = false;
    return PFSLAM_OK;
}

int pfslam_ipc_connect(pfslam_engine *e, int32_t rank, const void *handle)
{
    if (!e || !handle) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (!e->xreg) return set_error(PFSLAM_ERR_STATE, "single-rank engines have no exchange region");
    if (rank == e->xc_p2p.rank) return PFSLAM_OK;
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *base = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    return set_peer(e, rank, base, true);
}

int pfslam_connect_peer(pfslam_engine *e, int32_t rank, void *peer_region)
{
    if (!e || !peer_region) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (!e->xreg) return set_error(PFSLAM_ERR_STATE, "single-rank engines have no exchange region");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    cudaPointerAttributes at;
    CUDA_TRY(cudaPointerGetAttributes(&at, peer_region));
    if (at.type != cudaMemoryTypeDevice) return set_error(PFSLAM_ERR_ARG, "peer region is not device memory");
    if (at.device != e->cfg.device) {
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, e->cfg.device, at.device));
        if (!can) return set_error(PFSLAM_ERR_UNSUPPORTED, "device %d cannot access device %d", e->cfg.device, at.device);
        cudaError_t ce = cudaDeviceEnablePeerAccess(at.device, 0);
        if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled)
            return set_error(PFSLAM_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(ce));
        cudaGetLastError();
    }
    return set_peer(e, rank, peer_region, false);
}

// All peers known: from now on pfslam_step / pfslam_step_async run the sharded step with the
// exchange inside the kernels.  Callers barrier across ranks between this call and the first step.
int pfslam_exchange_ready(pfslam_engine *e)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    if (!e->xreg) return set_error(PFSLAM_ERR_STATE, "single-rank engines have no exchange region");
    for (int r = 0; r < e->n_ranks; r++)
        if (!e->peer_base[r]) return set_error(PFSLAM_ERR_STATE, "rank %d is not connected", r);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < e->n_ranks; r++) {
        e->xc_p2p.peer[r] = static_cast<unsigned char *>(e->peer_base[r]);
        // rank r's snapshot, parity 0: the copy pushed into this rank's own region, or (pull mode) r's own slot over NVLink
        unsigned char *from = e->snap_push ? e->xreg : e->xc_p2p.peer[r];
        e->xc_p2p.pose_src[r] = reinterpret_cast<const float *>(from + e->xoff_snap) + 4ll * e->n * r;
    }
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    if (e->graph) { cudaGraphDestroy(e->graph); e->graph = nullptr; }
    if (e->graph_io_exec) { cudaGraphExecDestroy(e->graph_io_exec); e->graph_io_exec = nullptr; }
    if (e->graph_io) { cudaGraphDestroy(e->graph_io); e->graph_io = nullptr; }
    e->graph_failed = false;
    e->p2p_ready = true;
    return PFSLAM_OK;
}

int pfslam_synchronize(pfslam_engine *e)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

// ------------------------------------------------------------------------------------------------
int pfslam_upload_scan(pfslam_engine *e, const float *scan_host)
{
    if (!e || !scan_host) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    // the pinned staging copy must not be overwritten while an earlier H2D is in flight
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    memcpy(e->h_scan, scan_host, sizeof(float) * e->cfg.n_beams);
    CUDA_TRY(cudaMemcpyAsync(e->scan, e->h_scan, sizeof(float) * e->cfg.n_beams,
                             cudaMemcpyHostToDevice, e->stream));
    return PFSLAM_OK;
}

// peer-memory shards: this rank's snapshot into every peer's region (pf_xchg.cuh)
static int push_snapshot(pfslam_engine *e, cudaStream_t st)
{
    if (!e->cur_xc->parity_mask || !e->snap_push) return PFSLAM_OK;
    k_snapshot_push<<<std::min(ceil_div(e->n, 128), std::max(e->n_sms, 1)), 128, 0, st>>>(*e->cur_xc, e->sp, e->n);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

// k_motion's by-value copy of the step parameters: real values inside a step graph (the launch updates the node's
// argument), "look in device memory" everywhere else (push_params / the host has put them there)
static StepParams motion_params(pfslam_engine *e)
{
    if (e->in_capture) return e->capture_params;
    StepParams p = {};
    p.seq = kParamsInDeviceMemory;
    return p;
}

static int ph_motion(pfslam_engine *e, int32_t frame)
{
    { int rc = push_params(e, e->cur.scan ? e->cur.scan : e->scan, frame); if (rc) return rc; }
    const Xchg &xc = *e->cur_xc;
    // the pre-resample snapshot is written by the same kernel (hosts that all-gather it pass none)
    k_motion<<<ceil_div(e->n, 256), 256, 0, e->stream>>>(e->x, e->y, e->th, e->n, e->sp, e->gidx0, e->twork->bounds,
                                                         xc.snap, xc.snap_stride, xc.parity_mask, xc.snap_aos, e->score_partial,
                                                         e->io_capture ? e->h_scan_dev : nullptr, e->scan, e->cfg.n_beams, e->pcs,
                                                         motion_params(e));
    if (e->laps_on) e->laps.mark(e->stream, kLapMotion);
    e->bounds_valid = true;
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

int pfslam_phase_motion(pfslam_engine *e, int32_t frame)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->cur_xc = &e->xc_host;
    return ph_motion(e, frame);
}

static int score_phase(pfslam_engine *e, const float *scan_dev, cudaEvent_t ev0, cudaEvent_t ev1, const Xchg &xc)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    { int rc = push_params(e, scan_dev ? scan_dev : e->scan, e->cur.frame); if (rc) return rc; }
    const StepParams *scan = e->sp;
    const bool use_aux = e->overlap && !e->laps_on;
    if (e->score_mode == PFSLAM_SCORE_EXACT) {
        if (ev0) cudaEventRecord(ev0, e->stream);
        k_score_exact<<<ceil_div(e->n, 32), 256, 0, e->stream>>>(
            e->grid, e->geom, e->x, e->y, e->th, e->n, e->gidx0, scan, e->angle, e->cfg.n_beams,
            e->fit, e->blk_min, e->blk_maxkey);
        if (ev1) cudaEventRecord(ev1, e->stream);
        e->launches++;
        CUDA_TRY(cudaGetLastError());
        k_extrema<<<1, 1024, 0, e->stream>>>(e->blk_min, e->blk_maxkey, ceil_div(e->n, 32), e->x, e->y,
                                             e->th, e->gidx0, e->ext_local, xc, e->sp);
        e->launches++;
    } else if (e->score_mode == PFSLAM_SCORE_TILED) {
        int nl = score_tiled_launch(e->tmap, e->grid, e->geom, e->x, e->y, e->th, e->n, e->gidx0, scan, e->angle,
                                    e->cfg.n_beams, e->fit, e->blk_min, e->blk_maxkey, e->ext_local, e->fwork,
                                    e->twork, e->angle_cs, e->bounds_valid, e->score_partial, e->counters, xc, e->tiled_grid, e->stream, ev0, ev1,
                                    use_aux ? e->aux : nullptr, e->ev_fork[0], e->ev_join[0], e->laps_on ? &e->laps : nullptr,
                                    e->staged ? staged_kernel() : nullptr, staged_threads(), staged_smem_bytes(),
                                    staged_threads() * staged_ppt(), e->pcs, staged_windows());
        e->bounds_valid = false;   // consumed (and reset) by k_tile_prep
        if (nl < 0) return set_error(PFSLAM_ERR_CUDA, "tiled scoring launch failed: %s",
                                     cudaGetErrorString(cudaGetLastError()));
        e->launches += nl;
    } else {
        int nl = score_filtered_launch(e->grid, e->geom, e->x, e->y, e->th, e->n, e->gidx0, scan, e->angle,
                                       e->cfg.n_beams, e->fit, e->blk_min, e->blk_maxkey, e->ext_local,
                                       e->fwork, e->score_partial, e->counters, xc, e->stream, ev0, ev1);
        if (nl < 0) return set_error(PFSLAM_ERR_CUDA, "filtered scoring launch failed: %s",
                                     cudaGetErrorString(cudaGetLastError()));
        e->launches += nl;
    }
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

static int ph_score(pfslam_engine *e, const float *scan_dev)
{
    if (e->prof_on && e->prof_n < (int)e->prof_ev.size() / 2) {
        const int k = e->prof_n++;
        return score_phase(e, scan_dev, e->prof_ev[2 * k], e->prof_ev[2 * k + 1], *e->cur_xc);
    }
    return score_phase(e, scan_dev, nullptr, nullptr, *e->cur_xc);
}

int pfslam_phase_score(pfslam_engine *e, const float *scan_dev)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->cur_xc = &e->xc_host;
    return ph_score(e, scan_dev);
}

int pfslam_profile_enable(pfslam_engine *e, int32_t on)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    if (on && e->prof_ev.empty()) {
        e->prof_ev.resize(2 * 4096);
        for (auto &ev : e->prof_ev) CUDA_TRY(cudaEventCreate(&ev));
    }
    e->prof_on = on != 0;
    e->prof_n = 0;
    return PFSLAM_OK;
}

int pfslam_profile_read(pfslam_engine *e, float *ms_kernel_mean, int32_t *n_launches)
{
    if (!e || !ms_kernel_mean || !n_launches) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    double tot = 0.0;
    for (int k = 0; k < e->prof_n; k++) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e->prof_ev[2 * k], e->prof_ev[2 * k + 1]));
        tot += ms;
    }
    *n_launches = e->prof_n;
    *ms_kernel_mean = e->prof_n ? (float)(tot / e->prof_n) : 0.f;
    e->prof_n = 0;
    return PFSLAM_OK;
}

int pfslam_profile_score(pfslam_engine *e, float *ms_kernel, float *ms_phase)
{
    if (!e || !ms_kernel || !ms_phase) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; i++) CUDA_TRY(cudaEventCreate(&ev[i]));
    CUDA_TRY(cudaEventRecord(ev[0], e->stream));
    int rc = score_phase(e, nullptr, ev[1], ev[2], e->xc_host);
    if (rc == PFSLAM_OK) {
        cudaEventRecord(ev[3], e->stream);
        cudaError_t ce = cudaStreamSynchronize(e->stream);
        if (ce == cudaSuccess) ce = cudaEventElapsedTime(ms_kernel, ev[1], ev[2]);
        if (ce == cudaSuccess) ce = cudaEventElapsedTime(ms_phase, ev[0], ev[3]);
        if (ce != cudaSuccess) rc = set_error(PFSLAM_ERR_CUDA, "profile_score: %s", cudaGetErrorString(ce));
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
    return rc;
}

static int ph_weights(pfslam_engine *e)
{
    const int n_sync = (e->cfg.quirks & PFSLAM_QUIRK_Q1_HALF_WEIGHT_SYNC) ? (e->n_global + 1) / 2 : e->n_global;
    // grid path: the prefix step rides in the same launch (the kd path needs ICP in between, and a host that
    // runs its own collectives needs the tiles gathered first)
    const bool local_or_peer = e->n_ranks == 1 || e->cur_xc->parity_mask;
    const int fuse = (local_or_peer && e->cfg.path == PFSLAM_PATH_GRID2D && e->n_tiles * e->n_ranks <= kFusedPrefixMaxTiles) ? 1 : 0;
    e->prefix_fused = fuse != 0;
    // dependent launch of the scorer's combine kernel on the grid path (the kd path runs k_extrema before it)
    launch_k(e->cfg.path == PFSLAM_PATH_GRID2D && e->score_mode == PFSLAM_SCORE_TILED, k_weights_scan, dim3(e->n_tiles), dim3(kScanThreads), 0, e->stream,
             *e->cur_xc, e->sp, e->fit, e->w, e->n, e->gidx0, n_sync, e->n_tiles, e->tiles_local, fuse,
             e->n_global, e->prefix, e->res, 1, e->counters + 5);
    if (e->laps_on) e->laps.mark(e->stream, kLapWeights);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

static const char *const kLapNames[kLapCount] = {"k_motion", "k_tile_prep", "k_score_tiled", "k_score_fast", "k_score_combine_rows",
                                                 "k_weights_scan", "k_prefix", "k_map_free", "k_map_wall", "k_resample"};

const char *pfslam_lap_name(int32_t id) { return id >= 0 && id < kLapCount ? kLapNames[id] : ""; }

int pfslam_profile_laps(pfslam_engine *e, int32_t on)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    if (on && e->laps.ev.empty()) {
        e->laps.ev.resize(16 * 1024);
        e->laps.id.assign(e->laps.ev.size(), 0);
        for (auto &ev : e->laps.ev) CUDA_TRY(cudaEventCreate(&ev));
    }
    e->laps_on = on != 0;
    e->laps.n = 0;
    return PFSLAM_OK;
}

int pfslam_profile_laps_read(pfslam_engine *e, float *ms_mean, int32_t *count)
{
    if (!e || !ms_mean || !count) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    double tot[kLapCount] = {};
    for (int i = 0; i < kLapCount; i++) count[i] = 0;
    for (int k = 1; k < e->laps.n; k++) {
        const int id = e->laps.id[k];
        if (id < 0) continue;                       // a step's start marker
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e->laps.ev[k - 1], e->laps.ev[k]));
        tot[id] += ms; count[id]++;
    }
    for (int i = 0; i < kLapCount; i++) ms_mean[i] = count[i] ? (float)(tot[i] / count[i]) : 0.f;
    e->laps.n = 0;
    return PFSLAM_OK;
}

int pfslam_phase_weights(pfslam_engine *e)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->cur_xc = &e->xc_host;
    return ph_weights(e);
}

static int launch_prefix(pfslam_engine *e)
{
    if (e->prefix_fused) { e->prefix_fused = false; return PFSLAM_OK; }   // done by k_weights_scan's last block
    const int nt = e->n_tiles * e->n_ranks;
    k_prefix<<<1, 1024, sizeof(float) * 2 * nt, e->stream>>>(*e->cur_xc, e->sp, e->n_tiles, e->n_global, e->prefix, e->res,
                                                             e->cfg.path == PFSLAM_PATH_KD ? 0 : 1);
    if (e->laps_on) e->laps.mark(e->stream, kLapPrefix);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

static int clear_map_masks(pfslam_engine *e, cudaStream_t st)
{
    CUDA_TRY(cudaMemsetAsync(e->wall_bits, 0, e->bits_bytes, st));      // (the free pass claims by epoch stamp)
    return PFSLAM_OK;
}

// pose_from_ext: robotPos = the best particle's pose read from the exchanged extrema instead of the frame
// result -- lets the map update run next to the weight / resample kernels (it needs nothing else from them)
static int launch_map(pfslam_engine *e, cudaStream_t st, int pose_from_ext)
{
    const StepParams *scan = e->sp;
    if (pose_from_ext && e->cur_xc->parity_mask) {      // this branch does not follow k_weights_scan's wait
        k_xc_wait<<<1, 32, 0, st>>>(*e->cur_xc, e->sp, kXcExt, e->res);
        e->launches++;
    }
    k_map_free<<<e->cfg.n_beams, 128, 0, st>>>(e->grid, e->geom, e->res, scan, e->angle, e->free_stamp,
                                               e->counters, *e->cur_xc, pose_from_ext);
    if (e->laps_on) e->laps.mark(st, kLapMapFree);
    // dependent launch of k_map_free (same stream, directly behind it) unless a lap marker sits in between
    launch_k(!e->laps_on, k_map_wall, dim3(1), dim3(kWallThreads), 0, st, e->grid, e->geom, e->res, scan, e->angle,
             e->cfg.n_beams, e->wall_bits, e->counters, *e->cur_xc, pose_from_ext);
    if (e->laps_on) e->laps.mark(st, kLapMapWall);
    e->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

// In the sharded step the prefix kernel (robotPos, Neff, decision) needs the all-gathered tile
// sums, so it runs at the head of the resample phase; the map update then uses robotPos.  The
// reference's order measurement -> map -> resample is preserved because the map update does not
// read particle weights and the resample does not read the map.
static int ph_map(pfslam_engine *e, const float *scan_dev)
{
    int rc = push_params(e, scan_dev ? scan_dev : e->scan, e->cur.frame);
    if (rc) return rc;
    if ((rc = launch_prefix(e))) return rc;
    if ((rc = clear_map_masks(e, e->stream))) return rc;
    return launch_map(e, e->stream, 0);
}

int pfslam_phase_map(pfslam_engine *e, const float *scan_dev)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->cur_xc = &e->xc_host;
    return ph_map(e, scan_dev);
}

static int ph_resample(pfslam_engine *e, int32_t frame)
{
    { int rc = push_params(e, e->cur.scan ? e->cur.scan : e->scan, frame); if (rc) return rc; }
    // dependent launch of k_weights_scan when that kernel is the one just before (prefix fused into it)
    launch_k(e->resample_follows_weights, k_resample, dim3(ceil_div(e->n, 256)), dim3(256), 0, e->stream,
             *e->cur_xc, e->res, e->prefix, e->n_tiles, e->n, e->n_global, e->gidx0, e->sp, e->x, e->y, e->th, e->w);
    e->resample_follows_weights = false;
    if (e->laps_on) e->laps.mark(e->stream, kLapResample);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

int pfslam_phase_resample(pfslam_engine *e, int32_t frame)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->cur_xc = &e->xc_host;
    return ph_resample(e, frame);
}


// ------------------------------------------------------------------------------------------------
// kd-tree path, host side.  Tree (re)builds run on the host like the reference's KDTree::Create /
// Balance (kdtree.cpp:25-67): the shape depends on std::sort's order among equal keys, and calling the
// same std::sort on the same sequence is what keeps it identical to the reference build.
struct KdPt { float x, y, z, w; };
static bool kd_less_x(const KdPt &a, const KdPt &b) { return a.x < b.x; }
static bool kd_less_y(const KdPt &a, const KdPt &b) { return a.y < b.y; }
static bool kd_less_z(const KdPt &a, const KdPt &b) { return a.z < b.z; }

static void kd_build_range(KdPt *first, KdPt *last, KdNode *list, int idx, int parent)
{
    const int axis = parent < 0 ? 0 : (list[parent].axis + 1) % 3;
    std::sort(first, last, axis == 0 ? kd_less_x : axis == 1 ? kd_less_y : kd_less_z);
    const int size = (int)(last - first), mid = size / 2;
    KdNode &n = list[idx];
    n.axis = axis; n.left = -1; n.right = -1; n.parent = parent;
    n.x = first[mid].x; n.y = first[mid].y; n.z = first[mid].z; n.w = first[mid].w;
    if (mid > 0) { n.left = idx + 1; kd_build_range(first, first + mid, list, idx + 1, idx); }
    if (mid < size - 1) { list[idx].right = idx + mid + 1; kd_build_range(first + mid + 1, last, list, idx + mid + 1, idx); }
}

static void kd_build(std::vector<KdPt> &pts, KdNode *list)
{
    if (pts.empty()) return;
    std::sort(pts.begin(), pts.end(), kd_less_x);          // Create sorts by x, then InsertList sorts again
    kd_build_range(pts.data(), pts.data() + pts.size(), list, 0, -1);
}

// rebuild the scorer's search shadow after a topology change (first build, insert, rebalance, set_kd)
static int kd_refresh_shadow(pfslam_engine *e, int n_nodes_hint)
{
    if (n_nodes_hint > 0) e->kd_size_ub = n_nodes_hint;
    const int cover = std::min(e->kd_size_ub, e->kd_cap);
    if (cover <= 0) return PFSLAM_OK;
    if (e->kd_walk == 2) k_kd_shadow<2><<<ceil_div(cover, 256), 256, 0, e->stream>>>(e->kd, e->ks, e->kds, e->kd_cap);
    else k_kd_shadow<1><<<ceil_div(cover, 256), 256, 0, e->stream>>>(e->kd, e->ks, e->kds, e->kd_cap);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

static void launch_score_kd(pfslam_engine *e)
{
    const dim3 grid(ceil_div(e->n, 32)), block(256);
#define PF_SCORE_KD(F) k_score_kd<F><<<grid, block, 0, e->stream>>>(e->kd, e->kds, e->x, e->y, e->th, e->n, e->gidx0, e->sp, e->angle, \
                                                                    e->cfg.n_beams, e->fit, e->blk_min, e->blk_maxkey)
    if (!e->kd_flat) PF_SCORE_KD(0);
    else if (e->kd_walk == 2) PF_SCORE_KD(2);
    else PF_SCORE_KD(1);
#undef PF_SCORE_KD
    e->launches++;
}

// frame % 100 == 5 (kernel.cu:1707-1711): device tree -> host rebuild -> device
static int kd_balance(pfslam_engine *e)
{
    KdState st;
    CUDA_TRY(cudaMemcpyAsync(&st, e->ks, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (st.size <= 0) return PFSLAM_OK;
    e->h_kd.resize(st.size);
    CUDA_TRY(cudaMemcpyAsync(e->h_kd.data(), e->kd, sizeof(KdNode) * st.size, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    std::vector<KdPt> pts(st.size);
    for (int i = 0; i < st.size; i++) { pts[i].x = e->h_kd[i].x; pts[i].y = e->h_kd[i].y; pts[i].z = e->h_kd[i].z; pts[i].w = e->h_kd[i].w; }
    kd_build(pts, e->h_kd.data());
    CUDA_TRY(cudaMemcpyAsync(e->kd, e->h_kd.data(), sizeof(KdNode) * st.size, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return kd_refresh_shadow(e, st.size);
}

// PFUpdateMapKD (kernel.cu:1406-1540)
static int kd_update_map(pfslam_engine *e)
{
    const int n_words = (int)(e->bits_bytes / 4);
    CUDA_TRY(cudaMemsetAsync(e->free_bits, 0, e->bits_bytes, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->wall_bits, 0, e->bits_bytes, e->stream));
    k_kd_mark<<<e->cfg.n_beams, 128, 0, e->stream>>>(e->geom, e->res, e->sp, e->angle, e->free_bits, e->wall_bits);
    k_bits_count<<<e->n_bits_blk, 256, 0, e->stream>>>(e->free_bits, e->wall_bits, n_words, e->bits_blk);
    k_bits_offsets<<<1, 32, 0, e->stream>>>(e->bits_blk, e->n_bits_blk, e->ks);
    k_bits_scatter<<<e->n_bits_blk, 256, 0, e->stream>>>(e->free_bits, e->wall_bits, n_words, e->bits_blk,
                                                         e->free_cells, e->wall_cells, e->pc_cap);
    k_kd_points_nn<<<ceil_div(2 * e->pc_cap, 128), 128, 0, e->stream>>>(e->kd, e->geom, e->res, e->ks, e->wall_cells,
                                                                       e->free_cells, e->pc_cap, e->kd_pts, e->kd_nn_idx);
    e->launches += 5;
    if (e->kd_empty) {
        // first scan: build the tree from the wall points on the host (kernel.cu:1532-1536)
        KdState st;
        CUDA_TRY(cudaMemcpyAsync(&st, e->ks, sizeof st, cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        const int nW = std::min(std::min(st.n_wall, e->pc_cap), e->kd_cap);
        if (nW > 0) {
            std::vector<float2> p(nW);
            CUDA_TRY(cudaMemcpy(p.data(), e->kd_pts, sizeof(float2) * nW, cudaMemcpyDeviceToHost));
            std::vector<KdPt> pts(nW);
            for (int i = 0; i < nW; i++) { pts[i].x = p[i].x; pts[i].y = p[i].y; pts[i].z = 0.0f; pts[i].w = 0.0f; }   // Q12: w = 0
            e->h_kd.assign(nW, KdNode());
            kd_build(pts, e->h_kd.data());
            CUDA_TRY(cudaMemcpy(e->kd, e->h_kd.data(), sizeof(KdNode) * nW, cudaMemcpyHostToDevice));
            st.size = nW; st.n_ins = nW;
            CUDA_TRY(cudaMemcpy(e->ks, &st, sizeof st, cudaMemcpyHostToDevice));
            e->kd_empty = false;
            int rc = kd_refresh_shadow(e, nW);
            if (rc) return rc;
        }
    } else {
        // claim stamps grow monotonically; node indices move at a rebuild, where stale stamps stay below every new one
        k_kd_weights<<<ceil_div(e->pc_cap, 128), 128, 0, e->stream>>>(e->kd, e->geom, e->ks, e->pc_cap, 0, e->kd_pts, e->kd_nn_idx, e->kd_claim, ++e->kd_stamp);
        k_kd_weights<<<ceil_div(e->pc_cap, 128), 128, 0, e->stream>>>(e->kd, e->geom, e->ks, e->pc_cap, 1, e->kd_pts, e->kd_nn_idx, e->kd_claim, ++e->kd_stamp);
        k_kd_insert<<<1, 1024, 0, e->stream>>>(e->kd, e->geom, e->ks, e->pc_cap, e->kd_cap, e->kd_pts, e->kd_nn_idx, e->kd_ins_index);
        e->launches += 3;
        // the tree size lives on the device: cover an upper bound (at most pc_cap inserts per frame)
        e->kd_size_ub = std::min(e->kd_cap, e->kd_size_ub + e->pc_cap);
        { int rc = kd_refresh_shadow(e, 0); if (rc) return rc; }
    }
    k_kd_finish<<<1, 1, 0, e->stream>>>(e->res, e->ks, e->counters, e->pc_cap, e->kd_cap);
    e->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return PFSLAM_OK;
}

// particleFilter, kd variant (kernel.cu:1702-1761)
static int kd_step(pfslam_engine *e, const float *scan_dev, int32_t frame)
{
    int rc = push_params(e, scan_dev ? scan_dev : e->scan, frame);
    if (rc) return rc;
    if (frame % 100 == 5 && !e->kd_empty && (rc = kd_balance(e))) return rc;
    if (e->kd_empty) {
        CUDA_TRY(cudaMemsetAsync(e->res, 0, sizeof(FrameResult), e->stream));      // robotPos = 0 (kernel.cu:1715)
        return kd_update_map(e);
    }
    if ((rc = ph_motion(e, frame))) return rc;
    if ((rc = push_snapshot(e, e->stream))) return rc;
    const bool prof = e->prof_on && e->prof_n < (int)e->prof_ev.size() / 2;
    if (prof) cudaEventRecord(e->prof_ev[2 * e->prof_n], e->stream);
    launch_score_kd(e);
    if (prof) { cudaEventRecord(e->prof_ev[2 * e->prof_n + 1], e->stream); e->prof_n++; }
    k_extrema<<<1, 1024, 0, e->stream>>>(e->blk_min, e->blk_maxkey, ceil_div(e->n, 32), e->x, e->y, e->th, e->gidx0,
                                         e->ext_local, *e->cur_xc, e->sp);
    e->launches += 1;
    e->bounds_valid = false;
    if ((rc = ph_weights(e))) return rc;
    k_icp<<<1, 1024, sizeof(float) * 5 * e->cfg.n_beams, e->stream>>>(e->kd, *e->cur_xc, e->sp, e->angle,
                                                                    e->cfg.n_beams, e->res);
    e->launches += 1;
    if ((rc = launch_prefix(e))) return rc;
    if ((rc = kd_update_map(e))) return rc;
    return ph_resample(e, frame);
}

// PFSLAM_TAIL=fused: one launch for weights / prefix / resample (k_weights_resample, a grid-wide barrier in the middle).
// It needs every one of its n_tiles blocks resident at the same time: allowed while that is at most two 256-thread
// blocks per SM -- a quarter of the thread slots -- so the map branch's kernels next to it can never lock it out.
// Not the default: with one block per tile the resampler's random probes run on n_tiles SMs only, and bench.py
// measures 110.9 us per step against 100.3 us for the three-kernel sequence.
static bool tail_fusable(pfslam_engine *e)
{
    const bool local_or_peer = e->n_ranks == 1 || e->cur_xc->parity_mask;
    return e->tail_fused && local_or_peer && e->cfg.path == PFSLAM_PATH_GRID2D &&
           e->n_tiles <= 2 * e->n_sms && e->n_tiles * e->n_ranks <= kFusedPrefixMaxTiles;
}

static int run_phases(pfslam_engine *e, const float *scan_dev, int32_t frame)
{
    int rc;
    if (e->laps_on) e->laps.mark(e->stream, kLapStart);
    if ((rc = ph_motion(e, frame))) return rc;
    if (!(e->overlap && !e->laps_on)) {
        if ((rc = push_snapshot(e, e->stream))) return rc;
        if ((rc = ph_score(e, scan_dev))) return rc;
        if ((rc = ph_weights(e))) return rc;
        if ((rc = ph_map(e, scan_dev))) return rc;
        return ph_resample(e, frame);
    }
    // Two branches after the scores are in: {weights + tile scans, prefix / Neff, resample} on the main
    // stream and {mask clears, free-cell and wall-cell map update} on the auxiliary one.  The map update
    // needs only robotPos (the best particle's pose, taken from the extrema), the resample never reads the
    // map, so the reference's order measurement -> map -> resample is preserved in effect.
    CUDA_TRY(cudaEventRecord(e->ev_fork[1], e->stream));
    CUDA_TRY(cudaStreamWaitEvent(e->aux, e->ev_fork[1], 0));
    if ((rc = clear_map_masks(e, e->aux))) return rc;      // off the critical path: overlaps the scoring
    if ((rc = push_snapshot(e, e->aux))) return rc;        // ... and so does the snapshot exchange (joined before the combine kernel)
    if ((rc = ph_score(e, scan_dev))) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_fork[2], e->stream));
    CUDA_TRY(cudaStreamWaitEvent(e->aux, e->ev_fork[2], 0));
    if ((rc = launch_map(e, e->aux, 1))) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_join[1], e->aux));
    if (tail_fusable(e)) {
        // weights + tile scans + prefix + resample as one launch with a grid-wide barrier (k_weights_resample)
        const int n_sync = (e->cfg.quirks & PFSLAM_QUIRK_Q1_HALF_WEIGHT_SYNC) ? (e->n_global + 1) / 2 : e->n_global;
        launch_k(e->score_mode == PFSLAM_SCORE_TILED, k_weights_resample, dim3(e->n_tiles), dim3(kScanThreads), 0, e->stream,
                 *e->cur_xc, e->sp, e->fit, e->w, e->n, e->gidx0, n_sync, e->n_tiles, e->tiles_local, e->n_global, e->prefix,
                 e->res, reinterpret_cast<unsigned long long *>(e->counters + 8), e->x, e->y, e->th);
        e->launches++;
        CUDA_TRY(cudaGetLastError());
    } else {
        if ((rc = ph_weights(e))) return rc;
        e->resample_follows_weights = e->prefix_fused;
        if ((rc = launch_prefix(e))) return rc;
        if ((rc = ph_resample(e, frame))) return rc;
    }
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join[1], 0));
    return PFSLAM_OK;
}

// Capture the whole single-GPU step once; every later step is one cudaGraphLaunch, with that step's {scan pointer,
// frame, sequence number, I/O slots} set as the last argument of the graph's first kernel node (k_motion) beforehand.
static int build_graph(pfslam_engine *e, bool with_io)
{
    const long long launches_before = e->launches;
    if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return -1; }
    e->capture_params = StepParams{};                    // placeholder: every launch sets the node's argument
    e->capture_params.scan = e->scan;
    e->in_capture = true; e->io_capture = with_io;       // with_io: k_motion pulls the scan from the mapped staging buffer
    int rc = run_phases(e, nullptr, 0);
    e->in_capture = false; e->io_capture = false;
    if (with_io) k_publish_result<<<1, 32, 0, e->stream>>>(e->res, e->sp);
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    e->launches = launches_before;
    if (rc != PFSLAM_OK || ce != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return -1; }
    size_t nn = 0;
    cudaGraphGetNodes(g, nullptr, &nn);
    std::vector<cudaGraphNode_t> nodes(nn);
    cudaGraphGetNodes(g, nodes.data(), &nn);
    cudaGraphNode_t pn = nullptr;
    int n_kernels = 0;
    for (auto nd : nodes) {
        cudaGraphNodeType ty;
        cudaGraphNodeGetType(nd, &ty);
        if (ty == cudaGraphNodeTypeKernel) {
            n_kernels++;
            cudaKernelNodeParams kp;
            if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && kp.func == (void *)k_motion && kp.kernelParams) pn = nd;
        }
    }
    if (!pn) { cudaGraphDestroy(g); return -1; }
    cudaGraphExec_t ge = nullptr;
    if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(g); return -1; }
    if (with_io) { e->graph_io = g; e->graph_io_exec = ge; e->graph_io_param_node = pn; e->graph_io_kernels = n_kernels; }
    else { e->graph = g; e->graph_exec = ge; e->graph_param_node = pn; e->graph_kernels = n_kernels; }
    return 0;
}

static bool graph_usable(const pfslam_engine *e)
{
    return e->use_graph && !e->prof_on && !e->laps_on && !e->graph_failed && e->cfg.path == PFSLAM_PATH_GRID2D;
}

// one replay of a captured step with this frame's parameters in its k_motion node; io_slot >= 0: the slot of the pinned
// scan / result rings the step pulls its scan from and publishes its result into
static int launch_graph(pfslam_engine *e, cudaGraphExec_t ge, cudaGraphNode_t pn, const float *scan, int32_t frame, int io_slot = -1)
{
    StepParams p;
    p.scan = scan; p.frame = frame; p.seq = e->seq;
    p.scan_src = io_slot >= 0 ? e->h_scan_dev + (size_t)io_slot * e->cfg.n_beams : nullptr;
    p.res_host = io_slot >= 0 ? e->h_res_dev + io_slot : nullptr;
    // the parameters are the last argument of the graph's k_motion node (copied by the call: nothing to keep alive)
    cudaKernelNodeParams kp;
    CUDA_TRY(cudaGraphKernelNodeGetParams(pn, &kp));
    void *args[kMotionArgs];
    for (int i = 0; i < kMotionArgs - 1; i++) args[i] = kp.kernelParams[i];
    args[kMotionArgs - 1] = &p;
    kp.kernelParams = args;
    CUDA_TRY(cudaGraphExecKernelNodeSetParams(ge, pn, &kp));
    CUDA_TRY(cudaGraphLaunch(ge, e->stream));
    e->cur = p; e->cur_valid = true;
    e->launches += ge == e->graph_io_exec ? e->graph_io_kernels : e->graph_kernels;
    return PFSLAM_OK;
}

int pfslam_step_async(pfslam_engine *e, const float *scan_dev, int32_t frame)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    if (e->n_ranks != 1 && !e->p2p_ready)
        return set_error(PFSLAM_ERR_STATE, "sharded engine without a peer exchange (pfslam_exchange_ready): "
                                           "step it phase by phase from a host that runs the collectives");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    e->cur_xc = e->p2p_ready ? &e->xc_p2p : &e->xc_host;
    e->seq++;
    if (e->cfg.path == PFSLAM_PATH_KD) return kd_step(e, scan_dev, frame);
    const float *scan = scan_dev ? scan_dev : e->scan;
    if (graph_usable(e)) {
        if (!e->graph_exec && build_graph(e, false) != 0) e->graph_failed = true;
        if (e->graph_exec) return launch_graph(e, e->graph_exec, e->graph_param_node, scan, frame);
    }
    // plain launches (profiling, or graph capture unavailable on this stream)
    int rc = push_params(e, scan, frame);
    if (rc) return rc;
    return run_phases(e, scan_dev, frame);
}

static void copy_result(const FrameResult *r, pfslam_frame_result *out)
{
    out->pose[0] = r->pose[0]; out->pose[1] = r->pose[1]; out->pose[2] = r->pose[2];
    out->fit_min = r->fit_min; out->fit_max = r->fit_max; out->best_index = r->best_index;
    out->sum_w = r->sum_w; out->sum_w2 = r->sum_w2; out->neff = r->neff;
    out->resampled = r->resampled; out->n_free_cells = r->n_free; out->n_wall_cells = r->n_wall;
    out->n_slow_evals = r->n_slow;
    out->kd_size = r->kd_size; out->kd_inserted = r->kd_ins;
    out->exchange_timeout = r->xchg_timeout;
    out->resample_count = r->resample_count;
    out->wait_extrema_ns = r->wait_ext_ns; out->wait_tiles_ns = r->wait_tiles_ns;
    out->n_windows = r->n_windows; out->n_wide_beams = r->n_wide;
}

static int finish_result(pfslam_engine *e, const FrameResult *r, pfslam_frame_result *out);

int pfslam_fetch_result(pfslam_engine *e, pfslam_frame_result *out)
{
    if (!e || !out) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(e->h_res, e->res, sizeof(FrameResult), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return finish_result(e, e->h_res, out);
}

static int finish_result(pfslam_engine *e, const FrameResult *r, pfslam_frame_result *out)
{
    pfslam_frame_result tmp;
    copy_result(r, out ? out : &tmp);
    if (r->xchg_timeout)
        return set_error(PFSLAM_ERR_STATE, "peer exchange timed out (a shard stopped stepping, or the shards are out of step)");
    if (r->kd_overflow & 1)
        return set_error(PFSLAM_ERR_STATE, "kd map update: a scan produced more wall points than the point lists hold (%d)", e->pc_cap);
    if (r->kd_overflow & 2)
        return set_error(PFSLAM_ERR_STATE, "kd tree full: kd_capacity = %d nodes", e->kd_cap);
    return PFSLAM_OK;
}

static bool io_graph_ready(pfslam_engine *e)
{
    if (!(graph_usable(e) && (e->n_ranks == 1 || e->p2p_ready))) return false;
    e->cur_xc = e->p2p_ready ? &e->xc_p2p : &e->xc_host;      // before the capture: the kernels take it by value
    if (!e->graph_io_exec && build_graph(e, true) != 0) e->graph_failed = true;
    return e->graph_io_exec != nullptr;
}

// Streaming input (the reference reads lidar->scans[frame] synchronously, main.cpp:199-206): the frame's scan goes into
// a slot of a pinned, device-mapped ring, the step is ONE graph launch that pulls it from there and publishes its
// result into the slot's result record; nothing blocks.  Up to PFSLAM_RING_DEPTH frames may be in flight, so a 40 Hz
// producer never waits for the device and the device never waits for the host.
int pfslam_submit(pfslam_engine *e, const float *scan, int32_t frame, int32_t *ticket)
{
    if (!e || !scan || !ticket) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    if (!io_graph_ready(e))
        return set_error(PFSLAM_ERR_UNSUPPORTED, "streaming needs the captured 2D step (a non-default stream, the grid path, no profiling)");
    const int t = e->next_ticket, sl = 1 + (t - 1) % PFSLAM_RING_DEPTH;
    if (e->ring_ticket[sl - 1] != 0)
        return set_error(PFSLAM_ERR_STATE, "scan ring full: %d frames in flight, pfslam_wait the oldest first", PFSLAM_RING_DEPTH);
    memcpy(e->h_scan + (size_t)sl * e->cfg.n_beams, scan, sizeof(float) * e->cfg.n_beams);
    e->seq++;
    int rc = launch_graph(e, e->graph_io_exec, e->graph_io_param_node, e->scan, frame, sl);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e->ring_ev[sl - 1], e->stream));
    e->ring_ticket[sl - 1] = t;
    e->next_ticket++;
    *ticket = t;
    return PFSLAM_OK;
}

int pfslam_wait(pfslam_engine *e, int32_t ticket, pfslam_frame_result *out)
{
    if (!e || ticket <= 0) return set_error(PFSLAM_ERR_ARG, "bad argument");
    const int sl = 1 + (ticket - 1) % PFSLAM_RING_DEPTH;
    if (e->ring_ticket[sl - 1] != ticket) return set_error(PFSLAM_ERR_STATE, "ticket %d is not in flight", ticket);
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaEventSynchronize(e->ring_ev[sl - 1]));
    e->ring_ticket[sl - 1] = 0;
    return finish_result(e, e->h_res + sl, out);
}

int pfslam_step(pfslam_engine *e, const float *scan, int32_t frame, pfslam_frame_result *out)
{
    int rc;
    if (!e || !scan) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    bool ring_idle = true;
    for (int t : e->ring_ticket) ring_idle = ring_idle && t == 0;
    if (ring_idle && io_graph_ready(e)) {
        // host scan in, host result out: one graph launch (scan pull and result publish inside) and one synchronisation
        // (earlier asynchronous work -- pfslam_step_async, phase calls -- may still be reading staging slot 0)
        if (cudaStreamQuery(e->stream) != cudaSuccess) { cudaGetLastError(); CUDA_TRY(cudaStreamSynchronize(e->stream)); }
        memcpy(e->h_scan, scan, sizeof(float) * e->cfg.n_beams);
        e->seq++;
        if ((rc = launch_graph(e, e->graph_io_exec, e->graph_io_param_node, e->scan, frame, 0))) return rc;
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        return finish_result(e, e->h_res, out);
    }
    pfslam_frame_result tmp;
    if ((rc = pfslam_upload_scan(e, scan))) return rc;
    if ((rc = pfslam_step_async(e, nullptr, frame))) return rc;
    if ((rc = pfslam_fetch_result(e, out ? out : &tmp))) return rc;
    return PFSLAM_OK;
}

int particleFilterStep(pfslam_engine *e, const float *scan, int32_t frame, float pose_out[3])
{
    pfslam_frame_result r;
    int rc = pfslam_step(e, scan, frame, &r);
    if (rc) return rc;
    if (pose_out) { pose_out[0] = r.pose[0]; pose_out[1] = r.pose[1]; pose_out[2] = r.pose[2]; }
    return PFSLAM_OK;
}

// ------------------------------------------------------------------------------------------------
int pfslam_update_grid(pfslam_engine *e, const float *scan_host, const float pose[3])
{
    if (!e || !scan_host || !pose) return set_error(PFSLAM_ERR_ARG, "null argument");
    int rc = pfslam_upload_scan(e, scan_host);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    memset(e->h_res, 0, sizeof(FrameResult));
    e->h_res->pose[0] = pose[0]; e->h_res->pose[1] = pose[1]; e->h_res->pose[2] = pose[2];
    CUDA_TRY(cudaMemcpyAsync(e->res, e->h_res, sizeof(FrameResult), cudaMemcpyHostToDevice, e->stream));
    if ((rc = push_params(e, e->scan, e->cur.frame))) return rc;
    e->cur_xc = &e->xc_host;
    if (e->cfg.path == PFSLAM_PATH_KD) {             // PFUpdateMapKD for an explicit robotPos
        if ((rc = kd_update_map(e))) return rc;
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        return PFSLAM_OK;
    }
    if ((rc = clear_map_masks(e, e->stream))) return rc;
    if ((rc = launch_map(e, e->stream, 0))) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_score_particles(pfslam_engine *e, const float *scan_host, int32_t *fit_out)
{
    if (!e || !scan_host || !fit_out) return set_error(PFSLAM_ERR_ARG, "null argument");
    int rc = pfslam_upload_scan(e, scan_host);
    if (rc) return rc;
    if (e->cfg.path == PFSLAM_PATH_KD) {
        if (e->kd_empty) return set_error(PFSLAM_ERR_STATE, "no kd tree yet");
        if ((rc = push_params(e, e->scan, e->cur.frame))) return rc;
        launch_score_kd(e);
    } else if ((rc = pfslam_phase_score(e, nullptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(fit_out, e->fit, sizeof(int) * e->n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_get_particles(pfslam_engine *e, float *x, float *y, float *theta, float *w)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    const size_t b = sizeof(float) * e->n;
    if (x) CUDA_TRY(cudaMemcpyAsync(x, e->x, b, cudaMemcpyDeviceToHost, e->stream));
    if (y) CUDA_TRY(cudaMemcpyAsync(y, e->y, b, cudaMemcpyDeviceToHost, e->stream));
    if (theta) CUDA_TRY(cudaMemcpyAsync(theta, e->th, b, cudaMemcpyDeviceToHost, e->stream));
    if (w) CUDA_TRY(cudaMemcpyAsync(w, e->w, b, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_set_particles(pfslam_engine *e, const float *x, const float *y, const float *theta, const float *w)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    e->bounds_valid = false;
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    const size_t b = sizeof(float) * e->n;
    if (x) CUDA_TRY(cudaMemcpyAsync(e->x, x, b, cudaMemcpyHostToDevice, e->stream));
    if (y) CUDA_TRY(cudaMemcpyAsync(e->y, y, b, cudaMemcpyHostToDevice, e->stream));
    if (theta) CUDA_TRY(cudaMemcpyAsync(e->th, theta, b, cudaMemcpyHostToDevice, e->stream));
    if (w) CUDA_TRY(cudaMemcpyAsync(e->w, w, b, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_get_grid(pfslam_engine *e, int8_t *grid_out)
{
    if (!e || !grid_out) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(grid_out, e->grid, (size_t)e->geom.w * e->geom.h, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_set_grid(pfslam_engine *e, const int8_t *grid_in)
{
    if (!e || !grid_in) return set_error(PFSLAM_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(e->grid, grid_in, (size_t)e->geom.w * e->geom.h, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

int pfslam_get_map_dim(pfslam_engine *e, int32_t *map_w, int32_t *map_h)
{
    if (!e) return set_error(PFSLAM_ERR_ARG, "null engine");
    if (map_w) *map_w = e->geom.w;
    if (map_h) *map_h = e->geom.h;
    return PFSLAM_OK;
}

int pfslam_get_pose(pfslam_engine *e, float pose[3])
{
    pfslam_frame_result r;
    int rc = pfslam_fetch_result(e, &r);
    if (rc) return rc;
    pose[0] = r.pose[0]; pose[1] = r.pose[1]; pose[2] = r.pose[2];
    return PFSLAM_OK;
}

int pfslam_device_buffer(pfslam_engine *e, int32_t which, void **dev_ptr, int64_t *bytes)
{
    if (!e || !dev_ptr || !bytes) return set_error(PFSLAM_ERR_ARG, "null argument");
    switch (which) {
    case PFSLAM_BUF_EXTREMA_LOCAL: *dev_ptr = e->ext_local; *bytes = sizeof(Extrema); break;
    case PFSLAM_BUF_EXTREMA_ALL: *dev_ptr = e->ext_all; *bytes = (int64_t)sizeof(Extrema) * e->n_ranks; break;
    case PFSLAM_BUF_TILES_LOCAL: *dev_ptr = e->tiles_local; *bytes = 4 * e->tiles_block; break;
    case PFSLAM_BUF_TILES_ALL: *dev_ptr = e->tiles_all; *bytes = 4 * e->tiles_block * e->n_ranks; break;
    case PFSLAM_BUF_POSE_LOCAL: *dev_ptr = e->x; *bytes = 12ll * e->n; break;
    case PFSLAM_BUF_POSE_ALL: *dev_ptr = e->pose_all; *bytes = 12ll * e->n * e->n_ranks; break;
    case PFSLAM_BUF_SCAN: *dev_ptr = e->scan; *bytes = 4ll * e->cfg.n_beams; break;
    default: return set_error(PFSLAM_ERR_ARG, "unknown buffer %d", which);
    }
    return PFSLAM_OK;
}

int pfslam_kd_nn(pfslam_engine *e, const float *q_xyz, int32_t n, int32_t *idx_out)
{
    if (!e || !q_xyz || !idx_out || n <= 0) return set_error(PFSLAM_ERR_ARG, "bad argument");
    if (e->cfg.path != PFSLAM_PATH_KD) return set_error(PFSLAM_ERR_STATE, "engine was not created with PFSLAM_PATH_KD");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    if (n > e->kd_q_cap) {
        cudaFree(e->kd_q); cudaFree(e->kd_qi); e->kd_q = nullptr; e->kd_qi = nullptr; e->kd_q_cap = 0;
        const int cap = std::max(n, 4096);
        CUDA_TRY(cudaMalloc(&e->kd_q, sizeof(float) * 3 * (size_t)cap));
        CUDA_TRY(cudaMalloc(&e->kd_qi, sizeof(int) * (size_t)cap));
        e->kd_q_cap = cap;
    }
    CUDA_TRY(cudaMemcpyAsync(e->kd_q, q_xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, e->stream));
    k_kd_nn<<<ceil_div(n, 128), 128, 0, e->stream>>>(e->kd, e->ks, e->kd_q, n, e->kd_qi);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(idx_out, e->kd_qi, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PFSLAM_OK;
}

// measurement hook (bench.py's kd roofline line): mean number of tree nodes one NN walk of the scorer loads, measured
// over the first n_sample particles x all in-range beams of the engine's current scan
int pfslam_kd_mean_visits(pfslam_engine *e, int32_t n_sample, double *mean_visits)
{
    if (!e || !mean_visits) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (e->cfg.path != PFSLAM_PATH_KD || e->kd_empty) return set_error(PFSLAM_ERR_STATE, "no kd tree");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    n_sample = std::max(1, std::min(n_sample, e->n));
    if (!e->kd_cnt) CUDA_TRY(cudaMalloc(&e->kd_cnt, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(e->kd_cnt, 0, 2 * sizeof(unsigned long long), e->stream));
    k_kd_count_visits<<<ceil_div(n_sample, 32), 256, 0, e->stream>>>(e->kd, e->x, e->y, e->th, n_sample, e->sp, e->angle, e->cfg.n_beams, e->kd_cnt);
    e->launches++;
    unsigned long long h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, e->kd_cnt, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    *mean_visits = h[1] ? (double)h[0] / (double)h[1] : 0.0;
    return PFSLAM_OK;
}

int pfslam_kd_icp(pfslam_engine *e, const float *scan_host, const float robot_prev[3], const float start[3], float pose_out[3])
{
    if (!e || !scan_host || !robot_prev || !start || !pose_out) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (e->cfg.path != PFSLAM_PATH_KD) return set_error(PFSLAM_ERR_STATE, "engine was not created with PFSLAM_PATH_KD");
    if (e->kd_empty) return set_error(PFSLAM_ERR_STATE, "no kd tree yet");
    int rc = pfslam_upload_scan(e, scan_host);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    memset(e->h_res, 0, sizeof(FrameResult));
    e->h_res->pose[0] = robot_prev[0]; e->h_res->pose[1] = robot_prev[1]; e->h_res->pose[2] = robot_prev[2];
    CUDA_TRY(cudaMemcpyAsync(e->res, e->h_res, sizeof(FrameResult), cudaMemcpyHostToDevice, e->stream));
    Extrema ex; memset(&ex, 0, sizeof ex);
    ex.best_gidx = e->gidx0; ex.x = start[0]; ex.y = start[1]; ex.th = start[2];
    CUDA_TRY(cudaMemcpyAsync(e->ext_local, &ex, sizeof ex, cudaMemcpyHostToDevice, e->stream));
    if ((rc = push_params(e, e->scan, e->cur.frame))) return rc;
    Xchg one = e->xc_host;                           // the start pose is read from a one-rank extrema record
    one.n_ranks = 1; one.rank = 0; one.ext_all = e->ext_local;
    k_icp<<<1, 1024, sizeof(float) * 5 * e->cfg.n_beams, e->stream>>>(e->kd, one, e->sp, e->angle, e->cfg.n_beams, e->res);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    pfslam_frame_result r;
    if ((rc = pfslam_fetch_result(e, &r))) return rc;
    pose_out[0] = r.pose[0]; pose_out[1] = r.pose[1]; pose_out[2] = r.pose[2];
    return PFSLAM_OK;
}

int pfslam_get_kd(pfslam_engine *e, void *nodes_out, int32_t cap, int32_t *n_nodes)
{
    if (!e || !n_nodes) return set_error(PFSLAM_ERR_ARG, "null argument");
    if (e->cfg.path != PFSLAM_PATH_KD) { *n_nodes = 0; return PFSLAM_OK; }
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    KdState st;
    CUDA_TRY(cudaMemcpyAsync(&st, e->ks, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    *n_nodes = st.size;
    if (nodes_out && cap > 0 && st.size > 0) {
        CUDA_TRY(cudaMemcpyAsync(nodes_out, e->kd, sizeof(KdNode) * (size_t)std::min(cap, st.size), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return PFSLAM_OK;
}

int pfslam_set_kd(pfslam_engine *e, const void *nodes_in, int32_t n_nodes)
{
    if (!e || !nodes_in || n_nodes <= 0) return set_error(PFSLAM_ERR_ARG, "bad argument");
    if (e->cfg.path != PFSLAM_PATH_KD) return set_error(PFSLAM_ERR_STATE, "engine was not created with PFSLAM_PATH_KD");
    if (n_nodes > e->kd_cap) return set_error(PFSLAM_ERR_ARG, "tree larger than kd_capacity");
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    KdState st; memset(&st, 0, sizeof st); st.size = n_nodes;
    CUDA_TRY(cudaMemcpyAsync(e->kd, nodes_in, sizeof(KdNode) * (size_t)n_nodes, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->ks, &st, sizeof st, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->kd_empty = false;
    // the scorer's shadow walk assumes a planar tree: z == 0 everywhere, axes in {0, 1, 2}
    const KdNode *nd = static_cast<const KdNode *>(nodes_in);
    e->kd_flat = true;
    for (int i = 0; i < n_nodes && e->kd_flat; i++)
        e->kd_flat = nd[i].z == 0.0f && nd[i].axis >= 0 && nd[i].axis <= 2;
    return kd_refresh_shadow(e, n_nodes);
}

int64_t pfslam_launch_count(pfslam_engine *e) { return e ? e->launches : 0; }

static const char *const kTraceNames[kTrCount] = {"k_motion", "k_tile_prep", "k_score_tiled", "k_score_fast", "k_score_combine_rows",
                                                  "k_weights_scan", "k_resample", "k_map_free", "k_map_wall", "k_publish_result",
                                                  "mark0", "mark1", "mark2", "mark3", "mark4", "mark5"};
const char *pfslam_trace_name(int32_t id) { return id >= 0 && id < kTrCount ? kTraceNames[id] : ""; }

// on != 0: switch the in-kernel timeline on and reset it; out (optional, 2 * PFSLAM_TRACE_COUNT words): the [first
// entry, last exit] %globaltimer pairs recorded since the last reset.  The caller synchronises around it.
int pfslam_debug_trace(int32_t on, uint64_t *out)
{
    unsigned long long tab[2 * 16];
    if (out) {
        CUDA_TRY(cudaMemcpyFromSymbol(tab, g_trace, sizeof tab));
        memcpy(out, tab, sizeof(uint64_t) * 2 * kTrCount);
    }
    for (int i = 0; i < 16; i++) { tab[2 * i] = ~0ull; tab[2 * i + 1] = 0ull; }
    CUDA_TRY(cudaMemcpyToSymbol(g_trace, tab, sizeof tab));
    const int v = on ? 1 : 0;
    CUDA_TRY(cudaMemcpyToSymbol(g_trace_on, &v, sizeof v));
    return PFSLAM_OK;
}

int pfslam_debug_staged_timing(uint64_t *out, int32_t n_words)
{
    if (!out || n_words <= 0 || n_words > 256 * 12) return set_error(PFSLAM_ERR_ARG, "bad argument");
    CUDA_TRY(cudaMemcpyFromSymbol(out, g_staged_ts, sizeof(uint64_t) * (size_t)n_words));
    return PFSLAM_OK;
}

int pfslam_debug_trig(int32_t device, const float *x_host, int64_t n, float *cos_out, float *sin_out)
{
    if (!x_host || !cos_out || !sin_out || n <= 0) return set_error(PFSLAM_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(device));
    float *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(float) * 3 * n));
    cudaError_t ce = cudaMemcpy(d, x_host, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        k_debug_trig<<<(unsigned)((n + 255) / 256), 256>>>(d, n, d + n, d + 2 * n);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(cos_out, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(sin_out, d + 2 * n, sizeof(float) * n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (ce != cudaSuccess) return set_error(PFSLAM_ERR_CUDA, "debug_trig: %s", cudaGetErrorString(ce));
    return PFSLAM_OK;
}

}  // extern "C"
