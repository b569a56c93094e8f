// pf_kernels2d.cuh -- sm_100a kernels of the 2D occupancy-grid particle-filter step.
//
// One frame = motion -> score -> extrema -> weights+tile scans -> prefix/Neff -> map update ->
// resample, all on one stream with no host round trip.  Reference semantics per kernel are cited
// inline (paths relative to michaelwillett/GPU-ICP-SLAM src/).
#pragma once
#include <cstdlib>
#include <vector>

#include "pf_arith.cuh"
#include "pf_xchg.cuh"

namespace pf {

// Programmatic dependent launch (sm_90+): a kernel launched with the stream-serialization attribute may
// have its blocks scheduled while the kernel before it is still running, once every block of that kernel
// has executed pdl_trigger(); it must execute pdl_wait() before touching anything the earlier kernel
// wrote (the wait returns when that grid has completed and flushed).  Both are no-ops in a normal launch.
// Used on the step's critical-path edges to hide launch latency (PFSLAM_PDL=0 turns it off).
// release / acquire fence at GPU scope for the "last block finishes the job" hand-overs (stores -> fence -> counter
// atomic | counter atomic -> fence -> loads).  __threadfence() is the sequentially consistent fence (MEMBAR.SC),
// which these do not need.
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PFSLAM_PDL"); v = e ? (atoi(e) != 0) : 1; }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(bool dependent, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// In-graph timeline (pfslam_debug_trace): when on, thread 0 of every block folds %globaltimer into its kernel's
// [first entry, last exit] pair -- the only way to see where the kernels of the captured step really run (programmatic
// dependent launches and graph branches overlap them; events cannot be recorded inside a graph replay).
enum { kTrMotion = 0, kTrTilePrep, kTrScore, kTrScoreFast, kTrCombine, kTrWeights, kTrResample, kTrMapFree, kTrMapWall, kTrPublish,
       kTrMark0, kTrMark1, kTrMark2, kTrMark3, kTrMark4, kTrMark5, kTrCount };   // marks: [first, last] block to reach a point
__device__ int g_trace_on = 0;
__device__ unsigned long long g_trace[2 * 16];
struct TraceScope {
    int id;
    __device__ __forceinline__ explicit TraceScope(int i) : id(i)
    {
        if (g_trace_on && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMin(&g_trace[2 * id], t);
        }
    }
    __device__ __forceinline__ ~TraceScope()
    {
        if (g_trace_on && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMax(&g_trace[2 * id + 1], t);
        }
    }
};

// diagnostic mark inside a kernel: block-wide (every thread calls it); [first block, last block] to get there
__device__ __forceinline__ void trace_mark(int id)
{
    if (!g_trace_on) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(&g_trace[2 * id], t);
        atomicMax(&g_trace[2 * id + 1], t);
    }
}

// host side: per-kernel lap events of the serialised profiling step (pfslam_profile_laps)
enum { kLapStart = -1, kLapMotion = 0, kLapTilePrep, kLapScoreTiled, kLapScoreFast, kLapCombine, kLapWeights,
       kLapPrefix, kLapMapFree, kLapMapWall, kLapResample, kLapCount };
struct LapRec {
    std::vector<cudaEvent_t> ev;
    std::vector<int> id;
    int n = 0;
    void mark(cudaStream_t s, int lap_id)
    {
        if (n < (int)ev.size()) { cudaEventRecord(ev[n], s); id[n] = lap_id; n++; }
    }
};

constexpr int kTile = 1024;            // reduction / scan tile (pfslam order)
constexpr int kScanThreads = 256;      // 8 warps x 32 lanes x 4 items
constexpr int kFusedPrefixMaxTiles = 1024;   // k_weights_scan folds the prefix step in up to 2^20 particles
constexpr int kMaxPrefixTiles = 4096;        // k_prefix: 2 floats per tile in (default-sized) dynamic shared memory
constexpr float kLidarRange = 20.0f;   // kernel.cu:44
constexpr int kFreeWeight = -1;        // kernel.cu:32
constexpr int kOccupiedWeight = 4;     // kernel.cu:33
constexpr int kClamp = 113;            // kernel.cu:518 (1<<7)-15

struct MapGeom {
    int   w, h;                // map_dim, kernel.cu:120
    float scale_x, scale_y;
    float res_x, res_y;
};

// device-resident per-frame result (mirrors pfslam_frame_result)
struct FrameResult {
    float pose[3];
    int   fit_min, fit_max, best_index;
    float sum_w, sum_w2, neff;
    int   resampled, n_free, n_wall, n_slow;
    int   kd_size, kd_ins;
    int   xchg_timeout;        // sharded engines: a peer-exchange wait ran into its time limit this frame
    int   resample_count;      // steps that resampled so far
    int   wait_ext_ns, wait_tiles_ns;   // accumulated peer-exchange wait times (block 0 / the last block of k_weights_scan)
    int   n_windows, n_wide;   // tiled scorer: windows placed / beams left to the global-memory path this frame
    int   kd_overflow;         // kd path: 1 = more wall points in a scan than the point lists hold, 2 = node array full (sticky)
};

// per-step inputs, read by the kernels from device memory so that a captured CUDA graph of the
// step can be replayed with new values (k_motion, the step's first kernel, receives them by value and files them here)
struct FrameResult;
constexpr int kParamsInDeviceMemory = (int)0x80000000;    // StepParams.seq of a by-value argument that is not one
struct StepParams {
    const float *scan;   // this frame's ranges (device)
    int frame;           // frame number (seeds, kernel.cu:380, :434)
    int seq;             // step sequence number (peer-exchange flags and buffer parity, pf_xchg.cuh)
    // host-API steps (pfslam_step / pfslam_submit): the slot of the pinned, device-mapped scan ring the first kernel
    // pulls this frame's ranges from, and the slot of the result ring the last node publishes into; null otherwise
    const float *scan_src;
    FrameResult *res_host;
};

// ---------------------------------------------------------------------------------------------
// motion: kernel.cu:375-397 ParticleAddNoise; device evaluation order of
// glm::vec3 noise(distx(e2), disty(e2), distt(e2)) is x, y, theta (read off the reference SASS).
__device__ __forceinline__ int float_order(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int k)
{
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

// Also reduces the post-noise pose bounds of the cloud (ordered-int min/max of x, y, theta) into
// bounds[6] for the tiled scorer's window placement, zeroes the scorer's accumulator row, and writes the pre-resample snapshot the
// resampler gathers from -- `snap` [parity][x | y | theta], null when a host all-gathers it.
__global__ void __launch_bounds__(256)
k_motion(float *__restrict__ x, float *__restrict__ y, float *__restrict__ th, int n,
         StepParams *sp, int gidx0, int *__restrict__ bounds,
         float *__restrict__ snap, long long snap_stride, int parity_mask, int snap_aos, int *__restrict__ acc_row,
         const float *__restrict__ scan_src, float *__restrict__ scan_dst, int n_beams, float4 *__restrict__ pcs,
         const StepParams spv)
{
    TraceScope trace_scope(kTrMotion);
    __shared__ int s_b[6];
    pdl_trigger();                              // k_tile_prep's blocks may be staged now; they wait for this grid
    // The step's parameters arrive as a kernel argument of this, the first kernel of the captured step (the host
    // updates the argument of the graph's kernel node before every launch -- no copy node in front of the graph), and
    // one thread files them in device memory for the kernels behind it (all of which start after this grid has
    // completed).  Plain launches and hosts that set the parameters themselves pass kParamsInDeviceMemory.
    const bool by_value = spv.seq != kParamsInDeviceMemory;
    const StepParams P = by_value ? spv : *sp;
    if (by_value && blockIdx.x == 0 && threadIdx.x == 0) *sp = spv;
    // host API step (pfslam_step): the frame's scan is pulled from the pinned, device-mapped staging buffer by the
    // first kernel of the step (nobody reads the device copy before this grid has completed), so the step graph needs
    // no copy node for it
    if (scan_src && blockIdx.x == gridDim.x - 1) {
        const float *__restrict__ src = P.scan_src;
        if (src) for (int j = threadIdx.x; j < n_beams; j += blockDim.x) scan_dst[j] = src[j];
    }
    if (threadIdx.x < 6) s_b[threadIdx.x] = (threadIdx.x & 1) ? (int)0x80000000 : 0x7fffffff;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    if (i < n) {
        const int frame = P.frame;
        uint32_t st = pf_minstd_seed(pf_seed(frame, gidx0 + i, 0));
        float nx = pf_normal(st, 0.015f);
        float ny = pf_normal(st, 0.015f);
        float nt = pf_normal(st, 0.01f);
        const float vx = __fadd_rn(x[i], nx), vy = __fadd_rn(y[i], ny), vt = __fadd_rn(th[i], nt);
        x[i] = vx; y[i] = vy; th[i] = vt;
        acc_row[i] = 0;                        // the tiled scorer adds into it (pf_score_tiled.cuh)
        if (pcs) { float sn, cs; sincosf(vt, &sn, &cs); pcs[i] = make_float4(vx, vy, cs, sn); }   // ... and reads the pose from here
        if (snap) {
            float *sn = snap + (long long)(P.seq & parity_mask) * snap_stride;
            if (snap_aos) reinterpret_cast<float4 *>(sn)[i] = make_float4(vx, vy, vt, 0.0f);
            else { sn[i] = vx; sn[n + i] = vy; sn[2 * n + i] = vt; }
        }
        lo[0] = hi[0] = float_order(vx); lo[1] = hi[1] = float_order(vy); lo[2] = hi[2] = float_order(vt);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_b[2 * c], lo[c]); atomicMax(&s_b[2 * c + 1], hi[c]); }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        if (threadIdx.x & 1) atomicMax(&bounds[threadIdx.x], s_b[threadIdx.x]);
        else atomicMin(&bounds[threadIdx.x], s_b[threadIdx.x]);
    }
}

// Peer-memory shards: copy this rank's pre-resample snapshot (written by k_motion into slot [parity][rank] of its own
// exchange region) into the same slot of every peer's region -- one coalesced 16-byte store per particle and peer.
// Runs on the side branch under the scoring kernels and completes before this rank's k_weights_scan raises its tile
// flags, so a rank that has seen everybody's tile flags (every k_resample has) holds everybody's snapshot.
// Small blocks (128 threads, few registers, no shared memory) on a grid-stride loop: they fit next to the three resident
// blocks per SM of the tiled scorer instead of displacing one of them.
__global__ void __launch_bounds__(128)
k_snapshot_push(const Xchg xc, const StepParams *__restrict__ sp, int n)
{
    const long long slot0 = ((long long)(sp->seq & 1) * xc.n_ranks + xc.rank) * n;          // float4 index in the snapshot area
    const float4 *__restrict__ mine = reinterpret_cast<const float4 *>(xc.peer[xc.rank] + xc.off_snap) + slot0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v = mine[i];
        for (int r = 0; r < xc.n_ranks; r++)
            if (r != xc.rank) reinterpret_cast<float4 *>(xc.peer[r] + xc.off_snap)[slot0 + i] = v;
    }
}

// host API step: the frame result goes to the pinned, device-mapped result record from the last node of the graph
__global__ void k_publish_result(const FrameResult *__restrict__ res, const StepParams *__restrict__ sp)
{
    TraceScope trace_scope(kTrPublish);
    const int *s = reinterpret_cast<const int *>(res);
    int *d = reinterpret_cast<int *>(sp->res_host);
    if (!d) return;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameResult) / 4); i += blockDim.x) d[i] = s[i];
}

__global__ void k_debug_trig(const float *__restrict__ x, long long n, float *__restrict__ c,
                             float *__restrict__ s)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { c[i] = cosf(x[i]); s[i] = sinf(x[i]); }
}

__global__ void k_init_beams(float *__restrict__ angle, int n_beams)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_beams) angle[j] = pf_lidar_angle(j);
}

// ---------------------------------------------------------------------------------------------
// One (particle, beam) evaluation exactly as the reference computes it on the GPU
// (kernel.cu:182-187, :262-270, :247-251; fused multiply-add and IEEE division per its SASS).
__device__ __forceinline__ int eval_exact(const int8_t *__restrict__ grid, const MapGeom &g,
                                          float c0x, float c0y, float px, float py, float pth,
                                          float angle, float r)
{
    float rot = __fadd_rn(angle, pth);
    float cs = cosf(rot), sn = sinf(rot);
    float wx = __fmaf_rn(r, cs, px);
    float wy = __fmaf_rn(r, sn, py);
    float gx = roundf(__fadd_rn(c0x, __fdiv_rn(wx, g.res_x)));
    float gy = roundf(__fadd_rn(c0y, __fdiv_rn(wy, g.res_y)));
    if (gx >= 0.0f && gx < (float)g.w && gy >= 0.0f && gy < (float)g.h)
        return (int)grid[(int)gx * g.w + (int)gy];
    return 0;
}

__device__ __forceinline__ long long extrema_key(int fit, int gidx)
{
    return (long long)fit * 4294967296ll + (long long)(0xFFFFFFFFu - (uint32_t)gidx);
}

// Scoring, exact mode: block = 8 warps, lane = particle (32 particles per block), warp w takes
// beams j = w, w+8, ...; partial sums meet in shared memory.  Writes fit[] and one partial
// (min, max-key) record per block for k_extrema.
__global__ void __launch_bounds__(256)
k_score_exact(const int8_t *__restrict__ grid, MapGeom g, const float *__restrict__ x,
              const float *__restrict__ y, const float *__restrict__ th, int n, int gidx0,
              const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams,
              int *__restrict__ fit, int *__restrict__ blk_min, long long *__restrict__ blk_maxkey)
{
    __shared__ int part[8][32];
    const float *__restrict__ scan = sp->scan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    int acc = 0;
    if (p < n) {
        const float px = x[p], py = y[p], pth = th[p];
        for (int j = warp; j < n_beams; j += 8)
            acc += eval_exact(grid, g, c0x, c0y, px, py, pth, __ldg(&angle[j]), __ldg(&scan[j]));
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        int s = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += part[w][lane];
        int mn = 0x7fffffff;
        long long mk = (long long)0x8000000000000000ull;
        if (p < n) { fit[p] = s; mn = s; mk = extrema_key(s, gidx0 + p); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            long long t = __shfl_xor_sync(0xffffffffu, mk, o);
            mk = t > mk ? t : mk;
        }
        if (lane == 0) { blk_min[blockIdx.x] = mn; blk_maxkey[blockIdx.x] = mk; }
    }
}

// thrust::minmax_element (kernel.cu:323-326): min, max and the FIRST arg-max, from the per-block
// partials; also records the best particle's pose for the exchange.
__global__ void __launch_bounds__(1024)
k_extrema(const int *__restrict__ blk_min, const long long *__restrict__ blk_maxkey, int n_blk,
          const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th,
          int gidx0, Extrema *__restrict__ out, const Xchg xc, const StepParams *__restrict__ sp)
{
    __shared__ int smin[32];
    __shared__ long long smax[32];
    int mn = 0x7fffffff;
    long long mk = (long long)0x8000000000000000ull;
    for (int i = threadIdx.x; i < n_blk; i += blockDim.x) {
        mn = min(mn, blk_min[i]);
        long long t = blk_maxkey[i];
        mk = t > mk ? t : mk;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        long long t = __shfl_xor_sync(0xffffffffu, mk, o);
        mk = t > mk ? t : mk;
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; w++) { mn = min(mn, smin[w]); mk = smax[w] > mk ? smax[w] : mk; }
        int best = (int)(0xFFFFFFFFu - (uint32_t)(mk & 0xFFFFFFFFll));
        Extrema e;
        e.fit_min = mn; e.fit_max = (int)(mk >> 32); e.best_gidx = best;
        e.x = x[best - gidx0]; e.y = y[best - gidx0]; e.th = th[best - gidx0];
        e.pad0 = 0; e.pad1 = 0;
        xc_publish_extrema(xc, out, e, sp->seq);
    }
}

// combine the per-rank extrema: global min, max, first arg-max and its pose
__device__ __forceinline__ void reduce_extrema(const Extrema *__restrict__ all, int n_ranks,
                                               int &gmin, int &gmax, int &best, float pose[3], bool wire)
{
    Extrema b = xc_load_extrema(all, wire);
    gmin = b.fit_min;
    long long mk = extrema_key(b.fit_max, b.best_gidx);
    for (int r = 1; r < n_ranks; r++) {
        const Extrema e = xc_load_extrema(all + r, wire);
        gmin = min(gmin, e.fit_min);
        long long t = extrema_key(e.fit_max, e.best_gidx);
        if (t > mk) { mk = t; b = e; }
    }
    gmax = (int)(mk >> 32);
    best = b.best_gidx;
    pose[0] = b.x; pose[1] = b.y; pose[2] = b.th;
}

// ---------------------------------------------------------------------------------------------
// pfslam-order tile scan of 4 items per thread over a 1024 tile (256 threads).  Returns the
// monotone inclusive values LM[0..3] for this thread's items and the tile total (in all threads).
// Order: thread sequential -> Kogge-Stone over the warp's 32 thread totals -> sequential over the
// 8 warp totals -> L = (warp_excl + lane_excl) + thread_incl -> running max (exact).
__device__ __forceinline__ float tile_scan4(const float e[4], float lm[4], float *s_wtot,
                                            float *s_wmax)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s0 = e[0];
    float s1 = __fadd_rn(s0, e[1]);
    float s2 = __fadd_rn(s1, e[2]);
    float s3 = __fadd_rn(s2, e[3]);
    float v = s3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = __fadd_rn(v, t);
    }
    float lane_excl = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) lane_excl = 0.0f;
    if (lane == 31) s_wtot[warp] = v;
    __syncthreads();
    float wexcl = 0.0f;
    for (int w = 0; w < warp; w++) wexcl = __fadd_rn(wexcl, s_wtot[w]);
    float b = __fadd_rn(wexcl, lane_excl);
    float l0 = __fadd_rn(b, s0), l1 = __fadd_rn(b, s1), l2 = __fadd_rn(b, s2), l3 = __fadd_rn(b, s3);
    // exclusive running max over earlier threads (values are >= 0; thread-local values ascend)
    float m = l3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, m, o);
        if (lane >= o) m = fmaxf(m, t);
    }
    float lane_pm = __shfl_up_sync(0xffffffffu, m, 1);
    if (lane == 0) lane_pm = 0.0f;
    if (lane == 31) s_wmax[warp] = m;
    __syncthreads();
    float pm = lane_pm;
    for (int w = 0; w < warp; w++) pm = fmaxf(pm, s_wmax[w]);
    lm[0] = fmaxf(pm, l0); lm[1] = fmaxf(pm, l1); lm[2] = fmaxf(pm, l2); lm[3] = fmaxf(pm, l3);
    float total = s_wmax[0];
    for (int w = 1; w < 8; w++) total = fmaxf(total, s_wmax[w]);
    __syncthreads();   // s_wtot / s_wmax are reused by the next call
    return total;
}

// Sequential fp32 sum of v[0..n) in index order (the arithmetic contract's global tile order), optionally leaving
// the inclusive prefix in place.  Loads are batched 16 ahead of the dependent adds, so the chain costs one FADD
// latency per tile instead of a shared-memory round trip per tile (512 tiles on 8 GPUs: ~9 us -> ~1.5 us).
__device__ __forceinline__ float chain_sum(float *v, int n, bool keep_prefix)
{
    float p = 0.0f;
    int t = 0;
    for (; t + 16 <= n; t += 16) {
        float r[16];
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = v[t + k];
#pragma unroll
        for (int k = 0; k < 16; k++) { p = __fadd_rn(p, r[k]); r[k] = p; }
        if (keep_prefix) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[t + k] = r[k];
        }
    }
    for (; t < n; t++) { p = __fadd_rn(p, v[t]); if (keep_prefix) v[t] = p; }
    return p;
}

// weights + tile scans.  kernel.cu:287-294 kernUpdateWeights: w = w*((float)fit - min)*c with
// c = 1/(float)(max-min) when max > min (kernel.cu:329-331); SURVEY Q1: only the first
// ceil(N/2) particles' new weights persist (kernel.cu:337).  Then the per-tile scans that feed
// Neff (kernel.cu:456-472) and the resampling CDF (kernel.cu:478).
// A tiles block is [tsum_w (n_tiles)][tsum_w2 (n_tiles)] at xc.sum_off and lm (n) at xc.lm_off.
// Sharded over peer memory (xc.parity_mask): the kernel first waits for every rank's extrema, and
// stores its tile results straight into every rank's exchange region; the last block raises the flags.
__global__ void __launch_bounds__(kScanThreads)
k_weights_scan(const Xchg xc, const StepParams *__restrict__ sp, const int *__restrict__ fit,
               float *__restrict__ w, int n, int gidx0, int n_sync, int n_tiles,
               float *tiles_local, int fuse_prefix, int n_global, float *__restrict__ prefix,
               FrameResult *__restrict__ res, int write_pose, int *__restrict__ done_counter)
{
    TraceScope trace_scope(kTrWeights);
    __shared__ float s_wtot[8], s_wmax[8];
    __shared__ int s_last;
    pdl_trigger();                              // k_resample's blocks may be staged (measured: 0.3 us per step better than
                                                // staging them after the tile scans)
    pdl_wait();                                 // scores and extrema of k_score_combine_rows
    const int seq = sp->seq;
    if (xc.parity_mask) {
        if (threadIdx.x < 32) {
            const unsigned long long t0 = xc_now_ns();
            const bool ok = xc_wait_warp(xc, kXcExt, seq);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (!ok) res->xchg_timeout = 1;
                res->wait_ext_ns += (int)(xc_now_ns() - t0);
            }
        }
        __syncthreads();
    }
    int gmin, gmax, best; float pose[3];
    reduce_extrema(xc_ext(xc, seq), xc.n_ranks, gmin, gmax, best, pose, xc.parity_mask != 0);
    const int rng = gmax - gmin;
    const float c = rng > 0 ? __fdiv_rn(1.0f, (float)rng) : 1.0f;
    const float fmin = (float)gmin;
    const int base = blockIdx.x * kTile + threadIdx.x * 4;
    float e[4], q[4], lm[4], lm2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int i = base + k;
        float we = 0.0f;
        if (i < n) {
            we = w[i];
            if (rng > 0) we = __fmul_rn(__fmul_rn(we, __fsub_rn((float)fit[i], fmin)), c);
            if (gidx0 + i < n_sync) w[i] = we;
        }
        e[k] = we;
        q[k] = __fmul_rn(we, we);
    }
    trace_mark(kTrMark3);                       // extrema and weights in registers
    float t2 = tile_scan4(q, lm2, s_wtot, s_wmax);
    float t1 = tile_scan4(e, lm, s_wtot, s_wmax);
    trace_mark(kTrMark4);                       // tile scans done
    if (!xc.parity_mask) {
        float *lm_out = tiles_local + xc.lm_off;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (base + k < n) lm_out[base + k] = lm[k];
        if (threadIdx.x == 0) { tiles_local[xc.sum_off + blockIdx.x] = t1; tiles_local[xc.sum_off + n_tiles + blockIdx.x] = t2; }
        if (!fuse_prefix) return;
    } else {
        // shards are tile-aligned (n % kTile == 0): every thread owns 4 valid, 16-byte-aligned items
        for (int r = 0; r < xc.n_ranks; r++) {
            float *blk = reinterpret_cast<float *>(xc.peer[r] + xc.off_tiles) +
                         ((long long)(seq & 1) * xc.n_ranks + xc.rank) * xc.tiles_block;
            *reinterpret_cast<float4 *>(blk + xc.lm_off + base) = make_float4(lm[0], lm[1], lm[2], lm[3]);
            if (threadIdx.x == 0) { blk[xc.sum_off + blockIdx.x] = t1; blk[xc.sum_off + n_tiles + blockIdx.x] = t2; }
        }
        __syncthreads();          // thread 0's system fence below then covers the whole block's stores
    }
    if (threadIdx.x == 0) {
        if (xc.parity_mask) xc_fence_sys(); else fence_gpu();
        s_last = atomicAdd(done_counter, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    trace_mark(kTrMark5);                       // the last block knows it is the last
    // The last block to finish also does k_prefix's job when asked to (global tile prefix in tile order,
    // Neff, resample decision, robotPos).  Sharded over peer memory it first raises this rank's tile
    // flags, then waits for everybody's: one spinning block, and no separate kernel on the critical path.
    if (xc.parity_mask) {
        if (threadIdx.x == 0) { *done_counter = 0; xc_signal(xc, kXcTiles, seq); }
        if (!fuse_prefix) return;
        if (threadIdx.x < 32) {
            const unsigned long long t0 = xc_now_ns();
            const bool ok = xc_wait_warp(xc, kXcTiles, seq);
            if (threadIdx.x == 0) {
                if (!ok) res->xchg_timeout = 1;
                res->wait_tiles_ns += (int)(xc_now_ns() - t0);
            }
        }
        __syncthreads();
    } else {
        fence_gpu();
    }
    __shared__ float s_t[2 * kFusedPrefixMaxTiles];
    const int nt = xc.n_ranks * n_tiles;
    {
        const float *all = xc.parity_mask ? xc_tiles(xc, seq) : tiles_local;
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
            const int r = t / n_tiles, tl = t - r * n_tiles;
            const float *blk = all + (long long)r * xc.tiles_block + xc.sum_off;
            s_t[t] = __ldcg(&blk[tl]);
            s_t[nt + t] = __ldcg(&blk[n_tiles + tl]);
        }
    }
    __syncthreads();
    // the two sequential chains (sum of w in global tile order = the CDF's tile prefix, sum of w^2) run in two
    // warps side by side, in place in shared memory; the prefix goes out to global memory in parallel afterwards
    __shared__ float s_tot[2];
    if (threadIdx.x == 0) s_tot[0] = chain_sum(s_t, nt, true);
    else if (threadIdx.x == 32) s_tot[1] = chain_sum(s_t + nt, nt, false);
    __syncthreads();
    for (int t = threadIdx.x; t <= nt; t += blockDim.x) prefix[t] = t ? s_t[t - 1] : 0.0f;
    if (threadIdx.x == 0) {
        const float p = s_tot[0], p2 = s_tot[1];
        const float neff = __fdiv_rn(__fmul_rn(p, p), p2);
        if (write_pose) { res->pose[0] = pose[0]; res->pose[1] = pose[1]; res->pose[2] = pose[2]; }
        res->fit_min = gmin; res->fit_max = gmax; res->best_index = best;
        res->sum_w = p; res->sum_w2 = p2; res->neff = neff;
        res->resampled = ((double)neff < 0.7 * (double)n_global) ? 1 : 0;
        if (!xc.parity_mask) *done_counter = 0;
    }
}

// global tile prefixes (sequential in global tile order), Neff (kernel.cu:472), the resample
// decision (kernel.cu:474) and robotPos (kernel.cu:338).  prefix: n_tiles_global+1 floats.
// Sharded over peer memory: waits for every rank's tile results first.
__global__ void __launch_bounds__(1024)
k_prefix(const Xchg xc, const StepParams *__restrict__ sp, int n_tiles_local, int n_global,
         float *__restrict__ prefix, FrameResult *__restrict__ res, int write_pose)
{
    extern __shared__ float s_t[];            // 2 * n_tiles_global
    const int seq = sp->seq;
    if (xc.parity_mask) {
        if (threadIdx.x < 32) {
            const bool ok = xc_wait_warp(xc, kXcTiles, seq);
            if (!ok && threadIdx.x == 0) res->xchg_timeout = 1;
        }
        __syncthreads();
    }
    const float *tiles_all = xc_tiles(xc, seq);
    const int nt = xc.n_ranks * n_tiles_local;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        int r = t / n_tiles_local, tl = t - r * n_tiles_local;
        const float *blk = tiles_all + (long long)r * xc.tiles_block + xc.sum_off;
        s_t[t] = __ldcg(&blk[tl]);
        s_t[nt + t] = __ldcg(&blk[n_tiles_local + tl]);
    }
    __syncthreads();
    __shared__ float s_tot[2];
    if (threadIdx.x == 0) s_tot[0] = chain_sum(s_t, nt, true);
    else if (threadIdx.x == 32) s_tot[1] = chain_sum(s_t + nt, nt, false);
    __syncthreads();
    for (int t = threadIdx.x; t <= nt; t += blockDim.x) prefix[t] = t ? s_t[t - 1] : 0.0f;
    if (threadIdx.x == 0) {
        const float p = s_tot[0], p2 = s_tot[1];
        float neff = __fdiv_rn(__fmul_rn(p, p), p2);
        int gmin, gmax, best; float pose[3];
        reduce_extrema(xc_ext(xc, seq), xc.n_ranks, gmin, gmax, best, pose, xc.parity_mask != 0);
        if (write_pose) { res->pose[0] = pose[0]; res->pose[1] = pose[1]; res->pose[2] = pose[2]; }
        res->fit_min = gmin; res->fit_max = gmax; res->best_index = best;
        res->sum_w = p; res->sum_w2 = p2; res->neff = neff;
        res->resampled = ((double)neff < 0.7 * (double)n_global) ? 1 : 0;
    }
}

// resample: kernel.cu:429-444 kernWeightedSample (seed (Neff, frame, i)), gathering
// from the pre-resample snapshot.  The CDF is prefix[tile] + lm[i]; it is monotone, so
// the two-level binary search returns the reference's linear-scan index.  Sharded over peer memory
// the drawn particle's pose is loaded from its owner's snapshot (xc.pose_src[owner], NVLink).
__global__ void __launch_bounds__(256)
k_resample(const Xchg xc, FrameResult *res, const float *__restrict__ prefix,
           int n_tiles_local, int n_local, int n_global, int gidx0,
           const StepParams *__restrict__ sp, float *__restrict__ x, float *__restrict__ y,
           float *__restrict__ th, float *__restrict__ w)
{
    TraceScope trace_scope(kTrResample);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();                                 // k_weights_scan's prefix, decision and tiles
    if (i >= n_local || !res->resampled) return;
    if (i == 0) res->resample_count++;
    const int frame = sp->frame, seq = sp->seq;
    const float *tiles_all = xc_tiles(xc, seq);
    const int nt = (n_global + kTile - 1) / kTile;   // == n_ranks * n_tiles_local when sharded
    uint32_t st = pf_minstd_seed(pf_seed((int)res->neff, frame, gidx0 + i));
    uint32_t u = pf_minstd_next(st) - 1u;
    float rnd = __fmul_rn(__fmul_rn((float)u, 4.656612873077392578125e-10f), res->sum_w);
    // tile: first t with cdf(last of t) = prefix[t+1] >= rnd
    int lo = 0, hi = nt;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (rnd > prefix[mid + 1]) lo = mid + 1; else hi = mid; }
    int src;
    if (lo >= nt) src = n_global - 1;
    else {
        const int r = (lo * kTile) / n_local;
        const int l0 = lo * kTile - r * n_local;
        const float *lm = tiles_all + (long long)r * xc.tiles_block + xc.lm_off + l0;
        const float pt = prefix[lo];
        int cnt = min(kTile, n_global - lo * kTile);
        int a = 0, b = cnt;
        while (a < b) { int mid = (a + b) >> 1; if (rnd > __fadd_rn(pt, lm[mid])) a = mid + 1; else b = mid; }
        src = lo * kTile + (a < cnt ? a : cnt - 1);
    }
    const int r = src / n_local, l = src - r * n_local;
    const float *pp = xc.pose_src[r] + (long long)(seq & xc.parity_mask) * xc.snap_stride;
    if (xc.snap_aos) {
        const float4 v = reinterpret_cast<const float4 *>(pp)[l];
        x[i] = v.x; y[i] = v.y; th[i] = v.z;
    } else {
        x[i] = pp[l]; y[i] = pp[n_local + l]; th[i] = pp[2 * n_local + l];
    }
    w[i] = 1.0f;
}

// ---------------------------------------------------------------------------------------------
// k_weights_scan + the prefix step + k_resample as ONE launch with one grid-wide barrier in the middle: block t owns
// tile t (1024 particles) in both halves.  After the barrier EVERY block sums the tile totals itself (the same
// sequential chain, so the same bits), keeps the prefix in shared memory and resamples its own tile -- no "last
// block" election, no prefix round trip through global memory, no dependent kernel launch on the step's critical
// path.  The barrier needs all n_tiles blocks resident at once; the host uses this kernel only while n_tiles is a
// fraction of what the device holds (otherwise the three-kernel sequence), and the spin is bounded like every
// other wait of the step.
//   single GPU : barrier = a monotone arrival counter (64-bit, never reset: a block's target is the next multiple
//                of the grid size above its own ticket)
//   peer memory: the last arrival raises this rank's tile flags; the barrier IS the wait for every rank's flags
//                (this rank's own included), so the local barrier and the exchange wait are one spin.
__global__ void __launch_bounds__(kScanThreads)
k_weights_resample(const Xchg xc, const StepParams *__restrict__ sp, const int *__restrict__ fit,
                   float *__restrict__ w, int n, int gidx0, int n_sync, int n_tiles, float *tiles_local,
                   int n_global, float *__restrict__ prefix, FrameResult *__restrict__ res,
                   unsigned long long *__restrict__ arrivals, float *__restrict__ x, float *__restrict__ y,
                   float *__restrict__ th)
{
    TraceScope trace_scope(kTrWeights);
    __shared__ float s_wtot[8], s_wmax[8];
    __shared__ __align__(16) float s_t[2 * kFusedPrefixMaxTiles];
    __shared__ float s_tot[2];
    __shared__ int s_ok;
    pdl_wait();                                 // scores and extrema of k_score_combine_rows
    const int seq = sp->seq, frame = sp->frame;
    if (threadIdx.x == 0) s_ok = 1;
    if (xc.parity_mask) {
        if (threadIdx.x < 32) {
            const unsigned long long t0 = xc_now_ns();
            const bool ok = xc_wait_warp(xc, kXcExt, seq);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (!ok) res->xchg_timeout = 1;
                res->wait_ext_ns += (int)(xc_now_ns() - t0);
            }
        }
        __syncthreads();
    }
    int gmin, gmax, best; float pose[3];
    reduce_extrema(xc_ext(xc, seq), xc.n_ranks, gmin, gmax, best, pose, xc.parity_mask != 0);
    const int rng = gmax - gmin;
    const float c = rng > 0 ? __fdiv_rn(1.0f, (float)rng) : 1.0f;
    const float fmin = (float)gmin;
    const int base = blockIdx.x * kTile + threadIdx.x * 4;
    float e[4], q[4], lm[4], lm2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int i = base + k;
        float we = 0.0f;
        if (i < n) {
            we = w[i];
            if (rng > 0) we = __fmul_rn(__fmul_rn(we, __fsub_rn((float)fit[i], fmin)), c);
            if (gidx0 + i < n_sync) w[i] = we;
        }
        e[k] = we;
        q[k] = __fmul_rn(we, we);
    }
    float t2 = tile_scan4(q, lm2, s_wtot, s_wmax);
    float t1 = tile_scan4(e, lm, s_wtot, s_wmax);
    if (!xc.parity_mask) {
        float *lm_out = tiles_local + xc.lm_off;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (base + k < n) lm_out[base + k] = lm[k];
        if (threadIdx.x == 0) { tiles_local[xc.sum_off + blockIdx.x] = t1; tiles_local[xc.sum_off + n_tiles + blockIdx.x] = t2; }
    } else {
        for (int r = 0; r < xc.n_ranks; r++) {
            float *blk = reinterpret_cast<float *>(xc.peer[r] + xc.off_tiles) +
                         ((long long)(seq & 1) * xc.n_ranks + xc.rank) * xc.tiles_block;
            *reinterpret_cast<float4 *>(blk + xc.lm_off + base) = make_float4(lm[0], lm[1], lm[2], lm[3]);
            if (threadIdx.x == 0) { blk[xc.sum_off + blockIdx.x] = t1; blk[xc.sum_off + n_tiles + blockIdx.x] = t2; }
        }
    }
    __syncthreads();              // thread 0's fence below then covers the whole block's stores
    // ---- grid-wide barrier -------------------------------------------------------------------------
    if (xc.parity_mask) {
        if (threadIdx.x == 0) {
            xc_fence_sys();
            const unsigned long long ticket = atomicAdd(arrivals, 1ull);
            if (ticket % gridDim.x == gridDim.x - 1) xc_signal(xc, kXcTiles, seq);
        }
        if (threadIdx.x < 32) {
            const unsigned long long t0 = xc_now_ns();
            const bool ok = xc_wait_warp(xc, kXcTiles, seq);
            if (threadIdx.x == 0) {
                if (!ok) s_ok = 0;
                if (blockIdx.x == 0) res->wait_tiles_ns += (int)(xc_now_ns() - t0);
            }
        }
    } else if (threadIdx.x == 0) {
        fence_gpu();
        const unsigned long long ticket = atomicAdd(arrivals, 1ull);
        const unsigned long long target = (ticket / gridDim.x + 1ull) * gridDim.x;
        const unsigned long long t0 = xc_now_ns();
        unsigned spins = 0;
        while (*reinterpret_cast<volatile unsigned long long *>(arrivals) < target) {
            if ((++spins & 255u) == 0u && xc_now_ns() - t0 > 2000000000ull) { s_ok = 0; break; }
        }
        fence_gpu();
    }
    __syncthreads();
    if (g_trace_on && threadIdx.x == 0) atomicMin(&g_trace[2 * kTrResample], xc_now_ns());
    if (!s_ok) { if (threadIdx.x == 0) res->xchg_timeout = 1; return; }    // bounded: nothing hangs, the frame is flagged
    pdl_trigger();                              // the step's last node may be staged
    // ---- every block: tile totals -> prefix in shared memory, Neff, decision ----------------------------
    const int nt = xc.n_ranks * n_tiles;
    const float *all = xc.parity_mask ? xc_tiles(xc, seq) : tiles_local;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        const int r = t / n_tiles, tl = t - r * n_tiles;
        const float *blk = all + (long long)r * xc.tiles_block + xc.sum_off;
        s_t[t] = __ldcg(&blk[tl]);
        s_t[nt + t] = __ldcg(&blk[n_tiles + tl]);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_tot[0] = chain_sum(s_t, nt, true);
    else if (threadIdx.x == 32) s_tot[1] = chain_sum(s_t + nt, nt, false);
    __syncthreads();
    const float sum_w = s_tot[0], sum_w2 = s_tot[1];
    const float neff = __fdiv_rn(__fmul_rn(sum_w, sum_w), sum_w2);
    const int resampled = ((double)neff < 0.7 * (double)n_global) ? 1 : 0;
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t <= nt; t += blockDim.x) prefix[t] = t ? s_t[t - 1] : 0.0f;
        if (threadIdx.x == 0) {
            res->pose[0] = pose[0]; res->pose[1] = pose[1]; res->pose[2] = pose[2];
            res->fit_min = gmin; res->fit_max = gmax; res->best_index = best;
            res->sum_w = sum_w; res->sum_w2 = sum_w2; res->neff = neff;
            res->resampled = resampled;
            if (resampled) res->resample_count++;
        }
    }
    if (!resampled) { if (g_trace_on && threadIdx.x == 0) atomicMax(&g_trace[2 * kTrResample + 1], xc_now_ns()); return; }
    // ---- resample this block's tile (k_resample; s_t[t-1] = prefix[t]) -----------------------------------
    // Four particles per thread, searched side by side with fixed trip counts (branch-free lower bounds), so the
    // dependent probes of the four overlap instead of running one particle after the other.
    const int n_local = n;
    float rnd[4]; int lo[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = blockIdx.x * kTile + k * kScanThreads + threadIdx.x;
        uint32_t st = pf_minstd_seed(pf_seed((int)neff, frame, gidx0 + i));
        uint32_t u = pf_minstd_next(st) - 1u;
        rnd[k] = __fmul_rn(__fmul_rn((float)u, 4.656612873077392578125e-10f), sum_w);
        lo[k] = 0;
    }
    // tile: number of leading tiles whose inclusive prefix s_t[t] is < rnd  (== first t with prefix[t+1] >= rnd)
    int top = 1;
    while (top * 2 <= nt) top *= 2;
    for (int step = top; step > 0; step >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int m = lo[k] + step;
            const bool take = (m <= nt) & (rnd[k] > s_t[min(m, nt) - 1]);
            lo[k] = take ? m : lo[k];
        }
    }
    const float *lmp[4]; float pt[4]; int cnt[4], a[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int t = min(lo[k], nt - 1);
        const int r = (t * kTile) / n_local;
        lmp[k] = all + (long long)r * xc.tiles_block + xc.lm_off + (t * kTile - r * n_local);
        pt[k] = t ? s_t[t - 1] : 0.0f;
        cnt[k] = min(kTile, n_global - t * kTile);
        a[k] = 0;
    }
    // in the tile: number of leading items with prefix + lm < rnd.  Plain loads: the barrier above ordered every
    // block's (and, through the flags, every rank's) lm stores before them, and the last probes share a cache line.
    // (loads are unconditional on a clamped index: a branch per probe would serialise the four particles again)
    for (int step = kTile; step > 0; step >>= 1) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = lmp[k][min(a[k] + step, cnt[k]) - 1];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int m = a[k] + step;
            const bool take = (m <= cnt[k]) & (rnd[k] > __fadd_rn(pt[k], v[k]));
            a[k] = take ? m : a[k];
        }
    }
    float4 got[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int src = lo[k] >= nt ? n_global - 1 : lo[k] * kTile + (a[k] < cnt[k] ? a[k] : cnt[k] - 1);
        const int r = src / n_local, l = src - r * n_local;
        const float *pp = xc.pose_src[r] + (long long)(seq & xc.parity_mask) * xc.snap_stride;
        if (xc.snap_aos) got[k] = reinterpret_cast<const float4 *>(pp)[l];
        else got[k] = make_float4(pp[l], pp[n_local + l], pp[2 * n_local + l], 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = blockIdx.x * kTile + k * kScanThreads + threadIdx.x;
        if (i < n_local) { x[i] = got[k].x; y[i] = got[k].y; th[i] = got[k].z; w[i] = 1.0f; }
    }
    if (g_trace_on && threadIdx.x == 0) atomicMax(&g_trace[2 * kTrResample + 1], xc_now_ns());
}

// ---------------------------------------------------------------------------------------------
// map update.  kernel.cu:551-555 center cell; :524-549 kernGetWalls; :190-240 traceRay;
// :513-522 kernUpdateMap (-1 once per free cell, then +4 once per wall cell, clamp +-113).
__device__ __forceinline__ void center_cell(const MapGeom &g, float rx, float ry, int &cx, int &cy)
{
    float fx = __fadd_rn(__fadd_rn(__fmul_rn(0.5f, (float)g.w), __fdiv_rn(rx, g.res_x)), __fdiv_rn(g.res_x, 2.0f));
    float fy = __fadd_rn(__fadd_rn(__fmul_rn(0.5f, (float)g.h), __fdiv_rn(ry, g.res_y)), __fdiv_rn(g.res_y, 2.0f));
    cx = (int)roundf(fx);
    cy = (int)roundf(fy);
}

// hit cell of beam j for the robot pose; false if the beam fails the +-20 m filter
__device__ __forceinline__ bool beam_hit(const MapGeom &g, const float *pose, int cx, int cy,
                                         float angle, float r, float &wx, float &wy)
{
    float rot = __fadd_rn(angle, pose[2]);
    wx = __fmul_rn(r, cosf(rot));
    wy = __fmul_rn(r, sinf(rot));
    if (!(fabsf(wx) < kLidarRange && fabsf(wy) < kLidarRange)) return false;
    wx = __fadd_rn(roundf(__fdiv_rn(wx, g.res_x)), (float)cx);
    wy = __fadd_rn(roundf(__fdiv_rn(wy, g.res_y)), (float)cy);
    return true;
}

__device__ __forceinline__ int8_t clamp_add(int8_t v, int d)
{
    int t = (int)v + d;
    return (int8_t)(t < -kClamp ? -kClamp : t > kClamp ? kClamp : t);
}

// robotPos for the map update: the frame result, or the best particle's pose from the exchanged extrema (below)
__device__ __forceinline__ void map_pose(const FrameResult *__restrict__ res, const Xchg &xc, int seq, int pose_from_ext,
                                         float pose[3])
{
    if (pose_from_ext) {
        int gmin, gmax, best;
        reduce_extrema(xc_ext(xc, seq), xc.n_ranks, gmin, gmax, best, pose, xc.parity_mask != 0);
    } else {
        pose[0] = res->pose[0]; pose[1] = res->pose[1]; pose[2] = res->pose[2];
    }
}

// One warp that waits for every rank's flag of `kind`: heads a graph branch that consumes exchanged data
// but does not follow a kernel that has already waited (the map branch).  A single spinning warp, so
// that waiting never occupies the SMs the other shards' kernels may need.
__global__ void k_xc_wait(const Xchg xc, const StepParams *__restrict__ sp, int kind, FrameResult *__restrict__ res)
{
    const bool ok = xc_wait_warp(xc, kind, sp->seq);
    if (!ok && threadIdx.x == 0) res->xchg_timeout = 1;
}

// One beam's ray as traceRay sets it up (kernel.cu:190-214): endpoints swapped so that x is the major axis and
// ascending; `ring` below is the distance from the robot's cell along the major axis (== the Chebyshev distance of
// the visited cell, because Bresenham's minor offset never exceeds the major one).
struct FreeRay {
    int  sx, sy, deltax, deltay, e0, ystep;
    bool steep, swapped, valid;
};
__device__ __forceinline__ FreeRay free_ray(const MapGeom &g, const float *pose, int cx, int cy, float angle, float r)
{
    FreeRay L;
    float wx, wy;
    L.valid = beam_hit(g, pose, cx, cy, angle, r, wx, wy);
    int sx = cx, sy = cy, ex = (int)wx, ey = (int)wy;
    L.steep = abs(ey - sy) > abs(ex - sx);
    int t;
    if (L.steep) { t = sx; sx = sy; sy = t; t = ex; ex = ey; ey = t; }
    L.swapped = sx > ex;
    if (L.swapped) { t = sx; sx = ex; ex = t; t = sy; sy = ey; ey = t; }
    L.sx = sx; L.sy = sy;
    L.deltax = ex - sx; L.deltay = abs(ey - sy); L.e0 = L.deltax / 2;
    L.ystep = ey > sy ? 1 : -1;
    return L;
}
// cell index of traceRay's step k (closed form, see the header comment of the map update), or -1 when the step does
// not exist or falls outside the map
__device__ __forceinline__ int free_ray_cell(const FreeRay &L, const MapGeom &g, int k)
{
    if (!L.valid || k < 0 || k >= L.deltax) return -1;
    const int num = k * L.deltay - L.e0;
    const int m = num > 0 ? (num + L.deltax - 1) / L.deltax : 0;
    const int xx = L.sx + k, yy = L.sy + L.ystep * m;
    const int id = L.steep ? yy * g.w + xx : xx * g.w + yy;
    return (xx < g.w && yy < g.h && xx >= 0 && yy >= 0 && id < g.w * g.h) ? id : -1;
}

// Map update, free cells: block = one beam, threads stride over the Bresenham steps (closed form, so every step is
// independent).  "-1 once per cell" (the reference's bool mask): a cell is claimed by exchanging this update's epoch
// into a 32-bit stamp per cell; whoever reads back an older stamp applies the -1.
//   * stamps, not a bitmap: atomics on one 32-byte sector serialise in L2 (~15 ns each, measured), and with one bit
//     per cell a sector covers 256 cells of a row, which several hundred x-major beams cross -- the bitmap version
//     waited 8 us for its atomics.  A stamp sector covers 8 cells; and the stamps need no clearing between frames.
//   * adjacent beams trace the same cells out to ~230 cells from the robot: a step is skipped when the previous beam
//     visits the same cell at the same ring (checked exactly), so the lowest beam of a run claims it -- 3x fewer
//     atomics, none of them piled on the cells around the robot.
// The cell's old value is loaded next to the claim (nobody else writes the cell during this kernel).
// robotPos for the map update: the frame result (explicit-pose entry point, phase-by-phase hosts), or --
// when the map update runs as a branch of the step graph next to the weight/resample kernels -- the
// best particle's pose straight from the exchanged extrema (== what k_prefix writes into the result).
constexpr int kFreeUnroll = 4;      // steps per thread in flight (4 x 128 = 512 steps per pass: one pass for rays up to 12.8 m)
__global__ void __launch_bounds__(128, 12)
k_map_free(int8_t *__restrict__ grid, MapGeom g, FrameResult *__restrict__ res,
           const StepParams *__restrict__ sp, const float *__restrict__ angle,
           unsigned *__restrict__ stamps, int *__restrict__ counters, const Xchg xc, int pose_from_ext)
{
    TraceScope trace_scope(kTrMapFree);
    pdl_trigger();                              // k_map_wall's blocks may be staged; they wait for this grid to complete
    __shared__ int s_cnt;
    __shared__ FreeRay s_ray[2];
    const float *__restrict__ scan = sp->scan;
    const int j = blockIdx.x;
    const unsigned epoch = (unsigned)__ldcg(&counters[10]) + 1u;      // k_map_wall closes the epoch
    if ((threadIdx.x & 31) == 0 && threadIdx.x < 64) {                // this beam's ray and the previous beam's, one warp each
        const int which = threadIdx.x >> 5, jj = j - which;
        if (which == 0) s_cnt = 0;
        FreeRay R; R.valid = false;
        if (jj >= 0) {
            float pose[3];
            map_pose(res, xc, sp->seq, pose_from_ext, pose);
            int cx, cy; center_cell(g, pose[0], pose[1], cx, cy);
            R = free_ray(g, pose, cx, cy, angle[jj], scan[jj]);
        }
        s_ray[which] = R;
    }
    __syncthreads();
    const FreeRay L = s_ray[0], P = s_ray[1];
    trace_mark(kTrMark0);                       // pose and rays known
    int mine = 0;
    if (L.valid) {
        for (int k0 = threadIdx.x; k0 < L.deltax; k0 += kFreeUnroll * blockDim.x) {
            int idx[kFreeUnroll]; unsigned old[kFreeUnroll]; int8_t v[kFreeUnroll];
#pragma unroll
            for (int u = 0; u < kFreeUnroll; u++) {
                const int k = k0 + u * blockDim.x;
                int id = free_ray_cell(L, g, k);
                if (id >= 0) {
                    const int ring = L.swapped ? L.deltax - k : k;
                    if (free_ray_cell(P, g, P.swapped ? P.deltax - ring : ring) == id) id = -1;   // the previous beam has it
                }
                idx[u] = id; old[u] = epoch; v[u] = 0;
                if (id >= 0) { old[u] = atomicExch(&stamps[id], epoch); v[u] = grid[id]; }
            }
            if (g_trace_on && threadIdx.x == 0) {            // thread 0's atomics have returned
                unsigned any = 0;
#pragma unroll
                for (int u = 0; u < kFreeUnroll; u++) any += old[u];
                if (any != 0xFFFFFFFFu) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    atomicMin(&g_trace[2 * kTrMark1], t);
                    atomicMax(&g_trace[2 * kTrMark1 + 1], t);
                }
            }
#pragma unroll
            for (int u = 0; u < kFreeUnroll; u++)
                if (old[u] != epoch) { grid[idx[u]] = clamp_add(v[u], kFreeWeight); mine++; }
        }
    }
    trace_mark(kTrMark2);                       // cells updated
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&counters[0], s_cnt);
}

// wall cells: ONE block, threads stride over the beams; runs after k_map_free has completed, so the +4 lands on top
// of the -1 exactly like the reference's two kernUpdateMap launches.  Only the grid update itself depends on the
// free pass: the pose, the hit cell and the claim in the wall mask are computed while k_map_free is still running
// (dependent launch).  Publishes the frame's cell counters.
constexpr int kWallThreads = 1024, kWallPerThread = 2;
__global__ void __launch_bounds__(kWallThreads)
k_map_wall(int8_t *__restrict__ grid, MapGeom g, FrameResult *__restrict__ res,
           const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams,
           unsigned *__restrict__ wall_bits, int *__restrict__ counters, const Xchg xc, int pose_from_ext)
{
    TraceScope trace_scope(kTrMapWall);
    const float *__restrict__ scan = sp->scan;
    float pose[3];
    if (pose_from_ext) map_pose(res, xc, sp->seq, 1, pose);   // extrema: final before either map kernel was launched
    else { pdl_wait(); map_pose(res, xc, sp->seq, 0, pose); }
    int cx, cy; center_cell(g, pose[0], pose[1], cx, cy);
    int mine = 0;
    for (int j0 = 0; j0 < n_beams; j0 += kWallThreads * kWallPerThread) {      // one pass for up to 2048 beams
        int idx[kWallPerThread]; bool first[kWallPerThread];
#pragma unroll
        for (int u = 0; u < kWallPerThread; u++) {
            const int j = j0 + threadIdx.x + u * kWallThreads;
            first[u] = false; idx[u] = 0;
            float wx, wy;
            if (j < n_beams && beam_hit(g, pose, cx, cy, angle[j], scan[j], wx, wy) &&
                wx >= 0.0f && wx < (float)g.w && wy >= 0.0f && wy < (float)g.h) {
                idx[u] = (int)__fmaf_rn(wx, (float)g.w, wy);
                const unsigned bit = 1u << (idx[u] & 31);
                const unsigned old = atomicOr(&wall_bits[idx[u] >> 5], bit);
                first[u] = !(old & bit);
            }
        }
        if (pose_from_ext) pdl_wait();          // every -1 of k_map_free has landed
        int8_t v[kWallPerThread];
#pragma unroll
        for (int u = 0; u < kWallPerThread; u++) v[u] = first[u] ? grid[idx[u]] : (int8_t)0;
#pragma unroll
        for (int u = 0; u < kWallPerThread; u++)
            if (first[u]) { grid[idx[u]] = clamp_add(v[u], kOccupiedWeight); mine++; }
    }
    __shared__ int s_wall;
    if (threadIdx.x == 0) s_wall = 0;
    __syncthreads();
    if (mine) atomicAdd(&s_wall, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        res->n_free = __ldcg(&counters[0]); res->n_wall = s_wall;
        res->n_slow = __ldcg(&counters[2]);
        res->n_wide = __ldcg(&counters[6]); res->n_windows = __ldcg(&counters[7]);
        counters[0] = 0; counters[1] = 0; counters[2] = 0;
        counters[10] = counters[10] + 1;         // closes this map update's epoch (k_map_free's cell stamps)
    }
}

}  // namespace pf
