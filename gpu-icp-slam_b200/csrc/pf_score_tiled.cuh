// pf_score_tiled.cuh -- scoring: TMA-staged occupancy-grid windows in shared memory.
//
// Same contract as pf_score_filtered.cuh (bit-identical to the reference's kernEvaluateParticles,
// src/kernel.cu:257-284), restructured for the B200 memory system:
//
//   * the scan is cut into groups of 32 consecutive beams; consecutive beams hit a contiguous wall
//     segment, so the hit cells of a group -- for EVERY particle of the cloud (conservative interval
//     bound from the cloud's pose bounds) -- fit a 128x128-cell window (a second window takes the
//     group's leftovers).  k_tile_prep places the windows, one warp per group;
//   * a window is staged into shared memory by one TMA 2D box load per (block, window) behind an
//     mbarrier, the next one in flight while the current one is scored; out-of-map parts are
//     zero-filled by the TMA unit, which IS the reference's bounds test (kernel.cu:248);
//   * 2^-16-cell fixed point relative to the window origin inside the float mantissa (magic 2^23):
//     two packed FFMA2 give both axes; the cell byte of each axis is byte 2 of the result, so ONE
//     PRMT builds the offset x*256 + y, one LEA.HI skews it to the conflict-avoiding pitch-272 layout
//     and one LDS.S8 fetches the cell;
//   * the rounding guard band (+-64 units = +-9.8e-4 cell, error bound 38 units) is
//     "bits 7..15 == 0"; uncertain pairs set a bit in a per-particle mask (one bit per beam of the
//     window), are queued in shared memory and re-evaluated once per particle group with the
//     reference's exact expression;
//   * the frame's (particle group, window) items are cut into one equal share per resident block
//     (grid = SMs x blocks/SM, a single full wave);
//   * beams whose conservative box does not fit a window (depth discontinuities, wide clouds) go to
//     the LDG kernel (k_score_fast) on a side branch of the step graph, out-of-domain beams to its
//     exact row.
#pragma once
#include <cuda.h>

#include "pf_score_filtered.cuh"

namespace pf {

constexpr int kTileX = 128;                 // window rows (x cells) and usable columns (y cells)
constexpr int kTileBytes = kTileX * kTileX; // dense TMA box: 128 (y, contiguous) x 128 (x) bytes
constexpr int kSkewPitch = 272;             // row pitch of the gather copy (bank = 4 x + y/4 mod 32, see TiledSmem)
constexpr int kSkewBytes = kTileX * kSkewPitch + 16;
constexpr int kChunkBeams = 32;
constexpr int kMaxGroups = 64;              // groups of 32 consecutive fast beams (2048 beams)
constexpr int kPrepPasses = 4;              // windows a group of 32 beams may get (staging a window is cheap for k_score_staged)
constexpr int kMaxChunks = kPrepPasses * kMaxGroups;
constexpr int kTiledGroup = 1024;           // particles per block: 256 threads x 4, or 512 threads x 2 (PFSLAM_TILED_THREADS)
constexpr int kTiledQueueCap = 1024;        // (particle, window) records with at least one uncertain beam, per group share
constexpr float kMagicT = 8388608.0f;       // 2^23
constexpr int kFracT = 16;
constexpr float kGuardT = 64.0f;            // units of 2^-16 cell; band test = bits 7..15 zero (error bound 38)
constexpr double kRotBudgetUnits = 26.0;    // share of the rounding of rot = angle + theta in the 64-unit guard band
constexpr int kBoxMargin = 2;               // cells added around the conservative hit box
constexpr int kWindowCostBeams = 8;         // fixed cost of a window (TMA wait, re-layout, two barriers) in beam units
constexpr int kShareFactor = 1;             // k_score_tiled: shares per block (first static, rest claimed on completion); 2 measured no better than 1: the tail is not block imbalance

// stage table of k_score_staged (pf_score_staged.cuh), built by k_tile_prep's last warp
constexpr int kStageWindows = 5;            // windows resident per stage (5 x 34.9 KB of shared memory)
constexpr int kMaxStages = 384;             // tiled + wide + slow stages of a frame (at most, with one window per stage: 256 + 64 + 64)
enum { kStTiled = 0, kStWide = 1, kStSlow = 2 };
// Units of the work line, calibrated with the in-kernel stamps (tools/score_probe.py, PFSLAM_STAGED_DEBUG=16): a tiled
// beam over one particle group = 1 (0.26 us); a wide beam (LDG path) = 4; a slow beam (exact expression) = 6; and every
// (stage, particle group) piece costs a fixed set-up -- pose loads, per-window offsets, queue pushes, the REDs, mostly
// latency -- of ~3 us on a tiled stage (5 windows) and ~1.5 us on a wide / slow one.
constexpr int kWideWeight = 4, kSlowWeight = 6;
__host__ __device__ constexpr int stage_weight(int kind) { return kind == kStTiled ? 1 : kind == kStWide ? kWideWeight : kSlowWeight; }
__host__ __device__ constexpr int stage_setup(int kind) { return kind == kStTiled ? 12 : 6; }

struct TileChunk { int x0, y0, count, pad; };

struct TiledWork {
    int n_chunks, pad0, pad1, pad2;
    TileChunk chunk[kMaxChunks];
    int order[kMaxChunks];                     // non-empty window slots, in beam order
    int cum[kMaxChunks + 1];                   // work (beams + per-window cost) before order[i]; cum[n_chunks] = total
    float4 tconst[kMaxChunks * kChunkBeams];   // {-Bx, Ay, Ax, By} in 2^-16-cell units (two FFMA2 operand pairs)
    int tbeam[kMaxChunks * kChunkBeams];       // original beam index
    int bounds[8];                             // cloud bounds as ordered ints: xmin,xmax,ymin,ymax,tmin,tmax
    int wide_run, slow_run;                    // running sizes of the wide / slow lists while k_tile_prep's warps append
    int done;                                  // warps of k_tile_prep that have finished this frame
    int stat_wide, stat_slow;                  // beam counts of the frame's wide / slow lists (for the frame result)
    // for k_score_staged: the non-empty windows compacted, their beam prefix, the stage table and the first block of
    // every stage (stage-aligned slices of the work line)
    int4 win[kMaxChunks];                      // {x0, y0, beam count, window slot} of order[i]
    int bcum[kMaxChunks + 1];                  // beams before window i
    int4 stage[kMaxStages];                    // {kind, first window / list index, beams, units of the line before it (per group)}
    int sfirst[kMaxStages + 1];                // first block of stage s; sfirst[n_stages] = blocks
    int n_stages, u_total, aligned;            // aligned: the slices are stage-aligned (enough blocks), else equal cuts
    int share_next;                            // k_score_tiled: next unclaimed share beyond the first wave (reset per frame)
};

__global__ void k_bounds_reset(TiledWork *__restrict__ tw)
{
    if (threadIdx.x < 3) { tw->bounds[2 * threadIdx.x] = 0x7fffffff; tw->bounds[2 * threadIdx.x + 1] = (int)0x80000000; }
}

// pose bounds of the local particle cloud (min/max of x, y, theta)
__global__ void __launch_bounds__(256)
k_cloud_bounds(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th, int n,
               TiledWork *__restrict__ tw, int *__restrict__ acc_row, float4 *__restrict__ pcs)
{
    __shared__ int s_b[6];
    if (threadIdx.x < 6) s_b[threadIdx.x] = (threadIdx.x & 1) ? (int)0x80000000 : 0x7fffffff;
    __syncthreads();
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int k0 = float_order(x[i]), k1 = float_order(y[i]), k2 = float_order(th[i]);
        acc_row[i] = 0;                        // k_motion's other jobs when it did not run this frame
        { float sn, cs; sincosf(th[i], &sn, &cs); pcs[i] = make_float4(x[i], y[i], cs, sn); }
        lo[0] = min(lo[0], k0); hi[0] = max(hi[0], k0);
        lo[1] = min(lo[1], k1); hi[1] = max(hi[1], k1);
        lo[2] = min(lo[2], k2); hi[2] = max(hi[2], k2);
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
        if (threadIdx.x & 1) atomicMax(&tw->bounds[threadIdx.x], s_b[threadIdx.x]);
        else atomicMin(&tw->bounds[threadIdx.x], s_b[threadIdx.x]);
    }
}

// range of cos / sin over the angle interval [p0, p1] (float; the caller pads the box by whole cells)
__device__ __forceinline__ void trig_range(float p0, float p1, float &cmin, float &cmax, float &smin, float &smax)
{
    const float PI = 3.14159265358979323846f, I2PI = 0.15915494309189535f;
    float s0, c0, s1, c1;
    sincosf(p0, &s0, &c0); sincosf(p1, &s1, &c1);
    const float e = 2e-6f;                                  // covers sincosf error and angle rounding
    cmin = fminf(c0, c1) - e; cmax = fmaxf(c0, c1) + e; smin = fminf(s0, s1) - e; smax = fmaxf(s0, s1) + e;
    if (!(p1 - p0 < 6.0f)) { cmin = smin = -1.0f; cmax = smax = 1.0f; return; }
    // extrema inside the interval: cos = +1 at 2k*pi, -1 at (2k+1)*pi; sin = +-1 at +-pi/2 + 2k*pi
    if (floorf(p1 * I2PI) > floorf(p0 * I2PI)) cmax = 1.0f;
    if (floorf((p1 - PI) * I2PI) > floorf((p0 - PI) * I2PI)) cmin = -1.0f;
    if (floorf((p1 - 0.5f * PI) * I2PI) > floorf((p0 - 0.5f * PI) * I2PI)) smax = 1.0f;
    if (floorf((p1 + 0.5f * PI) * I2PI) > floorf((p0 + 0.5f * PI) * I2PI)) smin = -1.0f;
}

// cos / sin of the (fixed) beam angles in double, computed once per engine
__global__ void k_init_beam_trig(const float *__restrict__ angle, int n_beams, double2 *__restrict__ cs)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_beams) { double a = (double)angle[j]; cs[j] = make_double2(cos(a), sin(a)); }
}

// Per-frame preparation, one WARP per group of 32 consecutive beams (up to 64 groups = 2048 beams), no
// block-wide step: classify every beam as
//   slow  (outside the fast domain: r >= 20 m, sentinel, NaN)        -> wk->slow   (k_score_fast, exact row)
//   tiled (conservative hit box fits its group's 128x128 window)    -> tw chunks  (k_score_tiled)
//   wide  (fast domain, but does not fit)                           -> wk->fconst (k_score_fast)
// A group gets a window placed on all of its beams; the beams that do not fit get a second window of
// their own; what still does not fit is "wide".  The last warp to finish compacts the non-empty windows
// (order / cum) and the groups' wide and slow lists, and resets the cloud bounds for the next frame.
constexpr int kPrepWarps = 8;
__global__ void __launch_bounds__(kPrepWarps * 32)
k_tile_prep(const StepParams *__restrict__ sp, const float *__restrict__ angle,
            const double2 *__restrict__ angle_cs, int n_beams, MapGeom g,
            ScoreFilteredWork *__restrict__ wk, TiledWork *tw, int staged_groups, int staged_blocks, int stage_windows)
{
    TraceScope trace_scope(kTrTilePrep);
    pdl_trigger();                              // k_score_tiled's blocks may be staged
    pdl_wait();                                 // k_motion's cloud bounds
    const float *__restrict__ scan = sp->scan;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);       // beam group
    const int n_groups = (n_beams + kChunkBeams - 1) / kChunkBeams;
    if (c >= n_groups) return;
    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    const double pxmin = (double)order_float(__ldcg(&tw->bounds[0])), pxmax = (double)order_float(__ldcg(&tw->bounds[1]));
    const double pymin = (double)order_float(__ldcg(&tw->bounds[2])), pymax = (double)order_float(__ldcg(&tw->bounds[3]));
    const double tmin = (double)order_float(__ldcg(&tw->bounds[4])), tmax = (double)order_float(__ldcg(&tw->bounds[5]));
    // the fixed-point error bound assumes poses within kFastMaxPoseCells of the map centre
    const double lim_x = (double)kFastMaxPoseCells * (double)g.res_x, lim_y = (double)kFastMaxPoseCells * (double)g.res_y;
    const bool cloud_ok = pxmin <= pxmax && pymin <= pymax && tmin <= tmax &&
                          fabs(pxmin) < lim_x && fabs(pxmax) < lim_x && fabs(pymin) < lim_y && fabs(pymax) < lim_y &&
                          fabs(tmin) < 1e3 && fabs(tmax) < 1e3;

    // ---- this lane's beam
    const int j = c * kChunkBeams + lane;
    bool todo = false, slow = false;
    int4 b = make_int4(0, 0, 0, 0);
    double rx = 0, ry = 0, ca = 0, sa = 0;
    if (j < n_beams) {
        const float r = scan[j];
        rx = (double)r / (double)g.res_x; ry = (double)r / (double)g.res_y;
        bool fast = fabs(rx) < (double)kFastMaxCells && fabs(ry) < (double)kFastMaxCells;   // false for NaN
        // The reference rounds rot = angle + theta to float before its trig (kernel.cu:183); the fast paths do
        // not, so their distance from the reference grows with ulp(rot) * r.  The error budget
        // leaves kRotBudgetUnits of 2^-16 cell for that term: beams that exceed it over this cloud's heading
        // range are scored exactly.  (|rot| < 8 never does; a robot that has turned several times does for
        // far beams.)
        if (fast && cloud_ok) {
            const double a = (double)angle[j];
            const double rotmax = fmax(fabs(a + tmin), fabs(a + tmax)) + 1e-5;
            const double half_ulp = ldexp(1.0, max(ilogb(rotmax), 1) - 24);
            if (half_ulp * fmax(fabs(rx), fabs(ry)) * 65536.0 > kRotBudgetUnits) fast = false;
        }
        slow = !fast;
        if (fast) {
            todo = true;
            const double2 t2 = angle_cs[j];
            ca = t2.x; sa = t2.y;
            if (cloud_ok) {
                float cmin, cmax, smin, smax;
                const float a = angle[j];
                // interval of rot = angle + theta over the cloud, padded for float rounding
                trig_range(a + (float)tmin - 4e-6f, a + (float)tmax + 4e-6f, cmin, cmax, smin, smax);
                const double xa = rx * (double)cmin, xb = rx * (double)cmax, ya = ry * (double)smin, yb = ry * (double)smax;
                const double vx0 = (double)c0x + pxmin / (double)g.res_x + fmin(xa, xb);
                const double vx1 = (double)c0x + pxmax / (double)g.res_x + fmax(xa, xb);
                const double vy0 = (double)c0y + pymin / (double)g.res_y + fmin(ya, yb);
                const double vy1 = (double)c0y + pymax / (double)g.res_y + fmax(ya, yb);
                b = make_int4((int)floor(vx0) - kBoxMargin, (int)ceil(vx1) + kBoxMargin,
                              (int)floor(vy0) - kBoxMargin, (int)ceil(vy1) + kBoxMargin);
            } else {
                b = make_int4(-(1 << 20), 1 << 20, -(1 << 20), 1 << 20);   // never fits -> wide
            }
        }
    }

    // ---- up to two windows for the group
    int cnt2[kPrepPasses] = {};
    for (int pass = 0; pass < kPrepPasses; pass++) {
        int x0 = todo ? b.x : 0x3fffffff, x1 = todo ? b.y : -0x3fffffff;
        int y0 = todo ? b.z : 0x3fffffff, y1 = todo ? b.w : -0x3fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        const unsigned tm = __ballot_sync(0xffffffffu, todo);
        const int slot = kPrepPasses * c + pass;
        if (!tm) { if (lane == 0) { TileChunk tc; tc.x0 = 0; tc.y0 = 0; tc.count = 0; tc.pad = 0; tw->chunk[slot] = tc; } continue; }
        // fallback anchor when the span is too large: the middle beam still to place
        const int mid_lane = __fns(tm, 0, (__popc(tm) + 1) / 2);
        const int mx = (__shfl_sync(0xffffffffu, b.x, mid_lane) + __shfl_sync(0xffffffffu, b.y, mid_lane)) / 2;
        const int my = (__shfl_sync(0xffffffffu, b.z, mid_lane) + __shfl_sync(0xffffffffu, b.w, mid_lane)) / 2;
        int ox, oy;
        if ((long long)x1 - x0 < kTileX - 1) ox = x0 - (kTileX - 1 - (x1 - x0)) / 2;
        else ox = mx - kTileX / 2;
        // y is the contiguous dimension of the grid: the TMA box must start on a 16-byte boundary
        // there (measured on B200: unaligned inner coordinates fault), so the y origin is aligned down
        if ((long long)y1 - y0 < kTileX - 1) { const int slack = kTileX - 2 - (y1 - y0); oy = y0 - 1 - max(0, slack - 15) / 2; }
        else oy = my - kTileX / 2;
        ox = max(-100000, min(100000, ox));
        oy = (max(-100000, min(100000, oy)) >> 4) << 4;
        const bool member = todo && b.x >= ox + 1 && b.y <= ox + kTileX - 2 && b.z >= oy + 1 && b.w <= oy + kTileX - 2;
        const unsigned mm = __ballot_sync(0xffffffffu, member);
        if (member) {
            const double u = (double)(1 << kFracT);
            const int k = __popc(mm & ((1u << lane) - 1));
            tw->tconst[slot * kChunkBeams + k] = make_float4((float)(-rx * sa * u), (float)(ry * ca * u),
                                                             (float)(rx * ca * u), (float)(ry * sa * u));
            tw->tbeam[slot * kChunkBeams + k] = j;
            todo = false;
        }
        cnt2[pass] = __popc(mm);
        if (lane == 0) { TileChunk tc; tc.x0 = ox; tc.y0 = oy; tc.count = cnt2[pass]; tc.pad = 0; tw->chunk[slot] = tc; }
    }
    // ---- what is left goes to k_score_fast: appended to its dense lists (the order across groups does not
    // matter, scores are integer sums)
    const unsigned wm = __ballot_sync(0xffffffffu, todo), sm_ = __ballot_sync(0xffffffffu, slow);
    int wbase = 0, sbase = 0;
    if (lane == 0) {
        if (wm) wbase = atomicAdd(&tw->wide_run, __popc(wm));
        if (sm_) sbase = atomicAdd(&tw->slow_run, __popc(sm_));
    }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    sbase = __shfl_sync(0xffffffffu, sbase, 0);
    if (todo) {
        const double u = (double)(1 << kFracBits);
        const int k = wbase + __popc(wm & ((1u << lane) - 1));
        wk->fconst[k] = make_float4((float)(rx * ca * u), (float)(-rx * sa * u), (float)(ry * ca * u), (float)(ry * sa * u));
        wk->fbeam[k] = j;
    }
    if (slow) wk->slow[sbase + __popc(sm_ & ((1u << lane) - 1))] = j;
    int last = 0;
    fence_gpu();                           // every lane's entries before the group is counted as done
    __syncwarp();
    if (lane == 0) last = atomicAdd(&tw->done, 1) == n_groups - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    fence_gpu();

    // ---- last warp: compact.  Non-empty windows in slot order with their work prefix ...
    int n_win = 0;
    {
        int base_m = 0, run = 0, runb = 0;
        for (int s0 = 0; s0 < kPrepPasses * n_groups; s0 += 32) {
            const int sl = s0 + lane;
            const int cv = sl < kPrepPasses * n_groups ? __ldcg(&tw->chunk[sl].count) : 0;
            const bool ne = cv > 0;
            const unsigned bm = __ballot_sync(0xffffffffu, ne);
            int inc = ne ? cv + kWindowCostBeams : 0;         // work weight of the window
            int incb = cv;
            const int wv = inc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o), ub = __shfl_up_sync(0xffffffffu, incb, o);
                if (lane >= o) { inc += u; incb += ub; }
            }
            if (ne) {
                const int pos = base_m + __popc(bm & ((1u << lane) - 1));
                tw->order[pos] = sl;
                tw->cum[pos] = run + inc - wv;
                tw->win[pos] = make_int4(__ldcg(&tw->chunk[sl].x0), __ldcg(&tw->chunk[sl].y0), cv, sl);
                tw->bcum[pos] = runb + incb - cv;
            }
            base_m += __popc(bm);
            run += __shfl_sync(0xffffffffu, inc, 31);
            runb += __shfl_sync(0xffffffffu, incb, 31);
        }
        if (lane == 0) { tw->cum[base_m] = run; tw->n_chunks = base_m; tw->bcum[base_m] = runb; }
        n_win = base_m;
    }
    const int nf = atomicAdd(&tw->wide_run, 0), ns = atomicAdd(&tw->slow_run, 0);
    // ... and the stage table of k_score_staged: tiled stages of kStageWindows windows, then the wide and the slow beams
    // in stages of kStageWindows * 32 beams; per stage the units of the work line before it, and its first block
    if (staged_groups > 0) {
        fence_gpu(); __syncwarp();
        const int SB = stage_windows * kChunkBeams;
        const int n_t = (n_win + stage_windows - 1) / stage_windows, n_w = (nf + SB - 1) / SB, n_s = (ns + SB - 1) / SB;
        const int n_st = min(n_t + n_w + n_s, kMaxStages);
        int run = 0;
        for (int s0 = 0; s0 < n_st; s0 += 32) {
            const int st = s0 + lane;
            int kind = 0, first = 0, nb = 0;
            if (st < n_st) {
                if (st < n_t) { kind = kStTiled; first = st * stage_windows; nb = __ldcg(&tw->bcum[min(first + stage_windows, n_win)]) - __ldcg(&tw->bcum[first]); }
                else if (st < n_t + n_w) { kind = kStWide; first = (st - n_t) * SB; nb = min(SB, nf - first); }
                else { kind = kStSlow; first = (st - n_t - n_w) * SB; nb = min(SB, ns - first); }
            }
            int units = st < n_st ? nb * stage_weight(kind) + stage_setup(kind) : 0;
            const int own = units;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, units, o); if (lane >= o) units += u; }
            if (st < n_st) tw->stage[st] = make_int4(kind, first, nb, run + units - own);
            run += __shfl_sync(0xffffffffu, units, 31);
        }
        fence_gpu(); __syncwarp();
        // Blocks per stage: proportional to the stage's units, at least one; the blocks left over by rounding down go,
        // one at a time, to the stage with the most units per block (lane l keeps stages l, l + 32, ...).
        const int aligned = staged_blocks >= 2 * n_st ? 1 : 0;
        constexpr int kPer = kMaxStages / 32;
        int su[kPer], sc[kPer];
        int used = 0;
#pragma unroll
        for (int q = 0; q < kPer; q++) {
            const int st = lane + 32 * q;
            su[q] = 0; sc[q] = 0;
            if (st < n_st) {
                const int4 sg = tw->stage[st];
                su[q] = sg.z * stage_weight(sg.x) + stage_setup(sg.x);
                sc[q] = max(1, (int)((long long)su[q] * staged_blocks / max(run, 1)));
            }
            used += sc[q];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) used += __shfl_xor_sync(0xffffffffu, used, o);
        // largest remainders: rank every stage's fractional part among all stages (broadcast by shuffle), the `left`
        // best-ranked stages take one more block.  (An iterative "give to the most loaded stage" loop here cost 10-30 us
        // of single-warp time before the scorer could start -- measured with the in-graph timeline.)
        int aligned_ok = aligned;
        if (aligned) {
            const int left = staged_blocks - used;
            if (left < 0) aligned_ok = 0;          // the "at least one block" floors alone exceed the grid: equal cuts instead
            else if (left > 0) {
                float rem[kPer];
                int rank[kPer];
#pragma unroll
                for (int q = 0; q < kPer; q++) {
                    const float ideal = (float)su[q] * (float)staged_blocks / (float)max(run, 1);
                    rem[q] = (lane + 32 * q < n_st) ? ideal - (float)sc[q] : -1.0e30f;   // stages floored up to 1 rank last
                    rank[q] = 0;
                }
                for (int st2 = 0; st2 < n_st; st2++) {
                    float r2 = 0.0f;
#pragma unroll
                    for (int q2 = 0; q2 < kPer; q2++) { const float v = __shfl_sync(0xffffffffu, rem[q2], st2 & 31); if ((st2 >> 5) == q2) r2 = v; }
#pragma unroll
                    for (int q = 0; q < kPer; q++) {
                        const int st = lane + 32 * q;
                        if (st < n_st && (r2 > rem[q] || (r2 == rem[q] && st2 < st))) rank[q]++;
                    }
                }
#pragma unroll
                for (int q = 0; q < kPer; q++) if (lane + 32 * q < n_st && rank[q] < left) sc[q]++;
                // left > n_st cannot happen: every floor is within one of its ideal, so at most n_st blocks are left
            }
        }
        // first block of every stage = exclusive prefix of the counts in stage order
        {
            int base = 0;
#pragma unroll
            for (int q = 0; q < kPer; q++) {
                int v = sc[q];
                const int own = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                const int st = lane + 32 * q;
                if (st < n_st) tw->sfirst[st] = base + v - own;
                base += __shfl_sync(0xffffffffu, v, 31);
            }
        }
        if (lane == 0) { tw->n_stages = n_st; tw->u_total = run; tw->aligned = aligned_ok; tw->sfirst[n_st] = staged_blocks; }
    }
    if (lane == 0) {
        wk->nf = nf; wk->ns = ns;
        tw->stat_wide = nf; tw->stat_slow = ns;
        // consumed: reset the cloud bounds for the next frame's k_motion, and the per-frame counters
        for (int q = 0; q < 3; q++) { tw->bounds[2 * q] = 0x7fffffff; tw->bounds[2 * q + 1] = (int)0x80000000; }
        tw->wide_run = 0; tw->slow_run = 0; tw->done = 0; tw->share_next = 0;
    }
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Shared-memory gather layout.  The PRMT offset is idx = x*256 + y; a row pitch of 256 B would put
// every row in the same banks (measured: ~7 wavefronts per LDS).  One LEA.HI turns it into
//     addr = idx + (idx >> 4) = x*272 + y + (y >> 4)
// i.e. pitch 272 with every 16-byte group of a row displaced by one more byte (group q starts at 17 q).
// bank = (4 x + y/4) mod 32: the 32 particles of a warp hit a footprint of a few rows x a few words,
// which this lattice keeps apart (simulated on the fixture's clouds: 1.2 wavefronts per gather against
// 2.2 for pitch 260 / shift 6, 2.0 for pitch 288 / shift 3).  The TMA box lands densely in `stage`;
// the block re-lays it out into `skew` once per window: a thread moves half a row (4 groups) with one
// PRMT per destination word (bytes 4m+b of the half row come from source byte 4m+b - q, q = (4m+b)/17).
__host__ __device__ constexpr uint32_t skew_selector(int m)
{
    uint32_t sel = 0;
    for (int b = 0; b < 4; b++) {
        const int k = 4 * m + b, q = k / 17;
        int n = 4 + b - q;                    // byte of the pair (w[m-1], w[m]); gap bytes take any valid source
        if (m == 16 && n > 3) n = 3;          // there is no w[16]: the last byte is a gap
        sel |= (uint32_t)n << (4 * b);
    }
    return sel;
}

struct TiledSmem {
    alignas(128) int8_t stage[kTileBytes];
    alignas(16) float4 cst[2][kChunkBeams];
    uint2 queue[kTiledQueueCap];              // {particle << 8 | window slot, mask of uncertain beams}
    int acc[kTiledGroup];
    alignas(16) int4 win[kMaxChunks];         // {x0, y0, beam count, window slot} of order[i]: the frame's window table
    int cum[kMaxChunks + 1];
    alignas(8) uint64_t bar;
    int qn, npairs;
    int item0, item1;
};

// Tiled scoring.  Work item = (group of 1024 particles, window); the items of a frame, in group-major
// order and weighted by their beam counts, are cut into gridDim.x equal shares -- one per resident block
// (grid = SMs x blocks/SM, a single full wave), so every SM carries the same load whatever the number
// of windows.  block = 256 threads x 4 particles.  A block's share spans at most a few particle
// groups; per group it keeps the sums in registers and adds them to acc_row[] (zeroed by k_motion) with
// one atomic per particle.  Uncertain pairs are queued in shared memory across windows and re-evaluated
// exactly once per group, so a window costs two block barriers (re-layout in, re-layout out).
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, (PPT == 4 ? 3 : 2))
k_score_tiled(const __grid_constant__ CUtensorMap tmap, const int8_t *__restrict__ grid, MapGeom g,
              const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th, int n,
              const StepParams *__restrict__ sp, const float *__restrict__ angle,
              const TiledWork *__restrict__ tw, const float4 *__restrict__ pcs, int *__restrict__ acc_row, int *__restrict__ counters)
{
    TraceScope trace_scope(kTrScore);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TiledSmem &sm = *reinterpret_cast<TiledSmem *>(smem_raw);
    // the gather copy of the window is STATIC shared memory: its address is a link-time constant, so the gather's LDS
    // carries it as an immediate and the loop needs no base add
    __shared__ __align__(16) int8_t s_skew[kSkewBytes];
    const int tid = threadIdx.x;
    pdl_wait();                                 // k_tile_prep's window table (and, before it, k_motion's step parameters)
    const float *__restrict__ scan = sp->scan;
    const int n_chunks = tw->n_chunks;

    // the frame's window table and work prefix into shared memory (one parallel round of global loads;
    // everything per-window afterwards is an LDS)
    for (int i = tid; i <= n_chunks; i += THREADS) {
        sm.cum[i] = tw->cum[i];
        if (i < n_chunks) {
            const int sl = tw->order[i];
            const TileChunk tc = tw->chunk[sl];
            sm.win[i] = make_int4(tc.x0, tc.y0, tc.count, sl);
        }
    }
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.qn = 0; sm.npairs = 0;
    }
#pragma unroll
    for (int k = 0; k < PPT; k++) sm.acc[tid + k * THREADS] = 0;
    __syncthreads();
    // The frame's work line is cut into kShareFactor x gridDim.x equal shares.  Every block starts on share blockIdx.x;
    // a block that finishes claims the next unclaimed one (tw->share_next), so the blocks whose windows turned out
    // cheap (few bank conflicts, few uncertain pairs) take work off the tail instead of idling.
    const int n_shares = kShareFactor * (int)gridDim.x;
    int share = blockIdx.x;
    unsigned tma_n = 0;                        // TMA loads this block has waited for so far (mbarrier phase)
    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    const float unit = (float)(1 << kFracT);
    const float irx = (float)(1.0 / (double)g.res_x), iry = (float)(1.0 / (double)g.res_y);
    const float mconst = kMagicT + 0.5f * unit + kGuardT;      // exact
    float px[PPT], py[PPT];
    float2 cc[PPT], ss[PPT];
    int acc[PPT];
    int grp = -1;
    int n_inline = 0;                          // pairs this thread re-evaluated inline (queue overflow)

    // add this group's sums (registers + exact re-evaluations) to acc_row[] and reset them
    auto flush = [&]() {
        __syncthreads();                       // every warp has queued its uncertain pairs of this group
        const int qn = min(sm.qn, kTiledQueueCap);
        int np = n_inline;
        n_inline = 0;
        for (int qi = tid; qi < qn; qi += THREADS) {
            const uint2 e = sm.queue[qi];
            const int pl = (int)(e.x >> 8), c = (int)(e.x & 0xffu);
            const int p = grp * kTiledGroup + pl;
            const float qx = x[p], qy = y[p], qt = th[p];
            int v = 0;
            np += __popc(e.y);
            for (unsigned m = e.y; m; m &= m - 1) {
                const int j = tw->tbeam[c * kChunkBeams + __ffs(m) - 1];
                v += eval_exact(grid, g, c0x, c0y, qx, qy, qt, angle[j], scan[j]);
            }
            if (v) atomicAdd(&sm.acc[pl], v);
        }
        if (np) atomicAdd(&sm.npairs, np);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int pl = tid + k * THREADS, p = grp * kTiledGroup + pl;
            const int v = acc[k] + sm.acc[pl];
            if (p < n && v) atomicAdd(&acc_row[p], v);
            sm.acc[pl] = 0;
        }
        if (tid == 0) { if (sm.npairs) atomicAdd(&counters[2], sm.npairs); sm.qn = 0; sm.npairs = 0; }
    };

    for (;;) {
    if (tid == 0) {
        // share [item0, item1): items whose first work unit falls into its slice of the (groups x frame work) line
        int i0 = 0, i1 = 0;
        const int bt = n_chunks > 0 ? sm.cum[n_chunks] : 0;
        if (bt > 0) {
            const int n_groups = (n + kTiledGroup - 1) / kTiledGroup;
            const long long total = (long long)n_groups * bt;
            const long long lo = total * share / n_shares, hi = total * (share + 1) / n_shares;
            for (int e = 0; e < 2; e++) {
                const long long u = e ? hi : lo;
                const int gq = (int)(u / bt), rem = (int)(u - (long long)gq * bt);
                int a = 0, b = n_chunks;                         // first c with cum[c] >= rem
                while (a < b) { const int mid = (a + b) >> 1; if (sm.cum[mid] < rem) a = mid + 1; else b = mid; }
                (e ? i1 : i0) = gq * n_chunks + a;
            }
        }
        sm.item0 = i0; sm.item1 = i1;
    }
    __syncthreads();
    const int item0 = sm.item0, item1 = sm.item1;
    if (item0 < item1) {
    // prologue: the first window of this share in flight, and its beam constants
    if (tid == 0) {
        mbar_expect_tx(&sm.bar, kTileBytes);
        const int4 w0 = sm.win[item0 % n_chunks];
        tma_load_2d(sm.stage, &tmap, w0.y, w0.x, &sm.bar);
    }
    if (tid < kChunkBeams) {
        const int4 w0 = sm.win[item0 % n_chunks];
        sm.cst[0][tid] = tid < w0.z ? tw->tconst[w0.w * kChunkBeams + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    grp = -1;
    for (int it = item0; it < item1; it++) {
        const int li = it - item0, s = li & 1;
        const int gi = it / n_chunks, ci = it - gi * n_chunks;
        if (gi != grp) {
            if (grp >= 0) flush();
            grp = gi;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                // lanes past the end take a copy of the last particle (results discarded), so every
                // evaluation stays inside the staged window
                const float4 v = pcs[min(grp * kTiledGroup + tid + k * THREADS, n - 1)];   // {x, y, cos, sin} from k_motion
                px[k] = v.x; py[k] = v.y;
                cc[k] = make_float2(v.z, v.z); ss[k] = make_float2(v.w, v.w);
                acc[k] = 0;
            }
        }
        const int4 wi = sm.win[ci];
        const int c = wi.w;
        TileChunk tc; tc.x0 = wi.x; tc.y0 = wi.y; tc.count = wi.z; tc.pad = 0;
        // the next window's beam constants start their trip from L2 now and land in the other buffer after
        // this window's gather loop (nobody reads that buffer until the next barrier pair)
        float4 cst_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < kChunkBeams && it + 1 < item1) {
            const int4 wn = sm.win[(it + 1) % n_chunks];
            if (tid < wn.z) cst_next = tw->tconst[wn.w * kChunkBeams + tid];
        }
        // window-relative fixed-point offsets of this thread's particles
        float2 P[PPT];
        const float offx = __fsub_rn(c0x, (float)tc.x0), offy = __fsub_rn(c0y, (float)tc.y0);
#pragma unroll
        for (int k = 0; k < PPT; k++)
            P[k] = make_float2(__fmaf_rn(__fmaf_rn(px[k], irx, offx), unit, mconst),
                               __fmaf_rn(__fmaf_rn(py[k], iry, offy), unit, mconst));
        mbar_wait(&sm.bar, (tma_n + (unsigned)li) & 1u);   // window landed in `stage`
        __syncthreads();                       // every warp has left the previous window's gather loop
        if (tid < 2 * kTileX) {   // re-lay the dense 128x128 box out: pitch 272, 16-byte group q of a row at byte 17 q
            const int r = tid >> 1, h = tid & 1;
            const uint4 *src = reinterpret_cast<const uint4 *>(sm.stage + r * kTileX + h * 64);
            uint32_t w[17];
#pragma unroll
            for (int i = 0; i < 4; i++) { const uint4 v = src[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
            w[16] = 0u;
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_skew + r * kSkewPitch + h * 68);
            dst[0] = w[0];
#pragma unroll
            for (int m = 1; m < 17; m++) dst[m] = prmt(w[m - 1], w[m], skew_selector(m));
        }
        __syncthreads();                       // skewed window + constants visible; `stage` is free again
        if (tid == 0 && it + 1 < item1) {      // next window streams in while this one is scored
            const int4 wn = sm.win[(it + 1) % n_chunks];
            mbar_expect_tx(&sm.bar, kTileBytes);
            tma_load_2d(sm.stage, &tmap, wn.y, wn.x, &sm.bar);
        }
        const int8_t *tile = s_skew;
        unsigned um[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) um[k] = 0u;
        const int cnt = tc.count;
        unsigned bit = 1u;
        // Main loop, ~12 instructions per evaluation: 2 FFMA2, PRMT, LEA.HI, LDS.S8, 4 for the guard-band
        // test, then either the add (certain) or the beam's bit in the particle's mask (uncertain).
#pragma unroll 4
        for (int b = 0; b < cnt; b++) {
            const float4 q = sm.cst[s][b];
            const float2 lo = make_float2(q.x, q.y), hi = make_float2(q.z, q.w);
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const float2 t2 = __ffma2_rn(hi, cc[k], __ffma2_rn(lo, ss[k], P[k]));
                const uint32_t bx = __float_as_uint(t2.x), by = __float_as_uint(t2.y);
                const uint32_t idx = prmt(bx, by, 0xBB26u);
                const int v = (int)tile[idx + (idx >> 4)];
                // guard band: bits 7..15 == 0 on either axis -> uncertain (the beam's bit; set at most once, so ADD == OR
                // and the instruction can go to either math pipe), else add the cell.  LOP3 with predicate output, the
                // second one and-ing into the first (LOP3.LUT.PAND): 2 instructions for both axes.
                asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                    "and.b32 t, %2, 0xFF80;\n\t"
                    "setp.ne.u32 p, t, 0;\n\t"
                    "lop3.and.b32 t|p, %3, 0xFF80, 0, 0xC0, p;\n\t"
                    "@!p add.u32 %0, %0, %4;\n\t"
                    "@p add.s32 %1, %1, %5;\n\t}"
                    : "+r"(um[k]), "+r"(acc[k]) : "r"(bx), "r"(by), "r"(bit), "r"(v));
            }
            bit <<= 1;
        }
        if (tid < kChunkBeams) sm.cst[s ^ 1][tid] = cst_next;
        // uncertain pairs (not added above): one record per (particle, window) for the group's exact pass.
        // One shared-memory atomic per warp: the lanes' record counts are prefix-summed with shuffles.
        {
            const int lane = tid & 31;
            unsigned mk[PPT];
            int cntm = 0;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                mk[k] = (grp * kTiledGroup + tid + k * THREADS < n) ? um[k] : 0u;
                cntm += mk[k] ? 1 : 0;
            }
            if (__any_sync(0xffffffffu, cntm)) {
                int inc = cntm;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                int base = 0;
                if (lane == 31) base = atomicAdd(&sm.qn, inc);
                base = __shfl_sync(0xffffffffu, base, 31);
                int qi = base + inc - cntm;
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    if (!mk[k]) continue;
                    const int pl = tid + k * THREADS;
                    if (qi < kTiledQueueCap) sm.queue[qi] = make_uint2(((unsigned)pl << 8) | (unsigned)c, mk[k]);
                    else {
                        const float qt = th[grp * kTiledGroup + pl];
                        n_inline += __popc(mk[k]);
                        for (unsigned m = mk[k]; m; m &= m - 1) {
                            const int j = tw->tbeam[c * kChunkBeams + __ffs(m) - 1];
                            acc[k] += eval_exact(grid, g, c0x, c0y, px[k], py[k], qt, angle[j], scan[j]);
                        }
                    }
                    qi++;
                }
            }
        }
    }
    flush();
    tma_n += (unsigned)(item1 - item0);
    }
    // the next unclaimed share, if any
    __syncthreads();
    if (tid == 0) sm.item0 = (int)gridDim.x + atomicAdd(const_cast<int *>(&tw->share_next), 1);
    __syncthreads();
    share = sm.item0;
    __syncthreads();                           // before thread 0 writes the item range of the next share
    if (share >= n_shares) break;
    }
}

// fit[p] = sum of n_rows partial rows; per-256-particle min / max-key partials
__global__ void __launch_bounds__(256)
k_score_combine_rows(const int *__restrict__ partial, int n_rows, int n, int gidx0, int *__restrict__ fit,
                     int *blk_min, long long *blk_maxkey, const float *__restrict__ x,
                     const float *__restrict__ y, const float *__restrict__ th, Extrema *__restrict__ ext_out,
                     int *__restrict__ done_counter, const Xchg xc, const StepParams *__restrict__ sp)
{
    TraceScope trace_scope(kTrCombine);
    __shared__ int smin[8];
    __shared__ long long smax[8];
    __shared__ int s_last;
    pdl_trigger();                              // k_weights_scan's blocks may be staged
    int mn = 0x7fffffff;
    long long mk = (long long)0x8000000000000000ull;
    {
        const int p = blockIdx.x * 256 + threadIdx.x;
        if (p < n) {
            int s = 0;
            for (int r = 0; r < n_rows; r++) s += partial[(size_t)r * n + p];
            fit[p] = s;
            mn = s;
            mk = extrema_key(s, gidx0 + p);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        long long t = __shfl_xor_sync(0xffffffffu, mk, o);
        mk = t > mk ? t : mk;
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { mn = min(mn, smin[w]); mk = smax[w] > mk ? smax[w] : mk; }
        blk_min[blockIdx.x] = mn; blk_maxkey[blockIdx.x] = mk;
        fence_gpu();
        s_last = atomicAdd(done_counter, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: thrust::minmax_element's result over all blocks (k_extrema folded in)
    fence_gpu();
    mn = 0x7fffffff; mk = (long long)0x8000000000000000ull;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        mn = min(mn, __ldcg(&blk_min[i]));
        const long long t = __ldcg(&blk_maxkey[i]);
        mk = t > mk ? t : mk;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        long long t = __shfl_xor_sync(0xffffffffu, mk, o);
        mk = t > mk ? t : mk;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { mn = min(mn, smin[w]); mk = smax[w] > mk ? smax[w] : mk; }
        const int best = (int)(0xFFFFFFFFu - (uint32_t)(mk & 0xFFFFFFFFll));
        Extrema e;
        e.fit_min = mn; e.fit_max = (int)(mk >> 32); e.best_gidx = best;
        e.x = x[best - gidx0]; e.y = y[best - gidx0]; e.th = th[best - gidx0];
        e.pad0 = 0; e.pad1 = 0;
        *done_counter = 0;
        xc_publish_extrema(xc, ext_out, e, sp->seq);
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of the occupancy grid: uint8 [map_w rows (x)][map_h cols (y)], box 128 (y) x 128 (x)
static int make_grid_tensor_map(CUtensorMap *out, const int8_t *grid, int map_w, int map_h)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    if (map_h % 16 != 0) return -2;                       // TMA global stride must be a multiple of 16 B
    cuuint64_t dims[2] = {(cuuint64_t)map_h, (cuuint64_t)map_w};
    cuuint64_t strides[1] = {(cuuint64_t)map_h};
    cuuint32_t box[2] = {(cuuint32_t)kTileX, (cuuint32_t)kTileX};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)grid, dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -3;
}

inline int score_tiled_rows() { return 1 + kFastSlices + 1; }   // tiled accumulator row, wide-beam rows, slow-beam row

// block shape of k_score_tiled: 256 threads x 4 particles (3 blocks / SM) or 512 x 2 (2 blocks / SM, more warps)
static int tiled_threads()
{
    static int v = 0;
    if (!v) { const char *e = getenv("PFSLAM_TILED_THREADS"); v = (e && atoi(e) == 512) ? 512 : 256; }
    return v;
}

// returns the grid size of k_score_tiled = SMs x resident blocks per SM (one full wave), or -1
static int score_tiled_setup(int device)
{
    int per_sm = 0, n_sm = 0;
    if (tiled_threads() == 512) {
        if (cudaFuncSetAttribute(k_score_tiled<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TiledSmem)) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_tiled<512, 2>, 512, sizeof(TiledSmem)) != cudaSuccess) return -1;
    } else {
        if (cudaFuncSetAttribute(k_score_tiled<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TiledSmem)) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_tiled<256, 4>, 256, sizeof(TiledSmem)) != cudaSuccess) return -1;
    }
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm > 0 && n_sm > 0 ? per_sm * n_sm : -1;
}

// the second-generation kernel (pf_score_staged.cuh) has the same signature
typedef void (*StagedKernel)(const CUtensorMap, const int8_t *, MapGeom, const float *, const float *, const float *, int,
                             const StepParams *, const float *, const TiledWork *, const ScoreFilteredWork *, const float4 *, int *, int *);

// returns the number of kernels launched, or -1.  partial: score_tiled_rows()*n ints.
// With an auxiliary stream the wide/slow-beam kernel (k_score_fast) runs NEXT TO the tiled kernel
// (fork after k_tile_prep, join before the row combine); inside a stream capture this becomes two
// parallel branches of the step graph.
static int score_tiled_launch(const CUtensorMap &tmap, const int8_t *grid, MapGeom g, const float *x, const float *y,
                              const float *th, int n, int gidx0, const StepParams *scan, const float *angle, int n_beams,
                              int *fit, int *blk_min, long long *blk_maxkey, Extrema *ext_local,
                              ScoreFilteredWork *wk, TiledWork *tw, const double2 *angle_cs, bool bounds_valid,
                              int *partial, int *counters, const Xchg &xc, int tiled_grid, cudaStream_t stream,
                              cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr,
                              cudaStream_t aux = nullptr, cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr,
                              LapRec *laps = nullptr, StagedKernel staged = nullptr, int staged_threads = 0, size_t staged_smem = 0,
                              int staged_particles = 1024, float4 *pcs = nullptr, int stage_windows = kStageWindows)
{
    int nl = 4;
    if (!bounds_valid) {   // poses were not produced by k_motion this frame (test hooks): recompute
        k_bounds_reset<<<1, 32, 0, stream>>>(tw);
        k_cloud_bounds<<<min(148, (n + 255) / 256), 256, 0, stream>>>(x, y, th, n, tw, partial, pcs);
        nl += 2;
    }
    const int n_prep_groups = (n_beams + kChunkBeams - 1) / kChunkBeams;
    // k_score_staged: one block per SM; small filters get fewer blocks (a few beams of one group each at least)
    const int sgroups = staged ? (n + staged_particles - 1) / staged_particles : 0;
    const int gs = staged ? min(tiled_grid, max(1, ((n + 1023) / 1024) * 8)) : 0;
    launch_k(bounds_valid, k_tile_prep, dim3((n_prep_groups + kPrepWarps - 1) / kPrepWarps), dim3(kPrepWarps * 32), 0, stream,
             scan, angle, angle_cs, n_beams, g, wk, tw, sgroups, gs, stage_windows);                  // dependent of k_motion when it ran just before
    if (laps) laps->mark(stream, kLapTilePrep);
    const int nblk = (n + 255) / 256;
    if (staged) {
        // one kernel scores the whole scan (tiled, wide and slow beams): stage-major slices, one block per SM; small filters
        // get fewer blocks (a few beams of one group each at least)
        if (ev0) cudaEventRecord(ev0, stream);
        launch_k(true, staged, dim3(gs), dim3(staged_threads), staged_smem, stream, tmap, grid, g, x, y, th, n, scan, angle, tw, wk, pcs, partial, counters);
        if (ev1) cudaEventRecord(ev1, stream);
        if (laps) laps->mark(stream, kLapScoreTiled);
        k_score_combine_rows<<<nblk, 256, 0, stream>>>(partial, 1, n, gidx0, fit, blk_min, blk_maxkey, x, y, th, ext_local, counters + 4, xc, scan);
        if (laps) laps->mark(stream, kLapCombine);
        if (cudaGetLastError() != cudaSuccess) return -1;
        return nl - 1;
    }
    if (aux) { cudaEventRecord(ev_fork, stream); cudaStreamWaitEvent(aux, ev_fork, 0); }
    // one full wave; small filters get fewer blocks (an item is the smallest share)
    const int gt = min(tiled_grid, ((n + kTiledGroup - 1) / kTiledGroup) * kMaxChunks);
    if (ev0) cudaEventRecord(ev0, stream);
    if (tiled_threads() == 512)
        launch_k(true, k_score_tiled<512, 2>, dim3(gt), dim3(512), sizeof(TiledSmem), stream, tmap, grid, g, x, y, th, n, scan, angle, tw, pcs, partial, counters);
    else
        launch_k(true, k_score_tiled<256, 4>, dim3(gt), dim3(256), sizeof(TiledSmem), stream, tmap, grid, g, x, y, th, n, scan, angle, tw, pcs, partial, counters);
    if (ev1) cudaEventRecord(ev1, stream);
    if (laps) laps->mark(stream, kLapScoreTiled);
    dim3 gf((n + kFastThreads - 1) / kFastThreads, kFastSlices + 1);          // last row = slow beams
    k_score_fast<<<gf, kFastThreads, 0, aux ? aux : stream>>>(grid, g, x, y, th, n, scan, angle, n_beams, wk,
                                                              partial + (size_t)n, counters);
    if (laps) laps->mark(stream, kLapScoreFast);
    if (aux) { cudaEventRecord(ev_join, aux); cudaStreamWaitEvent(stream, ev_join, 0); }
    k_score_combine_rows<<<nblk, 256, 0, stream>>>(partial, score_tiled_rows(), n, gidx0, fit, blk_min, blk_maxkey,
                                                   x, y, th, ext_local, counters + 4, xc, scan);
    if (laps) laps->mark(stream, kLapCombine);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return nl;
}

}  // namespace pf
