// pf_score_staged.cuh -- the scoring kernel, second generation: window STAGES resident in shared memory,
// work cut stage-major.
//
// Same contract and the same arithmetic as k_score_tiled (pf_score_tiled.cuh: fixed point in the float
// mantissa, guard band, exact re-evaluation of uncertain pairs -> bit-identical to the reference's
// kernEvaluateParticles, src/kernel.cu:257-284), same per-frame preparation (k_tile_prep).  What changes
// is how the frame's work meets the staged windows.  ncu on k_score_tiled showed 45 % of the stall samples
// and 20 % of the instructions OUTSIDE the gather loop: a block re-staged a 16 KB window (TMA wait, re-layout
// pass, two block barriers, queue flush) for every (window, 1024 particles) item, i.e. per ~18 beams.  Here
//
//   * a STAGE = kStageWindows consecutive windows, all resident at once in the conflict-avoiding pitch-272
//     layout (4 x 34.8 KB of the SM's 227 KB);
//   * the frame's work line is ordered stage-major: [stage][particle group][beam of the stage], and cut into
//     gridDim.x equal slices, one per SM.  A stage is ~1/15 of the frame, a slice 1/148: a block stages
//     one or two stages per frame instead of ~9 windows, and scores ~10 % of the cloud against each;
//   * inside a slice there is no block barrier: a piece = (group of THREADS x PPT particles, beam range of
//     the stage); a thread keeps its particles' sums in registers over all windows of the piece and adds
//     them to acc_row[] with one RED per particle; uncertain pairs go to a shared-memory queue that is
//     drained once, at the end of the block;
//   * address of a cell = ONE permute + ONE multiply-add: PRMT packs the two cell bytes as x<<24 | y<<16, and
//     mad.hi.u32(idx, 17 << 12, window base) = base + x*272 + y + (y >> 4) -- on the FMA pipe, where the old
//     LEA.HI + base add were two ALU-pipe instructions (the ALU pipe was the busiest one);
//   * guard band: two LOP3 with predicate output instead of two shifts + two compares.
#pragma once
#include "pf_score_tiled.cuh"

namespace pf {


// A window's buffer holds the gather layout (128 rows of pitch 272) in its first kSkewBytes.  The TMA box lands
// DENSE in the tail of the same buffer, at kLandOffset: re-laying row r out writes bytes [272 r, 272 r + 137),
// which ends at or before the start of dense row r (kLandOffset + 128 r) for every r <= 127 -- so with all
// dense rows read into registers first (one block barrier) the re-layout runs in place, no landing buffers are
// needed, and all windows of a stage are in flight at once.
constexpr int kLandOffset = 18432;                    // 144 * 128: TMA destinations are 128-byte aligned
constexpr int kWinBufBytes = 34944;                   // 273 * 128 >= kLandOffset + kTileBytes, kSkewBytes
static_assert(kLandOffset + kTileBytes <= kWinBufBytes && kSkewBytes <= kWinBufBytes, "window buffer");
static_assert(127 * kSkewPitch + 137 <= kLandOffset + 127 * kTileX, "in-place re-layout");

constexpr int kWarpQueueCap = 96;           // uncertain (particle, window) records per warp between two drains

template <int K, int NWARPS>
struct StagedSmem {
    alignas(128) int8_t buf[K][kWinBufBytes];         // the stage's windows
    alignas(16) float4 cst[K][kChunkBeams];           // beam constants of the stage's windows
    float2 bm[K][kChunkBeams];                        // {angle, range} of the same beams, for the exact re-evaluations
    uint2 queue[NWARPS][kWarpQueueCap];               // per warp: {particle << 3 | window of the stage, mask of uncertain beams}
    alignas(16) int4 win[kMaxChunks];                 // {x0, y0, beam count, window slot} of order[i]
    int bcum[kMaxChunks + 1];                         // beams before window i (pure counts)
    unsigned long long wbase[K];                      // shared-memory address of buf[k], in the high word
    alignas(16) int4 stage[kMaxStages];               // {kind, first window / list index, beams, units of the line before it (per group)}
    long long lo, hi;                                 // this block's slice of the work line
    alignas(8) uint64_t bar;
    int npairs;
};

// half a dense row (64 bytes) of a landed box -> registers; then registers -> the pitch-272 gather layout
// (see TiledSmem: 16-byte group q of a row at byte 17 q, one PRMT per destination word); `task` in [0, 256)
__device__ __forceinline__ void relayout_load(const int8_t *__restrict__ buf, int task, uint32_t w[17])
{
    const int r = task >> 1, h = task & 1;
    const uint4 *src = reinterpret_cast<const uint4 *>(buf + kLandOffset + r * kTileX + h * 64);
#pragma unroll
    for (int i = 0; i < 4; i++) { const uint4 v = src[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
    w[16] = 0u;
}
__device__ __forceinline__ void relayout_store(int8_t *__restrict__ buf, int task, const uint32_t w[17])
{
    const int r = task >> 1, h = task & 1;
    uint32_t *dst = reinterpret_cast<uint32_t *>(buf + r * kSkewPitch + h * 68);
    dst[0] = w[0];
#pragma unroll
    for (int m = 1; m < 17; m++) dst[m] = prmt(w[m - 1], w[m], skew_selector(m));
}

// timing experiments only (PFSLAM_STAGED_DEBUG, results are then WRONG): 1 = skip the exact drain, 2 = skip the TMA loads
// and the re-layout, 4 = skip the gather loops, 8 = skip the per-window mask / queue code
__device__ int g_staged_dbg = 0;
// dbg & 16: thread 0 of every block stamps %globaltimer at its phase boundaries: [block][0] entry, [1] after the
// dependency wait, [2] prologue done, [3] sum of stage loads, [4] sum of piece set-ups (loads, sincos), [5] sum of gather
// + queue code, [6] sum of REDs, [7] drain, [8] exit, [9] pieces, [10] stages  (nanoseconds)
__device__ unsigned long long g_staged_ts[256 * 12];
__device__ __forceinline__ unsigned long long staged_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Stage kinds.  The beams the tiled path cannot take ride in the same kernel as further stages of the same work line:
// "wide" beams (conservative hit box larger than a window; the 2^-11 fixed-point LDG path of pf_score_filtered.cuh) and
// "slow" beams (outside the fast domain: exact for every particle).  One kernel scores the whole scan: no side kernel
// competing for the SMs (k_score_fast could not even start while 148 x 768 threads x 80 registers were resident), no
// partial rows to combine.  A beam of a wide / slow stage counts kWideWeight / kSlowWeight units of the work line.
// (stage kinds, weights and the table itself: pf_score_tiled.cuh, built by k_tile_prep's last warp)

// VAR selects the address instructions of the gather loop (measured on the B200 with tools/probes/pipe_probe.cu:
// IMAD.HI issues at half the rate of IMAD / IMAD.WIDE / PRMT / LOP3 / LEA; the loop is issue-bound, so instructions
// are what counts):   1: PRMT + IMAD.WIDE + IADD     2: PRMT + IADD + LEA.HI
template <int THREADS, int PPT, int K, int VAR>
__global__ void __launch_bounds__(THREADS, (THREADS >= 512 ? 1 : 3))
k_score_staged(const __grid_constant__ CUtensorMap tmap, const int8_t *__restrict__ grid, MapGeom g,
               const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th, int n,
               const StepParams *__restrict__ sp, const float *__restrict__ angle,
               const TiledWork *__restrict__ tw, const ScoreFilteredWork *__restrict__ wk,
               const float4 *__restrict__ pcs, int *__restrict__ acc_row, int *__restrict__ counters)
{
    TraceScope trace_scope(kTrScore);
    constexpr int GP = THREADS * PPT;           // particles per group
    constexpr int NWARPS = THREADS / 32;
    constexpr int SB = K * kChunkBeams;         // beams per wide / slow stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StagedSmem<K, NWARPS> &sm = *reinterpret_cast<StagedSmem<K, NWARPS> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long ts_entry = staged_now();
    pdl_wait();                                 // k_tile_prep's window table and beam lists (and k_motion's step parameters)
    const float *__restrict__ scan = sp->scan;
    const int n_chunks = tw->n_chunks;
    const int dbg = g_staged_dbg;
    const bool stamp = (dbg & 16) && tid == 0 && blockIdx.x < 256;
    unsigned long long ts_wait = staged_now(), ts_acc[4] = {0, 0, 0, 0}, ts_last = 0;
    int n_pieces = 0, n_stage_loads = 0;

    // the frame's tables, prepared by k_tile_prep: windows, beam prefix, stages, first block of every stage
    const int n_stages = tw->n_stages;
    for (int i = tid; i <= n_chunks; i += THREADS) {
        sm.bcum[i] = tw->bcum[i];
        if (i < n_chunks) sm.win[i] = tw->win[i];
    }
    for (int i = tid; i < n_stages; i += THREADS) sm.stage[i] = tw->stage[i];
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.npairs = 0; sm.lo = 0; sm.hi = 0;
        if (blockIdx.x == 0) { counters[6] = wk->nf; counters[7] = n_chunks; }      // for the frame result (k_map_wall publishes)
    }
    if (tid < K) sm.wbase[tid] = (unsigned long long)smem_u32(sm.buf[tid]) << 32;
    __syncthreads();
    const int n_groups = (n + GP - 1) / GP;
    {
        // this block's slice [lo, hi) of the line: stage-aligned (every stage owns a whole number of blocks, proportional
        // to its work, so that no block pays for staging two stages) when there are enough blocks, else an equal cut
        const int me = (int)blockIdx.x, nblk = (int)gridDim.x;
        if (tw->aligned) {
            for (int st = tid; st < n_stages; st += THREADS) {
                const int first = tw->sfirst[st], next = tw->sfirst[st + 1];
                if (me >= first && me < next) {
                    const int4 sg = sm.stage[st];
                    const long long u0 = (long long)n_groups * sg.w, len = (long long)n_groups * (sg.z * stage_weight(sg.x) + stage_setup(sg.x));
                    sm.lo = u0 + len * (me - first) / (next - first);
                    sm.hi = u0 + len * (me - first + 1) / (next - first);
                }
            }
        } else if (tid == 0) {
            const long long total = (long long)n_groups * tw->u_total;
            sm.lo = total * me / nblk; sm.hi = total * (me + 1) / nblk;
        }
    }
    __syncthreads();
    const unsigned long long ts_prologue = staged_now();
    ts_last = ts_prologue;
    const long long lo = sm.lo, hi = sm.hi;
    if (lo >= hi) return;

    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    const float unit = (float)(1 << kFracT);
    const float irx = (float)(1.0 / (double)g.res_x), iry = (float)(1.0 / (double)g.res_y);
    const float mconst = kMagicT + 0.5f * unit + kGuardT;      // exact
    uint32_t ph = 0;                            // phase of the stage barrier
    int n_exact = 0;                            // pairs this thread re-evaluated with the exact expression
    int qcnt = 0;                               // records in this warp's queue (warp-uniform)
    int s = 0;

    // the warp's queued uncertain pairs, re-evaluated with the reference's exact expression (beam data of the
    // CURRENT stage from shared memory: called before the stage changes, and at the end)
    auto drain = [&]() {
        if (dbg & 1) { qcnt = 0; return; }
        for (int qi = lane; qi < qcnt; qi += 32) {
            const uint2 e = sm.queue[warp][qi];
            const int p = (int)(e.x >> 3), kw = (int)(e.x & 7u);
            const float qx = x[p], qy = y[p], qt = th[p];
            int v = 0;
            n_exact += __popc(e.y);
            for (unsigned m = e.y; m; m &= m - 1) {
                const float2 ar = sm.bm[kw][__ffs(m) - 1];
                v += eval_exact(grid, g, c0x, c0y, qx, qy, qt, ar.x, ar.y);
            }
            if (v) atomicAdd(&acc_row[p], v);
        }
        __syncwarp();
        qcnt = 0;
    };
    // one record per (particle, 32-beam chunk) with uncertain beams, in the warp's own queue
    auto push = [&](unsigned mk, int p, int kw) {
        const unsigned hot = __ballot_sync(0xffffffffu, mk != 0u);
        if (!hot) return;
        if (qcnt + __popc(hot) > kWarpQueueCap) drain();
        if (mk) sm.queue[warp][qcnt + __popc(hot & ((1u << lane) - 1u))] = make_uint2(((unsigned)p << 3) | (unsigned)kw, mk);
        qcnt += __popc(hot);
    };

    long long u = lo;
    int staged = -1;
    while (u < hi) {
        while (s + 1 < n_stages && (long long)n_groups * sm.stage[s + 1].w <= u) s++;
        const int4 st = sm.stage[s];            // {kind, first window / list index, beams, units before}
        const int kind = st.x, wgt = stage_weight(kind);
        const int w0 = st.y, w1 = min(w0 + K, n_chunks);     // tiled stages only
        // the piece: particle group gq against beams [b0, b1) of the stage.  A group's units are the set-up units
        // followed by the beams; a beam belongs to the slice that holds its first unit.
        const int kSetupUnits = stage_setup(kind);
        const int ug = st.z * wgt + kSetupUnits;
        const long long up = u - (long long)n_groups * st.w;
        const int gq = (int)(up / ug), r0 = (int)(up - (long long)gq * ug);
        const int rlen = (int)min((long long)(ug - r0), hi - u);
        const int b0 = (max(r0 - kSetupUnits, 0) + wgt - 1) / wgt, b1 = (max(r0 + rlen - kSetupUnits, 0) + wgt - 1) / wgt;
        u += rlen;
        if (b0 >= b1) continue;                 // set-up units only: the beams of this group are another block's
        if (s != staged) {
            // ---- stage s: window boxes in flight at once (tiled), the beam constants and the beam table meanwhile
            if (staged >= 0) drain();           // while the previous stage's beam table is still there
            __syncthreads();                    // every warp has left the previous stage's gather loops
            if (kind == kStTiled) {
                if (tid == 0 && !(dbg & 2)) {
                    mbar_expect_tx(&sm.bar, (uint32_t)(w1 - w0) * kTileBytes);
                    for (int k = w0; k < w1; k++) {
                        const int4 wi = sm.win[k];
                        tma_load_2d(sm.buf[k - w0] + kLandOffset, &tmap, wi.y, wi.x, &sm.bar);
                    }
                }
                for (int i = tid; i < (w1 - w0) * kChunkBeams; i += THREADS) {
                    const int k = i / kChunkBeams, b = i - k * kChunkBeams;
                    const int4 wi = sm.win[w0 + k];
                    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                    float2 ar = make_float2(0.f, 0.f);
                    if (b < wi.z) {
                        c = tw->tconst[wi.w * kChunkBeams + b];
                        const int j = tw->tbeam[wi.w * kChunkBeams + b];
                        ar = make_float2(angle[j], scan[j]);
                    }
                    sm.cst[k][b] = c; sm.bm[k][b] = ar;
                }
                if (!(dbg & 2)) {
                    mbar_wait(&sm.bar, ph); ph ^= 1u;
                    // in-place re-layout, 256 half-row tasks per window; per round of THREADS tasks: dense half rows into
                    // registers, barrier, then out (a round never shares a window with the next one's loads: THREADS % 256 == 0)
                    static_assert(THREADS % 256 == 0, "re-layout rounds must cover whole windows");
#pragma unroll 1
                    for (int t = tid; t < ((K * 256 + THREADS - 1) / THREADS) * THREADS; t += THREADS) {   // same trip count for every thread
                        const bool mine = t < (w1 - w0) * 256;
                        uint32_t w[17];
                        if (mine) relayout_load(sm.buf[t >> 8], t & 255, w);
                        __syncthreads();
                        if (mine) relayout_store(sm.buf[t >> 8], t & 255, w);
                    }
                }
            } else {
                for (int i = tid; i < st.z; i += THREADS) {
                    const int j = kind == kStWide ? wk->fbeam[st.y + i] : wk->slow[st.y + i];
                    if (kind == kStWide) sm.cst[i / kChunkBeams][i % kChunkBeams] = wk->fconst[st.y + i];
                    sm.bm[i / kChunkBeams][i % kChunkBeams] = make_float2(angle[j], scan[j]);
                }
            }
            __syncthreads();                    // windows, constants and the beam table visible
            staged = s;
            if (stamp) { const unsigned long long t = staged_now(); ts_acc[0] += t - ts_last; ts_last = t; n_stage_loads++; }
        }
        float px[PPT], py[PPT], cs[PPT], sn[PPT];
        int acc[PPT];
        bool valid[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            // lanes past the end take a copy of the last particle (results discarded), so every evaluation stays
            // inside the staged windows
            const int p = gq * GP + tid + k * THREADS;
            valid[k] = p < n;
            const float4 v = pcs[min(p, n - 1)];             // {x, y, cos(theta), sin(theta)} from k_motion
            px[k] = v.x; py[k] = v.y; cs[k] = v.z; sn[k] = v.w;
            acc[k] = 0;
            // the next piece is almost always the next group of the same stage: its poses on their way into L1
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pcs + min(p + GP, n - 1)));
        }
        if (stamp) { const unsigned long long t = staged_now(); ts_acc[1] += t - ts_last; ts_last = t; n_pieces++; }
        if (kind == kStTiled) {
            const int bc = sm.bcum[w0];
            for (int wi_ = w0; wi_ < w1; wi_++) {
                const int4 wi = sm.win[wi_];
                const int cb0 = sm.bcum[wi_] - bc;
                const int lb0 = max(b0 - cb0, 0), lb1 = min(b1 - cb0, wi.z);
                if (lb0 >= lb1) continue;
                const int kw = wi_ - w0;
                const uint32_t base = (uint32_t)(sm.wbase[kw] >> 32);
                const float offx = __fsub_rn(c0x, (float)wi.x), offy = __fsub_rn(c0y, (float)wi.y);
                float2 P[PPT];
                unsigned um[PPT];
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    P[k] = make_float2(__fmaf_rn(__fmaf_rn(px[k], irx, offx), unit, mconst),
                                       __fmaf_rn(__fmaf_rn(py[k], iry, offy), unit, mconst));
                    um[k] = 0u;
                }
                unsigned bit = 1u << lb0;
                // Gather loop, per evaluation: 2 FFMA2, PRMT, address (2), LDS.S8, 2 LOP3 (guard band), then the add (certain)
                // or the beam's bit in the particle's mask (uncertain; a bit is set at most once, so ADD == OR and the
                // instruction can go to either math pipe).
#pragma unroll 4
                for (int b = (dbg & 4) ? lb1 : lb0; b < lb1; b++) {
                    const float4 q = sm.cst[kw][b];
                    const float2 qlo = make_float2(q.x, q.y), qhi = make_float2(q.z, q.w);
#pragma unroll
                    for (int k = 0; k < PPT; k++) {
                        const float2 t2 = __ffma2_rn(qhi, make_float2(cs[k], cs[k]), __ffma2_rn(qlo, make_float2(sn[k], sn[k]), P[k]));
                        const uint32_t bx = __float_as_uint(t2.x), by = __float_as_uint(t2.y);
                        uint32_t addr;
                        if (VAR == 1) {
                            const uint32_t idx = prmt(bx, by, 0x26BBu);          // x << 24 | y << 16
                            unsigned long long wide;                             // high word: (x*256 + y) * 17 / 16
                            asm("mul.wide.u32 %0, %1, 0x11000;" : "=l"(wide) : "r"(idx));
                            addr = (uint32_t)(wide >> 32) + base;                //          = x*272 + y + (y >> 4)
                        } else {
                            const uint32_t idx = prmt(bx, by, 0xBB26u);          // x * 256 + y
                            addr = idx + (idx >> 4) + base;
                        }
                        int v;
                        asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr));
                        // guard band: bits 7..15 == 0 on either axis -> uncertain (the beam's bit), else add the cell
                        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                            "and.b32 t, %2, 0xFF80;\n\t"
                            "setp.ne.u32 p, t, 0;\n\t"
                            "lop3.and.b32 t|p, %3, 0xFF80, 0, 0xC0, p;\n\t"   // LOP3.LUT.PAND: p &= (by & mask) != 0
                            "@!p add.u32 %0, %0, %4;\n\t"
                            "@p add.s32 %1, %1, %5;\n\t}"
                            : "+r"(um[k]), "+r"(acc[k]) : "r"(bx), "r"(by), "r"(bit), "r"(v));
                    }
                    bit <<= 1;
                }
                if (!(dbg & 8)) {
#pragma unroll
                    for (int k = 0; k < PPT; k++) push(valid[k] ? um[k] : 0u, gq * GP + tid + k * THREADS, kw);
                    __syncwarp();
                    if (qcnt >= 32) drain();    // a full warp's worth: re-evaluate now, while the other warps gather
                }
            }
        } else if (kind == kStWide) {
            // wide beams: the filtered LDG path (pf_score_filtered.cuh k_score_fast): 2^-11-cell fixed point on absolute
            // cell coordinates, cells from global memory (L1 / L2); particles outside its domain take every beam exactly
            const float unitF = (float)(1 << kFracBits);
            const float kx = (float)((double)unitF / (double)g.res_x), ky = (float)((double)unitF / (double)g.res_y);
            const float mx = kMagic + 0.5f * unitF + (float)kGuard;
            const int basex = __float_as_int(kMagic) - ((int)c0x << kFracBits), basey = __float_as_int(kMagic) - ((int)c0y << kFracBits);
            const unsigned gmask = ((1u << kFracBits) - 1u) & ~(2u * kGuard - 1u);
            float PX[PPT], PY[PPT];
            bool pslow[PPT];
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                PX[k] = __fmaf_rn(px[k], kx, mx); PY[k] = __fmaf_rn(py[k], ky, mx);
                const float pth = th[min(gq * GP + tid + k * THREADS, n - 1)];
                pslow[k] = !(fabsf(px[k]) * kx < kFastMaxPoseCells * unitF && fabsf(py[k]) * ky < kFastMaxPoseCells * unitF &&
                             fabsf(pth) < kFastMaxTheta);
            }
            for (int c0 = b0 & ~(kChunkBeams - 1); c0 < b1; c0 += kChunkBeams) {
                const int kw = c0 / kChunkBeams, lb0 = max(b0, c0) - c0, lb1 = min(b1, c0 + kChunkBeams) - c0;
                unsigned um[PPT];
#pragma unroll
                for (int k = 0; k < PPT; k++) um[k] = 0u;
#pragma unroll 2
                for (int b = lb0; b < lb1; b++) {
                    const float4 c = sm.cst[kw][b];
#pragma unroll
                    for (int k = 0; k < PPT; k++) {
                        const float tx = __fmaf_rn(c.x, cs[k], __fmaf_rn(c.y, sn[k], PX[k]));
                        const float ty = __fmaf_rn(c.w, cs[k], __fmaf_rn(c.z, sn[k], PY[k]));
                        const int bx = __float_as_int(tx), by = __float_as_int(ty);
                        const bool unc = (((unsigned)bx & gmask) == 0u) | (((unsigned)by & gmask) == 0u) | pslow[k];
                        if (unc) um[k] |= 1u << b;
                        else {
                            const int cx = (bx - basex) >> kFracBits, cy = (by - basey) >> kFracBits;
                            if ((unsigned)cx < (unsigned)g.w && (unsigned)cy < (unsigned)g.h) acc[k] += (int)grid[cx * g.w + cy];
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < PPT; k++) push(valid[k] ? um[k] : 0u, gq * GP + tid + k * THREADS, kw);
                __syncwarp();
            }
        } else {
            // slow beams (r >= 20 m, the sentinel, NaN): the reference's exact expression for every particle
            float pth[PPT];
#pragma unroll
            for (int k = 0; k < PPT; k++) pth[k] = th[min(gq * GP + tid + k * THREADS, n - 1)];
            for (int b = b0; b < b1; b++) {
                const float2 ar = sm.bm[b / kChunkBeams][b % kChunkBeams];
#pragma unroll
                for (int k = 0; k < PPT; k++) acc[k] += eval_exact(grid, g, c0x, c0y, px[k], py[k], pth[k], ar.x, ar.y);
            }
        }
        if (stamp) { const unsigned long long t = staged_now(); ts_acc[2] += t - ts_last; ts_last = t; }
#pragma unroll
        for (int k = 0; k < PPT; k++)
            if (valid[k] && acc[k]) atomicAdd(&acc_row[gq * GP + tid + k * THREADS], acc[k]);
        if (stamp) { const unsigned long long t = staged_now(); ts_acc[3] += t - ts_last; ts_last = t; }
    }

    // ---- the rest of the warp's uncertain pairs
    drain();
    if (n_exact) atomicAdd(&sm.npairs, n_exact);
    __syncthreads();
    if (tid == 0 && sm.npairs) atomicAdd(&counters[2], sm.npairs);
    if (stamp) {
        unsigned long long *o = g_staged_ts + blockIdx.x * 12;
        const unsigned long long t = staged_now();
        o[0] = ts_entry; o[1] = ts_wait; o[2] = ts_prologue; o[3] = ts_acc[0]; o[4] = ts_acc[1]; o[5] = ts_acc[2]; o[6] = ts_acc[3];
        o[7] = t - ts_last; o[8] = t; o[9] = (unsigned long long)n_pieces; o[10] = (unsigned long long)n_stage_loads;
        o[11] = (unsigned long long)s | ((unsigned long long)sm.stage[s].x << 8) | ((unsigned long long)sm.stage[s].z << 16) | ((unsigned long long)(hi - lo) << 32);
    }
}

// block shape: 768 threads x 4 particles (default) or 1024 x 2 (PFSLAM_STAGED_THREADS=1024); gather-loop variant
// PFSLAM_STAGED_VARIANT (see k_score_staged)
static int staged_threads()
{
    static int v = 0;
    if (!v) { const char *e = getenv("PFSLAM_STAGED_THREADS"); v = (e && atoi(e) == 1024) ? 1024 : (e && atoi(e) == 256) ? 256 : 768; }
    return v;
}
// 768 x 4 particles x 5 resident windows, 1 block per SM (default); 1024 x 2 x 5; or 256 x 4 x 1 window, 3 blocks per SM
static int staged_windows() { return staged_threads() == 256 ? 1 : kStageWindows; }
static int staged_ppt() { return staged_threads() == 1024 ? 2 : 4; }
static int staged_variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PFSLAM_STAGED_VARIANT"); v = e ? atoi(e) : 2; if (v != 1) v = 2; }
    return v;
}

static StagedKernel staged_kernel()
{
    const int t = staged_threads(), v = staged_variant();
    if (t == 256) return v == 1 ? (StagedKernel)k_score_staged<256, 4, 1, 1> : (StagedKernel)k_score_staged<256, 4, 1, 2>;
    if (t == 1024) return v == 1 ? (StagedKernel)k_score_staged<1024, 2, kStageWindows, 1> : (StagedKernel)k_score_staged<1024, 2, kStageWindows, 2>;
    return v == 1 ? (StagedKernel)k_score_staged<768, 4, kStageWindows, 1> : (StagedKernel)k_score_staged<768, 4, kStageWindows, 2>;
}
static size_t staged_smem_bytes()
{
    const int t = staged_threads();
    return t == 256 ? sizeof(StagedSmem<1, 8>) : t == 1024 ? sizeof(StagedSmem<kStageWindows, 32>) : sizeof(StagedSmem<kStageWindows, 24>);
}

// returns the grid size of k_score_staged = SMs x resident blocks per SM (one full wave), or -1
static int score_staged_setup(int device)
{
    int per_sm = 0, n_sm = 0;
    const size_t smem = staged_smem_bytes();
    if (cudaFuncSetAttribute(staged_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, staged_kernel(), staged_threads(), smem) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm > 0 && n_sm > 0 ? per_sm * n_sm : -1;
}

}  // namespace pf
