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

constexpr int kStageWindows = 4;            // windows resident per stage
constexpr int kStagedQueueCap = 2048;       // (particle, window) records with at least one uncertain beam, per block

// A window's buffer holds the gather layout (128 rows of pitch 272) in its first kSkewBytes.  The TMA box lands
// DENSE in the tail of the same buffer, at kLandOffset: re-laying row r out writes bytes [272 r, 272 r + 137),
// which ends at or before the start of dense row r (kLandOffset + 128 r) for every r <= 127 -- so with all
// dense rows read into registers first (one block barrier) the re-layout runs in place, no landing buffers are
// needed, and all windows of a stage are in flight at once.
constexpr int kLandOffset = 18432;                    // 144 * 128: TMA destinations are 128-byte aligned
constexpr int kWinBufBytes = 34944;                   // 273 * 128 >= kLandOffset + kTileBytes, kSkewBytes
static_assert(kLandOffset + kTileBytes <= kWinBufBytes && kSkewBytes <= kWinBufBytes, "window buffer");
static_assert(127 * kSkewPitch + 137 <= kLandOffset + 127 * kTileX, "in-place re-layout");

template <int K>
struct StagedSmem {
    alignas(128) int8_t buf[K][kWinBufBytes];         // the stage's windows
    alignas(16) float4 cst[K][kChunkBeams];           // beam constants of the stage's windows
    uint2 queue[kStagedQueueCap];                     // {particle << 8 | window slot, mask of uncertain beams}
    alignas(16) int4 win[kMaxChunks];                 // {x0, y0, beam count, window slot} of order[i]
    int bcum[kMaxChunks + 1];                         // beams before window i (pure counts)
    alignas(8) uint64_t bar;
    int qn, npairs;
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

// VAR selects the address / mask instructions of the gather loop (measured on the B200 with tools/probes/pipe_probe.cu:
// IMAD.HI issues at half the rate of IMAD / IMAD.WIDE / PRMT / LOP3 / LEA; IADD is accepted by both math pipes):
//   0: mad.hi.u32 + predicated OR        1: mad.wide.u32 (high word) + predicated ADD for the mask
//   2: LEA.HI-style shift-add + base add + predicated ADD
template <int THREADS, int PPT, int K, int VAR>
__global__ void __launch_bounds__(THREADS, 1)
k_score_staged(const __grid_constant__ CUtensorMap tmap, const int8_t *__restrict__ grid, MapGeom g,
               const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th, int n,
               const StepParams *__restrict__ sp, const float *__restrict__ angle,
               const TiledWork *__restrict__ tw, int *__restrict__ acc_row, int *__restrict__ counters)
{
    constexpr int GP = THREADS * PPT;           // particles per group
    const float *__restrict__ scan = sp->scan;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StagedSmem<K> &sm = *reinterpret_cast<StagedSmem<K> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    pdl_wait();                                 // k_tile_prep's window table
    const int n_chunks = tw->n_chunks;
    const int dbg = g_staged_dbg;

    for (int i = tid; i < n_chunks; i += THREADS) {
        const int sl = tw->order[i];
        const TileChunk tc = tw->chunk[sl];
        sm.win[i] = make_int4(tc.x0, tc.y0, tc.count, sl);
    }
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.qn = 0; sm.npairs = 0;
    }
    __syncthreads();
    if (tid < 32) {                             // beams before each window
        int run = 0;
        for (int i0 = 0; i0 < n_chunks; i0 += 32) {
            const int i = i0 + lane;
            int v = i < n_chunks ? sm.win[i].z : 0;
            const int own = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (i < n_chunks) sm.bcum[i] = run + v - own;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) sm.bcum[n_chunks] = run;
    }
    __syncthreads();
    const int b_total = sm.bcum[n_chunks];
    if (b_total <= 0) return;
    const int n_groups = (n + GP - 1) / GP;
    const long long total = (long long)n_groups * b_total;
    const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;
    const int n_stages = (n_chunks + K - 1) / K;

    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    const float unit = (float)(1 << kFracT);
    const float irx = (float)(1.0 / (double)g.res_x), iry = (float)(1.0 / (double)g.res_y);
    const float mconst = kMagicT + 0.5f * unit + kGuardT;      // exact
    uint32_t ph = 0;                            // phase of the stage barrier
    int n_inline = 0;                           // pairs this thread re-evaluated inline (queue overflow)
    int s = 0;

    long long u = lo;
    int staged = -1;
    while (u < hi) {
        while (s + 1 < n_stages && (long long)n_groups * sm.bcum[min(K * (s + 1), n_chunks)] <= u) s++;
        const int w0 = K * s, w1 = min(w0 + K, n_chunks);
        if (s != staged) {
            // ---- stage the windows of stage s: all boxes in flight at once, the beam constants meanwhile
            __syncthreads();                    // every warp has left the previous stage's gather loops
            if (!(dbg & 2)) {
            if (tid == 0) {
                mbar_expect_tx(&sm.bar, (uint32_t)(w1 - w0) * kTileBytes);
                for (int k = w0; k < w1; k++) {
                    const int4 wi = sm.win[k];
                    tma_load_2d(sm.buf[k - w0] + kLandOffset, &tmap, wi.y, wi.x, &sm.bar);
                }
            }
            }
            for (int i = tid; i < (w1 - w0) * kChunkBeams; i += THREADS) {
                const int k = i / kChunkBeams, b = i - k * kChunkBeams;
                const int4 wi = sm.win[w0 + k];
                sm.cst[k][b] = b < wi.z ? tw->tconst[wi.w * kChunkBeams + b] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!(dbg & 2)) {
            mbar_wait(&sm.bar, ph); ph ^= 1u;
            // in-place re-layout, 256 half-row tasks per window: all dense rows into registers, barrier, then out
            constexpr int kRounds = (K * 256 + THREADS - 1) / THREADS;
            uint32_t w[kRounds][17];
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const int t = tid + r * THREADS;
                if (t < (w1 - w0) * 256) relayout_load(sm.buf[t >> 8], t & 255, w[r]);
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const int t = tid + r * THREADS;
                if (t < (w1 - w0) * 256) relayout_store(sm.buf[t >> 8], t & 255, w[r]);
            }
            }
            __syncthreads();                    // skewed windows and constants visible
            staged = s;
        }
        const int bc = sm.bcum[w0], bs = sm.bcum[w1] - bc;
        const long long up = u - (long long)n_groups * bc;
        const int gq = (int)(up / bs), b0 = (int)(up - (long long)gq * bs);
        const int len = (int)min((long long)(bs - b0), hi - u), b1 = b0 + len;
        u += len;

        // ---- one piece: particle group gq against beams [b0, b1) of the stage
        float px[PPT], py[PPT], cs[PPT], sn[PPT];
        int acc[PPT];
        bool valid[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            // lanes past the end take a copy of the last particle (results discarded), so every evaluation stays
            // inside the staged windows
            const int p = gq * GP + tid + k * THREADS;
            valid[k] = p < n;
            const int pc = min(p, n - 1);
            px[k] = x[pc]; py[k] = y[pc];
            sincosf(th[pc], &sn[k], &cs[k]);
            acc[k] = 0;
        }
        for (int wi_ = w0; wi_ < w1; wi_++) {
            const int4 wi = sm.win[wi_];
            const int cb0 = sm.bcum[wi_] - bc;
            const int lb0 = max(b0 - cb0, 0), lb1 = min(b1 - cb0, wi.z);
            if (lb0 >= lb1) continue;
            const int kw = wi_ - w0;
            const uint32_t base = smem_u32(sm.buf[kw]);
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
            // Gather loop, per evaluation: 2 FFMA2, PRMT, IMAD.HI (address), LDS.S8, 2 LOP3 (guard band), then the add
            // (certain) or the beam's bit in the particle's mask (uncertain).
#pragma unroll 4
            for (int b = (dbg & 4) ? lb1 : lb0; b < lb1; b++) {
                const float4 q = sm.cst[kw][b];
                const float2 qlo = make_float2(q.x, q.y), qhi = make_float2(q.z, q.w);
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    const float2 t2 = __ffma2_rn(qhi, make_float2(cs[k], cs[k]), __ffma2_rn(qlo, make_float2(sn[k], sn[k]), P[k]));
                    const uint32_t bx = __float_as_uint(t2.x), by = __float_as_uint(t2.y);
                    uint32_t addr;
                    if (VAR == 0) {
                        const uint32_t idx = prmt(bx, by, 0x26BBu);          // x << 24 | y << 16
                        addr = __umulhi(idx, 0x11000u) + base;               // (x*256 + y) * 17 / 16 = x*272 + y + (y >> 4)
                    } else if (VAR == 1) {
                        const uint32_t idx = prmt(bx, by, 0x26BBu);
                        unsigned long long wide;
                        asm("mul.wide.u32 %0, %1, 0x11000;" : "=l"(wide) : "r"(idx));
                        addr = (uint32_t)(wide >> 32) + base;
                    } else {
                        const uint32_t idx = prmt(bx, by, 0xBB26u);          // x * 256 + y
                        addr = idx + (idx >> 4) + base;
                    }
                    int v;
                    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr));
                    // guard band: bits 7..15 == 0 on either axis -> uncertain (the beam's bit), else add the cell
                    if (VAR == 0)
                        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                            "and.b32 t, %2, 0xFF80;\n\t"
                            "setp.ne.u32 p, t, 0;\n\t"
                            "lop3.and.b32 t|p, %3, 0xFF80, 0, 0xC0, p;\n\t"   // LOP3.LUT.PAND: p &= (by & mask) != 0
                            "@!p or.b32 %0, %0, %4;\n\t"
                            "@p add.s32 %1, %1, %5;\n\t}"
                            : "+r"(um[k]), "+r"(acc[k]) : "r"(bx), "r"(by), "r"(bit), "r"(v));
                    else                                                     // a beam's bit is set at most once: ADD == OR
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
            // uncertain pairs (not added above): one record per (particle, window); one shared-memory atomic per warp
            {
                unsigned mk[PPT];
                int cntm = 0;
#pragma unroll
                for (int k = 0; k < PPT; k++) { mk[k] = valid[k] ? um[k] : 0u; cntm += mk[k] ? 1 : 0; }
                if (!(dbg & 8) && __any_sync(0xffffffffu, cntm)) {
                    int inc = cntm;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                    int qb = 0;
                    if (lane == 31) qb = atomicAdd(&sm.qn, inc);
                    qb = __shfl_sync(0xffffffffu, qb, 31);
                    int qi = qb + inc - cntm;
#pragma unroll
                    for (int k = 0; k < PPT; k++) {
                        if (!mk[k]) continue;
                        const int p = gq * GP + tid + k * THREADS;
                        if (qi < kStagedQueueCap) sm.queue[qi] = make_uint2(((unsigned)p << 8) | (unsigned)wi.w, mk[k]);
                        else {
                            const float qt = th[p];
                            n_inline += __popc(mk[k]);
                            for (unsigned m = mk[k]; m; m &= m - 1) {
                                const int j = tw->tbeam[wi.w * kChunkBeams + __ffs(m) - 1];
                                acc[k] += eval_exact(grid, g, c0x, c0y, px[k], py[k], qt, angle[j], scan[j]);
                            }
                        }
                        qi++;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; k++)
            if (valid[k] && acc[k]) atomicAdd(&acc_row[gq * GP + tid + k * THREADS], acc[k]);
    }

    // ---- drain: the uncertain pairs of the whole slice, re-evaluated with the reference's exact expression
    __syncthreads();
    const int qn = (dbg & 1) ? 0 : min(sm.qn, kStagedQueueCap);
    int np = n_inline;
    for (int qi = tid; qi < qn; qi += THREADS) {
        const uint2 e = sm.queue[qi];
        const int p = (int)(e.x >> 8), c = (int)(e.x & 0xffu);
        const float qx = x[p], qy = y[p], qt = th[p];
        int v = 0;
        np += __popc(e.y);
        for (unsigned m = e.y; m; m &= m - 1) {
            const int j = tw->tbeam[c * kChunkBeams + __ffs(m) - 1];
            v += eval_exact(grid, g, c0x, c0y, qx, qy, qt, angle[j], scan[j]);
        }
        if (v) atomicAdd(&acc_row[p], v);
    }
    if (np) atomicAdd(&sm.npairs, np);
    __syncthreads();
    if (tid == 0 && sm.npairs) atomicAdd(&counters[2], sm.npairs);
}

// block shape: 768 threads x 4 particles (default) or 1024 x 2 (PFSLAM_STAGED_THREADS=1024); gather-loop variant
// PFSLAM_STAGED_VARIANT (see k_score_staged)
static int staged_threads()
{
    static int v = 0;
    if (!v) { const char *e = getenv("PFSLAM_STAGED_THREADS"); v = (e && atoi(e) == 1024) ? 1024 : 768; }
    return v;
}
static int staged_variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PFSLAM_STAGED_VARIANT"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
    return v;
}

static StagedKernel staged_kernel()
{
    const int t = staged_threads(), v = staged_variant();
    if (t == 1024) return v == 0 ? (StagedKernel)k_score_staged<1024, 2, kStageWindows, 0> : v == 1 ? (StagedKernel)k_score_staged<1024, 2, kStageWindows, 1>
                                                                                             : (StagedKernel)k_score_staged<1024, 2, kStageWindows, 2>;
    return v == 0 ? (StagedKernel)k_score_staged<768, 4, kStageWindows, 0> : v == 1 ? (StagedKernel)k_score_staged<768, 4, kStageWindows, 1>
                                                                            : (StagedKernel)k_score_staged<768, 4, kStageWindows, 2>;
}

// returns the grid size of k_score_staged = SMs x resident blocks per SM (one full wave), or -1
static int score_staged_setup(int device)
{
    int per_sm = 0, n_sm = 0;
    const size_t smem = sizeof(StagedSmem<kStageWindows>);
    if (cudaFuncSetAttribute(staged_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, staged_kernel(), staged_threads(), smem) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm > 0 && n_sm > 0 ? per_sm * n_sm : -1;
}

}  // namespace pf
