// pf_score_filtered.cuh -- the hot kernel: per-(particle x beam) occupancy-grid scoring that
// returns exactly the integers of the reference's kernEvaluateParticles (src/kernel.cu:257-284)
// at a fraction of its instruction count.
//
// Reference arithmetic per (particle p, beam j)  [kernel.cu:182-187, :262-270, SASS-verified]:
//     rot = angle_j + theta_p;  x = fma(r_j, cosf(rot), px);  gx = roundf(c0x + x / res_x)   (same for y)
//     score += grid[(int)gx * W + (int)gy]   if 0 <= gx < W and 0 <= gy < H
// i.e. ~75 instructions (two libdevice trig polynomials, two IEEE divisions) per evaluation.
//
// Filtered evaluation.  With A = r cos(a)/res, B = r sin(a)/res (per beam, per frame, computed
// once in double) and c = cos(theta), s = sin(theta) (per particle) the pre-rounding cell
// coordinate is  v - c0 = px/res + A c - B s.  It is evaluated in 2^-11-cell fixed point inside
// the float mantissa ("magic number" 1.5*2^23) with two FFMAs per axis:
//     t = fma(A', c, fma(-B', s, P'))     P' = 1.5*2^23 + 2048 (px/res + 1/2) + G
// so that  bits(t) - bits(1.5*2^23)  is  floor(2048 (v - c0 + 1/2)) + G.  Its error against the
// reference's pre-rounding value is bounded by < 3 units of 2^-11 cell for r < 20 m (derivation
// in DESIGN.md), so with the guard G = 4 units the rounded cell is PROVABLY the reference's
// whenever the low 11 bits are >= 8; otherwise ("uncertain", ~0.8 % of evaluations) the pair is
// queued in shared memory and re-evaluated with the reference's exact expression (eval_exact).
// Beams outside the fast path's domain (r >= 20 m, NaN, the 4294967.0 sentinel) are evaluated
// exactly for every particle.  The result is therefore bit-identical to the exact kernel.
#pragma once
#include "pf_kernels2d.cuh"

namespace pf {

constexpr int kFastThreads = 128;        // particles per block (one per thread)
constexpr int kFastSlices = 4;           // beam slices per particle group (grid.y)
constexpr int kMaxBeams = 4096;
constexpr int kQueueCap = 1024;          // uncertain (particle, beam) pairs per block
constexpr float kMagic = 12582912.0f;    // 1.5 * 2^23
constexpr int kFracBits = 11;
constexpr int kGuard = 4;                // units of 2^-11 cell
constexpr float kFastMaxCells = 800.0f;  // |r/res| limit of the fast path
constexpr float kFastMaxPoseCells = 1000.0f;
constexpr float kFastMaxTheta = 8.0f;    // |theta| limit of the fast path (ulp of angle + theta, see pslow)

struct ScoreFilteredWork {
    int nf, ns, pad0, pad1;              // fast / slow beam counts of the current frame
    float4 fconst[kMaxBeams];            // {Ax, -Bx, Ay, By} in 2^-11-cell units
    int fbeam[kMaxBeams];                // original beam index of dense fast beam d
    int slow[kMaxBeams];                 // original indices of the slow beams
};

inline int score_partial_count(int n) { return (n + 31) / 32 > (n + kTile - 1) / kTile ? (n + 31) / 32 : (n + kTile - 1) / kTile; }

// Per-frame beam preparation: classify, compact, and compute the per-beam constants in double.
__global__ void __launch_bounds__(1024)
k_beam_prep(const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams, MapGeom g,
            ScoreFilteredWork *__restrict__ wk)
{
    const float *__restrict__ scan = sp->scan;
    __shared__ int s_warp[32];
    __shared__ int s_base_f, s_base_s;
    if (threadIdx.x == 0) { s_base_f = 0; s_base_s = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j0 = 0; j0 < n_beams; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        const bool valid = j < n_beams;
        float r = valid ? scan[j] : 0.0f;
        const double rx = (double)r / (double)g.res_x, ry = (double)r / (double)g.res_y;
        const bool fast = valid && fabs(rx) < (double)kFastMaxCells && fabs(ry) < (double)kFastMaxCells;  // false for NaN
        const bool slow = valid && !fast;
        // dense positions: fast beams first-come in beam order, slow likewise
        unsigned bf = __ballot_sync(0xffffffffu, fast), bs = __ballot_sync(0xffffffffu, slow);
        if (lane == 0) s_warp[warp] = __popc(bf) | (__popc(bs) << 16);
        __syncthreads();
        int pre_f = 0, pre_s = 0, tot_f = 0, tot_s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            int v = s_warp[w];
            if (w < warp) { pre_f += v & 0xffff; pre_s += v >> 16; }
            tot_f += v & 0xffff; tot_s += v >> 16;
        }
        const int df = s_base_f + pre_f + __popc(bf & ((1u << lane) - 1));
        const int ds = s_base_s + pre_s + __popc(bs & ((1u << lane) - 1));
        if (fast) {
            double a = (double)angle[j];
            double ca = cos(a), sa = sin(a);
            const double u = (double)(1 << kFracBits);
            wk->fconst[df] = make_float4((float)(rx * ca * u), (float)(-rx * sa * u),
                                         (float)(ry * ca * u), (float)(ry * sa * u));
            wk->fbeam[df] = j;
        }
        if (slow) wk->slow[ds] = j;
        __syncthreads();
        if (threadIdx.x == 0) { s_base_f += tot_f; s_base_s += tot_s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { wk->nf = s_base_f; wk->ns = s_base_s; }
}

// Fast scoring: block = 128 particles (one per thread) x one slice of the dense fast-beam list.
// partial layout: [kFastSlices + 1][n]  (last row = slow beams: block row kFastSlices).
__global__ void __launch_bounds__(kFastThreads)
k_score_fast(const int8_t *__restrict__ grid, MapGeom g, const float *__restrict__ x,
             const float *__restrict__ y, const float *__restrict__ th, int n,
             const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams,
             const ScoreFilteredWork *__restrict__ wk, int *__restrict__ partial,
             int *__restrict__ counters)
{
    TraceScope trace_scope(kTrScoreFast);
    const float *__restrict__ scan = sp->scan;
    if (blockIdx.y == kFastSlices) {
        // extra block row: the slow beams (r >= 20 m, sentinel, NaN), exact for every particle; usually 0-3
        const int p = blockIdx.x * kFastThreads + threadIdx.x;
        if (p >= n) return;
        const int ns = wk->ns;
        int acc = 0;
        if (ns > 0) {
            const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
            const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
            const float px = x[p], py = y[p], pth = th[p];
            for (int k = 0; k < ns; k++) { const int j = wk->slow[k]; acc += eval_exact(grid, g, c0x, c0y, px, py, pth, angle[j], scan[j]); }
        }
        partial[(size_t)kFastSlices * n + p] = acc;
        return;
    }
    __shared__ float4 s_const[(kMaxBeams + kFastSlices - 1) / kFastSlices];
    __shared__ unsigned s_queue[kQueueCap];
    __shared__ float s_pose[3][kFastThreads];
    __shared__ int s_acc[kFastThreads];
    __shared__ int s_qn;

    const int tid = threadIdx.x;
    const int p = blockIdx.x * kFastThreads + tid;
    const int per = (n_beams + kFastSlices - 1) / kFastSlices;
    const int nf = wk->nf;
    const int d0 = blockIdx.y * per;
    const int cnt = max(0, min(nf, d0 + per) - d0);
    for (int d = tid; d < cnt; d += kFastThreads) s_const[d] = wk->fconst[d0 + d];
    if (tid == 0) s_qn = 0;
    s_acc[tid] = 0;

    const float c0x = __fdiv_rn(__fmul_rn(0.5f, g.scale_x), g.res_x);
    const float c0y = __fdiv_rn(__fmul_rn(0.5f, g.scale_y), g.res_y);
    const int ox = (int)c0x, oy = (int)c0y;                    // integral (checked on the host)
    const float unit = (float)(1 << kFracBits);
    const float kx = (float)((double)unit / (double)g.res_x), ky = (float)((double)unit / (double)g.res_y);
    const float mx = kMagic + 0.5f * unit + (float)kGuard;      // exact: integers below 2^24
    float px = 0.f, py = 0.f, pth = 0.f;
    if (p < n) { px = x[p]; py = y[p]; pth = th[p]; }
    s_pose[0][tid] = px; s_pose[1][tid] = py; s_pose[2][tid] = pth;
    float sn, cs;
    sincosf(pth, &sn, &cs);
    const float PX = __fmaf_rn(px, kx, mx), PY = __fmaf_rn(py, ky, mx);
    // particles outside the fast domain (never in practice) are scored exactly, beam by beam
    // ... and particles whose heading is large enough for the float rounding of rot = angle + theta (which the
    // reference has and the fast path has not) to eat the guard band: |theta| < 8 keeps that term below 26 units
    // of 2^-16 cell at 800 cells of range
    const bool pslow = !(fabsf(px) * kx < kFastMaxPoseCells * unit && fabsf(py) * ky < kFastMaxPoseCells * unit &&
                         fabsf(pth) < kFastMaxTheta);
    const int basex = __float_as_int(kMagic) - (ox << kFracBits);
    const int basey = __float_as_int(kMagic) - (oy << kFracBits);
    const unsigned gmask = ((1u << kFracBits) - 1u) & ~(2u * kGuard - 1u);
    __syncthreads();

    int acc = 0;
    if (p < n && !pslow) {
#pragma unroll 4
        for (int d = 0; d < cnt; d++) {
            const float4 c = s_const[d];
            const float tx = __fmaf_rn(c.x, cs, __fmaf_rn(c.y, sn, PX));
            const float ty = __fmaf_rn(c.w, cs, __fmaf_rn(c.z, sn, PY));
            const int bx = __float_as_int(tx), by = __float_as_int(ty);
            const bool unc = ((bx & gmask) == 0) | ((by & gmask) == 0);
            if (unc) {
                int q = atomicAdd(&s_qn, 1);
                if (q < kQueueCap) s_queue[q] = ((unsigned)tid << 16) | (unsigned)d;
                else {
                    const int j = wk->fbeam[d0 + d];
                    acc += eval_exact(grid, g, c0x, c0y, px, py, pth, angle[j], scan[j]);
                }
            } else {
                const int cx = (bx - basex) >> kFracBits, cy = (by - basey) >> kFracBits;
                if ((unsigned)cx < (unsigned)g.w && (unsigned)cy < (unsigned)g.h)
                    acc += (int)grid[cx * g.w + cy];
            }
        }
    } else if (p < n) {
        for (int d = 0; d < cnt; d++) {
            const int j = wk->fbeam[d0 + d];
            acc += eval_exact(grid, g, c0x, c0y, px, py, pth, angle[j], scan[j]);
        }
    }
    __syncthreads();
    const int qn = min(s_qn, kQueueCap);
    for (int q = tid; q < qn; q += kFastThreads) {
        const unsigned it = s_queue[q];
        const int pl = (int)(it >> 16), d = (int)(it & 0xffffu);
        const int j = wk->fbeam[d0 + d];
        int v = eval_exact(grid, g, c0x, c0y, s_pose[0][pl], s_pose[1][pl], s_pose[2][pl], angle[j], scan[j]);
        if (v) atomicAdd(&s_acc[pl], v);
    }
    __syncthreads();
    if (p < n) partial[(size_t)blockIdx.y * n + p] = acc + s_acc[tid];
    if (tid == 0 && s_qn) atomicAdd(&counters[2], s_qn);
}

// fit[p] = sum of the slice partials; per-block (1024 particles) min / max-key partials
__global__ void __launch_bounds__(256)
k_score_combine(const int *__restrict__ partial, int n, int gidx0, int *__restrict__ fit,
                int *__restrict__ blk_min, long long *__restrict__ blk_maxkey)
{
    __shared__ int smin[8];
    __shared__ long long smax[8];
    int mn = 0x7fffffff;
    long long mk = (long long)0x8000000000000000ull;
    for (int k = 0; k < 4; k++) {
        const int p = blockIdx.x * kTile + k * 256 + threadIdx.x;
        if (p < n) {
            int s = 0;
#pragma unroll
            for (int r = 0; r <= kFastSlices; r++) s += partial[(size_t)r * n + p];
            fit[p] = s;
            mn = min(mn, s);
            long long t = extrema_key(s, gidx0 + p);
            mk = t > mk ? t : mk;
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
    }
}

static int score_filtered_setup(int device) { (void)device; return 0; }

inline size_t score_partial_ints(int n) { return (size_t)(kFastSlices + 1) * n; }

// returns the number of kernels launched, or -1.  partial: score_partial_ints(n) ints of scratch.
static int score_filtered_launch(const int8_t *grid, MapGeom g, const float *x, const float *y,
                                 const float *th, int n, int gidx0, const StepParams *scan,
                                 const float *angle, int n_beams, int *fit, int *blk_min,
                                 long long *blk_maxkey, Extrema *ext_local, ScoreFilteredWork *wk,
                                 int *partial, int *counters, const Xchg &xc, cudaStream_t stream,
                                 cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr)
{
    k_beam_prep<<<1, 1024, 0, stream>>>(scan, angle, n_beams, g, wk);
    dim3 grid_fast((n + kFastThreads - 1) / kFastThreads, kFastSlices + 1);   // last row = slow beams
    if (ev0) cudaEventRecord(ev0, stream);
    k_score_fast<<<grid_fast, kFastThreads, 0, stream>>>(grid, g, x, y, th, n, scan, angle, n_beams, wk,
                                                         partial, counters);
    if (ev1) cudaEventRecord(ev1, stream);
    const int nblk = (n + kTile - 1) / kTile;
    k_score_combine<<<nblk, 256, 0, stream>>>(partial, n, gidx0, fit, blk_min, blk_maxkey);
    k_extrema<<<1, 1024, 0, stream>>>(blk_min, blk_maxkey, nblk, x, y, th, gidx0, ext_local, xc, scan);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 4;
}

}  // namespace pf
