// pf_kernels_kd.cuh -- kd-tree point-cloud path.
//
// Reference: src/kernel.cu:816-1540 (NN walk, KD scoring, ICP, KD map update), :1702-1761 (driver),
// src/kdtree.cpp:25-123 (host build / insert / balance).  Node layout is the reference's
// KDTree::Node (32 B: axis, left, right, parent, value xyzw), so getPCData can lend it unchanged.
//
// Same arithmetic contract as the grid path: explicit IEEE operations in the reference's evaluation
// order (read off its SASS): glm::distance = sqrt(fma(dz,dz, fma(dx,dx, dy*dy))), scan*cos unfused,
// walls + pos a separate add.  Definitions for the reference's undefined reads (Q9-Q11) and the
// closed-form planar ICP rotation are stated in DESIGN.md (kd path).
#pragma once
#include "pf_kernels2d.cuh"

namespace pf {

struct KdNode { int axis, left, right, parent; float x, y, z, w; };   // == KDTree::Node (kdtree.hpp:16-27)

struct KdState {
    int size;           // nodes in the tree
    int n_wall, n_free; // point-cloud sizes of the current frame's map update
    int n_ins;          // nodes inserted this frame
    int pad[4];
};

__device__ __forceinline__ KdNode kd_load(const KdNode *__restrict__ tree, int i)
{
    const int4 a = __ldg(reinterpret_cast<const int4 *>(tree + i));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(tree + i) + 1);
    KdNode n; n.axis = a.x; n.left = a.y; n.right = a.z; n.parent = a.w; n.x = b.x; n.y = b.y; n.z = b.z; n.w = b.w;
    return n;
}
__device__ __forceinline__ KdNode kd_load_cg(const KdNode *tree, int i)
{
    const int4 a = __ldcg(reinterpret_cast<const int4 *>(tree + i));
    const float4 b = __ldcg(reinterpret_cast<const float4 *>(tree + i) + 1);
    KdNode n; n.axis = a.x; n.left = a.y; n.right = a.z; n.parent = a.w; n.x = b.x; n.y = b.y; n.z = b.z; n.w = b.w;
    return n;
}

__device__ __forceinline__ float kd_dist(float qx, float qy, float qz, float nx, float ny, float nz)
{
    const float dx = __fsub_rn(nx, qx), dy = __fsub_rn(ny, qy), dz = __fsub_rn(nz, qz);
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
}

__device__ __forceinline__ float kd_dist2(float qx, float qy, float qz, float nx, float ny, float nz)
{
    const float dx = __fsub_rn(nx, qx), dy = __fsub_rn(ny, qy), dz = __fsub_rn(nz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// kernel.cu:924-972 findCorrespondenceIndexKD: descend by split plane tracking the best node, then
// while improved look across the best node's PARENT plane once.  Q9: stop at the root.
// The reference compares sqrt'ed distances (`d < bestDist`).  sqrt_rn is monotone, so d2 >= best2
// implies d >= bestDist and the test is false without taking the root; only when d2 < best2 is the
// root taken and the reference's comparison applied (two squared distances can round to one root).
template <bool kCoherent>
__device__ __forceinline__ int kd_nn(const KdNode *__restrict__ tree, float qx, float qy, float qz)
{
    KdNode t = kCoherent ? kd_load_cg(tree, 0) : kd_load(tree, 0);
    float best2 = kd_dist2(qx, qy, qz, t.x, t.y, t.z);
    float bestDist = __fsqrt_rn(best2);
    int bestIdx = 0, bestParent = t.parent, head = 0;
    bool explored = false;
    for (;;) {
        while (head >= 0) {
            t = kCoherent ? kd_load_cg(tree, head) : kd_load(tree, head);
            const float d2 = kd_dist2(qx, qy, qz, t.x, t.y, t.z);
            if (d2 < best2) {
                const float d = __fsqrt_rn(d2);
                if (d < bestDist) { bestDist = d; best2 = d2; bestIdx = head; bestParent = t.parent; explored = false; }
            }
            const bool branch = t.axis == 0 ? qx < t.x : t.axis == 1 ? qy < t.y : t.axis == 2 ? qz < t.z : false;
            head = branch ? t.left : t.right;
        }
        if (explored || bestParent < 0) break;
        const KdNode p = kCoherent ? kd_load_cg(tree, bestParent) : kd_load(tree, bestParent);
        bool branch = false; float hd = 0.0f;
        if (p.axis == 0) { branch = qx < p.x; hd = fabsf(__fsub_rn(qx, p.x)); }
        if (p.axis == 1) { branch = qy < p.y; hd = fabsf(__fsub_rn(qy, p.y)); }
        if (p.axis == 2) { branch = qz < p.z; hd = fabsf(__fsub_rn(qz, p.z)); }
        if (!(hd < bestDist)) break;
        head = !branch ? p.left : p.right;
        explored = true;
    }
    return bestIdx;
}

// ---- search shadow of the tree ---------------------------------------------------------------------------
// The scoring walk touches a node per tree level per (particle, beam), and the reference's insert-grown
// trees are deep (measured on train_lidar0: mean depth 20, max ~100 at 6 k nodes).  For the planar
// clouds this engine builds (every z == 0, queries with z == 0) the walk needs only x, y, the split axis
// and the two links of a node, so the scorer walks a 16-byte shadow {x, y, left | axis << 30, right}
// (one LDG.128 per visit) rebuilt from the 32-byte nodes after every topology change; parent and weight
// of the winning node come from the full node.  Arithmetic is unchanged: with dz == 0,
// fma(dz, dz, fma(dx, dx, dy*dy)) == fma(dx, dx, dy*dy) exactly, and the z-level test `qz < node.z` is
// false (always right).  Trees with any z != 0 (only reachable through pfslam_set_kd) use kd_nn<>.
struct KdSearch { float x, y; int lr0, right; };
constexpr int kKdNoLeft = 0x3fffffff;

// kFormat 1: lr0 = left | axis << 30.
// kFormat 2 (descent only, for the branch-free visit of kd_nn_flat2): lr0 = left' << 1 | (axis == 1), where
// left' = right on z-level nodes -- there the reference's test `qz < node.z` is 0 < 0, so the descent
// always turns right and needs no third case; the parent-plane step reads the 32-byte node instead.
template <int kFormat>
__global__ void __launch_bounds__(256)
k_kd_shadow(const KdNode *__restrict__ tree, const KdState *__restrict__ ks, KdSearch *__restrict__ out, int cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || i >= ks->size) return;
    const KdNode n = kd_load_cg(tree, i);
    int lr0;
    if (kFormat == 1) lr0 = (n.left & kKdNoLeft) | (n.axis << 30);
    else lr0 = ((n.axis == 2 ? n.right : n.left) << 1) | (n.axis == 1 ? 1 : 0);
    reinterpret_cast<int4 *>(out)[i] = make_int4(__float_as_int(n.x), __float_as_int(n.y), lr0, n.right);
}

// Same walk over the format-2 shadow with a branch-free visit: the coordinate of the split axis is picked
// with two selects, the left link is one arithmetic shift.  (kd_nn_flat compiles to ~26 instructions and
// six branches per visit; ncu shows the scorer issue-bound under divergence)
__device__ __forceinline__ int kd_nn_flat2(const KdSearch *__restrict__ sh, const KdNode *__restrict__ tree, float qx, float qy)
{
    float best2, bestDist;
    int bestIdx = 0, head = 0;
    {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(sh));
        const float dx = __fsub_rn(__int_as_float(v.x), qx), dy = __fsub_rn(__int_as_float(v.y), qy);
        best2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
        bestDist = __fsqrt_rn(best2);
    }
    bool explored = false;
    for (;;) {
        while (head >= 0) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(sh) + head);
            const float nx = __int_as_float(v.x), ny = __int_as_float(v.y);
            const float dx = __fsub_rn(nx, qx), dy = __fsub_rn(ny, qy);
            const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            if (d2 < best2) {
                const float d = __fsqrt_rn(d2);
                if (d < bestDist) { bestDist = d; best2 = d2; bestIdx = head; explored = false; }
            }
            const bool ay = (v.z & 1) != 0;
            const float qa = ay ? qy : qx, na = ay ? ny : nx;
            head = qa < na ? (v.z >> 1) : v.w;
        }
        if (explored) break;
        const KdNode b = kd_load(tree, bestIdx);
        if (b.parent < 0) break;
        const KdNode p = kd_load(tree, b.parent);
        bool branch = false; float hd = 0.0f;
        if (p.axis == 0) { branch = qx < p.x; hd = fabsf(__fsub_rn(qx, p.x)); }
        if (p.axis == 1) { branch = qy < p.y; hd = fabsf(__fsub_rn(qy, p.y)); }
        if (p.axis == 2) { branch = 0.0f < p.z; hd = fabsf(__fsub_rn(0.0f, p.z)); }
        if (!(hd < bestDist)) break;
        head = !branch ? p.left : p.right;
        explored = true;
    }
    return bestIdx;
}

__device__ __forceinline__ int kd_nn_flat(const KdSearch *__restrict__ sh, const KdNode *__restrict__ tree, float qx, float qy)
{
    float best2, bestDist;
    int bestIdx = 0, head = 0;
    {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(sh));
        const float dx = __fsub_rn(__int_as_float(v.x), qx), dy = __fsub_rn(__int_as_float(v.y), qy);
        best2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
        bestDist = __fsqrt_rn(best2);
    }
    bool explored = false;
    for (;;) {
        while (head >= 0) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(sh) + head);
            const float nx = __int_as_float(v.x), ny = __int_as_float(v.y);
            const float dx = __fsub_rn(nx, qx), dy = __fsub_rn(ny, qy);
            const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            if (d2 < best2) {
                const float d = __fsqrt_rn(d2);
                if (d < bestDist) { bestDist = d; best2 = d2; bestIdx = head; explored = false; }
            }
            const unsigned axis = (unsigned)v.z >> 30;
            const bool branch = axis == 0u ? qx < nx : axis == 1u ? qy < ny : false;
            head = branch ? (v.z << 2) >> 2 : v.w;
        }
        if (explored) break;
        const int bestParent = __ldg(&tree[bestIdx].parent);
        if (bestParent < 0) break;
        const int4 p = __ldg(reinterpret_cast<const int4 *>(sh) + bestParent);
        const unsigned axis = (unsigned)p.z >> 30;
        const float px = __int_as_float(p.x), py = __int_as_float(p.y);
        bool branch = false; float hd = 0.0f;
        if (axis == 0u) { branch = qx < px; hd = fabsf(__fsub_rn(qx, px)); }
        if (axis == 1u) { branch = qy < py; hd = fabsf(__fsub_rn(qy, py)); }
        if (axis == 2u) { branch = false; hd = 0.0f; }          // |qz - p.z| with both 0
        if (!(hd < bestDist)) break;
        head = !branch ? (p.z << 2) >> 2 : p.w;
        explored = true;
    }
    return bestIdx;
}

// kd NN lookup alone (the "kd-tree NN lookup" entry point): one thread per query
__global__ void k_kd_nn(const KdNode *__restrict__ tree, const KdState *__restrict__ ks, const float *__restrict__ q3,
                        int n, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ks->size > 0 ? kd_nn<false>(tree, q3[3 * i], q3[3 * i + 1], q3[3 * i + 2]) : -1;
}

// kernel.cu:1198-1308 kernEvaluateParticlesKD.  Node weights are integers (0, -100, +-1, +4, clamp
// +-113), so the float sum of the reference is an exact integer and the grid path's integer
// extrema / weight kernels are reused unchanged (the float overload's `int min`, Q5, is exact too).
// block = 8 warps, lane = particle, warp w takes beams w, w+8, ...  (Interleaving 4 walks per thread
// for memory-level parallelism was measured SLOWER, 5.5 vs 3.7 ms: the kernel is bound by divergent
// instruction issue, not by load latency.)
template <int kFlat>      // 0: generic 32-byte walk, 1: format-1 shadow, 2: format-2 shadow with the branch-free visit
__global__ void __launch_bounds__(256)
k_score_kd(const KdNode *__restrict__ tree, const KdSearch *__restrict__ sh, const float *__restrict__ x, const float *__restrict__ y,
           const float *__restrict__ th, int n, int gidx0, const StepParams *__restrict__ sp,
           const float *__restrict__ angle, int n_beams, int *__restrict__ fit, int *__restrict__ blk_min,
           long long *__restrict__ blk_maxkey)
{
    __shared__ int part[8][32];
    const float *__restrict__ scan = sp->scan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    int acc = 0;
    if (p < n) {
        const float px = x[p], py = y[p], pth = th[p];
        for (int j = warp; j < n_beams; j += 8) {
            const float rot = __fadd_rn(__ldg(&angle[j]), pth);
            const float r = __ldg(&scan[j]);
            const float wx = __fmul_rn(r, cosf(rot)), wy = __fmul_rn(r, sinf(rot));
            if (fabsf(wx) < kLidarRange && fabsf(wy) < kLidarRange) {
                const int k = kFlat == 2 ? kd_nn_flat2(sh, tree, __fadd_rn(wx, px), __fadd_rn(wy, py))
                            : kFlat == 1 ? kd_nn_flat(sh, tree, __fadd_rn(wx, px), __fadd_rn(wy, py))
                                         : kd_nn<false>(tree, __fadd_rn(wx, px), __fadd_rn(wy, py), 0.0f);
                acc += (int)__ldg(&tree[k].w);
            }
        }
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

// Measurement only (bench.py's kd roofline line, SURVEY 8d "report with measured mean visited-node count"): the same
// walk as kd_nn<>, counting the nodes it loads, over the first n_sample particles x all in-range beams.
__global__ void __launch_bounds__(256)
k_kd_count_visits(const KdNode *__restrict__ tree, const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ th,
                  int n_sample, const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams,
                  unsigned long long *__restrict__ out /* [0] visits, [1] walks */)
{
    const float *__restrict__ scan = sp->scan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    unsigned long long visits = 0, walks = 0;
    if (p < n_sample) {
        const float px = x[p], py = y[p], pth = th[p];
        for (int j = warp; j < n_beams; j += 8) {
            const float rot = __fadd_rn(__ldg(&angle[j]), pth);
            const float r = __ldg(&scan[j]);
            const float wx = __fmul_rn(r, cosf(rot)), wy = __fmul_rn(r, sinf(rot));
            if (!(fabsf(wx) < kLidarRange && fabsf(wy) < kLidarRange)) continue;
            const float qx = __fadd_rn(wx, px), qy = __fadd_rn(wy, py), qz = 0.0f;
            KdNode t = kd_load(tree, 0);
            float best2 = kd_dist2(qx, qy, qz, t.x, t.y, t.z), bestDist = __fsqrt_rn(best2);
            int bestParent = t.parent, head = 0;
            bool explored = false;
            visits++; walks++;
            for (;;) {
                while (head >= 0) {
                    t = kd_load(tree, head); visits++;
                    const float d2 = kd_dist2(qx, qy, qz, t.x, t.y, t.z);
                    if (d2 < best2) { const float d = __fsqrt_rn(d2); if (d < bestDist) { bestDist = d; best2 = d2; bestParent = t.parent; explored = false; } }
                    const bool branch = t.axis == 0 ? qx < t.x : t.axis == 1 ? qy < t.y : t.axis == 2 ? qz < t.z : false;
                    head = branch ? t.left : t.right;
                }
                if (explored || bestParent < 0) break;
                const KdNode pn = kd_load(tree, bestParent); visits++;
                bool branch = false; float hd = 0.0f;
                if (pn.axis == 0) { branch = qx < pn.x; hd = fabsf(__fsub_rn(qx, pn.x)); }
                if (pn.axis == 1) { branch = qy < pn.y; hd = fabsf(__fsub_rn(qy, pn.y)); }
                if (pn.axis == 2) { branch = qz < pn.z; hd = fabsf(__fsub_rn(qz, pn.z)); }
                if (!(hd < bestDist)) break;
                head = !branch ? pn.left : pn.right;
                explored = true;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { visits += __shfl_xor_sync(0xffffffffu, visits, o); walks += __shfl_xor_sync(0xffffffffu, walks, o); }
    if (lane == 0) { atomicAdd(&out[0], visits); atomicAdd(&out[1], walks); }
}

// IEEE-only asin (Cephes asinf scheme): theta = asin(R[0][1]), kernel.cu:1079
__device__ __forceinline__ float pf_asinf(float x)
{
    float a = fabsf(x);
    if (a > 1.0f) a = 1.0f;
    bool flag = false; float z, xx;
    if (a > 0.5f) { z = __fmul_rn(0.5f, __fsub_rn(1.0f, a)); xx = __fsqrt_rn(z); flag = true; }
    else { xx = a; z = __fmul_rn(xx, xx); }
    float p = 4.2163199048E-2f;
    p = __fmaf_rn(p, z, 2.4181311049E-2f);
    p = __fmaf_rn(p, z, 4.5470025998E-2f);
    p = __fmaf_rn(p, z, 7.4953002686E-2f);
    p = __fmaf_rn(p, z, 1.6666752422E-1f);
    float r = __fmaf_rn(__fmul_rn(p, z), xx, xx);
    if (flag) { r = __fadd_rn(r, r); r = __fsub_rn(1.570796326794896619f, r); }
    return x < 0.0f ? -r : r;
}

// "ICP order" sum over v[0..n): lane l adds v[l], v[l+32], ... sequentially, then an xor-butterfly
__device__ __forceinline__ float icp_warp_sum(const float *v, int n, int lane)
{
    float s = 0.0f;
    for (int i = lane; i < n; i += 32) s = __fadd_rn(s, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    return s;
}

// kernel.cu:993-1093 transformPointICP in one block: targets from the PREVIOUS robotPos (the global the
// reference reads in kernGetWallsKD), NN correspondences, means, cross-covariance, planar rotation,
// robotPos = best particle pose + (t.x, t.y, theta).
__global__ void __launch_bounds__(1024)
k_icp(const KdNode *__restrict__ tree, const Xchg xc,
      const StepParams *__restrict__ sp, const float *__restrict__ angle, int n_beams, FrameResult *__restrict__ res)
{
    extern __shared__ float s_f[];           // tar x|y, cor x|y, prod : 5 * n_beams
    __shared__ float s_m[4], s_h[4];
    const float *__restrict__ scan = sp->scan;
    float *tx = s_f, *ty = s_f + n_beams, *cx = s_f + 2 * n_beams, *cy = s_f + 3 * n_beams, *pr = s_f + 4 * n_beams;
    const float rx0 = res->pose[0], ry0 = res->pose[1], rt0 = res->pose[2];
    for (int i = threadIdx.x; i < n_beams; i += blockDim.x) {
        const float rot = __fadd_rn(angle[i], rt0);
        const float wx = __fmul_rn(scan[i], cosf(rot)), wy = __fmul_rn(scan[i], sinf(rot));
        float ax = 0.0f, ay = 0.0f;                                            // Q11
        if (fabsf(wx) < kLidarRange && fabsf(wy) < kLidarRange) { ax = __fadd_rn(rx0, wx); ay = __fadd_rn(ry0, wy); }
        const KdNode nn = kd_load(tree, kd_nn<false>(tree, ax, ay, 0.0f));
        tx[i] = ax; ty[i] = ay; cx[i] = nn.x; cy[i] = nn.y;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 4) {
        const float *v = warp == 0 ? tx : warp == 1 ? ty : warp == 2 ? cx : cy;
        const float s = icp_warp_sum(v, n_beams, lane);
        if (lane == 0) s_m[warp] = __fdiv_rn(s, (float)n_beams);
    }
    __syncthreads();
    // H[i][j] = sum (tar_i - mu_tar_i)(cor_j - mu_cor_j), one entry at a time (4 passes over 1081 values)
    for (int e = 0; e < 4; e++) {
        const int i = e >> 1, j = e & 1;
        for (int q = threadIdx.x; q < n_beams; q += blockDim.x)
            pr[q] = __fmul_rn(__fsub_rn(i ? ty[q] : tx[q], s_m[i]), __fsub_rn(j ? cy[q] : cx[q], s_m[2 + j]));
        __syncthreads();
        if (warp == 0) { const float s = icp_warp_sum(pr, n_beams, lane); if (lane == 0) s_h[e] = s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float s = __fsub_rn(s_h[1], s_h[2]), k = __fadd_rn(s_h[0], s_h[3]);     // H01 - H10, H00 + H11
        const float nrm = __fsqrt_rn(__fmaf_rn(s, s, __fmul_rn(k, k)));
        float sn = 0.0f, cs = 1.0f;
        if (nrm > 0.0f) { sn = __fdiv_rn(s, nrm); cs = __fdiv_rn(k, nrm); }
        const float t_x = __fsub_rn(s_m[2], __fmaf_rn(cs, s_m[0], -__fmul_rn(sn, s_m[1])));
        const float t_y = __fsub_rn(s_m[3], __fmaf_rn(sn, s_m[0], __fmul_rn(cs, s_m[1])));
        int gmin, gmax, best; float pose[3];
        reduce_extrema(xc_ext(xc, sp->seq), xc.n_ranks, gmin, gmax, best, pose, xc.parity_mask != 0);   // k_weights_scan already waited for it
        res->pose[0] = __fadd_rn(pose[0], t_x);
        res->pose[1] = __fadd_rn(pose[1], t_y);
        res->pose[2] = __fadd_rn(pose[2], pf_asinf(sn));
    }
}

// ---- map update (kernel.cu:1406-1540) -----------------------------------------------------------------
// Masks: kernGetWalls in the window centred on the map centre with the robot's heading only.
__global__ void __launch_bounds__(128)
k_kd_mark(MapGeom g, const FrameResult *__restrict__ res, const StepParams *__restrict__ sp,
          const float *__restrict__ angle, unsigned *__restrict__ free_bits, unsigned *__restrict__ wall_bits)
{
    const float *__restrict__ scan = sp->scan;
    const int j = blockIdx.x;
    int cx, cy; center_cell(g, 0.0f, 0.0f, cx, cy);                 // kernel.cu:1408-1411
    const float pose[3] = {0.0f, 0.0f, res->pose[2]};
    float wx, wy;
    if (!beam_hit(g, pose, cx, cy, angle[j], scan[j], wx, wy)) return;
    int sx = cx, sy = cy, ex = (int)wx, ey = (int)wy;
    if (threadIdx.x == 0 && wx >= 0.0f && wx < (float)g.w && wy >= 0.0f && wy < (float)g.h) {
        const int idx = (int)__fmaf_rn(wx, (float)g.w, wy);
        atomicOr(&wall_bits[idx >> 5], 1u << (idx & 31));
    }
    const bool steep = abs(ey - sy) > abs(ex - sx);
    int t;
    if (steep) { t = sx; sx = sy; sy = t; t = ex; ex = ey; ey = t; }
    if (sx > ex) { t = sx; sx = ex; ex = t; t = sy; sy = ey; ey = t; }
    const int deltax = ex - sx, deltay = abs(ey - sy), e0 = deltax / 2;
    const int ystep = ey > sy ? 1 : -1;
    for (int k = threadIdx.x; k < deltax; k += blockDim.x) {
        const int num = k * deltay - e0;
        const int m = num > 0 ? (num + deltax - 1) / deltax : 0;
        const int xx = sx + k, yy = sy + ystep * m;
        const int idx = steep ? yy * g.w + xx : xx * g.w + yy;
        if (xx < g.w && yy < g.h && xx >= 0 && yy >= 0 && idx < g.w * g.h) atomicOr(&free_bits[idx >> 5], 1u << (idx & 31));
    }
}

// Ordered compaction of the set bits of both masks (== the reference's x-major double loop over the
// bool masks, kernel.cu:1435-1461): per-block popcounts, then offsets, then scatter of cell indices.
constexpr int kBitsBlockWords = 1024;
__global__ void __launch_bounds__(256)
k_bits_count(const unsigned *__restrict__ fb, const unsigned *__restrict__ wb, int n_words, int *__restrict__ blk_cnt)
{
    __shared__ int s[2];
    if (threadIdx.x < 2) s[threadIdx.x] = 0;
    __syncthreads();
    int cf = 0, cw = 0;
    for (int k = 0; k < 4; k++) {
        const int wi = blockIdx.x * kBitsBlockWords + k * 256 + threadIdx.x;
        if (wi < n_words) { cf += __popc(fb[wi]); cw += __popc(wb[wi]); }
    }
    for (int o = 16; o > 0; o >>= 1) { cf += __shfl_xor_sync(0xffffffffu, cf, o); cw += __shfl_xor_sync(0xffffffffu, cw, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], cf); atomicAdd(&s[1], cw); }
    __syncthreads();
    if (threadIdx.x == 0) { blk_cnt[2 * blockIdx.x] = s[0]; blk_cnt[2 * blockIdx.x + 1] = s[1]; }
}

__global__ void k_bits_offsets(int *__restrict__ blk_cnt, int n_blk, KdState *__restrict__ ks)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int of = 0, ow = 0;
        for (int b = 0; b < n_blk; b++) { int cf = blk_cnt[2 * b], cw = blk_cnt[2 * b + 1]; blk_cnt[2 * b] = of; blk_cnt[2 * b + 1] = ow; of += cf; ow += cw; }
        ks->n_free = of; ks->n_wall = ow;
    }
}

__global__ void __launch_bounds__(256)
k_bits_scatter(const unsigned *__restrict__ fb, const unsigned *__restrict__ wb, int n_words,
               const int *__restrict__ blk_off, int *__restrict__ free_cells, int *__restrict__ wall_cells, int cap)
{
    __shared__ int s_w[2][8];
    // thread t owns 4 consecutive words: block-wide exclusive scan of their popcounts (both masks)
    const int w0 = blockIdx.x * kBitsBlockWords + threadIdx.x * 4;
    unsigned f[4], w[4];
    int cf = 0, cw = 0;
    for (int k = 0; k < 4; k++) { f[k] = w0 + k < n_words ? fb[w0 + k] : 0u; w[k] = w0 + k < n_words ? wb[w0 + k] : 0u; cf += __popc(f[k]); cw += __popc(w[k]); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int vf = cf, vw = cw;
    for (int o = 1; o < 32; o <<= 1) { int a = __shfl_up_sync(0xffffffffu, vf, o), b = __shfl_up_sync(0xffffffffu, vw, o); if (lane >= o) { vf += a; vw += b; } }
    if (lane == 31) { s_w[0][warp] = vf; s_w[1][warp] = vw; }
    __syncthreads();
    int bf = blk_off[2 * blockIdx.x], bw = blk_off[2 * blockIdx.x + 1];
    for (int q = 0; q < warp; q++) { bf += s_w[0][q]; bw += s_w[1][q]; }
    int pf_ = bf + vf - cf, pw = bw + vw - cw;
    for (int k = 0; k < 4; k++) {
        unsigned m = f[k];
        while (m) { int b = __ffs(m) - 1; m &= m - 1; if (pf_ < cap) free_cells[pf_] = (w0 + k) * 32 + b; pf_++; }
        m = w[k];
        while (m) { int b = __ffs(m) - 1; m &= m - 1; if (pw < cap) wall_cells[pw] = (w0 + k) * 32 + b; pw++; }
    }
}

// point of a cell: ROUND_FRAC(cell*res - scale/2 + robotPos, res), host float order (kernel.cu:1442-1445)
__device__ __forceinline__ void kd_cell_point(const MapGeom &g, int cell, const float *pose, float &px, float &py)
{
    const int x = cell / g.w, y = cell - x * g.w;
    float a = __fmul_rn((float)x, g.res_x); a = __fsub_rn(a, __fdiv_rn(g.scale_x, 2.0f)); a = __fadd_rn(a, pose[0]);
    float b = __fmul_rn((float)y, g.res_y); b = __fsub_rn(b, __fdiv_rn(g.scale_y, 2.0f)); b = __fadd_rn(b, pose[1]);
    px = __fmul_rn(roundf(__fdiv_rn(a, g.res_x)), g.res_x);
    py = __fmul_rn(roundf(__fdiv_rn(b, g.res_y)), g.res_y);
}

// NN of every wall point and of the first min(n_free, n_wall) free points (Q10), against the tree
// BEFORE this frame's updates.  pts: [wall 0..cap) | free 0..cap) as float2; nn likewise.
__global__ void __launch_bounds__(128)
k_kd_points_nn(const KdNode *__restrict__ tree, MapGeom g, const FrameResult *__restrict__ res,
               const KdState *__restrict__ ks, const int *__restrict__ wall_cells, const int *__restrict__ free_cells,
               int cap, float2 *__restrict__ pts, int *__restrict__ nn)
{
    const int nW = min(ks->n_wall, cap), nF = min(min(ks->n_free, nW), cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool is_free = i >= cap;
    const int k = is_free ? i - cap : i;
    if (k >= (is_free ? nF : nW)) return;
    float px, py;
    kd_cell_point(g, is_free ? free_cells[k] : wall_cells[k], res->pose, px, py);
    pts[i] = make_float2(px, py);
    nn[i] = ks->size > 0 ? kd_nn<false>(tree, px, py, 0.0f) : -1;
}

// kernel.cu:1350-1364 kernUpdateMapKD: w = clamp(w + val) when the point is within sqrt(2)*res of its NN.
// The reference updates with a plain load and store from one thread per point, so points that share a nearest
// node race: of k colliding threads any number between 1 and k takes effect (the B200 shows both collapsing and
// accumulating collisions on one frame, T3).  The engine defines the deterministic outcome "once per node per
// launch" -- always one of the race's legal results, and the kd counterpart of the grid path's once-per-cell
// bool masks: the first point to stamp the node's claim word with this (step, pass) applies the update.
__global__ void __launch_bounds__(128)
k_kd_weights(KdNode *__restrict__ tree, MapGeom g, const KdState *__restrict__ ks, int cap, int pass,
             const float2 *__restrict__ pts, const int *__restrict__ nn, int *__restrict__ claim, int stamp)
{
    if (ks->size <= 0) return;
    const int nW = min(ks->n_wall, cap), nF = min(min(ks->n_free, nW), cap);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (pass == 0 ? nF : nW)) return;
    const int i = pass == 0 ? cap + k : k;
    const float val = pass == 0 ? (float)kFreeWeight : (float)kOccupiedWeight;
    const float minDist = __fsqrt_rn(__fmaf_rn(g.res_y, g.res_y, __fmul_rn(g.res_x, g.res_x)));
    const int t = nn[i];
    const float2 p = pts[i];
    const KdNode nd = kd_load_cg(tree, t);
    if (kd_dist(p.x, p.y, 0.0f, nd.x, nd.y, nd.z) < minDist && atomicMax(&claim[t], stamp) < stamp) {
        float v = __fadd_rn(nd.w, val);
        v = v < -(float)kClamp ? -(float)kClamp : v > (float)kClamp ? (float)kClamp : v;
        tree[t].w = v;
    }
}

// kernel.cu:1367-1379 kernTestCorrespondance + :1504-1520 sequential InsertNode, in one block.
// New nodes get indices size + rank (wall-point order).  Insertion is done in rounds that reproduce the
// sequential result: every pending point walks to its empty slot; the lowest-order point claiming a
// slot wins it (atomicMin on the child link, claims encoded above 0x80000000); losers continue from
// the winner's node next round.
__global__ void __launch_bounds__(1024)
k_kd_insert(KdNode *tree, MapGeom g, KdState *__restrict__ ks, int cap, int kd_cap,
            const float2 *__restrict__ pts, const int *__restrict__ nn, int *__restrict__ ins_index)
{
    __shared__ int s_scan[32];
    __shared__ int s_base;
    const int size0 = ks->size;
    const int nW = min(ks->n_wall, cap);
    const float half = __fmul_rn(__fsqrt_rn(__fmaf_rn(g.res_y, g.res_y, __fmul_rn(g.res_x, g.res_x))), 0.5f);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    if (size0 <= 0) { if (threadIdx.x == 0) ks->n_ins = 0; return; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // pass 1: flags + ordered ranks (chunks of 1024 wall points)
    for (int i0 = 0; i0 < nW; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool d = false;
        if (i < nW) {
            const KdNode nd = kd_load_cg(tree, nn[i]);
            d = kd_dist(pts[i].x, pts[i].y, 0.0f, nd.x, nd.y, nd.z) > half;
        }
        const unsigned bm = __ballot_sync(0xffffffffu, d);
        if (lane == 0) s_scan[warp] = __popc(bm);
        __syncthreads();
        int pre = s_base, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < warp) pre += s_scan[w]; tot += s_scan[w]; }
        if (i < nW) {
            const int idx = size0 + pre + __popc(bm & ((1u << lane) - 1));
            ins_index[i] = (d && idx < kd_cap) ? idx : -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    const int n_ins = min(s_base, max(0, kd_cap - size0));
    // pass 2: insertion rounds.  pending points keep (start node) in registers; each thread owns
    // wall points threadIdx.x, +1024, ... (at most 2 in practice; loop for generality)
    constexpr int kOwn = 4;
    int start[kOwn]; bool pend[kOwn];
    for (int u = 0; u < kOwn; u++) { const int i = threadIdx.x + u * blockDim.x; start[u] = 0; pend[u] = i < nW && ins_index[i] >= 0; }
    for (int round = 0; round < 4096; round++) {
        int slot_node[kOwn], slot_side[kOwn], slot_axis[kOwn];
        for (int u = 0; u < kOwn; u++) {
            if (!pend[u]) continue;
            const int i = threadIdx.x + u * blockDim.x;
            const float qx = pts[i].x, qy = pts[i].y;
            int next = start[u], parent = next, axis = 0; bool less = false;
            do {                                                   // kdtree.cpp:72-82
                parent = next;
                const KdNode nd = kd_load_cg(tree, next);
                axis = nd.parent == -1 ? 0 : (__ldcg(&tree[nd.parent].axis) + 1) % 3;
                const float a = axis == 0 ? qx : axis == 1 ? qy : 0.0f;
                const float b = axis == 0 ? nd.x : axis == 1 ? nd.y : nd.z;
                less = a < b;
                next = less ? nd.left : nd.right;
            } while (next >= 0 && (unsigned)next < 0x80000000u);
            slot_node[u] = parent; slot_side[u] = less ? 0 : 1; slot_axis[u] = axis;
            unsigned *link = reinterpret_cast<unsigned *>(less ? &tree[parent].left : &tree[parent].right);
            atomicMin(link, 0x80000000u | (unsigned)i);
        }
        __threadfence_block();
        __syncthreads();
        int still = 0;
        for (int u = 0; u < kOwn; u++) {
            if (!pend[u]) continue;
            const int i = threadIdx.x + u * blockDim.x;
            int *link = slot_side[u] == 0 ? &tree[slot_node[u]].left : &tree[slot_node[u]].right;
            const unsigned v = (unsigned)__ldcg(link);
            if (v == (0x80000000u | (unsigned)i)) {                // this point owns the slot: kdtree.cpp:85-104
                KdNode nd; nd.axis = (slot_axis[u] + 1) % 3; nd.left = -1; nd.right = -1; nd.parent = slot_node[u];
                nd.x = pts[i].x; nd.y = pts[i].y; nd.z = 0.0f; nd.w = -100.0f;                       // kernel.cu:1514
                tree[ins_index[i]] = nd;
                __threadfence_block();
                *link = ins_index[i];
                pend[u] = false;
            } else {                                               // continue below the winner next round
                start[u] = v >= 0x80000000u ? ins_index[v & 0x7fffffffu] : (int)v;
                still = 1;
            }
        }
        __threadfence_block();
        if (!__syncthreads_or(still)) break;
    }
    if (threadIdx.x == 0) { ks->n_ins = n_ins; ks->size = size0 + n_ins; }
}

// publish the kd counters into the frame result
__global__ void k_kd_finish(FrameResult *__restrict__ res, const KdState *__restrict__ ks, int *__restrict__ counters, int pc_cap, int kd_cap)
{
    res->n_free = ks->n_free; res->n_wall = ks->n_wall; res->n_slow = 0;
    res->kd_size = ks->size; res->kd_ins = ks->n_ins;
    // capacities are hard limits, not silent clamps: more wall points in one scan than the point lists hold, or a full
    // node array, make the step fail (sticky, reported by pfslam_fetch_result / pfslam_step)
    if (ks->n_wall > pc_cap) res->kd_overflow |= 1;
    if (ks->size >= kd_cap) res->kd_overflow |= 2;
    counters[0] = 0; counters[1] = 0; counters[2] = 0; counters[3] = 0;
}

}  // namespace pf
