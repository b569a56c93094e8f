// pf_xchg.cuh -- how the shards of one particle filter (one engine per GPU) exchange the few words
// a frame needs, from INSIDE the step's kernels, over NVLink peer memory.
//
// The reference is single-GPU; sharding is new.  Per frame a shard needs from the others
//   (1) the score extrema {min, max, first arg-max, its pose}              32 B per rank
//   (2) per-tile weight sums and the tile-local CDF values                  ~4 B per particle
//   (3) the pre-resample pose of whatever particle its resampler draws     16 B per particle
// Every engine owns one "exchange region" (a single cudaMalloc, IPC-exportable) with the same layout
// on every rank.  Producers STORE (1) and (2) straight into every peer's region from the kernel that
// computes them (k_score_combine_rows / k_extrema, k_weights_scan), then raise a per-(kind, source
// rank) flag in the peer's region with the step's sequence number; the first consumer kernel of the
// step spins (bounded) on its own region's flags.  (3) is PUSHED whole: every rank copies its 16 B/particle
// pre-resample snapshot into every peer's region with coalesced stores from a side-branch kernel that runs under the
// scoring (k_snapshot_push), and the resampler gathers locally.  (Round 1 PULLED the drawn poses one by one over
// NVLink: ~30 k random 16-byte remote reads per GPU on a resampling step cost 9 us on 2 GPUs and more on 8 --
// PFSLAM_SNAPSHOT=pull keeps that mode.)  No NCCL call, no host round trip: the sharded step is the same single
// CUDA graph as the single-GPU step.
//
// Buffers are double-buffered by step parity.  That is enough because the shards run in lock step:
// a rank publishes extrema(s+1) only after it has seen every rank's tiles(s), and tiles(s+1) only
// after every rank's extrema(s+1) -- so nobody can be more than one publication ahead of a reader.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

constexpr int kMaxRanks = 16;
constexpr int kXcExt = 0, kXcTiles = 1;      // flag kinds

// extrema record exchanged between ranks: 8 words
struct Extrema {
    int   fit_min, fit_max, best_gidx;
    float x, y, th;
    int   pad0, pad1;
};

// Passed by value to the kernels that produce or consume exchanged data.  parity_mask == 0: single
// GPU, or a host that runs its own collectives between the phases (dist.py over torch.distributed);
// then ext_all / tiles_all are plain local buffers and nothing is published or awaited.
struct Xchg {
    int n_ranks, rank;
    int parity_mask;              // 1 = peer-memory exchange
    unsigned timeout_ms;          // bound on every flag wait
    long long tiles_block;        // floats per (parity, source rank) tiles block
    long long sum_off, lm_off;    // offsets of [tsum_w | tsum_w2] and lm[] inside a tiles block
    long long snap_stride;        // floats between the two parities of a pose snapshot
    int snap_aos;                 // snapshot layout: 1 = float4 {x, y, theta, 0} per particle (one 16-byte gather,
                                  // one NVLink transaction per pull), 0 = x[] | y[] | theta[] (what a host all-gathers)
    // this rank's view (own region when parity_mask, else the engine's local / host-gathered buffers)
    Extrema *ext_all;             // [parity][kMaxRanks]
    float   *tiles_all;           // [parity][n_ranks][tiles_block]
    float   *snap;                // own pre-resample snapshot [parity][...]; null: host gathers
    int     *flags;               // [kind][kMaxRanks] sequence numbers, written by the peers
    // every rank's region (peer memory) and the byte offsets of the parts inside a region
    unsigned char *peer[kMaxRanks];
    long long off_ext, off_tiles, off_flags, off_snap;
    const float *pose_src[kMaxRanks];   // resample gather source per owner rank (parity 0)
};

__device__ __forceinline__ unsigned long long xc_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Flags are polled with relaxed system-scope loads (an acquire load per poll would cost a fence per poll) and
// the consumer fences ONCE after the flag has arrived; producers fence once and then raise every peer's flag
// with relaxed stores (a release store per peer would repeat the fence -- one NVLink round trip -- per peer,
// which is what made the exchange cost grow with the rank count).
// release / acquire fence at system scope (the sequentially consistent __threadfence_system() is not needed for
// the store -> fence -> flag | flag -> fence -> load hand-overs)
__device__ __forceinline__ void xc_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ int xc_ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void xc_st_relaxed(int *p, int v)
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 xc_ld_relaxed16(const int4 *p)
{
    int4 v;
    asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void xc_st_relaxed16(int4 *p, int4 v)
{
    asm volatile("st.relaxed.sys.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ Extrema *xc_ext(const Xchg &xc, int seq)
{
    return xc.ext_all + (seq & xc.parity_mask) * kMaxRanks;
}
__device__ __forceinline__ float *xc_tiles(const Xchg &xc, int seq)
{
    return xc.tiles_all + (long long)(seq & xc.parity_mask) * xc.n_ranks * xc.tiles_block;
}

// A whole warp (all 32 lanes call it) waits until every rank's flag of `kind` has reached this step's
// sequence number: lane r polls rank r's flag, so the wait costs one round of latency whatever the rank
// count.  Returns false on timeout (the caller records it in the frame result; nothing hangs).
// kXcExt has no flag word: in a peer region the extrema record travels as two self-validating 16-byte halves
// {min, max, arg-max index, seq} and {x, y, theta, seq}; each half is one store, so a half that shows this
// step's number is complete, and no fence is needed on either side.
__device__ __forceinline__ bool xc_wait_warp(const Xchg &xc, int kind, int seq)
{
    if (!xc.parity_mask) return true;
    const int lane = threadIdx.x & 31;
    const unsigned long long t0 = xc_now_ns();
    const unsigned long long limit = (unsigned long long)xc.timeout_ms * 1000000ull;
    bool ok = true;
    if (lane < xc.n_ranks) {
        unsigned spins = 0;
        if (kind == kXcExt) {
            const int4 *rec = reinterpret_cast<const int4 *>(xc.ext_all + (seq & 1) * kMaxRanks + lane);
            for (;;) {
                const int4 a = xc_ld_relaxed16(rec), b = xc_ld_relaxed16(rec + 1);
                if (a.w == seq && b.w == seq) break;
                if ((++spins & 255u) == 0u && xc_now_ns() - t0 > limit) { ok = false; break; }
            }
        } else {
            const int *f = xc.flags + kind * kMaxRanks;
            while (xc_ld_relaxed(f + lane) - seq < 0) {
                if ((++spins & 255u) == 0u && xc_now_ns() - t0 > limit) { ok = false; break; }
            }
        }
    }
    if (kind != kXcExt) xc_fence_sys();   // acquire: what the producers wrote before their flags is visible now
    return __all_sync(0xffffffffu, ok);
}

// One thread, after the data stores (and a block barrier if other threads made them): make the
// stores visible system-wide ONCE, then raise this rank's flag in every peer's region.
__device__ __forceinline__ void xc_signal(const Xchg &xc, int kind, int seq)
{
    xc_fence_sys();
    for (int r = 0; r < xc.n_ranks; r++)
        xc_st_relaxed(reinterpret_cast<int *>(xc.peer[r] + xc.off_flags) + kind * kMaxRanks + xc.rank, seq);
}

// Extrema of this shard -> ext_local (local modes) or slot [parity][rank] of every peer's region (wire format,
// see xc_wait_warp): 2 stores per peer, no fence, no flag.
__device__ __forceinline__ void xc_publish_extrema(const Xchg &xc, Extrema *ext_local, const Extrema &e, int seq)
{
    if (!xc.parity_mask) { *ext_local = e; return; }
    const int4 a = make_int4(e.fit_min, e.fit_max, e.best_gidx, seq);
    const int4 b = make_int4(__float_as_int(e.x), __float_as_int(e.y), __float_as_int(e.th), seq);
    for (int r = 0; r < xc.n_ranks; r++) {
        int4 *dst = reinterpret_cast<int4 *>(reinterpret_cast<Extrema *>(xc.peer[r] + xc.off_ext) +
                                             (seq & 1) * kMaxRanks + xc.rank);
        xc_st_relaxed16(dst, a);
        xc_st_relaxed16(dst + 1, b);
    }
}

// wire: the record sits in a peer region (two stamped halves), else it is a plain Extrema
__device__ __forceinline__ Extrema xc_load_extrema(const Extrema *p, bool wire)
{
    const int4 a = __ldcg(reinterpret_cast<const int4 *>(p));
    const int4 b = __ldcg(reinterpret_cast<const int4 *>(p) + 1);
    Extrema e;
    e.fit_min = a.x; e.fit_max = a.y; e.best_gidx = a.z; e.pad0 = 0; e.pad1 = 0;
    if (wire) { e.x = __int_as_float(b.x); e.y = __int_as_float(b.y); e.th = __int_as_float(b.z); }
    else { e.x = __int_as_float(a.w); e.y = __int_as_float(b.x); e.th = __int_as_float(b.y); }
    return e;
}

}  // namespace pf
