#!/usr/bin/env python
"""bench.py -- LiDAR frames/s of the particle-filter step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles P]

A step = one LiDAR frame through the whole 2D particle-filter step (motion, scoring, extrema,
weights, map update, resample) at 65 536 particles per GPU (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# keep stdout to the one JSON line: NCCL prints its version banner (and any debug output) to stdout
# unless told otherwise
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG")

N_BEAMS = 1081
METRIC = "lidar_frames_per_sec_at_65536_particles_per_gpu"
UNIT = "frames/s"


def synthetic_scans(n_frames):
    """BASELINE configs[4]: the synthetic 1081-beam corridor workload (gpu-icp-slam_b200/synth.py), frames 1.."""
    from gpu_icp_slam_b200 import synth
    sc, _ = synth.generate(n_frames + 1)
    return (np.ascontiguousarray(sc[1:]), list(range(1, n_frames + 1)),
            "synthetic 1081-beam corridor scans @ 40 Hz (gpu-icp-slam_b200/synth.py, PCG64(565))")


def dataset_path(dataset):
    """(path, description): the converted full dataset under data/_cache when it travelled to the box, else
    the committed 256-frame train_lidar0 fixture"""
    full = os.path.join(ROOT, "data", "_cache", dataset + ".scans.u16")
    if os.path.exists(full):
        return full, "real %s scans (data/_cache, first 4000 frames of the .mat, ping-pong)" % dataset
    return (os.path.join(ROOT, "tests", "golden", "train_lidar0_first256.scans.u16"),
            "real train_lidar0 scans (committed 256-frame fixture, ping-pong; data/_cache/%s is not on this box)" % dataset)


def workload_scans(n_frames, dataset="train_lidar0"):
    """real train_lidar scans played 1..last..1.. (ping-pong keeps the motion physically continuous
    for any number of steps)"""
    from gpu_icp_slam_b200 import scans as S
    path, desc = dataset_path(dataset)
    fx = S.load(path)
    last = min(len(fx) - 1, 4000)
    order, f, d = [], 1, 1
    for _ in range(n_frames):
        order.append(f)
        if f + d > last or f + d < 1:
            d = -d
        f += d
    return np.ascontiguousarray(fx[order]), order, desc


def default_dataset(args):
    if args.data != "auto":
        return args.data
    if args.path == "kd":
        return "train_lidar3"                      # BASELINE configs[2]
    return "train_lidar2" if args.gpus == 4 else "train_lidar0"   # configs[3] / configs[1]


def make_config(args, world, dataset):
    """the `config` object; identical in both arms (ours / --impl reference) for the same command line"""
    n = args.particles
    kd = args.path == "kd"
    return {"workload": ("%s, %s, %d particles per GPU x %d" %
                         (dataset, "kd-tree point cloud @ 25 mm" if kd else "2D occupancy grid 1600x1600 @ 0.025 m", n, world)),
            "particles_per_gpu": n, "particles_total": world * n, "beams": N_BEAMS, "path": args.path,
            "value_is": "frames/s x particles_total / 65536 (65 536-particle frame equivalents)",
            "l2": "ours: 256 MiB L2 flush between timed steps; reference CPU arm: working set (grid 2.56 MB + scans) larger than the host L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.p = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            t = [v.strip() for v in ln.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def scorer_traffic():
    """DRAM bytes per k_score_tiled launch (dram__bytes_read + dram__bytes_write) from the committed ncu
    --set full capture (tools/ncu_summary.py writes the file); None when there is none"""
    p = os.path.join(ROOT, "profiles", "score_tiled_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["traffic_bytes_per_launch"]), d.get("source", "")
    except Exception:
        return None, ""


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_frames_per_sec(n_particles, n_frames, threads, dataset="train_lidar0", warmup=1):
    """The reference's CPU implementation of the path on the host cores: the reference's own
    EvaluateParticle / ParticleAddNoise (oracle/_ref/libref.so, unmodified sources) when that
    library was built, else the oracle port; particle ranges fanned over `threads` host threads
    for the scoring (ctypes releases the GIL); the O(N) remainder (extrema, weights, map update,
    resample) runs through the oracle port on one thread.  Returns (frames/s, kind, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from helpers import P
    o = helpers.load_oracle()
    ref = helpers.load_ref()
    kind = "reference" if ref is not None else "port"
    # reference-CPU flavour of the arithmetic: libm trig, separate multiply-add (host g++ build)
    cfg = helpers.ocfg(helpers.TRIG_LIBM, helpers.MAD_SEPARATE)
    of = helpers.OracleFilter(n_particles, cfg)
    scans, order, _ = workload_scans(n_frames + warmup + 1, dataset)
    bounds = np.linspace(0, n_particles, threads + 1).astype(int)
    fit = np.zeros(n_particles, np.int32)

    def score_range(k, sc):
        a, b = int(bounds[k]), int(bounds[k + 1])
        if b <= a:
            return
        if ref is not None:
            ref.ref_evaluate_particles(P(of.grid, helpers.bp), cfg.map_w, cfg.map_h, cfg.scale_x, cfg.scale_y,
                                       cfg.res_x, cfg.res_y, P(of.x[a:b]), P(of.y[a:b]), P(of.th[a:b]), b - a,
                                       P(sc), P(fit[a:b], helpers.ip))
        else:
            o.pfo_score2d_many(C.byref(cfg), P(of.grid, helpers.bp), P(of.x[a:b]), P(of.y[a:b]), P(of.th[a:b]),
                               b - a, P(sc), P(fit[a:b], helpers.ip))

    def frame(i):
        sc = np.ascontiguousarray(scans[i])
        f = order[i]
        o.pfo_motion(of.s, f)
        ths = [threading.Thread(target=score_range, args=(k, sc)) for k in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        # the remainder of PFMeasurementUpdate / PFUpdateMap / PFResample through the port, fed the
        # scores computed above
        np.copyto(of.fit, fit)
        s = of.s.contents
        mn, mx, am = C.c_int32(), C.c_int32(), C.c_int()
        o.pfo_minmax(of.s.contents.fit, n_particles, C.byref(mn), C.byref(mx), C.byref(am))
        rng = mx.value - mn.value
        if rng > 0:
            w = of.weff
            w *= (fit.astype(np.float32) - np.float32(mn.value)) * np.float32(1.0 / rng)
        np.copyto(of.w, of.weff)
        s.robot[0], s.robot[1], s.robot[2] = float(of.x[am.value]), float(of.y[am.value]), float(of.th[am.value])
        o.pfo_update_map(of.s, P(sc))
        o.pfo_resample(of.s, f)

    for i in range(warmup):                   # warm the map
        frame(i)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + n_frames):
        frame(i)
    dt = time.perf_counter() - t0
    of.close()
    what = ("%d frames of the workload at %d particles; scoring = %s fanned over %d host threads, "
            "O(N) remainder through the oracle port on 1 thread" %
            (n_frames, n_particles, "reference EvaluateParticle (oracle/_ref)" if ref is not None else "oracle port", threads))
    return n_frames / dt, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    dataset = default_dataset(args)
    # same workload as our arm at --gpus N: the whole filter (N x particles-per-GPU) on the host cores,
    # reported in the same unit (65 536-particle frame equivalents per second)
    n = args.particles * world
    W = max(args.warmup, 3)
    # each "step" = one frame; bound the whole run (warm-up included) to a few minutes
    per_frame_guess = n * N_BEAMS * 35e-9 / max(1, threads * 0.8) + n * 1.2e-6
    budget = 240.0
    k = max(1, min(args.steps, int(budget / max(per_frame_guess, 1e-3)) - W))
    fps, kind, what = cpu_reference_frames_per_sec(n, k, threads, dataset, W)
    fps_raw = fps
    fps = fps * n / 65536.0
    _, _, desc = workload_scans(1, dataset)
    line = {
        "impl": "reference", "metric": METRIC + ("_kd" if args.path == "kd" else ""), "value": fps, "unit": UNIT, "n_gpus": world,
        "steps": k, "warmup": W, "ms_per_step": 1e3 / fps_raw, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": desc,
        "config": make_config(args, world, dataset),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": kind, "sample": what},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.path == "kd":
        line["note"] = ("the reference has no CPU kd search (SURVEY 8d): this arm times the reference's CPU 2D scoring path "
                        "on the same scans and particle count")
    print(json.dumps(line), flush=True)


def reference_gpu_leg(scans, order, n_warm, n_timed, kd):
    """The "beat this kernel" bar: the reference's own kernel.cu, compiled for sm_100 with
    PARTICLE_COUNT patched to 65536 at build time (oracle/_ref/libref_t3_65536.so), on this GPU, on the same
    scans: wall clock per step the reference's own way (its calls block), and kernEvaluateParticles[KD] alone
    by CUDA events."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_t3_65536.so")
    scene = os.path.join(ROOT, "oracle", "_ref", "map_settings.txt")
    if not (os.path.exists(so) and os.path.exists(scene)):
        return {"unavailable": "oracle/_ref/libref_t3_65536.so not built (needs /root/reference at build time)"}
    # the reference prints to stdout (scene parser, per-100-frame timers); this process's stdout carries the one JSON
    # line, so the C-level stdout is pointed at stderr for the duration of the leg
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _reference_gpu_leg(so, scene, scans, order, n_warm, n_timed, kd)
    finally:
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)


def _reference_gpu_leg(so, scene, scans, order, n_warm, n_timed, kd):
    fp = C.POINTER(C.c_float)
    lib = C.CDLL(so)
    lib.t3_init.argtypes = [C.c_char_p]
    lib.t3_step2d_ms.restype = C.c_double
    lib.t3_step2d_ms.argtypes = [fp, C.c_int, C.POINTER(C.c_double)]
    lib.t3_step_kd_ms.restype = C.c_double
    lib.t3_step_kd_ms.argtypes = [fp, C.c_int]
    lib.t3_time_evaluate_ms.restype = C.c_float
    lib.t3_time_evaluate_ms.argtypes = [fp, C.c_int, C.c_int]
    lib.t3_set_alloc_fill_sized.argtypes = [C.c_int, C.c_size_t, C.c_int]
    if lib.t3_init(scene.encode()) != 0:
        return {"unavailable": "t3_init failed"}
    lib.t3_reset_kd()
    if kd:
        lib.t3_set_alloc_fill_sized(0xFF, N_BEAMS * 16, 0)      # defined memory for SURVEY Q10 / Q11
    n_part = lib.t3_particle_count()
    phases = np.zeros(4)
    tot, resampled_ms = [], []
    ph = (C.c_double * 4)()
    for i in range(n_warm + n_timed):
        sc = np.ascontiguousarray(scans[i], np.float32)
        ms = lib.t3_step_kd_ms(sc.ctypes.data_as(fp), order[i]) if kd else lib.t3_step2d_ms(sc.ctypes.data_as(fp), order[i], ph)
        if i >= n_warm:
            tot.append(ms)
            if not kd:
                phases += np.array(list(ph))
    sc = np.ascontiguousarray(scans[n_warm + n_timed - 1], np.float32)
    score_ms = float(lib.t3_time_evaluate_ms(sc.ctypes.data_as(fp), 5, 1 if kd else 0))
    out = {"kind": "reference kernel.cu, PARTICLE_COUNT patched to %d at build time, -arch=sm_100, same GPU, same scans" % n_part,
           "particles": n_part, "steps": n_timed, "warmup": n_warm,
           "ms_per_step": float(np.mean(tot)), "ms_per_step_median": float(np.median(tot)),
           "frames_per_sec": 1e3 / float(np.mean(tot)),
           "scoring_kernel": "kernEvaluateParticlesKD" if kd else "kernEvaluateParticles", "scoring_ms": score_ms,
           "timing": "wall clock around the reference's blocking step functions (its own method, kernel.cu:1727-1748); scoring kernel by CUDA events"}
    if not kd:
        out["phases_ms"] = dict(zip(["motion", "measurement", "map", "resample"], [float(v) / n_timed for v in phases]))
    else:
        out["kd_size"] = int(lib.t3_kd_size())
    lib.t3_init(scene.encode())            # frees the 320 MB device tree of this run; fresh state
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import gpu_icp_slam_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.particles
    K, W = args.steps, max(args.warmup, 3)
    kd = args.path == "kd"
    dataset = default_dataset(args)
    # kd path: the tree has to grow to a realistic size before anything is timed (the reference reports
    # 200-500 k points for whole runs): --prewarm frames of the trajectory, untimed
    PW = args.prewarm if args.prewarm >= 0 else (2000 if kd else 0)
    if dataset == "synthetic":
        scans, order, data_desc = synthetic_scans(PW + K + W + 8)
    else:
        scans, order, data_desc = workload_scans(PW + K + W + 8, dataset)
    # everything runs on one non-default stream (the legacy default stream cannot be graph-captured)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    if world > 1:
        from gpu_icp_slam_b200.dist import ShardedParticleFilter
        pf = ShardedParticleFilter(n, device=local_rank, exchange=args.exchange,
                                   path=g.PATH_KD if kd else g.PATH_GRID2D)
    else:
        pf = g.ParticleFilter(n, device=local_rank, path=g.PATH_KD if kd else g.PATH_GRID2D)
        pf.set_stream(stream.cuda_stream)

    scans_dev = torch.from_numpy(scans).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step_async(i):
        if world > 1:
            pf.step_device(scans_dev[i].data_ptr(), order[i])
        else:
            pf.step_async(order[i], scans_dev[i].data_ptr())

    # ---- warm-up (untimed): builds the map, first-launch costs
    for i in range(PW + W):
        step_async(i)
        if world > 1 and i == 2 and args.graph:
            sync_all()
            pf.enable_graph()
    sync_all()
    scans, order, scans_dev = scans[PW:], order[PW:], scans_dev[PW:]
    eng0 = pf.engine if world > 1 else pf
    r0 = eng0.fetch_result()
    resample_count0, wait0 = r0.resample_count, (r0.wait_extrema_ns, r0.wait_tiles_ns)

    # ---- `value`: K steps, inputs resident in HBM, device time per step by CUDA events on the launch
    # stream, L2 flushed (256 MiB memset, untimed) between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pf.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sync_all()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xff)
        ev[k][0].record(stream)
        step_async(W + k)
        ev[k][1].record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = pf.launch_count - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(step_ms)
    # the filter settles into resampling every other frame: even / odd steps separate the two kinds of step
    even_odd_ms = [float(np.mean(step_ms[0::2])), float(np.mean(step_ms[1::2]))] if K >= 2 else None
    r1 = eng0.fetch_result()
    resampled_steps = r1.resample_count - resample_count0
    # time this rank's kernels spent waiting for the peers' extrema / tile sums, per step (in-kernel %globaltimer stamps)
    wait_us = [((r1.wait_extrema_ns - wait0[0]) & 0xffffffff) / 1e3 / K, ((r1.wait_tiles_ns - wait0[1]) & 0xffffffff) / 1e3 / K]

    # back-to-back (no flush) for reference: the streaming steady state
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        step_async(W + k)
    e1.record(stream)
    sync_all()
    b2b_ms = e0.elapsed_time(e1)

    # ---- roofline of the dominant kernel (scoring): CUDA events around that kernel alone, recorded
    # on the launch stream INSIDE full steps of a second timed pass (L2 flushed between steps)
    eng = pf.engine if world > 1 else pf
    saved_graph = None
    if world > 1 and getattr(pf, "_graph", None) is not None:      # event pairs need plain launches
        saved_graph, pf._graph = pf._graph, None
        eng.set_external_params(False)
    eng.profile_enable(True)
    for k in range(K):
        flush.fill_(k & 0xff)
        step_async(W + k)
    ker, n_prof = eng.profile_read()
    eng.profile_enable(False)
    if saved_graph is not None:
        eng.set_external_params(True)
        pf._graph = saved_graph
    ker = max(ker, 1e-6)
    # the sampler has watched the timed region, the back-to-back pass and the roofline pass (GPU busy throughout)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel breakdown of serialised steps (plain launches, an event after every kernel), L2 flushed
    laps = {}
    if not kd and world == 1:
        eng.profile_laps(True)
        for k in range(min(K, 100)):
            flush.fill_(k & 0xff)
            step_async(W + k)
        laps = eng.profile_laps_read()
        eng.profile_laps(False)
    iso_ms = [eng.profile_score() for _ in range(10)] if not kd else [(0.0, 0.0)]
    r = eng.fetch_result()

    # ---- e2e: the public host API, host scan in, host pose out, every step
    sync_all()
    t0 = time.perf_counter()
    for k in range(K):
        if world > 1:
            pf.step(scans[W + k], order[W + k])
        else:
            pf.step(scans[W + k], order[W + k])
    sync_all()
    e2e_s = time.perf_counter() - t0
    # the streaming flavour of the same API (pfslam_submit / pfslam_wait, two frames in flight): every frame's scan still
    # comes from host memory and every frame's result is read on the host; the host's launch latency overlaps the device
    e2e_stream_s = None
    if not kd:
        try:
            sync_all()
            t0 = time.perf_counter()
            pending = []
            for k in range(K):
                pending.append(eng0.submit(scans[W + k], order[W + k]))
                if len(pending) >= 2:
                    eng0.wait(pending.pop(0))
            while pending:
                eng0.wait(pending.pop(0))
            sync_all()
            e2e_stream_s = time.perf_counter() - t0
        except Exception as ex:          # streaming needs the captured step
            print("streaming e2e skipped: %s" % ex, file=sys.stderr)

    # max over ranks
    t = torch.tensor([dev_ms, b2b_ms, e2e_s * 1e3, ker, wait_us[0], wait_us[1], (e2e_stream_s or 0.0) * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, b2b_ms, e2e_ms, ker, wait_ext_us, wait_tiles_us, e2e_stream_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        scale = world * n / 65536.0                      # 65 536-particle frame equivalents per frame
        alg_bytes = n * (N_BEAMS + 20) + 4 * N_BEAMS     # N*1101 + 4324 per launch (per GPU)
        kd_visits = None
        if kd:   # one node per visit per (particle, beam), with the MEASURED mean visited-node count; the
            # scorer's walk loads the 16-byte search shadow of a node (pf_kernels_kd.cuh), plus the 32-byte winner
            kd_visits = eng.kd_mean_visits(4096)
            alg_bytes = int(n * N_BEAMS * (16.0 * kd_visits + 32.0))
        achieved = alg_bytes / (ker * 1e-3) / 1e9
        traffic, traffic_src = (None, "") if kd else scorer_traffic()
        line = {
            "metric": METRIC + ("_kd" if kd else ""), "value": K / (dev_ms * 1e-3) * scale, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": data_desc,
            "config": make_config(args, world, dataset),
            "measurement": {
                       "score_mode": "tiled (TMA-staged smem windows, bit-exact)", "l2_flush": "256 MiB memset between timed steps (untimed)",
                       "timing": "sum of per-step CUDA-event durations on the launch stream, max over ranks",
                       "parallelism": "particles sharded %d-way, map replicated" % world,
                       "prewarm_frames": PW,
                       "exchange": ("none (single GPU)" if world == 1 else
                                    "in-kernel stores/loads over NVLink peer memory, whole step = 1 CUDA graph per rank" if args.exchange == "peer"
                                    else "3 NCCL all-gathers between the step's phases")},
            "resampled_steps": int(resampled_steps), "ms_per_step_even_odd": even_odd_ms,
            "exchange_wait_us_per_step": ({"extrema": wait_ext_us, "tiles": wait_tiles_us,
                                           "note": "max over ranks of the time the step's kernels spent in the two peer waits"} if world > 1 else None),
            "value_back_to_back": K / (b2b_ms * 1e-3) * scale,
            "wall_s_timed_region": t_wall,
            "e2e": {"value": K / (e2e_ms * 1e-3) * scale, "unit": UNIT,
                    "h2d_bytes_per_step": 4 * N_BEAMS, "d2h_bytes_per_step": C.sizeof(g.FrameResult),
                    "call": "pfslam_step (blocking): host scan in, host result out, one graph launch + one synchronisation per frame"},
            "e2e_streaming": ({"value": K / (e2e_stream_ms * 1e-3) * scale, "unit": UNIT, "frames_in_flight": 2,
                               "call": "pfslam_submit / pfslam_wait over the pinned scan ring (same copies, host launch latency overlapped)"}
                              if e2e_stream_ms > 0 else None),
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_score_kd" if kd else "k_score_tiled", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kd_mean_visited_nodes_per_walk": kd_visits,
                         "note": ("kd: node loads are served by L1/L2 (the tree is a few MB), so this algorithmic-byte rate is not "
                                  "DRAM traffic and may exceed the HBM peak; the kernel is bound by divergent instruction issue (DESIGN.md 5.3)" if kd else None),
                         "kernel_ms": ker, "kernel_launches_timed": n_prof,
                         "kernel_ms_isolated": float(np.mean([a for a, _ in iso_ms])),
                         "scoring_phase_ms_isolated": float(np.mean([b for _, b in iso_ms]))},
            "kernels_ms_serialised": {k: round(v[0], 5) for k, v in laps.items()},
            "clocks": clocks,
            "last_frame": {"neff": r.neff, "resampled": r.resampled, "n_slow_evals": r.n_slow_evals,
                           "slow_eval_frac": r.n_slow_evals / float(n * N_BEAMS), "kd_size": r.kd_size},
        }
        if not args.no_cpu and world == 1:
            fps, kind, what = cpu_reference_frames_per_sec(n, max(2, min(8, int(20.0 / (n * N_BEAMS * 35e-9 + 1e-3)))), 1,
                                                           "train_lidar0" if dataset == "synthetic" else dataset)
            if kd:
                what += " (the reference has no CPU kd search: its CPU 2D scoring path on the same scans)"
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": 1, "kind": kind, "sample": what}
        if not args.no_refgpu and world == 1 and n == 65536:
            pf.close()
            del flush
            torch.cuda.empty_cache()
            nt = min(K, 100 if not kd else 60)
            # kd: the reference's step costs tens of ms at this particle count, so its tree is grown over fewer
            # frames than ours (a smaller tree favours the reference)
            rpw = min(PW, args.refgpu_prewarm) if kd else 0
            wsc, word, _ = (synthetic_scans if dataset == "synthetic" else (lambda m: workload_scans(m, dataset)))(rpw + W + nt)
            rg = reference_gpu_leg(wsc, word, rpw + W, nt, kd)
            rg["prewarm_frames"] = rpw
            line["reference_gpu"] = rg
            if "ms_per_step" in rg:
                rg["ours_over_reference_step"] = rg["ms_per_step"] / (dev_ms / K)
                rg["ours_over_reference_scoring_kernel"] = rg["scoring_ms"] / ker
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=65536, help="particles per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", action="store_true",
                    help="multi-GPU: capture the sharded step (kernels + all-gathers) in one CUDA graph (experimental)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"],
                    help="multi-GPU transport: kernels over NVLink peer memory (default) or NCCL all-gathers between phases")
    ap.add_argument("--data", default="auto", choices=["auto", "train_lidar0", "train_lidar2", "train_lidar3", "synthetic"],
                    help="scans: auto = the dataset BASELINE.json names for this path / GPU count (configs[1..3]); "
                         "synthetic = the corridor of configs[4]")
    ap.add_argument("--prewarm", type=int, default=-1, help="untimed frames before the warm-up (default: 2000 on the kd path, else 0)")
    ap.add_argument("--refgpu-prewarm", type=int, default=300, help="kd path: frames the reference-GPU leg grows its tree over")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the reference-kernel-on-this-GPU leg")
    ap.add_argument("--path", default="grid2d", choices=["grid2d", "kd"], help="map representation (BASELINE configs 2 / 3)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
