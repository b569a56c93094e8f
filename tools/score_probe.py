#!/usr/bin/env python
"""Isolated timing of the scoring kernel on a warmed map (CUDA events inside the engine, pfslam_profile_score):
    [PFSLAM_STAGED_DEBUG=..] [PFSLAM_STAGED_VARIANT=..] python tools/score_probe.py [frames] [particles]
Prints the mean / min kernel time and the scoring-phase time over 20 launches (L2 not flushed)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_icp_slam_b200 as g
from gpu_icp_slam_b200 import scans as S

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = os.path.join(root, "data", "_cache", "train_lidar0.scans.u16")
sc = S.load(p)[1500:] if os.path.exists(p) else S.load(os.path.join(root, "tests", "golden", "train_lidar0_first256.scans.u16"))
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
with g.ParticleFilter(n) as pf:
    import time
    t0 = time.time()
    f = 0
    while f < frames or time.time() - t0 < 2.5:        # keep the GPU busy long enough for its clocks to reach the maximum
        f += 1
        pf.step(sc[1 + (f - 1) % frames], f)
    pf.phase_motion(f + 1)
    pf.upload_scan(sc[1 + f % frames])
    t = np.array([pf.profile_score() for _ in range(60)])
    print("dbg=%s var=%s thr=%s: kernel mean %.2f us, min %.2f us; phase mean %.2f us" % (
        os.environ.get("PFSLAM_STAGED_DEBUG", "0"), os.environ.get("PFSLAM_STAGED_VARIANT", "-"), os.environ.get("PFSLAM_STAGED_THREADS", "-"),
        t[20:, 0].mean() * 1e3, t[:, 0].min() * 1e3, t[20:, 1].mean() * 1e3))
    if int(os.environ.get("PFSLAM_STAGED_DEBUG", "0")) & 16:
        from gpu_icp_slam_b200 import engine
        ts = np.zeros((148, 12), np.uint64)
        engine.load_library().pfslam_debug_staged_timing(ts.ctypes.data, ts.size)
        t = ts.astype(np.float64)
        t0 = t[:, 0].min()
        names = ["entry(after first block)", "dependency wait", "prologue", "stage loads", "piece set-up", "gather+queue", "REDs", "drain"]
        vals = [t[:, 0] - t0, t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3], t[:, 4], t[:, 5], t[:, 6], t[:, 7]]
        for nm, v in zip(names, vals):
            print("  %-26s mean %7.2f us  max %7.2f us" % (nm, v.mean() / 1e3, v.max() / 1e3))
        print("  %-26s mean %7.2f us  max %7.2f us; pieces/block %.1f, stage loads/block %.1f" % (
            "block lifetime", (t[:, 8] - t[:, 0]).mean() / 1e3, (t[:, 8] - t[:, 0]).max() / 1e3, t[:, 9].mean(), t[:, 10].mean()))
        print("  kernel span (first entry -> last exit) %.2f us" % ((t[:, 8].max() - t0) / 1e3))
        life = (t[:, 8] - t[:, 0]) / 1e3
        order = np.argsort(-life)
        print("  slowest / fastest blocks: (block, lifetime us, gather us, pieces, last stage, kind, beams in stage, slice units)")
        for b in list(order[:8]) + list(order[-4:]):
            w = int(ts[b, 11])
            print("   %4d  %6.1f  %6.1f  %2d   stage %2d kind %d beams %3d units %d" % (b, life[b], t[b, 5] / 1e3, int(t[b, 9]), w & 255, (w >> 8) & 255, (w >> 16) & 0xffff, w >> 32))
