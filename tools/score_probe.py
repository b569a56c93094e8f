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
    for f in range(1, frames + 1):
        pf.step(sc[f], f)
    pf.phase_motion(frames + 1)
    pf.upload_scan(sc[frames + 1])
    t = np.array([pf.profile_score() for _ in range(20)])
    print("dbg=%s var=%s thr=%s: kernel mean %.2f us, min %.2f us; phase mean %.2f us" % (
        os.environ.get("PFSLAM_STAGED_DEBUG", "0"), os.environ.get("PFSLAM_STAGED_VARIANT", "-"), os.environ.get("PFSLAM_STAGED_THREADS", "-"),
        t[5:, 0].mean() * 1e3, t[:, 0].min() * 1e3, t[5:, 1].mean() * 1e3))
