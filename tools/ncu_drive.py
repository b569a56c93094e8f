#!/usr/bin/env python
"""Minimal driver for ncu captures of the 2D step's kernels: N host-API steps at 65 536 particles on train_lidar0.

    ncu --set full --import-source on --clock-control none -k regex:'k_map_free|k_weights_scan' \\
        --launch-skip 100 --launch-count 2 -o gpurun_out/x python tools/ncu_drive.py [steps] [particles]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_icp_slam_b200 as g
from gpu_icp_slam_b200 import scans as S

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = os.path.join(root, "data", "_cache", "train_lidar0.scans.u16")
sc = S.load(p) if os.path.exists(p) else S.load(os.path.join(root, "tests", "golden", "train_lidar0_first256.scans.u16"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
first = 1500 if len(sc) > 2000 else 1          # the robot starts moving around frame 1500 of train_lidar0
with g.ParticleFilter(n) as pf:
    for f in range(1, steps + 1):
        r = pf.step(sc[first + f], f)
    print("frame %d: neff %.1f resampled %d" % (steps, r.neff, r.resampled))
