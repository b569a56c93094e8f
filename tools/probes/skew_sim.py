#!/usr/bin/env python
"""Shared-memory bank-conflict model of the tiled scorer's gather layout ("One-instruction
addressing"): for the 32 particles of a warp and one beam, the gather addresses are
    addr = idx + (idx >> s),  idx = x*256 + y      (one PRMT + one LEA.HI)
i.e. row pitch 256 + 2^(8-s) with every 2^s-byte group of a row displaced by one more byte.  Counts the
wavefronts (max distinct 32-bit words per bank) on clouds drawn from the motion model over real scans.

    python tools/probes/skew_sim.py        # prints the average wavefronts per gather for s = 6, 5, 4, 3
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gpu_icp_slam_b200 import scans as S  # noqa: E402

RES = 0.025


def wavefronts(scans, shift, scale=1.5, frames=(20, 80, 150, 220), seed=1):
    rng = np.random.default_rng(seed)
    ang = np.radians(-135 + 0.25 * np.arange(1081))
    tot = cnt = 0
    for f in frames:
        sc = scans[f]
        for trial in range(6):
            px = rng.normal(0, 0.015 * scale, 32)
            py = rng.normal(0, 0.015 * scale, 32)
            th = rng.normal(0, 0.01 * scale, 32) + 0.3 * trial
            for j in np.flatnonzero(sc < 20)[::9]:
                x = np.round((px + sc[j] * np.cos(ang[j] + th)) / RES).astype(int)
                y = np.round((py + sc[j] * np.sin(ang[j] + th)) / RES).astype(int)
                x -= x.min() - 10
                y -= y.min() - (10 + j % 4)
                idx = x * 256 + y
                word = (idx + (idx >> shift)) >> 2
                bank = word % 32
                tot += max(len(np.unique(word[bank == b])) for b in np.unique(bank))
                cnt += 1
    return tot / cnt


if __name__ == "__main__":
    scans = S.load(os.path.join(ROOT, "tests", "golden", "train_lidar0_first256.scans.u16"))
    for s in (6, 5, 4, 3):
        print("shift %d  pitch %d  bank = %d x + y/4: %.3f wavefronts per gather" %
              (s, 256 + (256 >> s), (64 + (64 >> s)) % 32, wavefronts(scans, s)))
