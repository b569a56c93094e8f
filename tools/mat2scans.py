#!/usr/bin/env python
"""CLI wrapper: python tools/mat2scans.py <in.mat> <out.scans.u16> [max_frames]  (see gpu-icp-slam_b200/scans.py)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_icp_slam_b200.scans import *  # noqa: F401,F403
from gpu_icp_slam_b200.scans import encode, mat_to_f32, save

if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    s = mat_to_f32(src)
    if len(sys.argv) > 3:
        s = s[: int(sys.argv[3])]
    save(dst, encode(s))
    print("wrote %s: %d frames x %d beams" % (dst, s.shape[0], s.shape[1]))
