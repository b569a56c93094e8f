#!/usr/bin/env python
"""Headless map export (the reference renders through GL / PCL, src/draw.cu and
src/main.cpp:231-327): occupancy grid -> binary PGM, kd point cloud -> ASCII PCD.

    python tools/export_map.py <scans.u16 | .mat> <out prefix> [--particles N] [--frames F] [--kd]

Runs the CUDA engine over the scans and writes <prefix>.pgm (the reference's drawMap grey scale,
src/draw.cu:107-109: pixel = 1 - (cell + 128) / 256, so unknown (-100) is light grey, free lighter,
occupied dark) or <prefix>.pcd (x y z intensity = node weight).
The two writers take plain arrays, so they also serve getPCData() output of a running filter."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_pgm(path, grid):
    """grid: int8 [map_w, map_h] (idx = x*map_w + y, the reference's layout); -100 = never seen"""
    g = np.asarray(grid, dtype=np.int8)
    img = np.clip(127 - g.astype(np.int16), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(img.T[::-1])                       # x to the right, y up
    with open(path, "wb") as fh:
        fh.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        fh.write(img.tobytes())


def write_pcd(path, kd_nodes):
    """kd_nodes: int32 [n, 8] words in KDTree::Node layout (axis, left, right, parent, x, y, z, w)"""
    a = np.ascontiguousarray(kd_nodes, dtype=np.int32).reshape(-1, 8)
    xyzw = a[:, 4:8].copy().view(np.float32)
    with open(path, "w") as fh:
        fh.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                 "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA ascii\n" % (len(a), len(a)))
        for x, y, z, w in xyzw:
            fh.write("%.4f %.4f %.4f %.1f\n" % (x, y, z, w))


def main():
    import argparse
    import gpu_icp_slam_b200 as g
    ap = argparse.ArgumentParser()
    ap.add_argument("scans")
    ap.add_argument("prefix")
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--kd", action="store_true")
    a = ap.parse_args()
    lidar = g.Lidar(a.scans)
    last = min(len(lidar.scans) - 1, a.frames or len(lidar.scans) - 1)
    with g.ParticleFilter(a.particles, path=g.PATH_KD if a.kd else g.PATH_GRID2D) as pf:
        for f in range(1, last + 1):                              # main.cpp:199-206: scan 0 is never consumed
            pf.step(lidar.scans[f], f)
        if a.kd:
            write_pcd(a.prefix + ".pcd", pf.get_kd())
        else:
            write_pgm(a.prefix + ".pgm", pf.get_grid())
    print("wrote", a.prefix + (".pcd" if a.kd else ".pgm"))


if __name__ == "__main__":
    main()
