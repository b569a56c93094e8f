#!/usr/bin/env python
"""Generates tests/golden/ref_vectors.npz from the REFERENCE's own host functions
(oracle/_ref/libref.so, built by oracle/Makefile from /root/reference/src).  Run in the build
container (the reference does not travel to the GPU box); the committed vectors pin the oracle.

    python tools/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from helpers import P  # noqa: E402

TRACE_CASES = np.array([
    # sx, sy, ex, ey
    [800, 800, 900, 830], [800, 800, 700, 830], [800, 800, 830, 900], [800, 800, 830, 700],
    [800, 800, 700, 770], [800, 800, 770, 700], [800, 800, 800, 900], [800, 800, 900, 800],
    [800, 800, 800, 800], [800, 800, 801, 800], [5, 5, -40, 17], [1590, 1590, 1700, 1650],
    [10, 1595, 60, 1700], [800, 800, 1599, 1599], [800, 800, 0, 0], [803, 797, 1200, 1196],
    [100, 100, 107, 300], [100, 100, 93, -20], [-5, -5, 40, 30], [1599, 0, 1400, 300],
], dtype=np.int32)


def main():
    ref = helpers.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libref.so missing: run `make -C oracle ref` first")
    out = {}
    hin = np.array([0, 1, 2, 3, 12345, 0x7fffffff, 0x80000000, 0x80000001, 0xdeadbeef, 0xffffffff], dtype=np.uint32)
    out["hash_in"] = hin
    out["hash_out"] = np.array([ref.ref_utilhash(int(a)) for a in hin], dtype=np.uint32)
    # ParticleAddNoise (host build: g++ evaluates the three draws right-to-left)
    noise = []
    for frame, idx0 in ((1, 0), (2, 0), (777, 1000), (12047, 65000)):
        x = np.zeros(64, np.float32); y = x.copy(); th = x.copy()
        ref.ref_add_noise(P(x), P(y), P(th), 64, frame, idx0)
        noise.append(np.stack([x, y, th]))
    out["noise_cases"] = np.array([[1, 0], [2, 0], [777, 1000], [12047, 65000]], dtype=np.int32)
    out["noise_out"] = np.stack(noise)
    # CleanLidarScan
    cl_in = np.array([[0, 1.0, 0.0], [540, 2.5, 0.1], [1080, 19.99, -0.2], [300, 4294967.0, 3.0], [777, 0.001, -3.1]], dtype=np.float64)
    cl = np.zeros((len(cl_in), 2), np.float32)
    for k, (n, r, t) in enumerate(cl_in):
        tmp = np.zeros(2, np.float32)
        ref.ref_clean_lidar_scan(int(n), C.c_float(r), C.c_float(t), P(tmp))
        cl[k] = tmp
    out["clean_in"] = cl_in
    out["clean_out"] = cl
    # EvaluateParticle on the deterministic synthetic grid
    scans = helpers.fixture_scans()
    grid = helpers.synth_grid()
    ev = []
    ev_frames = [1, 60, 200]
    cases = [dict(salt=1, spread=0.3, spread_th=0.2), dict(salt=2, spread=8.0, spread_th=3.0),
             dict(salt=3, spread=1.0, spread_th=3.0, center=(19.5, -19.5, 0.0))]
    for cs in cases:
        x, y, th = helpers.synth_particles(256, **cs)
        for f in ev_frames:
            sc = np.ascontiguousarray(scans[f])
            fit = np.zeros(256, np.int32)
            ref.ref_evaluate_particles(P(grid, helpers.bp), 1600, 1600, 40.0, 40.0, np.float32(0.025), np.float32(0.025),
                                       P(x), P(y), P(th), 256, P(sc), P(fit, helpers.ip))
            ev.append(fit)
    out["eval_frames"] = np.array(ev_frames, dtype=np.int32)
    out["eval_out"] = np.stack(ev).reshape(len(cases), len(ev_frames), 256)
    # traceRay
    out["trace_cases"] = TRACE_CASES
    idxs, lens = [], []
    for sx, sy, ex, ey in TRACE_CASES:
        m = np.zeros(1600 * 1600, np.uint8)
        ref.ref_trace_ray(int(sx), int(sy), int(ex), int(ey), 1600, 1600, P(m, helpers.ubp))
        nz = np.flatnonzero(m).astype(np.int32)
        idxs.append(nz); lens.append(len(nz))
    out["trace_len"] = np.array(lens, dtype=np.int32)
    out["trace_idx"] = np.concatenate(idxs) if idxs else np.zeros(0, np.int32)
    # Scene parser
    m6 = np.zeros(6, np.float32)
    ref.ref_scene_map(os.path.join(ROOT, "oracle", "_ref", "map_settings.txt").encode(), P(m6))
    out["scene_map"] = m6
    dst = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
