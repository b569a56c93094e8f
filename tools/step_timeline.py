#!/usr/bin/env python
"""In-graph timeline of the 2D step: where each kernel of the captured step really runs (first block entry ->
last block exit, %globaltimer), averaged over steady-state frames at 65 536 particles.

    python tools/step_timeline.py [steps]        (prints a table; used for profiles/*_timeline.md)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_icp_slam_b200 as g
from gpu_icp_slam_b200 import engine, scans as S

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = os.path.join(root, "data", "_cache", "train_lidar0.scans.u16")
sc = S.load(p) if os.path.exists(p) else S.load(os.path.join(root, "tests", "golden", "train_lidar0_first256.scans.u16"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
last = min(len(sc) - 1, 4000)
with g.ParticleFilter(n) as pf:
    f, t0 = 0, time.time()
    while time.time() - t0 < 2.5:                      # map built, clocks at their maximum
        f += 1
        pf.step(sc[1 + (f - 1) % last], f)
    rows = {0: {}, 1: {}}                               # by "this step resampled"
    for k in range(steps):
        f += 1
        pf.synchronize()
        engine.debug_trace(True)
        r = pf.step(sc[1 + (f - 1) % last], f)        # the host-API graph (scan pull + result publish inside)
        pf.synchronize()
        tr = engine.debug_trace(True, read=True)
        t_first = min(a for a, _ in tr.values())
        for name, (a, b) in tr.items():
            rows[int(r.resampled)].setdefault(name, []).append(((a - t_first) / 1e3, (b - t_first) / 1e3))
    engine.debug_trace(False)
    print("last frame: %d windows, %d wide beams, %d exact re-evaluations" % (r.n_windows, r.n_wide_beams, r.n_slow_evals))
    for kind in (0, 1):
        rk = rows[kind]
        if not rk:
            continue
        cnt = max(len(v) for v in rk.values())
        print("\n%d steps that %s:" % (cnt, "resampled" if kind else "did not resample"))
        print("| kernel | first block in (us) | last block out (us) | span (us) |")
        print("|---|---|---|---|")
        for name, v in sorted(rk.items(), key=lambda kv: np.mean([a for a, _ in kv[1]])):
            a, b = np.mean([x for x, _ in v]), np.mean([y for _, y in v])
            print("| %s | %.1f | %.1f | %.1f |" % (name, a, b, b - a))
        print("step (first entry -> last exit): %.1f us" % np.mean([max(rk[nm][i][1] for nm in rk if len(rk[nm]) > i) for i in range(cnt)]))
