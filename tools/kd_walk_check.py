#!/usr/bin/env python
"""GPU check of the kd scorer's visit variants (PFSLAM_KD_WALK=1 default, =2 branch-free visit): both must give
identical frame results and scores; prints the mean k_score_kd time of each.  python tools/kd_walk_check.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_icp_slam_b200 as g  # noqa: E402
from gpu_icp_slam_b200 import scans as S  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    sc = S.load(os.path.join(ROOT, "tests", "golden", "train_lidar0_first256.scans.u16"))
    eng = {}
    for walk in (1, 2):
        os.environ["PFSLAM_KD_WALK"] = str(walk)
        eng[walk] = g.ParticleFilter(n, path=g.PATH_KD)
    ok, ms = True, {}
    for walk, pf in eng.items():
        pf.profile_enable(True)
    for f in range(1, 25):
        res = {w: pf.step(sc[f], f) for w, pf in eng.items()}
        a, b = res[1], res[2]
        same = (list(a.pose) == list(b.pose) and (a.fit_min, a.fit_max, a.best_index, a.kd_size, a.resampled) ==
                (b.fit_min, b.fit_max, b.best_index, b.kd_size, b.resampled) and a.neff == b.neff)
        if not same:
            ok = False
            print("MISMATCH at frame", f, a.as_dict(), b.as_dict())
            break
    for walk, pf in eng.items():
        ms[walk] = pf.profile_read()
    fit = {w: pf.score_particles(sc[30]) for w, pf in eng.items()}
    ok = ok and bool(np.array_equal(fit[1], fit[2]))
    print("kd_walk_check n=%d: %s; k_score_kd walk1 %.3f ms, walk2 %.3f ms (%d launches), kd_size %d" %
          (n, "IDENTICAL" if ok else "DIFFERENT", ms[1][0], ms[2][0], ms[2][1], res[1].kd_size), flush=True)
    for pf in eng.values():
        pf.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
