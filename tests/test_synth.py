"""Synthetic workload of BASELINE.json configs[4] (SURVEY 8(d) "Config 5"): reproducible, well-formed, and
usable by the scan codec and the oracle."""
import hashlib

import numpy as np

import helpers  # noqa: F401  (puts the repo root on sys.path)
from gpu_icp_slam_b200 import scans as S
from gpu_icp_slam_b200 import synth


def test_synthetic_scans_are_reproducible_and_well_formed():
    a, pa = synth.generate(64)
    b, pb = synth.generate(64)
    assert a.shape == (64, 1081) and a.dtype == np.float32
    assert np.array_equal(a, b) and np.array_equal(pa, pb)                    # PCG64(565): bit-for-bit
    assert hashlib.sha256(synth.generate(200)[0].tobytes()).hexdigest()[:16] == "91837442bab0936d"
    valid = a[a != synth.SENTINEL]
    assert valid.min() >= 0.02 and valid.max() <= 30.0
    assert 0.0 < (a == synth.SENTINEL).mean() < 0.005                          # ~0.1 % invalid returns
    assert np.array_equal(np.round(valid.astype(np.float64) * 1000) / 1000, valid.astype(np.float64).round(3))
    # corridor geometry: 4 m wide, so the nearest wall is about 2 m to either side of the centre line
    side = a[0, [180, 900]]                                                    # beams at -90 and +90 degrees
    assert np.all(np.abs(side - 2.0) < 0.06)
    # 0.5 m/s at 40 Hz along +x of the first pose's frame
    assert np.allclose(pa[0], 0.0) and abs(pa[40, 0] - 0.5) < 1e-9 and abs(pa[40, 1]) < 1e-9
    # the packed scan format holds it exactly
    assert np.array_equal(S.decode(S.encode(a)), a)


def test_oracle_runs_on_the_synthetic_corridor():
    """the 2D oracle filter digests the synthetic scans: it stays centred between the corridor walls and
    keeps the heading (motion along a straight corridor is not observable from its side walls, and the
    reference has no odometry, so the along-track position is not asserted)"""
    sc, poses = synth.generate(61)
    of = helpers.OracleFilter(512)
    for f in range(1, 61):
        s = of.step(sc[f], f)
    lateral, heading = abs(s.robot[1] - poses[60, 1]), abs(s.robot[2] - poses[60, 2])
    explored = int((of.grid != -100).sum())
    of.close()
    assert lateral < 0.05 and heading < 0.02, (lateral, heading)
    assert explored > 50000, explored            # the corridor floor around the robot has been cleared
