"""GPU parity, kd-tree point-cloud path: CUDA engine vs the oracle, bit for bit."""
import ctypes as C

import numpy as np
import pytest

import helpers
from helpers import P

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _gpu():
    import gpu_icp_slam_b200 as g
    return g


def _oracle_tree(scans, n_frames=20, n=64):
    of = helpers.OracleKdFilter(n)
    for f in range(1, n_frames + 1):
        of.step(scans[f], f)
    return of


def test_kd_nn_lookup_bit_exact(scans):
    """findCorrespondenceIndexKD alone, on a tree grown by the oracle over 20 real scans"""
    g = _gpu()
    of = _oracle_tree(scans)
    tree = of.tree.copy()
    o = helpers.load_oracle_kd()
    rng = np.random.default_rng(1)
    xy = tree[:, 4:6].copy().view(np.float32)
    q = np.zeros((4000, 3), np.float32)
    q[:2000, :2] = xy[rng.integers(0, len(tree), 2000)] + rng.normal(0, 0.03, (2000, 2)).astype(np.float32)
    q[2000:, :2] = rng.uniform(-12, 12, (2000, 2)).astype(np.float32)
    with g.ParticleFilter(32, path=g.PATH_KD) as pf:
        pf.set_kd(tree)
        got = pf.kd_nn(q)
    want = np.array([o.pfo_kd_nn(tree.ctypes.data, float(a), float(b), float(c)) for a, b, c in q], np.int32)
    assert np.array_equal(got, want)
    of.close()


def test_kd_scoring_bit_exact(scans):
    """kernEvaluateParticlesKD parity: scores of 2000 particles against an oracle-grown tree"""
    g = _gpu()
    of = _oracle_tree(scans)
    tree = of.tree.copy()
    o = helpers.load_oracle_kd()
    n = 2000
    x, y, th = helpers.synth_particles(n, salt=4, spread=0.2, spread_th=0.1,
                                       center=(float(of.s.contents.robot[0]), float(of.s.contents.robot[1]), float(of.s.contents.robot[2])))
    cfg = helpers.ocfg()
    sc = np.ascontiguousarray(scans[21])
    want = np.array([o.pfo_kd_score(C.byref(cfg), tree.ctypes.data, float(a), float(b), float(c), P(sc))
                     for a, b, c in zip(x, y, th)], np.int32)
    with g.ParticleFilter(n, path=g.PATH_KD) as pf:
        pf.set_kd(tree)
        pf.set_particles(x, y, th, np.ones(n, np.float32))
        got = pf.score_particles(sc)
    assert np.array_equal(got, want)
    of.close()


@pytest.mark.parametrize("n", [500, 1000])
def test_kd_free_running_step_bit_exact(scans, n):
    """the whole kd step free-running from an empty map: first-scan build, NN scoring, ICP pose,
    map update with inserts, resample, and the frame-105 rebalance; tree compared node for node"""
    g = _gpu()
    frames = 112
    of = helpers.OracleKdFilter(n)
    with g.ParticleFilter(n, path=g.PATH_KD) as pf:
        for f in range(1, frames + 1):
            r = pf.step(scans[f], f)
            s = of.step(scans[f], f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose differs at frame %d" % f
            assert r.kd_size == s.kd_size, "kd size differs at frame %d" % f
            if f > 1:
                assert (r.fit_min, r.fit_max, r.best_index) == (s.fit_min, s.fit_max, s.best), "extrema at frame %d" % f
                assert np.array_equal(bits([r.neff]), bits([s.neff])) and r.resampled == s.resampled
                assert (r.n_wall_cells, r.n_free_cells, r.kd_inserted) == (s.n_wall_pts, s.n_free_pts, s.n_inserted)
            if f in (1, 2, 30, 104, 105, frames):
                assert np.array_equal(pf.get_kd(), of.tree), "tree differs at frame %d" % f
        x, y, th, w = pf.get_particles()
        assert np.array_equal(bits(x), bits(of.x)) and np.array_equal(bits(th), bits(of.th)) and np.array_equal(bits(w), bits(of.w))
    of.close()


def test_kd_reference_named_interface(scans):
    g = _gpu()
    g.particleFilterFree()
    g.particleFilterInit(g.Scene(), n_particles=256, path=g.PATH_KD)
    lidar = g.Lidar(scans=scans)
    for frame in range(1, 5):
        g.particleFilter(None, frame, lidar)
    parts, grid, kd, npart, nkd, pos = g.getPCData()
    assert kd is not None and nkd == len(kd) and nkd > 300 and kd.shape[1] == 8
    g.particleFilterFree()
