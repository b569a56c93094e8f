"""GPU edge cases through the C ABI, each checked bit-exact against the oracle: tiny and ragged
particle counts, degenerate scans, other beam counts and map geometries (which exercise the
scorer fallbacks), and the headless reference-interface driver."""
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _run_pair(n, scans, frames, n_beams=1081, scale=40.0, res=0.025, **kw):
    import gpu_icp_slam_b200 as g
    cfg = helpers.ocfg(n_beams=n_beams, scale=scale, res=res)
    of = helpers.OracleFilter(n, cfg)
    with g.ParticleFilter(n, n_beams=n_beams, scene=g.Scene(size=(scale, scale), res=res), **kw) as pf:
        assert (pf.map_w, pf.map_h) == (cfg.map_w, cfg.map_h)
        for f, sc in frames:
            r = pf.step(sc, f)
            s = of.step(sc, f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose, frame %d" % f
            assert (r.fit_min, r.fit_max, r.best_index, r.resampled) == (s.fit_min, s.fit_max, s.best, s.resampled), "frame %d" % f
            assert np.array_equal(bits([r.sum_w, r.sum_w2]), bits([s.sum_w, s.sum_w2]))
            assert (r.n_free_cells, r.n_wall_cells) == (s.n_free, s.n_wall)
        assert np.array_equal(pf.get_grid().reshape(-1), of.grid)
        x, y, th, w = pf.get_particles()
        assert np.array_equal(bits(x), bits(of.x)) and np.array_equal(bits(w), bits(of.w))
    of.close()


@pytest.mark.parametrize("n", [1, 2, 33, 1023, 1025, 2049])
def test_tiny_and_ragged_particle_counts(scans, n):
    _run_pair(n, scans, [(f, scans[f]) for f in range(1, 25)])


def test_degenerate_scans(scans):
    """all-sentinel, all-zero, NaN-laced, all >= 20 m and constant scans between normal ones"""
    sent = np.full(1081, 4294967.0, np.float32)
    zero = np.zeros(1081, np.float32)
    nan = scans[7].copy(); nan[::5] = np.nan
    far = np.full(1081, 25.0, np.float32)
    const = np.full(1081, 2.0, np.float32)
    seq = [(1, scans[1]), (2, sent), (3, scans[3]), (4, zero), (5, nan), (6, far), (7, const), (8, scans[8]), (9, scans[9])]
    _run_pair(1500, scans, seq)


@pytest.mark.parametrize("theta0", [7.5, 13.0, 25.0, 40.0, -70.0, 2000.0])
def test_large_headings_stay_exact(scans, theta0):
    """heading is never wrapped (here or in the reference): after several turns the float rounding of
    rot = angle + theta grows, and the fast paths must hand the beams / particles it could flip to the exact
    expression.  Scores of all three scorers == oracle for clouds around |theta| = 7.5 .. 2000 rad."""
    import ctypes as C
    import gpu_icp_slam_b200 as g
    from helpers import P
    o = helpers.load_oracle()
    cfg = helpers.ocfg()
    n = 3000
    grid = helpers.synth_grid(salt=77)
    x, y, th = helpers.synth_particles(n, salt=5, spread=0.3, spread_th=0.15, center=(1.0, -2.0, theta0))
    ones = np.ones(n, np.float32)
    for f in (3, 120):
        sc = np.ascontiguousarray(scans[f])
        want = np.zeros(n, np.int32)
        o.pfo_score2d_many(C.byref(cfg), P(grid, helpers.bp), P(x), P(y), P(th), n, P(sc), P(want, helpers.ip))
        for mode in (g.SCORE_EXACT, g.SCORE_FILTERED, g.SCORE_TILED):
            with g.ParticleFilter(n, score_mode=mode) as pf:
                pf.set_grid(grid)
                pf.set_particles(x, y, th, ones)
                got = pf.score_particles(sc)
                assert np.array_equal(got, want), "theta0 %g mode %d frame %d: %d scores differ" % (theta0, mode, f, (got != want).sum())


def test_other_beam_count(scans):
    """a 720-beam sensor (LIDAR_ANGLE(i) over the first 720 beams)"""
    _run_pair(1000, scans, [(f, scans[f][:720]) for f in range(1, 20)], n_beams=720)


@pytest.mark.parametrize("scale,res", [(20.0, 0.05), (30.0, 0.04), (40.0, 0.03)])
def test_other_map_geometries(scans, scale, res):
    """400x400 cells (tiled scorer), 750x750 (row stride not a multiple of 16 -> filtered scorer),
    1333x1333 with a non-integral map centre (-> exact scorer): all must still match the oracle"""
    _run_pair(1200, scans, [(f, scans[f]) for f in range(1, 20)], scale=scale, res=res)


def test_headless_driver_over_reference_symbols(tmp_path, scans):
    """tools/pfslam_run: the reference's main.cpp call sequence over the reference's own
    symbols (libpfslam_kernelh.so), Scene parsed by the reference's scene.cpp; its trajectory must equal
    the ctypes mirror's bit for bit"""
    import gpu_icp_slam_b200 as g
    exe = os.path.join(helpers.ROOT, "tools", "pfslam_run")
    scene = os.path.join(helpers.ORACLE_DIR, "_ref", "map_settings.txt")
    if not (os.path.exists(exe) and os.path.exists(scene)):
        pytest.skip("pfslam_run not built (needs the reference headers at build time)")
    csv = tmp_path / "traj.csv"
    env = dict(os.environ, PFSLAM_PARTICLE_COUNT="2048")
    r = subprocess.run([exe, scene, os.path.join(helpers.GOLDEN, "train_lidar0_first256.scans.u16"), "40", str(csv)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "particles 2048" in r.stdout
    traj = np.loadtxt(csv, delimiter=",")
    with g.ParticleFilter(2048, scene=g.Scene(scene)) as pf:
        for f in range(1, 41):
            p = pf.step(scans[f], f).pose
            assert np.array_equal(bits(traj[f - 1, 1:4]), bits(list(p))), "frame %d" % f


def test_headless_driver_kd_path_over_reference_symbols(tmp_path, scans):
    """the same driver with PFSLAM_PATH=kd: particleFilter() runs the kd step (what the reference's HEAD runs,
    kernel.cu:1714-1745) and getPCData lends the tree (kernel.cu:810-811); trajectory, node count and weight
    sum equal the ctypes mirror's"""
    import gpu_icp_slam_b200 as g
    exe = os.path.join(helpers.ROOT, "tools", "pfslam_run")
    scene = os.path.join(helpers.ORACLE_DIR, "_ref", "map_settings.txt")
    if not (os.path.exists(exe) and os.path.exists(scene)):
        pytest.skip("pfslam_run not built (needs the reference headers at build time)")
    csv = tmp_path / "traj_kd.csv"
    env = dict(os.environ, PFSLAM_PARTICLE_COUNT="1024", PFSLAM_PATH="kd")
    r = subprocess.run([exe, scene, os.path.join(helpers.GOLDEN, "train_lidar0_first256.scans.u16"), "30", str(csv)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    traj = np.loadtxt(csv, delimiter=",")
    with g.ParticleFilter(1024, scene=g.Scene(scene), path=g.PATH_KD) as pf:
        for f in range(1, 31):
            p = pf.step(scans[f], f).pose
            assert np.array_equal(bits(traj[f - 1, 1:4]), bits(list(p))), "frame %d" % f
        kd = pf.get_kd()
    assert "kd_nodes %d " % len(kd) in r.stdout, r.stdout
    assert "kd_weight_sum %.0f" % float(kd[:, 7].copy().view(np.float32).astype(np.float64).sum()) in r.stdout, r.stdout


def test_streaming_ring_equals_blocking_step(scans):
    """pfslam_submit / pfslam_wait (pinned scan ring, up to 8 frames in flight) == pfslam_step, frame for frame, bit for bit;
    a full ring is refused, an unknown ticket too"""
    import gpu_icp_slam_b200 as g
    n = 3000
    with g.ParticleFilter(n) as a, g.ParticleFilter(n) as b:
        want = [a.step(scans[f], f) for f in range(1, 41)]
        got, pending = [], []
        for f in range(1, 41):
            pending.append(b.submit(scans[f], f))
            if len(pending) == g.RING_DEPTH:
                with pytest.raises(g.PfslamError, match="ring full"):
                    b.submit(scans[f], f)
                got.append(b.wait(pending.pop(0)))
        while pending:
            got.append(b.wait(pending.pop(0)))
        with pytest.raises(g.PfslamError):
            b.wait(12345)
        for f, (x, y) in enumerate(zip(want, got), 1):
            assert np.array_equal(bits(list(x.pose)), bits(list(y.pose))), "frame %d" % f
            assert (x.fit_min, x.fit_max, x.best_index, x.resampled, x.n_free_cells, x.n_wall_cells) == \
                   (y.fit_min, y.fit_max, y.best_index, y.resampled, y.n_free_cells, y.n_wall_cells), "frame %d" % f
        assert np.array_equal(a.get_grid(), b.get_grid())
        # and the blocking call still works after the ring has been used
        ra, rb = a.step(scans[41], 41), b.step(scans[41], 41)
        assert np.array_equal(bits(list(ra.pose)), bits(list(rb.pose)))


@pytest.mark.parametrize("threads", ["768", "256"])
def test_staged_scorer_generation_is_bit_exact_too(scans, monkeypatch, threads):
    """k_score_staged (PFSLAM_TILED_KERNEL=staged: window stages resident in shared memory, wide and slow beams in the
    same kernel) is the measured-but-not-default scorer; it must return the same integers as the oracle"""
    if threads == "256":
        pytest.skip("one process, one block shape: the shape is latched at first use (covered by tools/score_probe.py runs)")
    monkeypatch.setenv("PFSLAM_TILED_KERNEL", "staged")
    _run_pair(2500, scans, [(f, scans[f]) for f in range(1, 30)])
    sent = np.full(1081, 4294967.0, np.float32)
    seq = [(1, scans[1]), (2, sent), (3, scans[3]), (4, np.full(1081, 25.0, np.float32)), (5, scans[5])]
    _run_pair(1500, scans, seq)


def test_fused_tail_kernel_is_bit_exact_too(scans, monkeypatch):
    """k_weights_resample (PFSLAM_TAIL=fused: weights, tile scans, prefix and resampling in one launch around a
    grid-wide barrier) is the measured-but-not-default tail; same bits as the oracle, partial last tile included"""
    monkeypatch.setenv("PFSLAM_TAIL", "fused")
    _run_pair(2500, scans, [(f, scans[f]) for f in range(1, 40)])
    _run_pair(5000, scans, [(f, scans[f]) for f in range(1, 24)])
