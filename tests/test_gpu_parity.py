"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle, bit for bit.

Tolerances: none.  Every comparison below is exact -- integer scores, cell indices and grid bytes
are integers, and the engine's floating-point results (poses, weights, Neff) are specified as
fixed IEEE-754 operation sequences that the oracle repeats (oracle/pfo.h), so float32 values are
compared by their bit patterns.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from helpers import P, MAD_FUSED, TRIG_CUDA

pytestmark = pytest.mark.gpu


def _gpu():
    import gpu_icp_slam_b200 as g
    return g


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_trig_emulation_matches_libdevice(oracle):
    """oracle/pfo.c's restatement of CUDA cosf/sinf == the device functions, on 6M inputs
    covering the angles the filter produces (|rot| < ~3.5) densely and |x| < 1e5 sparsely."""
    from gpu_icp_slam_b200 import engine
    rng = np.random.default_rng(7)
    xs = np.concatenate([
        np.linspace(-4.0, 4.0, 4_000_001, dtype=np.float64).astype(np.float32),
        rng.uniform(-1.0e5, 1.0e5, 1_000_000).astype(np.float32),
        (rng.standard_normal(1_000_000) * 1e-3).astype(np.float32),
        np.array([0.0, -0.0, 1e-30, -1e-30, np.pi / 2, np.pi, -np.pi, 105614.0], dtype=np.float32)])
    c_dev, s_dev = engine.debug_trig(xs)
    cfn, sfn = oracle.pfo_cosf_cuda, oracle.pfo_sinf_cuda
    # vectorise the scalar oracle through a small C loop substitute: ctypes call per element is slow,
    # so check a 200k sample elementwise and the rest through pfo_get_walls-free bulk helper
    idx = np.concatenate([np.arange(0, xs.size, 31), np.arange(xs.size - 8, xs.size)])
    c_or = np.array([cfn(float(v)) for v in xs[idx]], dtype=np.float32)
    s_or = np.array([sfn(float(v)) for v in xs[idx]], dtype=np.float32)
    bad_c = np.flatnonzero(bits(c_or) != bits(c_dev[idx]))
    bad_s = np.flatnonzero(bits(s_or) != bits(s_dev[idx]))
    assert bad_c.size == 0, "cos differs at x=%r" % xs[idx][bad_c[:5]]
    assert bad_s.size == 0, "sin differs at x=%r" % xs[idx][bad_s[:5]]


def test_noise_bit_exact(oracle):
    g = _gpu()
    n = 5000
    with g.ParticleFilter(n) as pf:
        for frame in (1, 2, 999):
            x0, y0, t0 = helpers.synth_particles(n, salt=frame)
            pf.set_particles(x0, y0, t0, np.ones(n, np.float32))
            pf.phase_motion(frame)
            x, y, th, _ = pf.get_particles()
            xo, yo, to = x0.copy(), y0.copy(), t0.copy()
            oracle.pfo_add_noise(P(xo), P(yo), P(to), n, frame, 0)
            assert np.array_equal(bits(x), bits(xo))
            assert np.array_equal(bits(y), bits(yo))
            assert np.array_equal(bits(th), bits(to))


MODES = {"exact": 0, "filtered": 1, "tiled": 2}


@pytest.mark.parametrize("mode", ["exact", "filtered", "tiled"])
@pytest.mark.parametrize("case", ["near_origin", "spread", "edge_of_map"])
def test_score_bit_exact(oracle, scans, mode, case):
    """kernEvaluateParticles parity on a dense pseudo-random grid (every cell nonzero-ish, so any
    cell-index disagreement changes the score)."""
    g = _gpu()
    n = 3000
    grid = helpers.synth_grid()
    if case == "near_origin":
        x, y, th = helpers.synth_particles(n, salt=1, spread=0.05, spread_th=0.02)
    elif case == "spread":
        x, y, th = helpers.synth_particles(n, salt=2, spread=8.0, spread_th=3.0)
    else:
        x, y, th = helpers.synth_particles(n, salt=3, spread=1.0, spread_th=3.0, center=(19.5, -19.5, 0.0))
    cfg = helpers.ocfg(TRIG_CUDA, MAD_FUSED)
    with g.ParticleFilter(n, score_mode=MODES[mode]) as pf:
        pf.set_grid(grid)
        pf.set_particles(x, y, th, np.ones(n, np.float32))
        for f in (1, 60, 200):
            sc = np.ascontiguousarray(scans[f])
            fit = pf.score_particles(sc)
            fo = np.zeros(n, np.int32)
            oracle.pfo_score2d_many(C.byref(cfg), P(grid, helpers.bp), P(x), P(y), P(th), n, P(sc), P(fo, helpers.ip))
            assert np.array_equal(fit, fo), "frame %d: %d of %d scores differ" % (f, (fit != fo).sum(), n)


def test_score_special_ranges(oracle):
    """sentinel 4294967.0, 0.001, > 20 m, exactly 20 m, zeros: slow-beam path of the filtered scorer"""
    g = _gpu()
    n = 1024
    grid = helpers.synth_grid(salt=5)
    x, y, th = helpers.synth_particles(n, salt=9, spread=2.0, spread_th=1.0)
    sc = np.full(1081, 3.0, np.float32)
    sc[::7] = 4294967.0
    sc[1::7] = 0.001
    sc[2::7] = 25.0
    sc[3::7] = 20.0
    sc[4::7] = 19.999
    sc[5::7] = 0.0
    cfg = helpers.ocfg(TRIG_CUDA, MAD_FUSED)
    fo = np.zeros(n, np.int32)
    oracle.pfo_score2d_many(C.byref(cfg), P(grid, helpers.bp), P(x), P(y), P(th), n, P(sc), P(fo, helpers.ip))
    for sm in (g.SCORE_EXACT, g.SCORE_FILTERED, g.SCORE_TILED):
        with g.ParticleFilter(n, score_mode=sm) as pf:
            pf.set_grid(grid)
            pf.set_particles(x, y, th, np.ones(n, np.float32))
            assert np.array_equal(pf.score_particles(sc), fo)


def test_update_grid_bit_exact(oracle, scans):
    """PFUpdateMap parity: Bresenham free cells, wall cells, clamped +-113 updates, cell counts"""
    g = _gpu()
    cfg = helpers.ocfg(TRIG_CUDA, MAD_FUSED)
    nc = cfg.map_w * cfg.map_h
    poses = [(0.0, 0.0, 0.0), (1.234, -2.5, 0.7), (-19.9, 19.9, 2.0), (19.99, 0.0, -3.0), (25.0, 3.0, 0.3)]
    with g.ParticleFilter(64) as pf:
        grid = helpers.synth_grid(salt=11)
        pf.set_grid(grid)
        go = grid.copy()
        for k, pose in enumerate(poses):
            sc = np.ascontiguousarray(scans[10 + 37 * k])
            pf.update_grid(sc, pose)
            cx, cy = C.c_int(), C.c_int()
            oracle.pfo_center_cell(C.byref(cfg), pose[0], pose[1], C.byref(cx), C.byref(cy))
            fm = np.zeros(nc, np.uint8); wm = np.zeros(nc, np.uint8)
            oracle.pfo_get_walls(C.byref(cfg), P(sc), cx.value, cy.value, C.c_float(pose[2]), P(fm, helpers.ubp), P(wm, helpers.ubp))
            oracle.pfo_apply_masks(P(go, helpers.bp), nc, P(fm, helpers.ubp), P(wm, helpers.ubp))
            r = pf.fetch_result()
            assert (r.n_free_cells, r.n_wall_cells) == (int(fm.sum()), int(wm.sum()))
            assert np.array_equal(pf.get_grid().reshape(-1), go), "pose %r" % (pose,)


@pytest.mark.parametrize("n,mode,q1", [(1000, "tiled", 1), (1000, "filtered", 1), (1000, "exact", 1), (4096, "tiled", 0), (777, "tiled", 1), (5000, "tiled", 1)])
def test_free_running_step_bit_exact(scans, n, mode, q1):
    """The whole 2D step, free-running from the initial state over the fixture frames: pose, score
    extrema, arg-max, Neff, resample decision, map-cell counts every frame; particles, weights and
    the full grid at checkpoints.  No teacher forcing: one differing bit anywhere would diverge."""
    g = _gpu()
    frames = 120 if n <= 1000 else 50
    of = helpers.OracleFilter(n, helpers.ocfg(TRIG_CUDA, MAD_FUSED, q1=q1))
    with g.ParticleFilter(n, score_mode=MODES[mode], quirks=(g.QUIRK_Q1 if q1 else 0)) as pf:
        n_resampled = 0
        for f in range(1, frames + 1):
            r = pf.step(scans[f], f)
            s = of.step(scans[f], f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose differs at frame %d" % f
            assert (r.fit_min, r.fit_max, r.best_index) == (s.fit_min, s.fit_max, s.best), "extrema at frame %d" % f
            assert np.array_equal(bits([r.sum_w, r.sum_w2, r.neff]), bits([s.sum_w, s.sum_w2, s.neff])), "Neff at frame %d" % f
            assert r.resampled == s.resampled
            assert (r.n_free_cells, r.n_wall_cells) == (s.n_free, s.n_wall)
            n_resampled += r.resampled
            if f % 40 == 0 or f == frames:
                x, y, th, w = pf.get_particles()
                assert np.array_equal(bits(x), bits(of.x)) and np.array_equal(bits(y), bits(of.y))
                assert np.array_equal(bits(th), bits(of.th)) and np.array_equal(bits(w), bits(of.w))
                assert np.array_equal(pf.get_grid().reshape(-1), of.grid)
        assert n_resampled > 0, "the run never exercised the resampler"
    of.close()


def test_graph_and_plain_launch_paths_agree(scans):
    """the captured-graph step (default) and the plain-launch step (used while profiling) are the
    same computation"""
    g = _gpu()
    n = 2048
    with g.ParticleFilter(n) as a, g.ParticleFilter(n) as b:
        b.profile_enable(True)                  # forces plain launches
        for f in range(1, 40):
            ra, rb = a.step(scans[f], f), b.step(scans[f], f)
            assert np.array_equal(bits(list(ra.pose)), bits(list(rb.pose))) and ra.best_index == rb.best_index
            assert ra.neff == rb.neff and ra.resampled == rb.resampled
        ms, cnt = b.profile_read()
        assert cnt == 39 and ms > 0
        assert np.array_equal(a.get_grid(), b.get_grid())
        xa, xb = a.get_particles(), b.get_particles()
        assert all(np.array_equal(bits(u), bits(v)) for u, v in zip(xa, xb))
        assert a.launch_count > 39 * 5


def test_reference_named_interface(scans):
    """particleFilterInit / particleFilter / getPCData / particleFilterFree call contract (main.cpp:175-237)"""
    g = _gpu()
    g.particleFilterFree()                       # harmless before Init (main.cpp:194)
    g.particleFilterInit(g.Scene(), n_particles=512)
    lidar = g.Lidar(scans=scans)
    for frame in range(1, 6):
        g.particleFilter(None, frame, lidar)
    parts, grid, kd, npart, nkd, pos = g.getPCData()
    assert parts.shape == (512, 4) and grid.shape == (1600, 1600) and npart == 512 and nkd == 0
    assert (grid != -100).sum() > 10000 and len(pos) == 3
    g.particleFilterFree()


def test_particle_filter_step_alias(scans):
    g = _gpu()
    lib = g.load_library()
    with g.ParticleFilter(256) as a, g.ParticleFilter(256) as b:
        pose = (C.c_float * 3)()
        sc = np.ascontiguousarray(scans[1])
        assert lib.particleFilterStep(a._h, sc.ctypes.data, 1, pose) == 0
        r = b.step(sc, 1)
        assert list(pose) == list(r.pose)


def test_errors_are_reported_not_fatal():
    g = _gpu()
    with pytest.raises(g.PfslamError):
        g.ParticleFilter(0)
    with pytest.raises(g.PfslamError):
        g.ParticleFilter(128, path=7)
    with pytest.raises(g.PfslamError):
        g.ParticleFilter(128, device=99)
    with g.ParticleFilter(128) as pf:
        with pytest.raises(g.PfslamError):
            pf.step(np.zeros(5, np.float32), 1)


def test_full_size_properties_65536(scans):
    """BASELINE configs[1] size (65 536 particles): size-independent properties through the C ABI.
    (i) the three scoring kernels -- exact reference expression, filtered LDG, TMA-tiled -- agree on
    every particle; (ii) free-running invariants: weights in [0,1], reset to 1 by a resample, grid
    within the +-113 clamp, arg-max consistent with the scores, Neff <= N; (iii) the trajectory does
    not depend on the scoring kernel."""
    g = _gpu()
    n = 65536
    grid = helpers.synth_grid(salt=41)
    x, y, th = helpers.synth_particles(n, salt=7, spread=0.15, spread_th=0.08)
    sc = np.ascontiguousarray(scans[33])
    fits = []
    for mode in (g.SCORE_EXACT, g.SCORE_FILTERED, g.SCORE_TILED):
        with g.ParticleFilter(n, score_mode=mode) as pf:
            pf.set_grid(grid)
            pf.set_particles(x, y, th, np.ones(n, np.float32))
            fits.append(pf.score_particles(sc))
    assert np.array_equal(fits[0], fits[1]) and np.array_equal(fits[0], fits[2])
    with g.ParticleFilter(n, score_mode=g.SCORE_TILED) as a, g.ParticleFilter(n, score_mode=g.SCORE_EXACT) as b:
        for f in range(1, 16):
            ra, rb = a.step(scans[f], f), b.step(scans[f], f)
            assert np.array_equal(bits(list(ra.pose)), bits(list(rb.pose))) and ra.best_index == rb.best_index
            assert ra.neff == rb.neff and 0 < ra.neff <= n and ra.fit_min <= ra.fit_max
            _, _, _, w = a.get_particles()
            assert w.min() >= 0.0 and w.max() <= 1.0
            if ra.resampled:
                assert np.all(w == 1.0)
        ga = a.get_grid()
        assert ga.min() >= -113 and ga.max() <= 113 and np.array_equal(ga, b.get_grid())


# ---- long horizons and the headline size, against the oracle (real datasets from data/_cache) -------------
def _dataset(name):
    sc = helpers.full_scans(name)
    if sc is None:
        pytest.skip("data/_cache/%s.scans.u16 is not on this box" % name)
    return sc


def test_free_running_2400_frames_train_lidar0():
    """BASELINE configs[0]/[1] data: 2400 frames of train_lidar0 (scans 1500..3900, the robot driving) free-running at 4096 particles, engine vs
    oracle: pose, extrema, Neff and the resample decision EVERY frame, the whole grid and particle cloud every
    250 frames -- all bit for bit."""
    g = _gpu()
    scans = _dataset("train_lidar0")[1500:]    # the robot stands still for the first ~1500 scans; from here on it drives
    n, frames = 4096, 2400
    of = helpers.OracleFilter(n)
    n_resampled = 0
    with g.ParticleFilter(n) as pf:
        for f in range(1, frames + 1):
            r = pf.step(scans[f], f)
            s = of.step_threaded(scans[f], f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose differs at frame %d" % f
            assert (r.fit_min, r.fit_max, r.best_index, r.resampled) == (s.fit_min, s.fit_max, s.best, s.resampled), "frame %d" % f
            assert np.array_equal(bits([r.sum_w, r.sum_w2, r.neff]), bits([s.sum_w, s.sum_w2, s.neff])), "frame %d" % f
            n_resampled += r.resampled
            if f % 250 == 0:
                assert np.array_equal(pf.get_grid().reshape(-1), of.grid), "grid differs at frame %d" % f
                x, y, th, w = pf.get_particles()
                assert np.array_equal(bits(x), bits(of.x)) and np.array_equal(bits(y), bits(of.y))
                assert np.array_equal(bits(th), bits(of.th)) and np.array_equal(bits(w), bits(of.w))
        assert r.resample_count == n_resampled and n_resampled > 50
    # the robot has driven and mapped
    assert abs(s.robot[0]) + abs(s.robot[1]) > 1.5 and (of.grid > 0).sum() > 2000, (list(s.robot), int((of.grid > 0).sum()))
    of.close()


def test_headline_size_65536_against_oracle():
    """BASELINE configs[1] size against the ORACLE (not scorer against scorer): 12 free-running frames of
    train_lidar0 at 65 536 particles -- every particle's pose and weight, every grid byte, every frame's
    extrema -- then a teacher-forced wide cloud (many beams off their windows, uncertain-pair queue under
    load) scored against the oracle particle by particle."""
    g = _gpu()
    scans = _dataset("train_lidar0")
    n = 65536
    of = helpers.OracleFilter(n)
    with g.ParticleFilter(n) as pf:
        for f in range(1, 13):
            r = pf.step(scans[f], f)
            s = of.step_threaded(scans[f], f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose differs at frame %d" % f
            assert (r.fit_min, r.fit_max, r.best_index, r.resampled) == (s.fit_min, s.fit_max, s.best, s.resampled), "frame %d" % f
            assert np.array_equal(bits([r.sum_w, r.sum_w2, r.neff]), bits([s.sum_w, s.sum_w2, s.neff])), "frame %d" % f
            x, y, th, w = pf.get_particles()
            assert np.array_equal(bits(x), bits(of.x)) and np.array_equal(bits(y), bits(of.y)), "frame %d" % f
            assert np.array_equal(bits(th), bits(of.th)) and np.array_equal(bits(w), bits(of.w)), "frame %d" % f
        assert np.array_equal(pf.get_grid().reshape(-1), of.grid)
        # teacher-forced: map after 1500 frames of the 4096-particle oracle run would be better still, but the
        # dense pseudo-random grid makes every cell disagreement visible
        grid = helpers.synth_grid(salt=43)
        pf.set_grid(grid)
        o = helpers.load_oracle()
        for salt, spread, spread_th in ((3, 0.02, 0.01), (4, 0.6, 0.35), (5, 6.0, 3.0)):
            x, y, th = helpers.synth_particles(n, salt=salt, spread=spread, spread_th=spread_th, center=(2.0, -3.0, 0.4))
            pf.set_particles(x, y, th, np.ones(n, np.float32))
            sc = np.ascontiguousarray(scans[700 + salt])
            got = pf.score_particles(sc)
            want = np.zeros(n, np.int32)
            of.x[:], of.y[:], of.th[:] = x, y, th
            of.grid[:] = grid
            import threading
            b = np.linspace(0, n, 33).astype(int)
            def work(k):
                a, e = int(b[k]), int(b[k + 1])
                o.pfo_score2d_many(C.byref(of.cfg), P(grid, helpers.bp), P(x[a:e]), P(y[a:e]), P(th[a:e]), e - a, P(sc), P(want[a:e], helpers.ip))
            ts = [threading.Thread(target=work, args=(k,)) for k in range(32)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            assert np.array_equal(got, want), "spread %g: %d of %d scores differ from the oracle" % (spread, (got != want).sum(), n)
    of.close()


def test_kd_free_running_400_frames_train_lidar3():
    """BASELINE configs[2] data: the kd step free-running on train_lidar3 (scans 2600..3000, robot driving) for 400 frames (four rebalances,
    a tree of tens of thousands of nodes), engine vs oracle: pose / extrema / Neff / tree size every frame,
    the whole tree node for node every 100 frames"""
    g = _gpu()
    scans = _dataset("train_lidar3")[2600:]    # the robot starts driving around scan 2500
    n, frames = 256, 400
    of = helpers.OracleKdFilter(n)
    with g.ParticleFilter(n, path=g.PATH_KD) as pf:
        for f in range(1, frames + 1):
            r = pf.step(scans[f], f)
            s = of.step(scans[f], f)
            assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose differs at frame %d" % f
            assert r.kd_size == s.kd_size, "kd size differs at frame %d" % f
            if f > 1:
                assert (r.fit_min, r.fit_max, r.best_index, r.resampled) == (s.fit_min, s.fit_max, s.best, s.resampled), "frame %d" % f
                assert np.array_equal(bits([r.neff]), bits([s.neff]))
            if f % 100 == 0 or f in (5, 6):
                assert np.array_equal(pf.get_kd(), of.tree), "tree differs at frame %d" % f
        assert r.kd_size > 3000
    of.close()
