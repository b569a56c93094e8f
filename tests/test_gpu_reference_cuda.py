"""T3: the engine against the REFERENCE'S OWN CUDA KERNELS, run on the same GPU in the same process.

oracle/_ref/libref_t3.so is the reference's unmodified src/kernel.cu (compiled for sm_100 by
oracle/Makefile) plus test-only accessors to its file-static device state (oracle/ref_t3_append.cu).
PARTICLE_COUNT is the reference's #define (1000).  Integer results must be identical; the
float-tolerance cases are the ones DESIGN.md section 2 lists (normal variates: CUDA erfcinvf vs the
engine's IEEE-only erfcinv; free-running: thrust scan order + the reference's racy resample)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from helpers import P, fp

pytestmark = pytest.mark.gpu

T3_SO = os.path.join(helpers.ORACLE_DIR, "_ref", "libref_t3.so")
SCENE = os.path.join(helpers.ORACLE_DIR, "_ref", "map_settings.txt")


@pytest.fixture(scope="module")
def t3():
    if not (os.path.exists(T3_SO) and os.path.exists(SCENE)):
        pytest.skip("oracle/_ref/libref_t3.so not built")
    lib = C.CDLL(T3_SO)
    lib.t3_init.argtypes = [C.c_char_p]
    lib.t3_set_particles.argtypes = [fp, fp, fp, fp]
    lib.t3_get_particles.argtypes = [fp, fp, fp, fp]
    lib.t3_set_grid.argtypes = [helpers.bp]
    lib.t3_get_grid.argtypes = [helpers.bp]
    lib.t3_set_robot.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.t3_get_fit.argtypes = [helpers.ip]
    lib.t3_measure.argtypes = [fp, fp]
    lib.t3_update_map.argtypes = [fp]
    assert lib.t3_init(SCENE.encode()) == 0
    assert lib.t3_particle_count() == 1000
    return lib


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


N = 1000


@pytest.mark.parametrize("case", ["near_origin", "spread", "edge_of_map"])
def test_scores_and_weights_equal_reference_kernels(t3, scans, case):
    """kernEvaluateParticles + minmax_element + kernUpdateWeights on the B200 == the engine, exactly,
    for all three scoring kernels of the engine"""
    import gpu_icp_slam_b200 as g
    grid = helpers.synth_grid(salt=21)
    kw = dict(near_origin=dict(salt=1, spread=0.05, spread_th=0.02), spread=dict(salt=2, spread=8.0, spread_th=3.0),
              edge_of_map=dict(salt=3, spread=1.0, spread_th=3.0, center=(19.5, -19.5, 0.0)))[case]
    x, y, th = helpers.synth_particles(N, **kw)
    ones = np.ones(N, np.float32)
    for f in (1, 77, 200):
        sc = np.ascontiguousarray(scans[f])
        t3.t3_set_grid(P(grid, helpers.bp))
        t3.t3_set_particles(P(x), P(y), P(th), P(ones))
        pose = np.zeros(3, np.float32)
        t3.t3_measure(P(sc), P(pose))
        fit_ref = np.zeros(N, np.int32)
        t3.t3_get_fit(P(fit_ref, helpers.ip))
        xr, yr, tr, wr = (np.zeros(N, np.float32) for _ in range(4))
        t3.t3_get_particles(P(xr), P(yr), P(tr), P(wr))
        for mode in (g.SCORE_EXACT, g.SCORE_FILTERED, g.SCORE_TILED):
            with g.ParticleFilter(N, score_mode=mode) as pf:
                pf.set_grid(grid)
                pf.set_particles(x, y, th, ones)
                fit = pf.score_particles(sc)
                assert np.array_equal(fit, fit_ref), "mode %d frame %d: %d scores differ" % (mode, f, (fit != fit_ref).sum())
                pf.phase_weights()
                pf.phase_map()
                r = pf.fetch_result()
                assert np.array_equal(bits(list(r.pose)), bits(pose)), "best pose"
                _, _, _, w = pf.get_particles()
                # reference: all device weights updated; engine keeps the persistent half (Q1)
                assert np.array_equal(bits(w[: N // 2]), bits(wr[: N // 2]))


def test_map_update_equals_reference_kernels(t3, scans):
    """kernGetWalls + traceRay + kernUpdateMap x2 on the B200 == k_map_free / k_map_wall, every grid byte"""
    import gpu_icp_slam_b200 as g
    grid = helpers.synth_grid(salt=31)
    poses = [(0.0, 0.0, 0.0), (1.234, -2.5, 0.7), (-19.9, 19.9, 2.0), (19.99, 0.0, -3.0), (3.3, 3.3, 3.1)]
    t3.t3_set_grid(P(grid, helpers.bp))
    with g.ParticleFilter(64) as pf:
        pf.set_grid(grid)
        for k, pose in enumerate(poses):
            sc = np.ascontiguousarray(scans[5 + 41 * k])
            t3.t3_set_robot(*[C.c_float(v) for v in pose])
            t3.t3_update_map(P(sc))
            pf.update_grid(sc, pose)
            gr = np.zeros(1600 * 1600, np.int8)
            t3.t3_get_grid(P(gr, helpers.bp))
            assert np.array_equal(pf.get_grid().reshape(-1), gr), "pose %r" % (pose,)


def test_noise_within_tolerance_of_reference_kernel(t3):
    """kernAddNoise on the B200 (CUDA erfcinvf) vs k_motion (IEEE-only erfcinv): same seeds, same
    variate order (x, y, theta); per-variate tolerance 2e-6 relative + 1e-9 (3e-4 beyond 5 sigma)"""
    import gpu_icp_slam_b200 as g
    zeros, ones = np.zeros(N, np.float32), np.ones(N, np.float32)
    with g.ParticleFilter(N) as pf:
        for frame in (1, 2, 999, 12000):
            t3.t3_set_particles(P(zeros), P(zeros), P(zeros), P(ones))
            t3.t3_motion(frame)
            xr, yr, tr, wr = (np.zeros(N, np.float32) for _ in range(4))
            t3.t3_get_particles(P(xr), P(yr), P(tr), P(wr))
            pf.set_particles(zeros, zeros, zeros, ones)
            pf.phase_motion(frame)
            x, y, th, _ = pf.get_particles()
            for got, ref_v, sig in ((x, xr, 0.015), (y, yr, 0.015), (th, tr, 0.01)):
                tol = np.where(np.abs(ref_v) < 5 * sig, 2e-6, 3e-4) * np.abs(ref_v) + 1e-9
                assert (np.abs(got - ref_v) <= tol).all(), "frame %d: max rel %g" % (frame, (np.abs(got - ref_v) / np.abs(ref_v)).max())
            assert abs(float(np.std(xr)) - 0.015) < 0.002 and abs(float(np.std(tr)) - 0.01) < 0.0015


def _free_run_against_reference(t3, scans):
    """one free run of both systems from the initial state; returns (max pose distance, explored IoU,
    occupied-cell agreement both ways, occupied cell counts)"""
    import gpu_icp_slam_b200 as g
    from scipy.ndimage import binary_dilation
    assert t3.t3_init(SCENE.encode()) == 0        # fresh reference state: grid -100, particles at 0
    d = []
    with g.ParticleFilter(N) as pf:
        for f in range(1, 81):
            sc = np.ascontiguousarray(scans[f])
            t3.t3_motion(f)
            pose = np.zeros(3, np.float32)
            t3.t3_measure(P(sc), P(pose))
            t3.t3_set_robot(C.c_float(pose[0]), C.c_float(pose[1]), C.c_float(pose[2]))
            t3.t3_update_map(P(sc))
            t3.t3_resample(f)
            r = pf.step(sc, f)
            d.append(np.hypot(r.pose[0] - pose[0], r.pose[1] - pose[1]))
        gr = np.zeros(1600 * 1600, np.int8)
        t3.t3_get_grid(P(gr, helpers.bp))
        gm = pf.get_grid().reshape(-1)
    a, b = gr != -100, gm != -100
    iou = (a & b).sum() / float((a | b).sum())
    # walls are one cell thick and the two trajectories differ by a few cells: compare the occupied
    # cells with a 3-cell tolerance
    occ_a, occ_b = (gr > 0).reshape(1600, 1600), (gm > 0).reshape(1600, 1600)
    k = np.ones((7, 7), bool)
    near_ab = (occ_a & binary_dilation(occ_b, k)).sum() / float(max(occ_a.sum(), 1))
    near_ba = (occ_b & binary_dilation(occ_a, k)).sum() / float(max(occ_b.sum(), 1))
    return max(d), iou, near_ab, near_ba, int(occ_a.sum()), int(occ_b.sum())


def test_free_running_against_reference_cuda_path(t3, scans):
    """Both systems run the 2D step free from the initial state (README.md:41-50 order).  Not bit
    comparable (see module docstring); the trajectories and maps must agree closely.  The reference's
    in-place resample is racy (kernel.cu:441), so its run is not repeatable and an unlucky interleaving can
    send its trajectory elsewhere: up to three attempts, one must agree."""
    seen = []
    for _ in range(3):
        dmax, iou, near_ab, near_ba, na, nb = _free_run_against_reference(t3, scans)
        seen.append((round(float(dmax), 4), round(float(iou), 4), round(float(near_ab), 3), round(float(near_ba), 3), na, nb))
        if dmax < 0.08 and na > 200 and nb > 200 and iou > 0.97 and near_ab > 0.9 and near_ba > 0.9:
            return
    raise AssertionError("no attempt agreed with the reference CUDA path: %s" % (seen,))


# ---- T3-kd: the kd-tree point-cloud path against the reference's own kd kernels ------------------------------
# The reference reads device memory it never wrote.  oracle/ref_t3_alloc.cu defines that
# memory through a link-time wrap of cudaMalloc (the reference's source is untouched): a "stop" record in
# front of dev_kd (Q9), and a chosen fill byte for fresh blocks -- 0xFF (NaN points, which update nothing)
# for the tail of dev_free (Q10), 0x00 for the unwritten ICP targets (Q11).  Those are exactly the
# definitions oracle/pfo_kd.cpp states, so every comparison below is on defined inputs.
ICP_BUF_BYTES = 1081 * 16


@pytest.fixture(scope="module")
def t3kd(t3):
    for name in ("t3_kd_set", "t3_kd_get", "t3_kd_nn", "t3_measure_kd", "t3_icp", "t3_update_map_kd", "t3_particle_filter"):
        if not hasattr(t3, name):
            pytest.skip("libref_t3.so predates the kd accessors")
    t3.t3_kd_set.argtypes = [C.c_void_p, C.c_int]
    t3.t3_kd_get.argtypes = [C.c_void_p]
    t3.t3_kd_nn.argtypes = [fp, C.c_int, helpers.ip]
    t3.t3_get_fitf.argtypes = [fp]
    t3.t3_measure_kd.argtypes = [fp, fp]
    t3.t3_icp.argtypes = [fp, fp, fp]
    t3.t3_update_map_kd.argtypes = [fp]
    t3.t3_particle_filter.argtypes = [fp, C.c_int]
    t3.t3_set_alloc_fill.argtypes = [C.c_int]
    t3.t3_set_alloc_fill_sized.argtypes = [C.c_int, C.c_size_t, C.c_int]
    t3.t3_get_robot.argtypes = [fp]
    return t3


def _grown_tree(scans, n_frames):
    """(tree int32[n, 8], robot pose) of the oracle's kd filter after n_frames of the fixture"""
    of = helpers.OracleKdFilter(64)
    for f in range(1, n_frames + 1):
        of.step(scans[f], f)
    tree = of.tree.copy()
    robot = np.array(list(of.s.contents.robot), np.float32)
    of.close()
    return tree, robot


def _ref_tree(t3):
    n = t3.t3_kd_size()
    out = np.zeros((max(n, 1), 8), np.int32)
    t3.t3_kd_get(out.ctypes.data)
    return out[:n]


@pytest.mark.parametrize("n_frames", [4, 60, 104, 106])
def test_kd_nn_equals_reference_kernel(t3kd, scans, n_frames):
    """findCorrespondenceIndexKD on the B200 == pfslam_kd_nn == the oracle walk, index for index, on trees
    grown by first-scan build + inserts (4, 60, 104 frames: deep insert chains) and just after the frame-105
    rebalance; queries near the walls, exactly on nodes, on the root, and far away"""
    import gpu_icp_slam_b200 as g
    tree, _ = _grown_tree(scans, n_frames)
    o = helpers.load_oracle_kd()
    rng = np.random.default_rng(n_frames)
    xy = tree[:, 4:6].copy().view(np.float32)
    q = np.zeros((6000, 4), np.float32)
    q[:2500, :2] = xy[rng.integers(0, len(tree), 2500)] + rng.normal(0, 0.03, (2500, 2)).astype(np.float32)
    q[2500:4500, :2] = rng.uniform(-12, 12, (2000, 2)).astype(np.float32)
    q[4500:5500, :2] = xy[rng.integers(0, len(tree), 1000)]                    # exactly on nodes
    q[5500:, :2] = xy[0] + rng.normal(0, 0.01, (500, 2)).astype(np.float32)    # around the root (Q9)
    t3kd.t3_kd_set(tree.ctypes.data, len(tree))
    ref_idx = np.zeros(len(q), np.int32)
    t3kd.t3_kd_nn(P(q), len(q), P(ref_idx, helpers.ip))
    want = np.array([o.pfo_kd_nn(tree.ctypes.data, float(a), float(b), 0.0) for a, b in q[:, :2]], np.int32)
    assert (want == 0).sum() > 50, "the root-is-best case (Q9) must be exercised"
    assert np.array_equal(ref_idx, want), "%d of %d NN indices differ from the reference kernel" % ((ref_idx != want).sum(), len(q))
    with g.ParticleFilter(32, path=g.PATH_KD) as pf:
        pf.set_kd(tree)
        got = pf.kd_nn(np.ascontiguousarray(q[:, :3]))
    assert np.array_equal(got, ref_idx)


@pytest.mark.parametrize("n_frames,case", [(30, "tight"), (104, "tight"), (104, "spread")])
def test_kd_scores_equal_reference_kernel(t3kd, scans, n_frames, case):
    """kernEvaluateParticlesKD + minmax_element + kernUpdateWeights(float) on the B200: scores identical to
    the engine's k_score_kd (the reference's float sums are exact integers), same best particle, weights
    equal to the reference's expression evaluated in IEEE float32"""
    import gpu_icp_slam_b200 as g
    tree, robot = _grown_tree(scans, n_frames)
    kw = dict(tight=dict(spread=0.08, spread_th=0.04), spread=dict(spread=2.0, spread_th=1.0))[case]
    x, y, th = helpers.synth_particles(N, salt=11, center=tuple(float(v) for v in robot), **kw)
    ones = np.ones(N, np.float32)
    t3kd.t3_set_alloc_fill(0)                                  # Q11: unwritten ICP targets are (0,0,0)
    for f in (n_frames + 1, n_frames + 40):
        sc = np.ascontiguousarray(scans[f])
        t3kd.t3_kd_set(tree.ctypes.data, len(tree))
        t3kd.t3_set_particles(P(x), P(y), P(th), P(ones))
        t3kd.t3_set_robot(*[C.c_float(float(v)) for v in robot])
        pose = np.zeros(3, np.float32)
        t3kd.t3_measure_kd(P(sc), P(pose))
        fitf = np.zeros(N, np.float32)
        t3kd.t3_get_fitf(P(fitf))
        xr, yr, tr, wr = (np.zeros(N, np.float32) for _ in range(4))
        t3kd.t3_get_particles(P(xr), P(yr), P(tr), P(wr))
        with g.ParticleFilter(N, path=g.PATH_KD) as pf:
            pf.set_kd(tree)
            pf.set_particles(x, y, th, ones)
            fit = pf.score_particles(sc)
            assert np.array_equal(fit.astype(np.float32), fitf) and np.array_equal(fit, fitf.astype(np.int64)), \
                "frame %d: %d kd scores differ from kernEvaluateParticlesKD" % (f, (fit != fitf).sum())
            best = int(np.argmax(fit))                          # first maximum == thrust::minmax_element
            mn, mx = int(fit.min()), int(fit.max())
            if mx > mn:
                c = np.float32(1.0) / np.float32(mx - mn)
                w_want = (ones * (fit.astype(np.float32) - np.float32(mn))) * c
                assert np.array_equal(bits(wr), bits(w_want)), "kernUpdateWeights(float)"
            # the returned pose is the ICP-corrected best particle: engine == oracle bitwise, reference within
            # the stated tolerance (thrust::reduce order, svd3.h's approximate Jacobi SVD)
            mine = pf.kd_icp(sc, robot, [x[best], y[best], th[best]])
            assert abs(mine[0] - pose[0]) < 2e-4 and abs(mine[1] - pose[1]) < 2e-4 and abs(mine[2] - pose[2]) < 1e-4, \
                "ICP pose %r vs reference %r" % (mine, pose)
    t3kd.t3_set_alloc_fill(-1)


def test_kd_icp_matches_reference(t3kd, scans):
    """transformPointICP on the B200 vs pfslam_kd_icp (== the oracle bit for bit): tolerance 2e-4 m / 1e-4 rad,
    the sum of thrust::reduce's float association over 1081 terms and svd3.h's 4-sweep Jacobi approximation
    (DESIGN.md 5.3); frames with out-of-range beams included (Q11 defined as zeros on both sides)"""
    import gpu_icp_slam_b200 as g
    o = helpers.load_oracle_kd()
    cfg = helpers.ocfg()
    t3kd.t3_set_alloc_fill(0)
    worst = np.zeros(3)
    for n_frames in (20, 104):
        tree, robot = _grown_tree(scans, n_frames)
        t3kd.t3_kd_set(tree.ctypes.data, len(tree))
        with g.ParticleFilter(32, path=g.PATH_KD) as pf:
            pf.set_kd(tree)
            for k, f in enumerate(range(n_frames + 1, n_frames + 25, 3)):
                sc = np.ascontiguousarray(scans[f])
                prev = robot + np.float32(0.01 * k) * np.array([1, -1, 0.5], np.float32)
                start = prev + np.array([0.02, -0.015, 0.01], np.float32)
                t3kd.t3_set_robot(*[C.c_float(float(v)) for v in prev])
                ref_out = np.zeros(3, np.float32)
                t3kd.t3_icp(P(start), P(sc), P(ref_out))
                mine = pf.kd_icp(sc, prev, start)
                want = np.zeros(3, np.float32)
                o.pfo_kd_icp(C.byref(cfg), tree.ctypes.data, P(prev), P(start), P(sc), P(want))
                assert np.array_equal(bits(mine), bits(want)), "engine ICP != oracle at frame %d" % f
                worst = np.maximum(worst, np.abs(mine.astype(np.float64) - ref_out))
    t3kd.t3_set_alloc_fill(-1)
    assert worst[0] < 2e-4 and worst[1] < 2e-4 and worst[2] < 1e-4, "worst |engine - reference| = %r" % (worst,)


def _oracle_update_map(tree, robot, scan, kd_cap=1 << 18, with_hits=False):
    """pfo_kd_update_map on a given tree and robotPos; returns the tree afterwards (and, with_hits, the number of
    points in range per node in the free and the wall weight pass)"""
    of = helpers.OracleKdFilter(8, kd_cap=kd_cap)
    s = of.s.contents
    if len(tree):
        C.memmove(s.tree, tree.ctypes.data, tree.nbytes)
    s.kd_size = len(tree)
    s.robot[0], s.robot[1], s.robot[2] = float(robot[0]), float(robot[1]), float(robot[2])
    sc = np.ascontiguousarray(scan, np.float32)
    hf, hw = np.zeros(max(len(tree), 1), np.int32), np.zeros(max(len(tree), 1), np.int32)
    of.o.pfo_kd_update_map_hits(of.s, P(sc), P(hf, helpers.ip), P(hw, helpers.ip))
    out = of.tree.copy()
    of.close()
    return (out, hf, hw) if with_hits else out


def _clamp113(v):
    return np.clip(v, -113.0, 113.0)


def test_kd_first_scan_build_equals_reference(t3kd, scans):
    """kdSize == 0: PFUpdateMapKD builds the tree with KDTree::Create from the first scan's wall points
    (kernel.cu:1532-1536); node for node == the engine's first-scan build"""
    import gpu_icp_slam_b200 as g
    assert t3kd.t3_init(SCENE.encode()) == 0
    t3kd.t3_reset_kd()
    t3kd.t3_set_robot(C.c_float(0), C.c_float(0), C.c_float(0))
    sc = np.ascontiguousarray(scans[1])
    t3kd.t3_update_map_kd(P(sc))
    ref_tree = _ref_tree(t3kd)
    assert len(ref_tree) > 300
    with g.ParticleFilter(32, path=g.PATH_KD) as pf:
        pf.update_grid(sc, [0.0, 0.0, 0.0])
        assert np.array_equal(pf.get_kd(), ref_tree)
    assert np.array_equal(_oracle_update_map(np.zeros((0, 8), np.int32), [0, 0, 0], sc), ref_tree)


@pytest.mark.parametrize("n_frames", [3, 50, 104])
def test_kd_map_update_equals_reference(t3kd, scans, n_frames):
    """PFUpdateMapKD on the B200 (kernGetWalls masks, host point lists, findCorrespondenceIndexKD x2,
    kernUpdateMapKD x2, kernTestCorrespondance, sequential KDTree::InsertNode) against the engine's device-side
    map update and the oracle, teacher-forced from the reference's tree over 8 consecutive frames:
      * topology, coordinates and new nodes: identical, node for node;
      * weights of nodes hit by at most one point per pass: identical;
      * weights of nodes hit by several points in a pass: the reference's plain load/store RACES there
        (kernel.cu:1361) -- its value must be one of the race's legal outcomes (between one and all of the
        colliding updates applied), and the engine's must be the defined one: once per node per pass."""
    import gpu_icp_slam_b200 as g
    tree, robot = _grown_tree(scans, n_frames)
    t3kd.t3_set_alloc_fill(0xFF)                               # Q10: the tail of dev_free is NaN points
    t3kd.t3_kd_set(tree.ctypes.data, len(tree))
    n_raced = n_collapsed = 0
    with g.ParticleFilter(32, path=g.PATH_KD) as pf:
        for k, f in enumerate(range(n_frames + 1, n_frames + 9)):
            before = _ref_tree(t3kd)                            # teacher forcing: both start from the reference's tree
            pf.set_kd(before)
            sc = np.ascontiguousarray(scans[f])
            pose = robot + np.float32(k) * np.array([0.013, -0.008, 0.004], np.float32)
            t3kd.t3_set_robot(*[C.c_float(float(v)) for v in pose])
            t3kd.t3_update_map_kd(P(sc))
            ref_tree = _ref_tree(t3kd)
            pf.update_grid(sc, pose)
            mine = pf.get_kd()
            want, hf, hw = _oracle_update_map(before, pose, sc, with_hits=True)
            assert np.array_equal(mine, want), "engine tree != oracle tree at frame %d" % f
            assert len(mine) == len(ref_tree), "frame %d: %d nodes vs reference %d" % (f, len(mine), len(ref_tree))
            assert np.array_equal(mine[:, :7], ref_tree[:, :7]), "frame %d: topology / coordinates differ" % f
            nb = len(before)
            assert np.array_equal(mine[nb:], ref_tree[nb:]), "frame %d: inserted nodes differ" % f
            w0 = before[:, 7].copy().view(np.float32).astype(np.float64)
            wm = mine[:nb, 7].copy().view(np.float32).astype(np.float64)
            wr = ref_tree[:nb, 7].copy().view(np.float32).astype(np.float64)
            single = (hf[:nb] <= 1) & (hw[:nb] <= 1)
            assert np.array_equal(wm[single], wr[single]), "frame %d: %d un-raced nodes differ" % (f, (wm[single] != wr[single]).sum())
            for i in np.flatnonzero(~single):
                legal = {float(_clamp113(_clamp113(w0[i] - a) + 4.0 * b)) for a in (range(1, hf[i] + 1) if hf[i] else [0])
                         for b in (range(1, hw[i] + 1) if hw[i] else [0])}
                assert float(wr[i]) in legal, "frame %d node %d: reference weight %g is not a legal race outcome of %g (%d free, %d wall hits)" % (
                    f, i, wr[i], w0[i], hf[i], hw[i])
                assert wm[i] == float(_clamp113(_clamp113(w0[i] - (1 if hf[i] else 0)) + (4.0 if hw[i] else 0.0)))
                n_raced += 1
                n_collapsed += int(wm[i] == wr[i])
    t3kd.t3_set_alloc_fill(-1)
    print("raced nodes: %d, of which the reference collapsed to the once-per-pass value: %d" % (n_raced, n_collapsed))


def test_kd_free_running_against_reference_driver(t3kd, scans):
    """the reference's particleFilter() itself (kernel.cu:1702-1768: the kd step at HEAD, including the
    frame%100==5 rebalance) against the engine's kd step, both free-running from an empty map at the
    reference's PARTICLE_COUNT.  Not bit comparable: thrust scan / reduce order, the racy in-place resample
    (kernel.cu:441), the racy weight update (kernel.cu:1361) and the approximate SVD make the reference's own
    runs differ from each other.  The yardstick is therefore the reference against itself: three reference
    runs give the run-to-run spread of trajectory and tree size; the engine's run must lie within 1.5x that
    spread (+2 cm) of the nearest reference run."""
    import gpu_icp_slam_b200 as g
    frames = 120
    ref_traj, ref_size = [], []
    for _ in range(3):
        assert t3kd.t3_init(SCENE.encode()) == 0
        t3kd.t3_reset_kd()
        t3kd.t3_set_alloc_fill_sized(0xFF, ICP_BUF_BYTES, 0)
        tr = np.zeros((frames, 3), np.float32)
        for f in range(1, frames + 1):
            sc = np.ascontiguousarray(scans[f])
            t3kd.t3_particle_filter(P(sc), f)
            t3kd.t3_get_robot(P(tr[f - 1]))
        ref_traj.append(tr)
        ref_size.append(t3kd.t3_kd_size())
        t3kd.t3_set_alloc_fill(-1)
    mine = np.zeros((frames, 3), np.float32)
    with g.ParticleFilter(N, path=g.PATH_KD) as pf:
        for f in range(1, frames + 1):
            r = pf.step(np.ascontiguousarray(scans[f]), f)
            mine[f - 1] = list(r.pose)
        my_size = r.kd_size

    def dist(a, b):
        return float(np.hypot(a[:, 0] - b[:, 0], a[:, 1] - b[:, 1]).max())
    spread = max(dist(ref_traj[i], ref_traj[j]) for i in range(3) for j in range(i + 1, 3))
    mine_d = min(dist(mine, t) for t in ref_traj)
    size_lo, size_hi = min(ref_size), max(ref_size)
    msg = "engine-to-nearest-reference %.4f m, reference run-to-run %.4f m; sizes %d vs reference %r" % (mine_d, spread, my_size, ref_size)
    print(msg)
    assert np.isfinite(mine).all() and all(np.isfinite(t).all() for t in ref_traj), msg
    assert mine_d <= 1.5 * spread + 0.02, msg
    assert size_lo - 0.15 * size_hi <= my_size <= size_hi + 0.15 * size_hi, msg
