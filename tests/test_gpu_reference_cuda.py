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
