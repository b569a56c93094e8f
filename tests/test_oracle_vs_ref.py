"""Oracle vs the reference's own host functions, live (oracle/_ref/libref.so built from
/root/reference/src by oracle/Makefile).  Skipped where the reference build is absent; the
committed golden vectors (test_oracle_golden.py) carry the same pins everywhere."""
import ctypes as C

import numpy as np

import helpers
from helpers import P, TRIG_LIBM, MAD_SEPARATE


def test_struct_sizes(ref):
    assert ref.ref_sizeof_particle() == 32 and ref.ref_sizeof_kdnode() == 32


def test_hash_random(ref, oracle):
    rng = np.random.default_rng(0)
    for a in rng.integers(0, 2**32, 2000, dtype=np.uint64):
        assert ref.ref_utilhash(int(a)) == oracle.pfo_utilhash(int(a))


def test_evaluate_particle_on_evolved_map(ref, oracle, scans):
    """run the oracle filter for 40 frames to get a realistic map + particle cloud, then compare
    EvaluateParticle on every particle for several scans: exact"""
    cfg = helpers.ocfg(TRIG_LIBM, MAD_SEPARATE)
    of = helpers.OracleFilter(300, cfg)
    for f in range(1, 41):
        of.step(scans[f], f)
    grid = of.grid.copy(); x = of.x.copy(); y = of.y.copy(); th = of.th.copy()
    for f in (41, 100, 255):
        sc = np.ascontiguousarray(scans[f])
        a = np.zeros(300, np.int32); b = np.zeros(300, np.int32)
        oracle.pfo_score2d_many(C.byref(cfg), P(grid, helpers.bp), P(x), P(y), P(th), 300, P(sc), P(a, helpers.ip))
        ref.ref_evaluate_particles(P(grid, helpers.bp), 1600, 1600, 40.0, 40.0, np.float32(0.025), np.float32(0.025),
                                   P(x), P(y), P(th), 300, P(sc), P(b, helpers.ip))
        assert np.array_equal(a, b)
    of.close()


def test_trace_ray_random(ref, oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        sx, sy = rng.integers(-20, 1620, 2)
        ex, ey = sx + rng.integers(-800, 801), sy + rng.integers(-800, 801)
        a = np.zeros(1600 * 1600, np.uint8); b = np.zeros(1600 * 1600, np.uint8)
        oracle.pfo_trace_ray(int(sx), int(sy), int(ex), int(ey), 1600, 1600, P(a, helpers.ubp))
        ref.ref_trace_ray(int(sx), int(sy), int(ex), int(ey), 1600, 1600, P(b, helpers.ubp))
        assert np.array_equal(a, b)
