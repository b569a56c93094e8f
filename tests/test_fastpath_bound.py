"""CPU model of the tiled scorer's fixed-point fast path (csrc/pf_score_tiled.cuh): the float32 operation
sequence of k_tile_prep's constants and k_score_tiled's two FFMA2 is replayed in numpy and compared with
real arithmetic.  DESIGN.md 5.1 budgets the fast side's error at 3 + 6 + 6 + 5 = 20 units of 2^-16 cell
(constants, cos/sin of theta, fixed-point roundings, pose term) out of the 38 that the +-64-unit guard
band must cover; this pins that share on random poses and real scans."""
import numpy as np

import helpers

RES = np.float32(0.025)
UNIT = np.float32(65536.0)
MAGIC = np.float32(8388608.0)
GUARD = np.float32(64.0)


def fma32(a, b, c):
    # float32 fused multiply-add: the product of two float32 is exact in float64
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def test_fast_path_error_share_is_within_budget():
    scans = helpers.fixture_scans()
    rng = np.random.default_rng(7)
    c0 = np.float32(np.float32(0.5) * np.float32(40.0)) / RES                     # 800, exact
    irx = np.float32(1.0 / np.float64(RES))
    mconst = np.float32(MAGIC + np.float32(0.5) * UNIT + GUARD)
    ang = ((np.float32(-135.0) + np.float32(0.25) * np.arange(1081, dtype=np.float32)) * np.float32(np.pi) / np.float32(180.0)).astype(np.float32)
    worst = 0.0
    for f in (5, 60, 140, 230):
        sc = scans[f]
        ok = sc < 19.9
        r, a = sc[ok].astype(np.float64), ang[ok].astype(np.float64)
        # k_tile_prep: per-beam constants in double, rounded once to float32
        rx = r / np.float64(RES)
        ax, bx = (rx * np.cos(a) * 65536.0).astype(np.float32), (rx * np.sin(a) * 65536.0).astype(np.float32)
        ay, by = ax.copy(), bx.copy()                                            # y axis: r sin(a + th) = B cos th + A sin th
        for _ in range(40):
            px, py = np.float32(rng.uniform(-12, 12)), np.float32(rng.uniform(-12, 12))
            th = np.float32(rng.uniform(-3.1, 3.1))
            cs, sn = np.float32(np.cos(np.float64(th))), np.float32(np.sin(np.float64(th)))
            # window origin per beam: an integer that puts the hit somewhere inside a 128-cell window
            vx_true = np.float64(c0) + (np.float64(px) + r * np.cos(a + np.float64(th))) / np.float64(RES)
            vy_true = np.float64(c0) + (np.float64(py) + r * np.sin(a + np.float64(th))) / np.float64(RES)
            x0 = (np.floor(vx_true) - rng.integers(1, 126, vx_true.size)).astype(np.float32)
            y0 = (np.floor(vy_true) - rng.integers(1, 126, vy_true.size)).astype(np.float32)
            offx, offy = (c0 - x0).astype(np.float32), (c0 - y0).astype(np.float32)
            Px = fma32(fma32(px, irx, offx), UNIT, mconst)
            Py = fma32(fma32(py, irx, offy), UNIT, mconst)
            tx = fma32(ax, cs, fma32(-bx, sn, Px))                                # k_score_tiled's two FFMA2 (x lanes)
            ty = fma32(by, cs, fma32(ay, sn, Py))                                 # (y lanes)
            ex = (tx.astype(np.float64) - 8388608.0) - (65536.0 * (vx_true - x0.astype(np.float64) + 0.5) + 64.0)
            ey = (ty.astype(np.float64) - 8388608.0) - (65536.0 * (vy_true - y0.astype(np.float64) + 0.5) + 64.0)
            worst = max(worst, float(np.abs(ex).max()), float(np.abs(ey).max()))
    # budget 20 units for the fast side (worst-case analysis; observed on this sample: 4.4 units)
    assert worst <= 20.0, worst
    assert worst > 0.5            # the model does exercise rounding
