"""Test-side access to the CPU oracle (oracle/) and, when built, the reference's own host functions
(oracle/_ref/libref.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use these."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

TRIG_LIBM, TRIG_CUDA = 0, 1
MAD_SEPARATE, MAD_FUSED = 0, 1

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
bp = C.POINTER(C.c_int8)
ubp = C.POINTER(C.c_uint8)


class OCfg(C.Structure):
    _fields_ = [("n_beams", C.c_int), ("map_w", C.c_int), ("map_h", C.c_int),
                ("scale_x", C.c_float), ("scale_y", C.c_float), ("res_x", C.c_float), ("res_y", C.c_float),
                ("trig", C.c_int), ("mad", C.c_int), ("quirk_q1", C.c_int)]


class OState(C.Structure):
    _fields_ = [("cfg", OCfg), ("n", C.c_int), ("x", fp), ("y", fp), ("th", fp), ("w", fp), ("weff", fp),
                ("fit", ip), ("cdf", fp), ("grid", bp), ("free_mask", ubp), ("wall_mask", ubp),
                ("robot", C.c_float * 3), ("fit_min", C.c_int32), ("fit_max", C.c_int32), ("best", C.c_int),
                ("sum_w", C.c_float), ("sum_w2", C.c_float), ("neff", C.c_float), ("resampled", C.c_int),
                ("n_free", C.c_int), ("n_wall", C.c_int)]


def ocfg(trig=TRIG_CUDA, mad=MAD_FUSED, q1=1, n_beams=1081, scale=40.0, res=0.025):
    r = np.float32(res)
    w = int(np.float32(scale) / r)
    return OCfg(n_beams, w, w, scale, scale, r, r, trig, mad, q1)


def build_oracle():
    src = [os.path.join(ORACLE_DIR, f) for f in ("pfo.c", "pfo_kd.cpp", "pfo.h")]
    if os.path.exists(ORACLE_SO) and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in src):
        return
    subprocess.run(["make", "-C", ORACLE_DIR, "_build/liboracle.so"], check=True, stdout=subprocess.DEVNULL)


_oracle = None


def load_oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    build_oracle()
    o = C.CDLL(ORACLE_SO)
    o.pfo_utilhash.restype = C.c_uint32
    o.pfo_utilhash.argtypes = [C.c_uint32]
    o.pfo_seed.restype = C.c_uint32
    o.pfo_seed.argtypes = [C.c_int, C.c_int, C.c_int]
    o.pfo_minstd_seed.restype = C.c_uint32
    o.pfo_minstd_seed.argtypes = [C.c_uint32]
    o.pfo_minstd_next.restype = C.c_uint32
    o.pfo_minstd_next.argtypes = [C.POINTER(C.c_uint32)]
    for f in ("pfo_cosf_cuda", "pfo_sinf_cuda", "pfo_logf", "pfo_erfcinvf"):
        getattr(o, f).restype = C.c_float
        getattr(o, f).argtypes = [C.c_float]
    o.pfo_normal.restype = C.c_float
    o.pfo_normal.argtypes = [C.POINTER(C.c_uint32), C.c_float]
    o.pfo_lidar_angle.restype = C.c_float
    o.pfo_lidar_angle.argtypes = [C.c_int]
    o.pfo_add_noise.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int]
    o.pfo_score2d.restype = C.c_int
    o.pfo_score2d.argtypes = [C.POINTER(OCfg), bp, C.c_float, C.c_float, C.c_float, fp]
    o.pfo_score2d_many.argtypes = [C.POINTER(OCfg), bp, fp, fp, fp, C.c_int, fp, ip]
    o.pfo_minmax.argtypes = [ip, C.c_int, ip, ip, C.POINTER(C.c_int)]
    o.pfo_center_cell.argtypes = [C.POINTER(OCfg), C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    o.pfo_trace_ray.restype = C.c_int
    o.pfo_trace_ray.argtypes = [C.c_int] * 6 + [ubp]
    o.pfo_get_walls.argtypes = [C.POINTER(OCfg), fp, C.c_int, C.c_int, C.c_float, ubp, ubp]
    o.pfo_apply_masks.argtypes = [bp, C.c_int, ubp, ubp]
    o.pfo_scan.restype = C.c_float
    o.pfo_scan.argtypes = [fp, C.c_int, fp]
    o.pfo_scan_tiles.argtypes = [fp, C.c_int, fp, fp]
    o.pfo_resample_src.restype = C.c_int
    o.pfo_resample_src.argtypes = [fp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
    o.pfo_create.restype = C.POINTER(OState)
    o.pfo_create.argtypes = [C.POINTER(OCfg), C.c_int]
    o.pfo_destroy.argtypes = [C.POINTER(OState)]
    for f in ("pfo_motion", "pfo_resample"):
        getattr(o, f).argtypes = [C.POINTER(OState), C.c_int]
    for f in ("pfo_measure", "pfo_update_map"):
        getattr(o, f).argtypes = [C.POINTER(OState), fp]
    o.pfo_step2d.argtypes = [C.POINTER(OState), fp, C.c_int]
    o.pfo_measure_scored.argtypes = [C.POINTER(OState)]
    _oracle = o
    return o


def load_ref():
    if not os.path.exists(REF_SO):
        return None
    r = C.CDLL(REF_SO)
    r.ref_utilhash.restype = C.c_uint32
    r.ref_utilhash.argtypes = [C.c_uint32]
    r.ref_clean_lidar_scan.argtypes = [C.c_int, C.c_float, C.c_float, fp]
    r.ref_trace_ray.argtypes = [C.c_int] * 6 + [ubp]
    r.ref_evaluate_particles.argtypes = [bp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                         fp, fp, fp, C.c_int, fp, ip]
    r.ref_add_noise.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int]
    r.ref_scene_map.argtypes = [C.c_char_p, fp]
    return r


def P(a, t=fp):
    return a.ctypes.data_as(t)


def fixture_scans():
    from gpu_icp_slam_b200 import scans
    return scans.load(os.path.join(GOLDEN, "train_lidar0_first256.scans.u16"))


def full_scans(name="train_lidar0"):
    from gpu_icp_slam_b200 import scans
    p = os.path.join(ROOT, "data", "_cache", name + ".scans.u16")
    return scans.load(p) if os.path.exists(p) else None


def np_utilhash(a):
    """vectorised kernel.cu:89-97"""
    a = np.asarray(a, dtype=np.uint64) & 0xFFFFFFFF
    m = np.uint64(0xFFFFFFFF)
    a = ((a + np.uint64(0x7ed55d16)) + (a << np.uint64(12))) & m
    a = ((a ^ np.uint64(0xc761c23c)) ^ (a >> np.uint64(19))) & m
    a = ((a + np.uint64(0x165667b1)) + (a << np.uint64(5))) & m
    a = ((a + np.uint64(0xd3a2646c)) ^ (a << np.uint64(9))) & m
    a = ((a + np.uint64(0xfd7046c5)) + (a << np.uint64(3))) & m
    a = ((a ^ np.uint64(0xb55a4f09)) ^ (a >> np.uint64(16))) & m
    return a.astype(np.uint32)


def synth_grid(w=1600, h=1600, salt=0):
    """deterministic pseudo-random occupancy grid in [-113, 113] (no RNG library dependence)"""
    i = np.arange(w * h, dtype=np.uint64) + np.uint64(salt * 7919)
    return ((np_utilhash(i) % np.uint32(227)).astype(np.int32) - 113).astype(np.int8)


def synth_particles(n, salt=0, spread=0.5, spread_th=0.3, center=(0.0, 0.0, 0.0)):
    i = np.arange(n, dtype=np.uint64)
    def u(k):
        return (np_utilhash(i * np.uint64(3) + np.uint64(k + 101 * salt)) % np.uint32(200001)).astype(np.float64) / 100000.0 - 1.0
    x = (center[0] + spread * u(0)).astype(np.float32)
    y = (center[1] + spread * u(1)).astype(np.float32)
    th = (center[2] + spread_th * u(2)).astype(np.float32)
    return x, y, th


class OracleFilter:
    """thin OO wrapper over pfo_state for tests"""

    def __init__(self, n, cfg=None):
        self.o = load_oracle()
        self.cfg = cfg or ocfg()
        self.n = n
        self.s = self.o.pfo_create(C.byref(self.cfg), n)
        self.nc = self.cfg.map_w * self.cfg.map_h

    def close(self):
        if self.s:
            self.o.pfo_destroy(self.s)
            self.s = None

    def arr(self, name, count, dtype):
        return np.ctypeslib.as_array(getattr(self.s.contents, name), shape=(count,)).view(dtype)

    @property
    def x(self): return self.arr("x", self.n, np.float32)
    @property
    def y(self): return self.arr("y", self.n, np.float32)
    @property
    def th(self): return self.arr("th", self.n, np.float32)
    @property
    def w(self): return self.arr("w", self.n, np.float32)
    @property
    def weff(self): return self.arr("weff", self.n, np.float32)
    @property
    def fit(self): return self.arr("fit", self.n, np.int32)
    @property
    def grid(self): return self.arr("grid", self.nc, np.int8)

    def step(self, scan, frame):
        sc = np.ascontiguousarray(scan, dtype=np.float32)
        self.o.pfo_step2d(self.s, P(sc), int(frame))
        return self.s.contents

    def step_threaded(self, scan, frame, threads=None):
        """pfo_step2d with the scoring loop fanned over host threads by particle range (ctypes releases the
        GIL); the same functions in the same order, so the result is identical to step()"""
        import threading
        threads = threads or min(32, os.cpu_count() or 1)
        sc = np.ascontiguousarray(scan, dtype=np.float32)
        o, s = self.o, self.s
        o.pfo_motion(s, int(frame))
        b = np.linspace(0, self.n, threads + 1).astype(int)
        x, y, th, fit = self.x, self.y, self.th, self.fit

        def work(k):
            a, e = int(b[k]), int(b[k + 1])
            if e > a:
                o.pfo_score2d_many(C.byref(self.cfg), s.contents.grid, P(x[a:e]), P(y[a:e]), P(th[a:e]), e - a, P(sc),
                                   P(fit[a:e], ip))
        ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        o.pfo_measure_scored(s)
        o.pfo_update_map(s, P(sc))
        o.pfo_resample(s, int(frame))
        return s.contents


# ---- kd path ---------------------------------------------------------------------------------------
class OKdState(C.Structure):
    _fields_ = [("cfg", OCfg), ("n", C.c_int), ("x", fp), ("y", fp), ("th", fp), ("w", fp), ("weff", fp),
                ("fit", ip), ("cdf", fp), ("tree", C.c_void_p), ("kd_size", C.c_int), ("kd_cap", C.c_int),
                ("free_mask", ubp), ("wall_mask", ubp), ("robot", C.c_float * 3),
                ("fit_min", C.c_int32), ("fit_max", C.c_int32), ("best", C.c_int),
                ("sum_w", C.c_float), ("sum_w2", C.c_float), ("neff", C.c_float), ("resampled", C.c_int),
                ("n_free_pts", C.c_int), ("n_wall_pts", C.c_int), ("n_inserted", C.c_int)]


def load_oracle_kd():
    o = load_oracle()
    if getattr(o, "_kd_typed", False):
        return o
    o.pfo_kd_create.argtypes = [fp, C.c_int, C.c_void_p]
    o.pfo_kd_insert.argtypes = [fp, C.c_void_p, C.c_int]
    o.pfo_kd_balance.argtypes = [C.c_void_p, C.c_int]
    o.pfo_kd_nn.restype = C.c_int
    o.pfo_kd_nn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    o.pfo_kd_score.restype = C.c_int
    o.pfo_kd_score.argtypes = [C.POINTER(OCfg), C.c_void_p, C.c_float, C.c_float, C.c_float, fp]
    o.pfo_asinf.restype = C.c_float
    o.pfo_asinf.argtypes = [C.c_float]
    o.pfo_kd_icp.argtypes = [C.POINTER(OCfg), C.c_void_p, fp, fp, fp, fp]
    o.pfo_kd_create_state.restype = C.POINTER(OKdState)
    o.pfo_kd_create_state.argtypes = [C.POINTER(OCfg), C.c_int, C.c_int]
    o.pfo_kd_destroy_state.argtypes = [C.POINTER(OKdState)]
    o.pfo_kd_step.argtypes = [C.POINTER(OKdState), fp, C.c_int]
    o.pfo_kd_update_map.argtypes = [C.POINTER(OKdState), fp]
    o.pfo_kd_update_map_hits.argtypes = [C.POINTER(OKdState), fp, ip, ip]
    o._kd_typed = True
    return o


class OracleKdFilter:
    def __init__(self, n, cfg=None, kd_cap=1 << 20):
        self.o = load_oracle_kd()
        self.cfg = cfg or ocfg()
        self.n = n
        self.s = self.o.pfo_kd_create_state(C.byref(self.cfg), n, kd_cap)

    def close(self):
        if self.s:
            self.o.pfo_kd_destroy_state(self.s)
            self.s = None

    def arr(self, name, count, dtype):
        return np.ctypeslib.as_array(getattr(self.s.contents, name), shape=(count,)).view(dtype)

    x = property(lambda s: s.arr("x", s.n, np.float32))
    y = property(lambda s: s.arr("y", s.n, np.float32))
    th = property(lambda s: s.arr("th", s.n, np.float32))
    w = property(lambda s: s.arr("w", s.n, np.float32))
    fit = property(lambda s: s.arr("fit", s.n, np.int32))

    @property
    def tree(self):
        n = self.s.contents.kd_size
        buf = (C.c_int32 * (8 * max(n, 1))).from_address(self.s.contents.tree)
        return np.ctypeslib.as_array(buf).reshape(-1, 8)[:n]

    def step(self, scan, frame):
        sc = np.ascontiguousarray(scan, dtype=np.float32)
        self.o.pfo_kd_step(self.s, P(sc), int(frame))
        return self.s.contents
