"""Host-side mirror of the reference's particle-filter interface (src/kernel.h:14-24) over the
C ABI of libpfslam.so (include/pfslam.h).

Two layers:
  * `ParticleFilter` -- one engine handle; `step(scan, frame)` is `particleFilter(pbo, frame, lidar)`
    (src/kernel.cu:1702) for the 2D occupancy-grid path, in the step order of README.md:41-50.
  * module-level `particleFilterInit / particleFilter / particleFilterFree / getPCData` with the
    reference's names and argument meaning, over one process-wide engine like the reference's
    file-static state (src/kernel.cu:55-83).

No CPU fallback: a missing libpfslam.so raises PfslamError with the build command.
"""
import ctypes as C
import os

import numpy as np

from . import scans as _scans

PATH_GRID2D, PATH_KD = 0, 1
SCORE_EXACT, SCORE_FILTERED, SCORE_TILED = 0, 1, 2
QUIRK_Q1 = 1
QUIRKS_REFERENCE = QUIRK_Q1
IPC_HANDLE_BYTES = 64
RING_DEPTH = 8
LAP_COUNT = 10
BUF_EXTREMA_LOCAL, BUF_EXTREMA_ALL, BUF_TILES_LOCAL, BUF_TILES_ALL, BUF_POSE_LOCAL, BUF_POSE_ALL, BUF_SCAN = range(7)

_HERE = os.path.dirname(os.path.abspath(__file__))


class PfslamError(RuntimeError):
    pass


class Config(C.Structure):
    """pfslam_config (include/pfslam.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("n_particles", C.c_int32), ("n_particles_global", C.c_int32),
        ("particle_offset", C.c_int32), ("n_ranks", C.c_int32), ("n_beams", C.c_int32),
        ("map_scale_x", C.c_float), ("map_scale_y", C.c_float),
        ("map_res_x", C.c_float), ("map_res_y", C.c_float),
        ("device", C.c_int32), ("path", C.c_int32), ("score_mode", C.c_int32), ("quirks", C.c_uint32),
        ("kd_capacity", C.c_int32),
    ]


class FrameResult(C.Structure):
    """pfslam_frame_result (include/pfslam.h)."""
    _fields_ = [
        ("pose", C.c_float * 3), ("fit_min", C.c_int32), ("fit_max", C.c_int32),
        ("best_index", C.c_int32), ("sum_w", C.c_float), ("sum_w2", C.c_float), ("neff", C.c_float),
        ("resampled", C.c_int32), ("n_free_cells", C.c_int32), ("n_wall_cells", C.c_int32),
        ("n_slow_evals", C.c_int32), ("kd_size", C.c_int32), ("kd_inserted", C.c_int32),
        ("exchange_timeout", C.c_int32), ("resample_count", C.c_int32),
        ("wait_extrema_ns", C.c_int32), ("wait_tiles_ns", C.c_int32),
        ("n_windows", C.c_int32), ("n_wide_beams", C.c_int32),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "pose"}
        d["pose"] = [float(v) for v in self.pose]
        return d


def lib_path():
    return os.path.join(_HERE, "libpfslam.so")


_lib = None

_SIGS = {
    "pfslam_default_config": (None, [C.POINTER(Config)]),
    "pfslam_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "pfslam_destroy": (C.c_int, [C.c_void_p]),
    "pfslam_last_error": (C.c_char_p, []),
    "pfslam_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(FrameResult)]),
    "particleFilterStep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float)]),
    "pfslam_step_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pfslam_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pfslam_wait": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(FrameResult)]),
    "pfslam_fetch_result": (C.c_int, [C.c_void_p, C.POINTER(FrameResult)]),
    "pfslam_upload_scan": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_phase_motion": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_phase_score": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_phase_weights": (C.c_int, [C.c_void_p]),
    "pfslam_phase_map": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_phase_resample": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_set_external_params": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_set_params": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pfslam_update_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfslam_score_particles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfslam_get_particles": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4),
    "pfslam_set_particles": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4),
    "pfslam_get_grid": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_set_grid": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_get_map_dim": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pfslam_get_pose": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "pfslam_synchronize": (C.c_int, [C.c_void_p]),
    "pfslam_device_buffer": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pfslam_exchange_region": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pfslam_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pfslam_ipc_connect": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pfslam_connect_peer": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pfslam_exchange_ready": (C.c_int, [C.c_void_p]),
    "pfslam_kd_nn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "pfslam_get_kd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pfslam_set_kd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pfslam_kd_mean_visits": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "pfslam_kd_icp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfslam_launch_count": (C.c_int64, [C.c_void_p]),
    "pfslam_profile_score": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pfslam_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "pfslam_profile_laps": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_profile_laps_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "pfslam_lap_name": (C.c_char_p, [C.c_int32]),
    "pfslam_debug_trig": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pfslam_debug_staged_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "pfslam_debug_trace": (C.c_int, [C.c_int32, C.c_void_p]),
    "pfslam_trace_name": (C.c_char_p, [C.c_int32]),
}


def load_library(path=None):
    """dlopen libpfslam.so and type its entry points.  Raises if the CUDA library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or lib_path()
    if not os.path.exists(p):
        raise PfslamError(
            "CUDA engine library not built: %s is missing. Run `python gpu-icp-slam_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def debug_trig(x, device=0):
    """libdevice (cosf, sinf) of a float32 array, evaluated on the GPU (test hook)."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32)
    c, s = np.empty_like(x), np.empty_like(x)
    rc = lib.pfslam_debug_trig(int(device), x.ctypes.data, x.size, c.ctypes.data, s.ctypes.data)
    if rc != 0:
        raise PfslamError("pfslam error %d: %s" % (rc, lib.pfslam_last_error().decode()))
    return c, s


TRACE_COUNT = 16


def debug_trace(on=True, read=False):
    """in-graph timeline of the 2D step's kernels: {kernel: (first entry, last exit)} in ns of %globaltimer for the
    launches since the last call (None when read is False); resets the table and switches tracing on / off"""
    lib = load_library()
    buf = np.zeros(2 * TRACE_COUNT, np.uint64)
    rc = lib.pfslam_debug_trace(1 if on else 0, buf.ctypes.data if read else None)
    if rc != 0:
        raise PfslamError("pfslam error %d: %s" % (rc, lib.pfslam_last_error().decode()))
    if not read:
        return None
    return {lib.pfslam_trace_name(i).decode(): (int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(TRACE_COUNT)
            if buf[2 * i + 1] > 0}


def exported_symbols():
    return sorted(_SIGS)


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and a.size != n:
        raise PfslamError("expected %d float32 values, got %d" % (n, a.size))
    return a


class Scene:
    """Mirror of the reference's Scene (src/scene.cpp:7-69): parses the MAP block of the text scene
    file (data/map_settings.txt) into maps[0] = {scale: (sx, sy, 0), resolution: (r, r, 1)}."""

    def __init__(self, filename=None, size=(40.0, 40.0), res=0.025):
        self.maps = []
        if filename is None:
            self.maps.append({"scale": (float(size[0]), float(size[1]), 0.0),
                              "resolution": (float(res), float(res), 1.0)})
            return
        with open(filename) as fh:
            lines = [ln.strip() for ln in fh.read().splitlines()]
        i = 0
        while i < len(lines):
            tok = lines[i].split()
            i += 1
            if tok and tok[0] == "MAP":
                sx = sy = r = None
                while i < len(lines) and lines[i]:
                    t = lines[i].split()
                    if t[0] == "SIZE":
                        sx, sy = float(t[1]), float(t[2])
                    elif t[0] == "RES":
                        r = float(t[1])
                    i += 1
                if sx is None or r is None:
                    raise PfslamError("MAP block without SIZE/RES in %s" % filename)
                self.maps.append({"scale": (sx, sy, 0.0), "resolution": (r, r, 1.0)})
        if not self.maps:
            raise PfslamError("no MAP block in %s" % filename)


class Lidar:
    """Mirror of the reference's Lidar (src/lidar.cpp:8-49): `scans[frame]` is the frame's
    float32[1081] range vector.  Reads the packed .scans.u16 format, or a .mat through scipy."""

    def __init__(self, filename=None, scans=None):
        if scans is not None:
            self.scans = np.ascontiguousarray(scans, dtype=np.float32)
        elif filename.endswith(".mat"):
            self.scans = _scans.mat_to_f32(filename)
        else:
            self.scans = _scans.load(filename)


class ParticleFilter:
    """One engine (one GPU's shard of the particle cloud + a replica of the map)."""

    def __init__(self, n_particles=1000, scene=None, n_beams=1081, device=0,
                 score_mode=SCORE_TILED, quirks=QUIRKS_REFERENCE, path=PATH_GRID2D,
                 n_particles_global=None, particle_offset=0, n_ranks=1, kd_capacity=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        scene = scene or Scene()
        m = scene.maps[0]
        cfg = Config()
        self._lib.pfslam_default_config(C.byref(cfg))
        cfg.n_particles = int(n_particles)
        cfg.n_particles_global = int(n_particles_global if n_particles_global is not None else n_particles)
        cfg.particle_offset = int(particle_offset)
        cfg.n_ranks = int(n_ranks)
        cfg.n_beams = int(n_beams)
        cfg.map_scale_x, cfg.map_scale_y = m["scale"][0], m["scale"][1]
        cfg.map_res_x, cfg.map_res_y = m["resolution"][0], m["resolution"][1]
        cfg.device = int(device)
        cfg.path = int(path)
        cfg.score_mode = int(score_mode)
        cfg.quirks = int(quirks)
        cfg.kd_capacity = int(kd_capacity)
        self.cfg = cfg
        self.n = cfg.n_particles
        self.n_beams = cfg.n_beams
        self._check(self._lib.pfslam_create(C.byref(cfg), C.byref(self._h)))
        w, h = C.c_int32(), C.c_int32()
        self._check(self._lib.pfslam_get_map_dim(self._h, C.byref(w), C.byref(h)))
        self.map_w, self.map_h = w.value, h.value

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise PfslamError("pfslam error %d: %s" % (rc, self._lib.pfslam_last_error().decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pfslam_destroy(self._h)
            self._h = C.c_void_p()

    free = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the step -------------------------------------------------------------------------------
    def step(self, scan, frame):
        """particleFilter(pbo, frame, lidar) with lidar->scans[frame] = scan.  Returns FrameResult."""
        s = _f32(scan, self.n_beams)
        out = FrameResult()
        self._check(self._lib.pfslam_step(self._h, s.ctypes.data, int(frame), C.byref(out)))
        return out

    def submit(self, scan, frame):
        """streaming step: enqueue one frame (scan copied into the pinned ring), returns a ticket; up to RING_DEPTH in flight"""
        s = _f32(scan, self.n_beams)
        t = C.c_int32()
        self._check(self._lib.pfslam_submit(self._h, s.ctypes.data, int(frame), C.byref(t)))
        return t.value

    def wait(self, ticket):
        """blocks until the submitted step's result is on the host"""
        out = FrameResult()
        self._check(self._lib.pfslam_wait(self._h, int(ticket), C.byref(out)))
        return out

    def step_async(self, frame, scan_dev_ptr=None):
        self._check(self._lib.pfslam_step_async(self._h, scan_dev_ptr, int(frame)))

    def fetch_result(self):
        out = FrameResult()
        self._check(self._lib.pfslam_fetch_result(self._h, C.byref(out)))
        return out

    def upload_scan(self, scan):
        s = _f32(scan, self.n_beams)
        self._check(self._lib.pfslam_upload_scan(self._h, s.ctypes.data))

    def phase_motion(self, frame):
        self._check(self._lib.pfslam_phase_motion(self._h, int(frame)))

    def phase_score(self, scan_dev_ptr=None):
        self._check(self._lib.pfslam_phase_score(self._h, scan_dev_ptr))

    def phase_weights(self):
        self._check(self._lib.pfslam_phase_weights(self._h))

    def phase_map(self, scan_dev_ptr=None):
        self._check(self._lib.pfslam_phase_map(self._h, scan_dev_ptr))

    def phase_resample(self, frame):
        self._check(self._lib.pfslam_phase_resample(self._h, int(frame)))

    def set_external_params(self, on=True):
        self._check(self._lib.pfslam_set_external_params(self._h, 1 if on else 0))

    def set_params(self, scan_dev_ptr, frame):
        self._check(self._lib.pfslam_set_params(self._h, scan_dev_ptr, int(frame)))

    def synchronize(self):
        self._check(self._lib.pfslam_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.pfslam_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # -- function-level entry points ----------------------------------------------------------------
    def score_particles(self, scan):
        """kernEvaluateParticles alone (src/kernel.cu:277): int32 score per particle."""
        s = _f32(scan, self.n_beams)
        fit = np.empty(self.n, dtype=np.int32)
        self._check(self._lib.pfslam_score_particles(self._h, s.ctypes.data, fit.ctypes.data))
        return fit

    def profile_score(self):
        """(ms of the dominant scoring kernel alone, ms of the whole scoring phase), CUDA events."""
        a, b = C.c_float(), C.c_float()
        self._check(self._lib.pfslam_profile_score(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_enable(self, on=True):
        self._check(self._lib.pfslam_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """(mean ms of the dominant scoring kernel over the steps since enable/read, launches)"""
        ms, n = C.c_float(), C.c_int32()
        self._check(self._lib.pfslam_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_laps(self, on=True):
        self._check(self._lib.pfslam_profile_laps(self._h, 1 if on else 0))

    def profile_laps_read(self):
        """{kernel name: (mean ms, launches)} of the serialised steps since profile_laps(True)"""
        ms, cnt = (C.c_float * LAP_COUNT)(), (C.c_int32 * LAP_COUNT)()
        self._check(self._lib.pfslam_profile_laps_read(self._h, ms, cnt))
        return {self._lib.pfslam_lap_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(LAP_COUNT) if cnt[i]}

    def update_grid(self, scan, pose):
        """PFUpdateMap alone (src/kernel.cu:551) for an explicit robot pose."""
        s = _f32(scan, self.n_beams)
        p = _f32(pose, 3)
        self._check(self._lib.pfslam_update_grid(self._h, s.ctypes.data, p.ctypes.data))

    # -- state ----------------------------------------------------------------------------------
    def get_particles(self):
        x, y, th, w = (np.empty(self.n, dtype=np.float32) for _ in range(4))
        self._check(self._lib.pfslam_get_particles(self._h, x.ctypes.data, y.ctypes.data, th.ctypes.data, w.ctypes.data))
        return x, y, th, w

    def set_particles(self, x=None, y=None, theta=None, w=None):
        arrs = [None if a is None else _f32(a, self.n) for a in (x, y, theta, w)]
        ptrs = [None if a is None else a.ctypes.data for a in arrs]
        self._check(self._lib.pfslam_set_particles(self._h, *ptrs))

    def get_grid(self):
        g = np.empty(self.map_w * self.map_h, dtype=np.int8)
        self._check(self._lib.pfslam_get_grid(self._h, g.ctypes.data))
        return g.reshape(self.map_w, self.map_h)

    def set_grid(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.int8).reshape(-1)
        if g.size != self.map_w * self.map_h:
            raise PfslamError("grid size mismatch")
        self._check(self._lib.pfslam_set_grid(self._h, g.ctypes.data))

    def get_pose(self):
        p = (C.c_float * 3)()
        self._check(self._lib.pfslam_get_pose(self._h, p))
        return [float(v) for v in p]

    # -- kd-tree path ---------------------------------------------------------------------------
    def kd_nn(self, q_xyz):
        """findCorrespondenceIndexKD alone (src/kernel.cu:924): node index of each (x, y, z) query."""
        q = np.ascontiguousarray(q_xyz, dtype=np.float32).reshape(-1, 3)
        out = np.empty(q.shape[0], dtype=np.int32)
        self._check(self._lib.pfslam_kd_nn(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def get_kd(self):
        """the tree as int32[n, 8] words in the reference's KDTree::Node layout"""
        n = C.c_int32()
        self._check(self._lib.pfslam_get_kd(self._h, None, 0, C.byref(n)))
        nodes = np.zeros((max(n.value, 1), 8), dtype=np.int32)
        if n.value:
            self._check(self._lib.pfslam_get_kd(self._h, nodes.ctypes.data, n.value, C.byref(n)))
        return nodes[: n.value]

    def kd_icp(self, scan, robot_prev, start):
        """transformPointICP alone (src/kernel.cu:993): pose = start + one ICP step of `scan` against the
        tree, targets built from robot_prev (the reference's global robotPos)."""
        s = _f32(scan, self.n_beams)
        a, b, out = _f32(robot_prev, 3), _f32(start, 3), np.zeros(3, np.float32)
        self._check(self._lib.pfslam_kd_icp(self._h, s.ctypes.data, a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def kd_mean_visits(self, n_sample=2048):
        """mean tree nodes loaded per NN walk of the scorer (measured on the current scan and cloud)"""
        v = C.c_double()
        self._check(self._lib.pfslam_kd_mean_visits(self._h, int(n_sample), C.byref(v)))
        return v.value

    def set_kd(self, nodes):
        a = np.ascontiguousarray(nodes, dtype=np.int32).reshape(-1, 8)
        self._check(self._lib.pfslam_set_kd(self._h, a.ctypes.data, a.shape[0]))

    # -- shard exchange over peer memory (include/pfslam.h "multi-GPU") ---------------------------
    def exchange_region(self):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._check(self._lib.pfslam_exchange_region(self._h, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def ipc_export(self):
        """the 64-byte CUDA IPC handle of this engine's exchange region"""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.pfslam_ipc_export(self._h, buf))
        return bytes(buf.raw)

    def ipc_connect(self, rank, handle):
        buf = C.create_string_buffer(bytes(handle), IPC_HANDLE_BYTES)
        self._check(self._lib.pfslam_ipc_connect(self._h, int(rank), buf))

    def connect_peer(self, rank, region_ptr):
        self._check(self._lib.pfslam_connect_peer(self._h, int(rank), C.c_void_p(region_ptr)))

    def exchange_ready(self):
        self._check(self._lib.pfslam_exchange_ready(self._h))

    def device_buffer(self, which):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._check(self._lib.pfslam_device_buffer(self._h, int(which), C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    @property
    def launch_count(self):
        return int(self._lib.pfslam_launch_count(self._h))


# ---- the reference's free-function interface over one process-wide engine (src/kernel.h:14-24) ----
_engine = None
_robot_pos = [0.0, 0.0, 0.0]
PARTICLE_COUNT = 1000   # src/kernel.cu:30; a run-time value here


def particleFilterInit(scene, n_particles=None, **kw):
    """void particleFilterInit(Scene *scene) (src/kernel.cu:107)."""
    global _engine, _robot_pos
    if _engine is not None:
        raise PfslamError("particleFilterInit called twice without particleFilterFree (SURVEY Q15)")
    _engine = ParticleFilter(n_particles or PARTICLE_COUNT, scene=scene, **kw)
    _robot_pos = [0.0, 0.0, 0.0]


def particleFilterFree():
    """void particleFilterFree() (src/kernel.cu:163); harmless before Init (src/main.cpp:194)."""
    global _engine
    if _engine is not None:
        _engine.close()
        _engine = None


def particleFilter(pbo, frame, lidar):
    """void particleFilter(uchar4 *pbo, int frame, Lidar *lidar) (src/kernel.cu:1702).
    `pbo` is unused, as at the reference's HEAD."""
    global _robot_pos
    if _engine is None:
        raise PfslamError("particleFilter before particleFilterInit")
    r = _engine.step(lidar.scans[frame], frame)
    _robot_pos = [float(v) for v in r.pose]
    return r


def getPCData():
    """getPCData(...) (src/kernel.cu:803-813): (particles[N,4] = x,y,theta,w; occupancy grid
    int8[map_w,map_h]; kd nodes (None on the 2D path); nParticles; nKD; robotPos)."""
    if _engine is None:
        raise PfslamError("getPCData before particleFilterInit")
    x, y, th, w = _engine.get_particles()
    kd = _engine.get_kd() if _engine.cfg.path == PATH_KD else None
    return (np.stack([x, y, th, w], axis=1), _engine.get_grid(), kd, _engine.n,
            0 if kd is None else len(kd), list(_robot_pos))
