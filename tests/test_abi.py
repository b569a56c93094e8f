"""The C-ABI library loads without a GPU and exports every symbol include/pfslam.h declares."""
import ctypes as C
import os
import re

import pytest

import helpers


def header_functions():
    src = open(os.path.join(helpers.ROOT, "include", "pfslam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfslam_\w+|particleFilterStep)\s*\(", src)))


def test_header_declares_the_path_entry_points():
    names = header_functions()
    for must in ("pfslam_create", "pfslam_destroy", "pfslam_step", "particleFilterStep", "pfslam_update_grid",
                 "pfslam_score_particles", "pfslam_get_grid", "pfslam_get_particles", "pfslam_device_buffer"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from gpu_icp_slam_b200 import engine
    lib = C.CDLL(engine.lib_path())
    for name in header_functions():
        assert hasattr(lib, name), "libpfslam.so does not export %s" % name
    assert set(engine.exported_symbols()) == set(header_functions())


def test_default_config_and_arg_errors_without_gpu():
    from gpu_icp_slam_b200 import engine
    lib = engine.load_library()
    cfg = engine.Config()
    lib.pfslam_default_config(C.byref(cfg))
    assert (cfg.n_particles, cfg.n_beams, cfg.abi_version) == (1000, 1081, 1)
    assert abs(cfg.map_res_x - 0.025) < 1e-9 and cfg.map_scale_x == 40.0
    h = C.c_void_p()
    cfg.abi_version = 99
    assert lib.pfslam_create(C.byref(cfg), C.byref(h)) == 1          # PFSLAM_ERR_ARG, no CUDA call made
    assert b"abi_version" in lib.pfslam_last_error()
    assert lib.pfslam_destroy(None) == 0                              # Free before Init is harmless


def test_unsupported_geometries_are_rejected_without_gpu():
    """non-square maps (grid indexed x*w+y everywhere, like the reference) and filters beyond the prefix
    kernel's tile limit are refused at create time, before any CUDA call"""
    from gpu_icp_slam_b200 import engine
    lib = engine.load_library()
    h = C.c_void_p()
    cfg = engine.Config()
    lib.pfslam_default_config(C.byref(cfg))
    cfg.map_scale_x, cfg.map_scale_y = 40.0, 20.0
    assert lib.pfslam_create(C.byref(cfg), C.byref(h)) == 4          # PFSLAM_ERR_UNSUPPORTED
    assert b"non-square" in lib.pfslam_last_error()
    lib.pfslam_default_config(C.byref(cfg))
    cfg.n_particles = 1024
    cfg.n_particles_global = 4096 * 1024 + 1024
    cfg.n_ranks = 4097
    assert lib.pfslam_create(C.byref(cfg), C.byref(h)) == 4
    assert b"at most" in lib.pfslam_last_error()


def test_missing_library_fails_loudly(tmp_path):
    from gpu_icp_slam_b200 import engine
    with pytest.raises(engine.PfslamError):
        engine.load_library(str(tmp_path / "nope.so"))


def test_product_does_not_touch_the_oracle():
    """nothing under gpu-icp-slam_b200/ or include/ may import, link, open or call anything of oracle/:
    the word itself must not occur in any product source (so no path can be assembled from pieces
    like os.path.join("oracle", ...)), nor the checker's symbol prefix or library names"""
    banned = ("oracle", "pfo_", "libref", "_ref/", '"_ref"', "'_ref'")
    for top in ("gpu-icp-slam_b200", "include"):
        for dp, _, files in os.walk(os.path.join(helpers.ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c", ".txt", ".cmake")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    for b in banned:
                        assert b not in txt, "%s mentions %r" % (os.path.join(dp, f), b)
    # and the shipped libraries do not link them
    import subprocess
    for so in ("libpfslam.so", "libpfslam_kernelh.so"):
        p = os.path.join(helpers.ROOT, "gpu-icp-slam_b200", so)
        if os.path.exists(p):
            needed = subprocess.run(["readelf", "-d", p], capture_output=True, text=True).stdout
            assert "oracle" not in needed and "libref" not in needed, so


def test_engine_fails_loudly_without_a_gpu():
    """no CPU fallback: creating an engine on a box without a CUDA device is an error, not a slow path"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import gpu_icp_slam_b200 as g
    with pytest.raises(g.PfslamError, match="cuda"):
        g.ParticleFilter(128)


def test_exchange_api_rejects_bad_arguments_without_gpu():
    from gpu_icp_slam_b200 import engine
    lib = engine.load_library()
    buf = C.create_string_buffer(engine.IPC_HANDLE_BYTES)
    assert lib.pfslam_ipc_export(None, buf) == 1                      # PFSLAM_ERR_ARG
    assert lib.pfslam_ipc_connect(None, 0, buf) == 1
    assert lib.pfslam_connect_peer(None, 0, None) == 1
    assert lib.pfslam_exchange_ready(None) == 1
    assert lib.pfslam_lap_name(2) == b"k_score_tiled" and lib.pfslam_lap_name(99) == b""
