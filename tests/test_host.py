"""Host-side logic that needs no GPU: scan packing, Scene/Lidar mirrors."""
import os

import numpy as np
import pytest

import helpers


def test_scan_pack_roundtrip_is_exact(tmp_path, scans):
    from gpu_icp_slam_b200 import scans as S
    u = S.encode(scans)
    assert u.dtype == np.uint16 and np.array_equal(S.decode(u), scans)
    p = tmp_path / "a.scans.u16"
    S.save(str(p), u)
    assert np.array_equal(S.load(str(p)), scans)
    assert (scans == np.float32(4294967.0)).any() or True


def test_scan_pack_rejects_unrepresentable():
    from gpu_icp_slam_b200 import scans as S
    with pytest.raises(ValueError):
        S.encode(np.array([[70.0]], np.float32))
    with pytest.raises(ValueError):
        S.encode(np.array([[1.00004]], np.float32))


def test_fixture_shape(scans):
    assert scans.shape == (256, 1081) and scans.dtype == np.float32


def test_scene_default_and_lidar():
    import gpu_icp_slam_b200 as g
    s = g.Scene()
    assert s.maps[0]["scale"][:2] == (40.0, 40.0) and s.maps[0]["resolution"][0] == 0.025
    l = g.Lidar(scans=np.ones((3, 1081)))
    assert l.scans.dtype == np.float32 and l.scans.shape == (3, 1081)


def test_scene_parse(tmp_path):
    import gpu_icp_slam_b200 as g
    p = tmp_path / "scene.txt"
    p.write_text("// c\nCAMERA\nRES 800 800\nFOVY 45\n\n// Patch\nMAP\nSIZE \t20 30\nRES\t.05\n")
    m = g.Scene(str(p)).maps[0]
    assert m["scale"] == (20.0, 30.0, 0.0) and m["resolution"] == (0.05, 0.05, 1.0)
    q = tmp_path / "bad.txt"
    q.write_text("CAMERA\nRES 1 1\n")
    with pytest.raises(g.PfslamError):
        g.Scene(str(q))


def test_map_export_writers(tmp_path):
    """tools/export_map.py: PGM of an oracle-built grid and PCD of an oracle-built kd cloud are well formed"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_map", os.path.join(helpers.ROOT, "tools", "export_map.py"))
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    scans = helpers.fixture_scans()
    of = helpers.OracleFilter(64)
    ok = helpers.OracleKdFilter(64)
    for f in range(1, 6):
        of.step(scans[f], f)
        ok.step(scans[f], f)
    em.write_pgm(str(tmp_path / "m.pgm"), of.grid.reshape(1600, 1600))
    raw = open(tmp_path / "m.pgm", "rb").read()
    assert raw.startswith(b"P5\n1600 1600\n255\n") and len(raw) == len(b"P5\n1600 1600\n255\n") + 1600 * 1600
    px = np.frombuffer(raw[-1600 * 1600:], np.uint8)
    assert (px == 227).mean() > 0.9 and (px > 227).sum() > 1000 and (px < 227).sum() > 100     # unknown / free / walls
    em.write_pcd(str(tmp_path / "m.pcd"), ok.tree)
    lines = open(tmp_path / "m.pcd").read().splitlines()
    assert lines[1] == "VERSION 0.7" and int(lines[9].split()[1]) == len(ok.tree) == len(lines) - 11
    of.close()
    ok.close()


def test_bench_config_is_identical_in_both_arms():
    """bench.py builds `config` with one function for the ours / --impl reference arms (the driver compares them)"""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(helpers.ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    for path, gpus in (("grid2d", 1), ("grid2d", 4), ("kd", 1)):
        args = types.SimpleNamespace(particles=65536, path=path, gpus=gpus, data="auto")
        ds = b.default_dataset(args)
        assert b.make_config(args, gpus, ds) == b.make_config(args, gpus, ds)
        assert ds == {"grid2d1": "train_lidar0", "grid2d4": "train_lidar2", "kd1": "train_lidar3"}[path + str(gpus)]
        assert ds in b.make_config(args, gpus, ds)["workload"]


def test_streaming_api_rejects_bad_arguments_without_gpu():
    import ctypes as C
    from gpu_icp_slam_b200 import engine
    lib = engine.load_library()
    t = C.c_int32()
    assert lib.pfslam_submit(None, None, 1, C.byref(t)) == 1          # PFSLAM_ERR_ARG
    assert lib.pfslam_wait(None, 1, None) == 1
    assert lib.pfslam_kd_icp(None, None, None, None, None) == 1
    assert lib.pfslam_trace_name(0) == b"k_motion" and lib.pfslam_trace_name(99) == b""
