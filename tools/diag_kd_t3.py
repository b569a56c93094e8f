import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers
from helpers import P, fp
import test_gpu_reference_cuda as T
import gpu_icp_slam_b200 as g
lib = C.CDLL(T.T3_SO)
lib.t3_init.argtypes = [C.c_char_p]
lib.t3_set_robot.argtypes = [C.c_float]*3
lib.t3_kd_set.argtypes = [C.c_void_p, C.c_int]; lib.t3_kd_get.argtypes = [C.c_void_p]
lib.t3_update_map_kd.argtypes = [fp]; lib.t3_set_alloc_fill.argtypes = [C.c_int]
assert lib.t3_init(T.SCENE.encode()) == 0
scans = helpers.fixture_scans()
tree, robot = T._grown_tree(scans, 50)
for fill in (0xFF, 0x00):
    lib.t3_set_alloc_fill(fill)
    lib.t3_kd_set(tree.ctypes.data, len(tree))
    sc = np.ascontiguousarray(scans[51])
    lib.t3_set_robot(*[C.c_float(float(v)) for v in robot])
    lib.t3_update_map_kd(P(sc))
    ref = T._ref_tree(lib)
    orc = T._oracle_update_map(tree, robot, sc)
    with g.ParticleFilter(32, path=g.PATH_KD) as pf:
        pf.set_kd(tree); pf.update_grid(sc, robot); mine = pf.get_kd()
    print("fill %x sizes ref %d oracle %d engine %d (before %d)" % (fill, len(ref), len(orc), len(mine), len(tree)))
    n = min(len(ref), len(orc))
    topo = (ref[:n, :7] != orc[:n, :7]).any(axis=1)
    print(" topology/coords differ on", topo.sum(), "nodes; engine==oracle:", np.array_equal(mine, orc))
    wr, wo, w0 = ref[:n, 7].copy().view(np.float32), orc[:n, 7].copy().view(np.float32), None
    nb = len(tree)
    w0 = tree[:, 7].copy().view(np.float32)
    d = np.flatnonzero(wr != wo)
    print(" weights differ on", len(d), "nodes")
    for i in d[:25]:
        print("   node %d: before %s ref %s oracle %s" % (i, w0[i] if i < nb else None, wr[i], wo[i]))
    # histogram of (oracle delta, ref delta)
    if len(d):
        dr, do = wr[:nb] - w0, wo[:nb] - w0
        import collections
        print(" (oracle delta, ref delta) counts:", collections.Counter(zip(do[d[d < nb]].tolist(), dr[d[d < nb]].tolist())).most_common(12))
    # repeatability of the reference
    lib.t3_kd_set(tree.ctypes.data, len(tree))
    lib.t3_update_map_kd(P(sc))
    ref2 = T._ref_tree(lib)
    print(" reference repeatable:", np.array_equal(ref, ref2))
