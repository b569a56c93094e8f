"""world_size-2 (and 1) runs of the multi-GPU orchestration (gpu-icp-slam_b200/dist.py) over gloo on
CPU, with oracle-backed shard engines: the sharded filter must reproduce the single-rank oracle
trajectory bit for bit (global-index seeding, fixed tile order, first arg-max across ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers

FRAMES = 12
N_TOTAL = 2048


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpu_icp_slam_b200.dist import ShardedParticleFilter
    from fake_engine import OracleShardEngine
    scans = helpers.fixture_scans()
    pf = ShardedParticleFilter(N_TOTAL // world, engine_factory=OracleShardEngine)
    assert pf.exchange == "collective"            # a stand-in engine cannot use the peer-memory transport
    try:
        ShardedParticleFilter(N_TOTAL // world, engine_factory=OracleShardEngine, exchange="peer")
        raise AssertionError("peer exchange accepted without the CUDA engine")
    except Exception as ex:
        assert "CUDA engine" in str(ex), ex
    poses = []
    for f in range(1, FRAMES + 1):
        r = pf.step(scans[f], f)
        poses.append(r["pose"] + [r["fit_min"], r["fit_max"], r["best_index"], r["neff"], r["resampled"]])
    np.save(os.path.join(out_dir, "traj_%d_%d.npy" % (world, rank)), np.array(poses, np.float64))
    np.save(os.path.join(out_dir, "grid_%d_%d.npy" % (world, rank)), pf.engine.grid)
    np.save(os.path.join(out_dir, "pose_%d_%d.npy" % (world, rank)), pf.engine.pose_local)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_filter_matches_single_rank_oracle(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    scans = helpers.fixture_scans()
    of = helpers.OracleFilter(N_TOTAL)
    want = []
    for f in range(1, FRAMES + 1):
        s = of.step(scans[f], f)
        want.append([s.robot[0], s.robot[1], s.robot[2], s.fit_min, s.fit_max, s.best, s.neff, s.resampled])
    want = np.array(want, np.float64)
    assert want[:, 7].sum() > 0, "no resample in the test window"
    for r in range(world):
        got = np.load(tmp_path / ("traj_%d_%d.npy" % (world, r)))
        assert np.array_equal(got, want), "rank %d trajectory differs" % r
        assert np.array_equal(np.load(tmp_path / ("grid_%d_%d.npy" % (world, r))), of.grid)
    n = N_TOTAL // world
    pose = np.concatenate([np.load(tmp_path / ("pose_%d_%d.npy" % (world, r))).reshape(3, n) for r in range(world)], axis=1)
    assert np.array_equal(pose[0].view(np.uint32), of.x.view(np.uint32))
    assert np.array_equal(pose[2].view(np.uint32), of.th.view(np.uint32))
    of.close()
