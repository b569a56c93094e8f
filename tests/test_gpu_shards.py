"""Sharded filter over the PEER-MEMORY exchange (csrc/pf_xchg.cuh), exercised on one GPU: R engines in
one process, each holding N/R particles, wired to each other's exchange regions with
pfslam_connect_peer and stepped concurrently on their own streams.  The kernels publish extrema and
weight tiles and their pose snapshots into each other's regions and spin on the sequence flags
exactly as they do across GPUs (only the IPC handle plumbing differs; that part is covered
by tools/dist_check.py under torchrun on a multi-GPU box).  Every shard must reproduce the
single-engine oracle bit for bit: SURVEY 8(e) "identical trajectories for 1/2/4/8 GPUs"."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_shards(g, n_total, n_ranks, **kw):
    n = n_total // n_ranks
    engines = [g.ParticleFilter(n, n_particles_global=n_total, particle_offset=r * n, n_ranks=n_ranks, **kw)
               for r in range(n_ranks)]
    regions = [e.exchange_region()[0] for e in engines]
    for e in engines:
        for r, ptr in enumerate(regions):
            e.connect_peer(r, ptr)
        e.exchange_ready()
    return engines


def step_all(engines, scan, frame):
    for e in engines:
        e.upload_scan(scan)
    for e in engines:            # every launch is asynchronous: the shards run concurrently
        e.step_async(frame)
    return [e.fetch_result() for e in engines]


@pytest.mark.parametrize("n_ranks,n_total,switches", [
    (2, 4096, {}), (4, 8192, {}), (8, 8192, {}),
    (4, 8192, {"PFSLAM_SNAPSHOT": "pull"}),        # resampler loads the drawn poses from the owner (round-2 mid scheme)
    (4, 8192, {"PFSLAM_TAIL": "fused"}),           # weights + prefix + resample as one kernel with a grid-wide barrier
])
def test_grid_shards_match_single_oracle(scans, n_ranks, n_total, switches, monkeypatch):
    import gpu_icp_slam_b200 as g
    for k, v in switches.items():                  # read when an engine is created
        monkeypatch.setenv(k, v)
    frames = 30
    of = helpers.OracleFilter(n_total)
    engines = make_shards(g, n_total, n_ranks)
    n = n_total // n_ranks
    resamples = 0
    try:
        for f in range(1, frames + 1):
            res = step_all(engines, scans[f], f)
            s = of.step(scans[f], f)
            resamples += s.resampled
            for k, r in enumerate(res):
                assert r.exchange_timeout == 0
                assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose, shard %d frame %d" % (k, f)
                assert (r.fit_min, r.fit_max, r.best_index) == (s.fit_min, s.fit_max, s.best), "extrema, shard %d frame %d" % (k, f)
                assert np.array_equal(bits([r.neff, r.sum_w]), bits([s.neff, s.sum_w])) and r.resampled == s.resampled
                assert (r.n_free_cells, r.n_wall_cells) == (s.n_free, s.n_wall)
        assert resamples >= 5, "the window must exercise the cross-shard resample"
        for k, e in enumerate(engines):
            assert np.array_equal(e.get_grid().reshape(-1), of.grid), "grid of shard %d" % k
            x, y, th, w = e.get_particles()
            sl = slice(k * n, (k + 1) * n)
            assert np.array_equal(bits(x), bits(of.x[sl])) and np.array_equal(bits(y), bits(of.y[sl]))
            assert np.array_equal(bits(th), bits(of.th[sl])) and np.array_equal(bits(w), bits(of.w[sl]))
    finally:
        for e in engines:
            e.close()
        of.close()


def test_kd_shards_match_single_oracle(scans):
    """kd-tree point-cloud path sharded 2-way: every shard keeps an identical replica of the tree"""
    import gpu_icp_slam_b200 as g
    n_total, n_ranks, frames = 2048, 2, 108          # includes the frame-105 rebalance
    of = helpers.OracleKdFilter(n_total)
    engines = make_shards(g, n_total, n_ranks, path=g.PATH_KD)
    n = n_total // n_ranks
    try:
        for f in range(1, frames + 1):
            res = step_all(engines, scans[f], f)
            s = of.step(scans[f], f)
            for k, r in enumerate(res):
                assert np.array_equal(bits(list(r.pose)), bits(list(s.robot))), "pose, shard %d frame %d" % (k, f)
                assert r.kd_size == s.kd_size
                if f > 1:
                    assert (r.fit_min, r.fit_max, r.best_index) == (s.fit_min, s.fit_max, s.best)
                    assert np.array_equal(bits([r.neff]), bits([s.neff])) and r.resampled == s.resampled
        for k, e in enumerate(engines):
            assert np.array_equal(e.get_kd(), of.tree), "tree of shard %d" % k
            x, y, th, w = e.get_particles()
            sl = slice(k * n, (k + 1) * n)
            assert np.array_equal(bits(x), bits(of.x[sl])) and np.array_equal(bits(th), bits(of.th[sl])) and np.array_equal(bits(w), bits(of.w[sl]))
    finally:
        for e in engines:
            e.close()
        of.close()


def test_unconnected_shard_refuses_to_step(scans):
    import gpu_icp_slam_b200 as g
    with g.ParticleFilter(1024, n_particles_global=2048, particle_offset=0, n_ranks=2) as e:
        with pytest.raises(g.PfslamError):
            e.exchange_ready()                      # rank 1 not connected
        with pytest.raises(g.PfslamError):
            e.step(scans[1], 1)


def test_missing_peer_times_out_instead_of_hanging(scans, monkeypatch):
    """a shard whose peer never steps must report the exchange timeout, not hang the GPU"""
    import gpu_icp_slam_b200 as g
    monkeypatch.setenv("PFSLAM_PEER_TIMEOUT_MS", "200")
    engines = make_shards(g, 2048, 2)
    try:
        with pytest.raises(g.PfslamError, match="timed out"):
            engines[0].step(scans[1], 1)            # engine 1 never publishes
    finally:
        for e in engines:
            e.close()
