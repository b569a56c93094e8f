#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
The sharded CUDA filter (gpu-icp-slam_b200/dist.py; PF_EXCHANGE=peer|collective, PF_PATH=grid2d|kd) must
reproduce the single-rank oracle trajectory bit for bit, for any rank count."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    frames = int(os.environ.get("PF_FRAMES", "30"))
    n_total = int(os.environ.get("PF_PARTICLES", "8192"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import helpers
    from gpu_icp_slam_b200.dist import ShardedParticleFilter
    scans = helpers.fixture_scans()
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    import gpu_icp_slam_b200 as g
    kd = os.environ.get("PF_PATH", "grid2d") == "kd"
    exchange = os.environ.get("PF_EXCHANGE", "peer")          # peer-memory kernels (default) or NCCL all-gathers
    pf = ShardedParticleFilter(n_total // world, device=local, exchange=exchange, path=g.PATH_KD if kd else g.PATH_GRID2D)
    got = []
    for f in range(1, frames + 1):
        if f == 4 and os.environ.get("PF_GRAPH", "0") == "1":
            pf.enable_graph()                 # frames 4.. run as one captured graph incl. the all-gathers
        r = pf.step(scans[f], f)
        got.append(list(r.pose) + [r.fit_min, r.fit_max, r.best_index, r.neff, r.resampled, r.n_free_cells, r.n_wall_cells])
    grid = pf.engine.get_grid().reshape(-1)
    tree = pf.engine.get_kd() if kd else None
    x, y, th, w = pf.engine.get_particles()
    ok = True
    if rank == 0:
        of = helpers.OracleKdFilter(n_total) if kd else helpers.OracleFilter(n_total)
        want = []
        for f in range(1, frames + 1):
            s = of.step(scans[f], f)
            want.append([s.robot[0], s.robot[1], s.robot[2], s.fit_min, s.fit_max, s.best, s.neff, s.resampled] +
                        ([s.n_free_pts, s.n_wall_pts] if kd else [s.n_free, s.n_wall]))
        a, b = np.array(got, np.float64), np.array(want, np.float64)
        if kd:
            a, b = a[1:], b[1:]                     # frame 1 only builds the tree (no scoring)
            ok = bool(np.array_equal(a, b)) and bool(np.array_equal(tree, of.tree))
        else:
            ok = bool(np.array_equal(a, b)) and bool(np.array_equal(grid, of.grid))
        n = n_total // world
        ok = ok and np.array_equal(x.view(np.uint32), of.x[:n].view(np.uint32)) and np.array_equal(w.view(np.uint32), of.w[:n].view(np.uint32))
        if not ok:
            bad = np.flatnonzero((a != b).any(axis=1))
            print("MISMATCH first bad frame", bad[:1] + 1 if bad.size else "grid/particles", flush=True)
        print("dist_check world=%d path=%s exchange=%s particles=%d frames=%d resamples=%d: %s" %
              (world, "kd" if kd else "grid2d", pf.exchange, n_total, frames, int(b[:, 7].sum()), "OK bit-exact" if ok else "FAILED"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    pf.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
