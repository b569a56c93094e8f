"""CPU stand-in for one shard engine, built on the oracle, implementing the phase protocol that
gpu-icp-slam_b200/dist.py drives (same buffers, same layouts as include/pfslam.h PFSLAM_BUF_*).
Lets the multi-rank orchestration be tested over gloo without a GPU."""
import ctypes as C

import numpy as np
import torch

import helpers
from helpers import P

TILE = 1024


class OracleShardEngine:
    def __init__(self, n, n_particles_global=None, particle_offset=0, n_ranks=1, quirks=1):
        self.o = helpers.load_oracle()
        self.cfg = helpers.ocfg(q1=1 if quirks & 1 else 0)
        self.n, self.ng, self.off, self.R = n, n_particles_global or n, particle_offset, n_ranks
        self.q1 = bool(quirks & 1)
        self.nt = (n + TILE - 1) // TILE
        self.block = 2 * self.nt + n
        self.pose_local = np.zeros(3 * n, np.float32)
        self.pose_all = np.zeros(3 * n * n_ranks, np.float32)
        self.ext_local = np.zeros(8, np.int32)
        self.ext_all = np.zeros(8 * n_ranks, np.int32)
        self.tiles_local = np.zeros(self.block, np.float32)
        self.tiles_all = np.zeros(self.block * n_ranks, np.float32)
        self.w = np.ones(n, np.float32)
        self.fit = np.zeros(n, np.int32)
        nc = self.cfg.map_w * self.cfg.map_h
        self.grid = np.full(nc, -100, np.int8)
        self.scan = np.zeros(1081, np.float32)
        self.res = {}
        self.launch_count = 0

    x = property(lambda s: s.pose_local[: s.n])
    y = property(lambda s: s.pose_local[s.n: 2 * s.n])
    th = property(lambda s: s.pose_local[2 * s.n:])

    def exchange_tensors(self):
        return {k: torch.from_numpy(getattr(self, k)) for k in
                ("ext_local", "ext_all", "tiles_local", "tiles_all", "pose_local", "pose_all")}

    def upload_scan(self, scan):
        self.scan[:] = scan

    def phase_motion(self, frame):
        self.o.pfo_add_noise(P(self.x), P(self.y), P(self.th), self.n, frame, self.off)
        if self.R == 1:
            self.pose_all[:] = self.pose_local

    def phase_score(self, _ptr=None):
        self.o.pfo_score2d_many(C.byref(self.cfg), P(self.grid, helpers.bp), P(self.x), P(self.y), P(self.th),
                                self.n, P(self.scan), P(self.fit, helpers.ip))
        b = int(np.argmax(self.fit))              # first maximum
        self.ext_local[:3] = (self.fit.min(), self.fit.max(), self.off + b)
        self.ext_local[3:6] = np.array([self.x[b], self.y[b], self.th[b]], np.float32).view(np.int32)
        if self.R == 1:
            self.ext_all[:] = self.ext_local

    def _extrema(self):
        e = self.ext_all.reshape(self.R, 8)
        gmin = int(e[:, 0].min())
        key = [(int(e[r, 1]), -int(e[r, 2])) for r in range(self.R)]
        br = max(range(self.R), key=lambda r: key[r])
        pose = e[br, 3:6].copy().view(np.float32)
        return gmin, int(e[br, 1]), int(e[br, 2]), pose

    def phase_weights(self):
        gmin, gmax, _, _ = self._extrema()
        weff = self.w.copy()
        if gmax > gmin:
            c = np.float32(1.0) / np.float32(gmax - gmin)
            weff = (weff * (self.fit.astype(np.float32) - np.float32(gmin))) * c
        n_sync = (self.ng + 1) // 2 if self.q1 else self.ng
        keep = (self.off + np.arange(self.n)) < n_sync
        self.w[keep] = weff[keep]
        lm = np.zeros(self.n, np.float32); tt = np.zeros(self.nt, np.float32)
        lm2 = np.zeros(self.n, np.float32); tt2 = np.zeros(self.nt, np.float32)
        sq = (weff * weff).astype(np.float32)
        self.o.pfo_scan_tiles(P(weff), self.n, P(lm), P(tt))
        self.o.pfo_scan_tiles(P(sq), self.n, P(lm2), P(tt2))
        self.tiles_local[: self.nt] = tt
        self.tiles_local[self.nt: 2 * self.nt] = tt2
        self.tiles_local[2 * self.nt:] = lm
        if self.R == 1:
            self.tiles_all[:] = self.tiles_local

    def phase_map(self, _ptr=None):
        ta = self.tiles_all.reshape(self.R, self.block)
        p = np.float32(0.0); p2 = np.float32(0.0)
        prefix = [p]
        for r in range(self.R):
            for t in range(self.nt):
                p = np.float32(p + ta[r, t]); p2 = np.float32(p2 + ta[r, self.nt + t])
                prefix.append(p)
        self.prefix = np.array(prefix, np.float32)
        gmin, gmax, best, pose = self._extrema()
        with np.errstate(all="ignore"):
            neff = np.float32(np.float32(p * p) / p2)
        self.res = dict(pose=[float(v) for v in pose], fit_min=gmin, fit_max=gmax, best_index=best,
                        sum_w=float(p), sum_w2=float(p2), neff=float(neff),
                        resampled=int(float(neff) < 0.7 * self.ng))
        cx, cy = C.c_int(), C.c_int()
        self.o.pfo_center_cell(C.byref(self.cfg), C.c_float(pose[0]), C.c_float(pose[1]), C.byref(cx), C.byref(cy))
        nc = self.grid.size
        fm = np.zeros(nc, np.uint8); wm = np.zeros(nc, np.uint8)
        self.o.pfo_get_walls(C.byref(self.cfg), P(self.scan), cx.value, cy.value, C.c_float(pose[2]), P(fm, helpers.ubp), P(wm, helpers.ubp))
        self.o.pfo_apply_masks(P(self.grid, helpers.bp), nc, P(fm, helpers.ubp), P(wm, helpers.ubp))

    def phase_resample(self, frame):
        if not self.res["resampled"]:
            return
        ta = self.tiles_all.reshape(self.R, self.block)
        cdf = np.concatenate([ta[r, 2 * self.nt:] for r in range(self.R)]).astype(np.float32)
        for t in range(self.R * self.nt):
            cdf[t * TILE:(t + 1) * TILE] = self.prefix[t] + cdf[t * TILE:(t + 1) * TILE]
        pa = self.pose_all.reshape(self.R, 3, self.n)
        neff = C.c_float(self.res["neff"]); tot = C.c_float(self.res["sum_w"])
        for i in range(self.n):
            src = self.o.pfo_resample_src(P(cdf), self.ng, tot, neff, frame, self.off + i)
            r, l = divmod(src, self.n)
            self.pose_local[i] = pa[r, 0, l]; self.pose_local[self.n + i] = pa[r, 1, l]; self.pose_local[2 * self.n + i] = pa[r, 2, l]
        self.w[:] = 1.0

    def fetch_result(self):
        return self.res

    def close(self):
        pass
