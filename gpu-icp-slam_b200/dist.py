"""Multi-GPU particle filter: particles sharded N/R per GPU, map replicated.

One process per GPU (torchrun); `torch.distributed` is the plumbing (rendezvous, the one-off exchange
of IPC handles, barriers).  Per frame the shards need from each other
  1. per-rank score extrema {min, max, first arg-max, its pose} (32 B/rank),
  2. per-tile weight sums + tile-local CDF values (~4 B/particle),
  3. the pre-resample poses the resampler draws (12 B/particle),
and every rank then derives the identical global min/max/robotPos, Neff, resample decision, CDF
and map update.  Random streams are seeded by GLOBAL particle index and every floating-point
reduction has a fixed global tile order, so the trajectory is bit-identical for any rank count.

Two transports:
  * exchange="peer" (default on GPUs): the engines' kernels store 1 and 2 straight into the peers'
    exchange regions over NVLink and pull 3 from the owner (csrc/pf_xchg.cuh); a frame is one CUDA
    graph launch per rank, with no collective call and no host round trip.
  * exchange="collective": the step's phases with three `all_gather_into_tensor` calls between them.
    Any engine that implements the phase protocol of `engine.ParticleFilter` works (tests drive it
    with a CPU stand-in over gloo); on GPUs it is the NCCL baseline the peer transport is measured
    against.
"""
import ctypes as C

import numpy as np

from . import engine as _engine


class _CudaArray:
    """minimal __cuda_array_interface__ holder so torch can alias an engine-owned device buffer"""

    def __init__(self, ptr, nbytes, typestr, itemsize):
        self.__cuda_array_interface__ = {
            "shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def cuda_exchange_tensors(pf, device):
    """torch tensors aliasing the engine's exchange buffers (include/pfslam.h PFSLAM_BUF_*)."""
    import torch
    out = {}
    for name, which, typestr in (("ext_local", _engine.BUF_EXTREMA_LOCAL, "<i4"), ("ext_all", _engine.BUF_EXTREMA_ALL, "<i4"),
                                 ("tiles_local", _engine.BUF_TILES_LOCAL, "<f4"), ("tiles_all", _engine.BUF_TILES_ALL, "<f4"),
                                 ("pose_local", _engine.BUF_POSE_LOCAL, "<f4"), ("pose_all", _engine.BUF_POSE_ALL, "<f4")):
        ptr, nbytes = pf.device_buffer(which)
        out[name] = torch.as_tensor(_CudaArray(ptr, nbytes, typestr, 4), device=torch.device("cuda", device))
    return out


def _all_gather(dist, out, inp, group):
    try:
        dist.all_gather_into_tensor(out, inp, group=group)
    except (RuntimeError, NotImplementedError):
        chunks = list(out.view(dist.get_world_size(group), -1).unbind(0))
        dist.all_gather(chunks, inp.view(-1), group=group)


class ShardedParticleFilter:
    """The whole filter = world_size shards of `n_per_rank` particles, one per process."""

    def __init__(self, n_per_rank, device=0, group=None, engine_factory=None, exchange=None, **engine_kw):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n = int(n_per_rank)
        self.n_global = self.n * self.world
        if self.world > 1 and self.n % 1024 != 0:
            raise _engine.PfslamError("sharded filters need n_per_rank % 1024 == 0 (tile-aligned shards)")
        kw = dict(n_particles_global=self.n_global, particle_offset=self.rank * self.n, n_ranks=self.world)
        kw.update(engine_kw)
        self.exchange = exchange or ("peer" if engine_factory is None and self.world > 1 else "collective")
        if self.exchange not in ("peer", "collective"):
            raise _engine.PfslamError("exchange must be 'peer' or 'collective'")
        if engine_factory is None:
            import torch
            self.engine = _engine.ParticleFilter(self.n, device=device, **kw)
            # The engine's step is one CUDA graph, which cannot be captured on the legacy default stream:
            # when the caller's current stream is the default one, the engine keeps its own non-default
            # stream (callers order against it with engine.synchronize() / fetch_result()).
            # (The collective transport interleaves torch collectives with the engine's phases, so there the
            # engine must share torch's stream, whatever it is.)
            cur = torch.cuda.current_stream(device)
            if cur.cuda_stream != 0 or (self.world > 1 and self.exchange == "collective"):
                self.engine.set_stream(cur.cuda_stream)
            if self.world == 1:
                self.exchange = "peer"          # a single shard is just the engine's own graph step
            elif self.exchange == "peer":
                self._connect_peers(device)
            else:
                self.t = cuda_exchange_tensors(self.engine, device)
        else:
            if self.exchange == "peer":
                raise _engine.PfslamError("the peer-memory exchange needs the CUDA engine")
            self.engine = engine_factory(self.n, **kw)
            self.t = self.engine.exchange_tensors()

        self._graph = None
        self._graph_launches = 0
        self._replays = 0

    def _connect_peers(self, device):
        """one-off: all-gather the 64-byte IPC handles of the exchange regions, map every peer's region"""
        import torch
        mine = torch.frombuffer(bytearray(self.engine.ipc_export()), dtype=torch.uint8).to(torch.device("cuda", device))
        allh = torch.empty(self.world * _engine.IPC_HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
        self.dist.all_gather_into_tensor(allh, mine, group=self.group)
        raw = bytes(allh.cpu().numpy().tobytes())
        for r in range(self.world):
            if r != self.rank:
                self.engine.ipc_connect(r, raw[r * _engine.IPC_HANDLE_BYTES:(r + 1) * _engine.IPC_HANDLE_BYTES])
        self.engine.exchange_ready()
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)      # nobody publishes before every region is mapped and zeroed

    # -- one frame; nothing here synchronises the host ---------------------------------------------
    def enable_graph(self):
        """Capture the whole sharded step -- the engine's kernels AND the three all-gathers -- into one
        CUDA graph (CUDA engines only; run a few eager steps first so NCCL is initialised).  Each
        later step is a 16-byte parameter copy + one graph replay."""
        import torch
        if self.exchange == "peer":
            return                      # the engine's own step is already one graph
        e = self.engine
        e.set_external_params(True)
        e.set_params(None, 0)
        torch.cuda.synchronize()
        before = e.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=torch.cuda.current_stream()):
            self._phases(None, 0)
        self._graph_launches = e.launch_count - before
        self._graph = graph

    def step_device(self, scan_dev_ptr, frame):
        if self.exchange == "peer":
            self.engine.step_async(frame, scan_dev_ptr)
            return
        if self._graph is not None:
            self.engine.set_params(scan_dev_ptr, frame)
            self._graph.replay()
            self._replays += 1
            return
        self._phases(scan_dev_ptr, frame)

    def _phases(self, scan_dev_ptr, frame):
        e, t, d, g = self.engine, self.t, self.dist, self.group
        e.phase_motion(frame)
        _all_gather(d, t["pose_all"], t["pose_local"], g)       # pre-resample snapshot of every shard
        e.phase_score(scan_dev_ptr)
        _all_gather(d, t["ext_all"], t["ext_local"], g)         # min / max / first arg-max / best pose
        e.phase_weights()
        _all_gather(d, t["tiles_all"], t["tiles_local"], g)     # tile sums of w, w^2 and the tile-local CDF
        e.phase_map(scan_dev_ptr)                                # prefix + Neff + robotPos, then the map update
        e.phase_resample(frame)

    def step(self, scan_host, frame):
        """particleFilter(pbo, frame, lidar) for the sharded filter: host scan in, result out."""
        if self.exchange == "peer":
            return self.engine.step(scan_host, frame)
        self.engine.upload_scan(scan_host)
        self.step_device(None, frame)
        return self.engine.fetch_result()

    @property
    def launch_count(self):
        return self.engine.launch_count + self._replays * self._graph_launches

    def close(self):
        if self.exchange == "peer" and self.world > 1:
            # peers may still be pulling poses from this rank's snapshot
            self.engine.synchronize()
            self.dist.barrier(group=self.group)
        self.engine.close()
