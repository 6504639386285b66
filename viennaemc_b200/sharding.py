"""Multi-GPU plumbing: block partition of the particle ids; for bulk runs the one collective they need (sum of the
per-step observable series); for device runs the per-step all-reduce callback of emcgpu_device_set_sharding (reservoir
particles per rank and cell, carriers per grid point) and the host mirrors of the rules the kernels use to split the
contact handling between the ranks.  torch.distributed only; works with nccl (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """[first, last) global particle ids of a rank: contiguous blocks, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    base, rem = divmod(int(n_total), world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def allreduce_observables(series, group=None):
    """Sum the [steps][valleys][3] = {sum E, sum v.E_dir, count} partial series of all ranks in place
    (torch tensor; one collective per run or per chunk -- bulk runs need no per-step communication)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(series, op=dist.ReduceOp.SUM, group=group)
    return series


def finalize_observables(series, n_total=None):
    """{sum E, sum v.E, count} -> per-valley <E>, <v.E_dir>, occupation, as the reference's getAvgEnergy /
    getAvgDriftVelocity / getValleyOccupationProbability return them (valleys without particles -> 0)."""
    s = np.asarray(series, dtype=np.float64)
    cnt = s[..., 2]
    total = cnt.sum(axis=-1, keepdims=True) if n_total is None else float(n_total)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(cnt > 0, s[..., 0] / cnt, 0.0)
        v = np.where(cnt > 0, s[..., 1] / cnt, 0.0)
        occ = cnt / total
    return e, v, occ


# ---- device runs: particles sharded, grids replicated -------------------------------------------------
def inject_share_of_rank(missing: int, rank: int, world: int, cell: int, step: int) -> int:
    """How many of the `missing` particles of a reservoir cell a rank injects (injectShareOfRank in
    viennaemc_b200/csrc/emc_device_run.cuh): an even split whose remainder rotates with cell and step."""
    rotated = (rank + world - (cell + step) % world) % world
    return missing // world + (1 if rotated < missing % world else 0)


def reservoir_decisions(share, expected, nr_carriers, rank, step):
    """What rank `rank` does in the reservoir cells given share[r][cell] = reservoir particles of rank r in that cell (the
    table the ranks all-reduce every step): (kept, dropped, injected) per cell.  Global index order = rank 0's particles,
    then rank 1's, ...; the first `slots` of a cell survive (handleOhmicContacts, emcBasicParticleHandler.hpp:158-192)."""
    share = np.asarray(share, dtype=np.int64)
    world, cells = share.shape
    slots = np.where(np.asarray(expected) > 0, np.ceil(np.asarray(expected) / nr_carriers), 0).astype(np.int64)
    lower = share[:rank].sum(axis=0)
    mine = share[rank]
    kept = np.clip(slots - lower, 0, mine)
    total_kept = np.minimum(share.sum(axis=0), slots)
    diff = np.asarray(expected) - total_kept * nr_carriers
    missing = np.where(diff > 0, np.ceil(diff / nr_carriers), 0).astype(np.int64)
    injected = np.array([inject_share_of_rank(int(m), rank, world, c, step) for c, m in enumerate(missing)], dtype=np.int64)
    return kept, mine - kept, injected


class DeviceRunSharding:
    """Binds a capi.Context to a torch.distributed process group for a sharded device run: installs the all-reduce callback
    (torch.distributed.all_reduce on a zero-copy view of the library's device buffer, on the same stream)."""

    def __init__(self, ctx, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.calls = 0

        class _View:  # __cuda_array_interface__ of a raw device pointer
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=2)

        device = torch.device("cuda", getattr(ctx, "device", torch.cuda.current_device()))

        def _allreduce(user, ptr, count, stream):
            # on the LIBRARY's stream (the kernels before and after the reduce are ordered on it), whatever torch's
            # current stream is; the default stream (0) is torch's default stream of that device
            t = torch.as_tensor(_View(ptr, int(count)), device=device)
            ext = torch.cuda.ExternalStream(int(stream), device=device) if stream else torch.cuda.default_stream(device)
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1

        self._cb = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)(_allreduce)  # keep alive
        ctx.device_set_sharding(self.rank, self.world, self._cb)

    def sum_counters(self, counters):
        """per-contact counters of a run are per rank: the terminal currents need their sum"""
        import torch
        import torch.distributed as dist

        t = torch.as_tensor(np.ascontiguousarray(counters, dtype=np.int64)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class NcclDeviceRunSharding:
    """Sharded device run with the library's exchange going straight to ncclAllReduce (libemcnccl.so, include/emcnccl.h):
    the all-reduce callback is a C function, no Python between the step kernels.  torch.distributed only carries the 128-byte
    unique id from rank 0 to the other ranks (any backend)."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        import os

        import torch.distributed as dist

        lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libemcnccl.so")
        if not os.path.exists(lib_path):
            raise ImportError(f"{lib_path} is missing: build it with `python -m viennaemc_b200.build`")
        self.L = C.CDLL(lib_path)
        self.L.emcnccl_last_error.restype = C.c_char_p
        self.L.emcnccl_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        self.L.emcnccl_calls.restype = C.c_int64
        self.L.emcnccl_calls.argtypes = [C.c_void_p]
        self.L.emcnccl_bytes.restype = C.c_int64
        self.L.emcnccl_bytes.argtypes = [C.c_void_p]
        self.L.emcnccl_destroy.argtypes = [C.c_void_p]
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = C.create_string_buffer(128)
        if self.rank == 0 and self.L.emcnccl_unique_id(ident) != 0:
            raise RuntimeError(self.L.emcnccl_last_error().decode())
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0, group=group)
        self.comm = C.c_void_p()
        if self.L.emcnccl_init(box[0], self.rank, self.world, ctx.device, C.byref(self.comm)) != 0:
            raise RuntimeError(self.L.emcnccl_last_error().decode())
        fn = C.cast(self.L.emcnccl_allreduce_sum_f64, C.c_void_p)
        ctx.device_set_sharding(self.rank, self.world, fn, self.comm)

    @property
    def calls(self):
        return int(self.L.emcnccl_calls(self.comm))

    @property
    def bytes(self):
        return int(self.L.emcnccl_bytes(self.comm))

    def sum_counters(self, counters):
        import torch
        import torch.distributed as dist

        t = torch.as_tensor(np.ascontiguousarray(counters, dtype=np.int64)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def close(self):
        if self.comm:
            self.ctx.device_set_sharding(0, 1, None)
            self.L.emcnccl_destroy(self.comm)
            self.comm = None
