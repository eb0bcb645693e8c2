"""Flat gradient buffer in NVLink symmetric memory + the fused all-reduce / gradient-norm kernel (csrc/tce_p2p.cu).

The data-parallel exchange of the reference-equivalent update (one flat all-reduce per optimiser step) as
ONE hand-written kernel over peer memory instead of ncclAllReduce + a reduction kernel.  ``torch.distributed`` is only
the plumbing: ``torch.distributed._symmetric_memory`` allocates the buffer and exchanges the peer mappings once."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .. import _lib


def available(group=None) -> bool:
    """NCCL process group on CUDA with symmetric-memory support and a world that fits the kernel's peer table."""
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:
        return False
    if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
        return False
    return dist.get_backend(group) == "nccl" and 1 < dist.get_world_size(group) <= 16


class P2PGradBuffer:
    """``buffer``: this rank's flat fp32 gradient buffer (give the parameters' ``.grad`` views into it);
    ``allreduce_sumsq(stats)``: average over the ranks into ``avg`` + squared norm, one launch."""

    def __init__(self, numel: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.numel = int(numel)
        padded = (self.numel + 3) // 4 * 4
        self.storage = symm_mem.empty(padded, dtype=torch.float32, device=device)
        self.storage.zero_()
        self.handle = symm_mem.rendezvous(self.storage, self.group)
        if int(self.handle.signal_pad_size) < 16 * self.world:
            raise _lib.TceError("symmetric-memory signal pad is too small for the peer table")
        self.buffer = self.storage[:self.numel]
        self.avg = torch.zeros(padded, dtype=torch.float32, device=device)
        self.local = torch.zeros(2, dtype=torch.int64, device=device)          # {sequence number, block ticket}
        self._bufs = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self._pads = (C.c_void_p * self.world)(*[int(p) for p in self.handle.signal_pad_ptrs])
        # the signal pad may hold values of earlier symmetric-memory users: clear the slots this kernel uses, then make
        # sure every rank has done so before the first launch
        pad = self.handle.get_signal_pad(self.rank, (2 * self.world,), dtype=torch.int64)
        pad.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    def allreduce_sumsq(self, stats: torch.Tensor) -> torch.Tensor:
        """stats [>= 3] fp64 = {step, sum g^2, error flag} as ``FlatAdam.stats``; returns the averaged gradient."""
        _lib.call("tce_p2p_allreduce_sumsq", self.world, self.rank, C.cast(self._bufs, C.c_void_p),
                  C.cast(self._pads, C.c_void_p), self.numel, self.avg.data_ptr(), self.local.data_ptr(),
                  stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return self.avg[:self.numel]
