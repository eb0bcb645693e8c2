"""Flat gradient buffer + the fused all-reduce / gradient-norm kernel over NVLink peer memory (csrc/tce_p2p.cu).

The data-parallel exchange of the reference-equivalent update (one flat all-reduce per optimiser step) as
ONE hand-written kernel over peer memory instead of ncclAllReduce + a reduction kernel.  ``torch.distributed`` is only
the plumbing: ``torch.distributed._symmetric_memory`` allocates the buffers and exchanges the peer mappings once.

Two schemes (``mode``, default from ``TCE_P2P_MODE`` or "push"):

* ``push`` -- every block stores its slice of the local gradient into a receive slot of every peer, releases a per-block
  flag, waits for the peers' flags and reduces from LOCAL memory: one NVLink one-way trip, no barriers
  (``tce_p2p_push_allreduce_sumsq``);
* ``pull`` -- arrive barrier, 128-bit loads through the peer mappings, departure barrier
  (``tce_p2p_allreduce_sumsq[_range]``): three round trips, kept as the cross-check.
"""
from __future__ import annotations

import ctypes as C
import os

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
    ``allreduce_sumsq(stats)``: average over the ranks into ``avg`` + squared norm, one launch;
    ``allreduce_sumsq_range``: the same for a 16-byte aligned range (an update may exchange its gradient in two ranges)."""

    def __init__(self, numel: int, device, group=None, mode: str | None = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.numel = int(numel)
        self.mode = mode or os.environ.get("TCE_P2P_MODE", "push")
        if self.mode not in ("push", "pull"):
            raise _lib.TceError(f"unknown peer-memory exchange mode {self.mode!r}")
        self.padded = padded = (self.numel + 3) // 4 * 4
        self.avg = torch.zeros(padded, dtype=torch.float32, device=device)
        self.local = torch.zeros(4, dtype=torch.int64, device=device)          # {sequence number, block ticket} per phase
        if self.mode == "pull":
            self.storage = symm_mem.empty(padded, dtype=torch.float32, device=device)
            self.storage.zero_()
            self.handle = symm_mem.rendezvous(self.storage, self.group)
            if int(self.handle.signal_pad_size) < 32 * self.world:              # two phases x 2 W slots of 8 bytes
                raise _lib.TceError("symmetric-memory signal pad is too small for the peer table")
            self._bufs = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
            self._pads = (C.c_void_p * self.world)(*[int(p) for p in self.handle.signal_pad_ptrs])
            # the signal pad may hold values of earlier symmetric-memory users: clear the slots this kernel uses
            self.handle.get_signal_pad(self.rank, (4 * self.world,), dtype=torch.int64).zero_()
        else:
            self.storage = torch.zeros(padded, dtype=torch.float32, device=device)     # gradients stay in local memory
            nbytes = int(_lib.load().tce_p2p_push_xchg_bytes(self.world, self.numel))
            self.xchg = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)  # flags + receive slots
            self.xchg.zero_()
            self.handle = symm_mem.rendezvous(self.xchg, self.group)
            self._xchg = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.buffer = self.storage[:self.numel]
        # every rank has cleared its flags before anybody's first launch
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    def allreduce_sumsq_range(self, stats: torch.Tensor, offset: int, n: int, phase: int, bump_step: bool) -> None:
        """Elements ``[offset, offset + n)`` only (``offset`` a multiple of 4; a range that ends at ``numel`` is extended
        over the zero padding), on the flags of ``phase`` (0 / 1): the mean network's slice can be exchanged while the
        covariance chain is still in its backward (``rl/fast_epoch.py``), the rest last.  ``bump_step``: this call
        advances the optimiser's step counter."""
        offset, n = int(offset), int(n)
        if offset + n == self.numel:
            n = self.padded - offset
        st = torch.cuda.current_stream().cuda_stream
        if self.mode == "push":
            _lib.call("tce_p2p_push_allreduce_sumsq", self.world, self.rank, C.cast(self._xchg, C.c_void_p),
                      self.storage.data_ptr(), self.numel, offset, n, int(phase), int(bool(bump_step)),
                      self.avg.data_ptr(), self.local.data_ptr(), stats.data_ptr(), st)
        else:
            _lib.call("tce_p2p_allreduce_sumsq_range", self.world, self.rank, C.cast(self._bufs, C.c_void_p),
                      C.cast(self._pads, C.c_void_p), offset, n, int(phase), int(bool(bump_step)), self.avg.data_ptr(),
                      self.local.data_ptr(), stats.data_ptr(), st)

    def allreduce_sumsq(self, stats: torch.Tensor) -> torch.Tensor:
        """stats [>= 3] fp64 = {step, sum g^2, error flag} as ``FlatAdam.stats``; returns the averaged gradient."""
        self.allreduce_sumsq_range(stats, 0, self.numel, 0, True)
        return self.avg[:self.numel]
