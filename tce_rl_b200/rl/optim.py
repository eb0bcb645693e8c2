"""Flat-buffer Adam for the policy network: gradient norm, clipping and the update in two launches.

``torch.optim.Adam`` semantics (L2 weight decay, bias correction, no amsgrad; ``clip_grad_norm_`` in front) as used at
temporal_correlated_agent.py:561-589, on parameters whose ``.grad`` tensors are views of ONE flat buffer
(``TemporalCorrelatedAgent.ensure_flat_grads``).  It is a ``torch.optim.Optimizer`` (param_groups / lr schedulers work);
the moments live in two flat fp32 buffers, the step counter and the squared gradient norm in ``stats`` on the device,
so a step is CUDA-graph capturable and needs no host synchronisation.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import TceError


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, flat_grad: torch.Tensor, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        if not params or len(params) > 32:
            raise TceError("FlatAdam takes 1..32 parameter tensors")
        if not flat_grad.is_cuda or flat_grad.dtype != torch.float32:
            raise TceError("FlatAdam needs a CUDA float32 flat gradient buffer (there is no CPU path)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat_grad = flat_grad
        n = flat_grad.numel()
        if n != sum(p.numel() for p in params):
            raise TceError("flat gradient buffer does not match the parameters")
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=flat_grad.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.stats = torch.zeros(3, dtype=torch.float64, device=flat_grad.device)     # {step, sum g^2, p2p error flag}
        self._early = 0            # elements already exchanged by exchange_early in the current step
        self._xchg_stream = self._early_done = None
        self.reducer = None        # rl.p2p.P2PGradBuffer: all-reduce fused with the gradient norm (data parallel)
        self._params = params
        self._sizes = (C.c_int64 * len(params))(*[p.numel() for p in params])

    def _check_views(self):
        off, base, es = 0, self.flat_grad.data_ptr(), self.flat_grad.element_size()
        for p in self._params:
            if p.grad is None or p.grad.data_ptr() != base + off * es or not p.is_contiguous():
                raise TceError("FlatAdam: parameter gradients must be views of the flat buffer (ensure_flat_grads)")
            off += p.numel()

    # ---- checkpoints in torch.optim.Adam's format (abstract_agent.py:109-174 saves / loads optimizer.state_dict()) ----
    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` (views
        of the flat moment buffers are cloned), so a checkpoint moves freely between the two optimisers."""
        if float(self.stats[0].item()) > 0:
            off = 0
            for p in self._params:
                n = p.numel()
                self.state[p] = {"step": self.stats[0].detach().to(torch.float32).clone(),
                                 "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
                off += n
        try:
            return super().state_dict()
        finally:
            self.state.clear()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        off = 0
        step = 0.0
        for p in self._params:
            n = p.numel()
            st = self.state.get(p)
            if st:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = float(st["step"])
            else:
                self.exp_avg[off:off + n].zero_()
                self.exp_avg_sq[off:off + n].zero_()
            off += n
        self.stats.zero_()
        self.stats[0] = step
        self.state.clear()

    def exchange_early(self, n_first: int) -> bool:
        """Data parallel: exchange the first ``n_first`` elements of the flat gradient NOW (they are final: the mean
        network's slice once its backward has run); ``step`` then only exchanges the remainder.  No-op (False) without a
        peer-memory reducer or when either part would be empty."""
        n = self.flat_grad.numel()
        n_first -= n_first % 4               # 128-bit words: up to three trailing elements travel with the remainder
        if self.reducer is None or n_first <= 0 or n_first >= n:
            return False
        # on a stream of its own: the remainder's exchange (issued by ``step`` on the caller's stream once the rest of
        # the gradient is final) must not queue behind this one -- the two overlap and ``step`` joins them
        main = torch.cuda.current_stream()
        if self._xchg_stream is None:
            self._xchg_stream = torch.cuda.Stream(device=self.flat_grad.device)
            self._early_done = torch.cuda.Event()
        self._xchg_stream.wait_stream(main)
        with torch.cuda.stream(self._xchg_stream):
            self.reducer.allreduce_sumsq_range(self.stats, 0, n_first, 0, False)
            self._early_done.record(self._xchg_stream)
        self._early = int(n_first)
        return True

    def begin(self):
        """Clear the gradients and the norm accumulator (call before backward; cheap, can be issued early)."""
        self.flat_grad.zero_()
        self.stats[1:2].zero_()

    def zero_grad(self, set_to_none: bool = False):
        self.begin()

    def grad_norm(self) -> torch.Tensor:
        """2-norm of the (unclipped) flat gradient of the last ``step`` (device scalar, fp64)."""
        return self.stats[1].sqrt()

    @torch.no_grad()
    def step(self, max_norm: float = 0.0, after_norm=None):
        """``after_norm``: called once the gradient norm (and the data-parallel exchange) is queued and before the update
        kernel -- a consumer of the norm (the epoch's metrics kernel) can be forked onto another stream there."""
        self._check_views()
        g = self.param_groups[0]
        st = torch.cuda.current_stream().cuda_stream
        ptrs = (C.c_void_p * len(self._params))(*[p.data_ptr() for p in self._params])
        if self.reducer is not None:           # averaged over the ranks through NVLink peer memory, norm in the same launch
            if self._early:                    # [0, _early) went ahead (exchange_early): only the rest trails
                self.reducer.allreduce_sumsq_range(self.stats, self._early, self.flat_grad.numel() - self._early, 1, True)
                grad = self.reducer.avg[:self.flat_grad.numel()]
                torch.cuda.current_stream().wait_event(self._early_done)
                self._early = 0
            else:
                grad = self.reducer.allreduce_sumsq(self.stats)
        else:
            grad = self.flat_grad
            _lib.call("tce_grad_sumsq", grad.data_ptr(), grad.numel(), self.stats.data_ptr(), st)
        if after_norm is not None:
            after_norm()
        _lib.call("tce_adam_step", len(self._params), C.cast(ptrs, C.c_void_p), C.cast(self._sizes, C.c_void_p),
                  grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                  self.stats.data_ptr(), float(max_norm), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                  float(g["eps"]), float(g["weight_decay"]), st)
