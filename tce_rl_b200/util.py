"""Host-side helpers with the names and semantics of ``mprl.util`` (only what the update path touches).

Reference: mprl/util/util_matrix.py, util_learning.py, util_numerical.py, util_data_structure.py,
util_hyperparams.py, util_nn.py.  These are tensor plumbing; the arithmetic of the hot path lives in
the CUDA kernels behind ``tce_rl_b200.ops``.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import nn


def parse_dtype_device(dtype, device):
    """util_data_structure.py:59-78 (fp32 is the kernels' arithmetic type; fp64 is refused loudly)."""
    if isinstance(dtype, torch.dtype):
        dt = dtype
    elif dtype in ("float32", "torch.float32"):
        dt = torch.float32
    elif dtype in ("float64", "torch.float64"):
        dt = torch.float64
    else:
        raise NotImplementedError(dtype)
    return dt, torch.device(device)


def add_expand_dim(data: torch.Tensor, add_dim_indices, add_dim_sizes) -> torch.Tensor:
    """util_matrix.py:71-111: new axes at the given RESULT positions, expanded (views, no copies)."""
    total = data.ndim + len(add_dim_indices)
    where = sorted(i % total for i in add_dim_indices)
    shape = [-1] * total
    for pos, size in zip(where, add_dim_sizes):
        data = data.unsqueeze(pos)
        shape[pos] = size
    return data.expand(*shape)


def build_lower_matrix(param_diag: torch.Tensor, param_off_diag: Optional[torch.Tensor]) -> torch.Tensor:
    """util_matrix.py:12-33 (plain torch; the policy head uses the fused ``tce::policy_head`` kernel)."""
    n = param_diag.shape[-1]
    L = param_diag.diag_embed()
    if param_off_diag is not None:
        rows, cols = torch.tril_indices(n, n, -1)
        L[..., rows, cols] = param_off_diag
    return L


def reverse_build_matrix(L: torch.Tensor, has_off_diag: bool):
    """util_matrix.py:36-55."""
    diag = torch.diagonal(L, dim1=-2, dim2=-1)
    if not has_off_diag:
        return diag, None
    rows, cols = torch.tril_indices(L.shape[-1], L.shape[-1], -1)
    return diag, L[..., rows, cols]


def tensor_linspace(start, end, steps: int) -> torch.Tensor:
    """util_matrix.py:139-192: [*a, d] endpoints -> [*a, steps, d]."""
    if not torch.is_tensor(start) and not torch.is_tensor(end):
        return torch.linspace(start, end, steps)
    if not torch.is_tensor(end):
        end = torch.zeros_like(start) + end
    if not torch.is_tensor(start):
        start = torch.zeros_like(end) + start
    assert start.shape == end.shape
    w_start = torch.linspace(1, 0, steps).to(start)
    w_end = torch.linspace(0, 1, steps).to(start)
    out = w_start * start[..., None] + w_end * end[..., None]
    return out.transpose(-1, -2)


def to_softplus_space(data, lower_bound: Optional[float]):
    """util_numerical.py:44-68."""
    return nn.functional.softplus(data) + (1e-2 if lower_bound is None else lower_bound)


def reverse_from_softplus_space(data, lower_bound: Optional[float]):
    """util_numerical.py:71-95."""
    return torch.log(torch.exp(data - (1e-2 if lower_bound is None else lower_bound)) - 1)


def select_ctx_pred_pts(**kwargs):
    """util_learning.py:74-141.  Index sampling stays on the HOST torch generator (bit-exact parity)."""
    num_ctx = kwargs.get("num_ctx", None)
    first_index = kwargs.get("first_index", None)
    fixed_interval = kwargs.get("fixed_interval", False)
    num_all, num_select = kwargs.get("num_all", None), kwargs.get("num_select", None)
    ctx_before_pred = kwargs.get("ctx_before_pred", False)
    if num_select is None:
        assert fixed_interval is False and first_index is None
        num_select = num_all
    else:
        assert num_select <= num_all
    if num_ctx is None:
        num_ctx = torch.randint(low=kwargs.get("num_ctx_min"), high=kwargs.get("num_ctx_max"), size=(1,))
    assert num_ctx < num_select
    if fixed_interval:
        interval, residual = num_all // num_select, num_all % num_select
        if first_index is None:
            first_index = torch.randint(low=0, high=interval + residual, size=[]).item()
        else:
            assert 0 <= first_index < interval + residual
        selected = torch.arange(start=first_index, end=num_all, step=interval, dtype=torch.long)
    else:
        selected = torch.sort(torch.randperm(n=num_all)[:num_select])[0]
    if num_ctx == 0:
        return [], selected
    if ctx_before_pred:
        return selected[:num_ctx], selected[num_ctx:]
    perm = torch.randperm(n=num_select)
    return selected[perm[:num_ctx]], selected[perm[num_ctx:]]


def select_pred_pairs(**kwargs) -> torch.Tensor:
    """util_learning.py:144-150: consecutive pairs of the selected indices (float tensor, as the reference)."""
    idx = select_ctx_pred_pts(num_ctx=0, **kwargs)[1]
    pairs = torch.zeros([idx.shape[0] - 1, 2])
    pairs[:, 0] = idx[:-1]
    pairs[:, 1] = idx[1:]
    return pairs


def mlp_arch_3_params(avg_neuron: int, num_hidden: int, shape: float):
    """util_hyperparams.py:7-46."""
    assert avg_neuron >= 0 and -1.0 <= shape <= 1.0 and num_hidden >= 1
    slope = shape * avg_neuron
    arch = []
    for i in range(num_hidden):
        x = 2 * i / (num_hidden - 1) - 1 if num_hidden != 1 else 0.0
        arch.append(max(int(np.floor(slope * x + avg_neuron)), 1))
    return arch


_ACT = {"tanh": torch.tanh, "relu": nn.functional.relu, "leaky_relu": nn.functional.leaky_relu,
        "softplus": nn.functional.softplus, None: None}


_SIDE_GRAD_STREAMS = set()          # streams with parameter-gradient work still to be joined
_LAYER_STREAMS = {}                 # id(nn.Linear) -> its side stream (kept OFF the module: deepcopy / pickle of an MLP
#                                     must not meet CUDA stream objects)


def _layer_stream(lin, device):
    st = _LAYER_STREAMS.get(id(lin))
    if st is None:
        st = _LAYER_STREAMS[id(lin)] = torch.cuda.Stream(device=device)
    return st


def join_side_grads():
    """Make the current stream wait for every weight/bias gradient that ``MLP(side_wgrad=True)`` queued on side
    streams during the last backward pass.  Call it after ``loss.backward()`` and before the gradients are read."""
    cur = torch.cuda.current_stream() if _SIDE_GRAD_STREAMS else None
    for st in _SIDE_GRAD_STREAMS:
        cur.wait_stream(st)
    _SIDE_GRAD_STREAMS.clear()


class _SideGradLinear(torch.autograd.Function):
    """y = x W^T + b.  Backward: only d/dx stays on the critical chain of the backward pass; the weight and bias
    gradients (two library GEMM calls per layer, leaves of the graph) are ACCUMULATED into the pre-allocated
    ``.grad`` buffers on a per-layer side stream and autograd is handed no gradient for them.  The caller joins
    the streams (``join_side_grads``) before it reads the gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, owner):
        ctx.save_for_backward(x, weight)
        ctx.owner = owner
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        owner = ctx.owner
        g = g.contiguous()
        main = torch.cuda.current_stream()
        side = _layer_stream(owner, g.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
            owner.weight.grad.addmm_(g2.t(), x2)
            owner.bias.grad.addmv_(g2.t(), g2.new_ones(g2.shape[0]))
            g.record_stream(side)
            x.record_stream(side)
        _SIDE_GRAD_STREAMS.add(side)
        gx = g @ weight if ctx.needs_input_grad[0] else None
        return gx, None, None, None


class MLP(nn.Module):
    """Dense net of util_nn.py:75-246 (orthogonal init with gain sqrt(2), output layer gain configurable).

    Stays on cuBLAS through torch (not one of the four hand-written subsystems).
    """

    def __init__(self, name, dim_in, dim_out, hidden_layers, init_method, out_layer_gain, act_func_hidden,
                 act_func_last, dtype=torch.float32, device="cpu"):
        super().__init__()
        self.mlp_name = name + "_mlp"
        dims = [dim_in, *hidden_layers, dim_out]
        self.layers = nn.ModuleList(nn.Linear(a, b, dtype=dtype, device=device) for a, b in zip(dims[:-1], dims[1:]))
        self.act_hidden, self.act_last = _ACT[act_func_hidden], _ACT[act_func_last]
        self.act_hidden_name, self.act_last_name = act_func_hidden, act_func_last
        for i, lin in enumerate(self.layers):
            gain = out_layer_gain if i == len(self.layers) - 1 else math.sqrt(2)
            if init_method == "orthogonal":
                nn.init.orthogonal_(lin.weight, gain=gain)
            elif init_method == "xavier":
                nn.init.xavier_normal_(lin.weight, gain=gain)
            else:
                raise ValueError(f"unsupported init_method {init_method}")
            nn.init.zeros_(lin.bias)

    side_wgrad = False      # opt-in (the agent's update loops): see _SideGradLinear / join_side_grads

    def _linear(self, lin, x):
        if (self.side_wgrad and x.is_cuda and torch.is_grad_enabled() and lin.weight.grad is not None
                and lin.bias.grad is not None):
            return _SideGradLinear.apply(x, lin.weight, lin.bias, lin)
        return lin(x)

    def forward(self, x):
        for lin in self.layers[:-1]:
            x = self.act_hidden(self._linear(lin, x))
        x = self._linear(self.layers[-1], x)
        return x if self.act_last is None else self.act_last(x)


class TrainableVariable:
    """util_nn.py:449-520: a bare parameter standing in for a network (non-contextual covariance)."""

    def __init__(self, name, data):
        self.name = name
        self.variable = nn.Parameter(data=data)

    @property
    def data(self):
        return self.variable.data

    def parameters(self):
        return [self.variable]
