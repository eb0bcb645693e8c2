"""Oracle restatement of the ``mprl.util`` helpers that sit on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).  Each function cites the
reference lines it follows; behaviour is pinned by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import torch


def build_lower_matrix(param_diag: torch.Tensor, param_off_diag: torch.Tensor | None) -> torch.Tensor:
    """diag + strictly-lower entries (row-major ``tril_indices(n, n, -1)`` order) -> L.

    Reference: mprl/util/util_matrix.py:12-33.
    """
    n = param_diag.shape[-1]
    L = torch.diag_embed(param_diag)
    if param_off_diag is not None:
        r, c = torch.tril_indices(n, n, -1)
        L[..., r, c] = param_off_diag
    return L


def reverse_build_matrix(L: torch.Tensor, has_off_diag: bool):
    """Inverse of :func:`build_lower_matrix`.  Reference: util_matrix.py:36-55."""
    d = torch.diagonal(L, dim1=-2, dim2=-1)
    if not has_off_diag:
        return d, None
    n = L.shape[-1]
    r, c = torch.tril_indices(n, n, -1)
    return d, L[..., r, c]


def add_expand_dim(data: torch.Tensor, add_dim_indices, add_dim_sizes) -> torch.Tensor:
    """Insert new axes at ``add_dim_indices`` (positions in the RESULT, negatives
    allowed) and expand them to ``add_dim_sizes``.  Reference: util_matrix.py:71-111.
    """
    nd = data.ndim + len(add_dim_indices)
    pos = sorted(i % nd for i in add_dim_indices)
    # sizes are consumed in increasing result-position order, as in the reference
    out = data
    for p in pos:
        out = out.unsqueeze(p)
    sizes = [-1] * nd
    for p, s in zip(pos, add_dim_sizes):
        sizes[p] = s
    return out.expand(*sizes)


def tensor_linspace(start, end, steps: int) -> torch.Tensor:
    """Vectorised linspace: out[..., k, :] style of util_matrix.py:139-192.

    For tensor ``start``/``end`` of shape [*a, d] the result is [*a, steps, d]
    (the reference transposes the last two axes of ``w0*start + w1*end``).
    """
    if not isinstance(start, torch.Tensor) and not isinstance(end, torch.Tensor):
        return torch.linspace(start, end, steps)
    if not isinstance(end, torch.Tensor):
        end = end + torch.zeros_like(start)
    if not isinstance(start, torch.Tensor):
        start = start + torch.zeros_like(end)
    w0 = torch.linspace(1, 0, steps).to(start)
    w1 = torch.linspace(0, 1, steps).to(start)
    out = w0 * start.unsqueeze(-1) + w1 * end.unsqueeze(-1)
    return out.transpose(-1, -2) if out.ndim >= 2 else out


def get_times(init_time: torch.Tensor, num_times: int, dt: float) -> torch.Tensor:
    """times[b, k] of an episode.  Reference: temporal_correlated_sampler.py:64-78."""
    return tensor_linspace(init_time + dt, init_time + num_times * dt, num_times).T


def indexing_interpolate(data: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """Float-index lookup with linear interpolation.  Reference: util_matrix.py:195-227."""
    i0 = torch.clip(indices.floor().long(), 0, data.shape[0] - 2)
    w = indices - i0
    for _ in range(data.ndim - 1):
        w = w.unsqueeze(-1)
    return torch.lerp(data[i0], data[i0 + 1], w.expand(*indices.shape, *data.shape[1:]))


def to_softplus_space(data: torch.Tensor, lower_bound: float | None) -> torch.Tensor:
    """softplus(x) + bound (default 1e-2).  Reference: util_numerical.py:44-68."""
    return torch.nn.functional.softplus(data) + (1e-2 if lower_bound is None else lower_bound)


def reverse_from_softplus_space(data: torch.Tensor, lower_bound: float | None) -> torch.Tensor:
    """Inverse of :func:`to_softplus_space`.  Reference: util_numerical.py:71-95."""
    return torch.log(torch.exp(data - (1e-2 if lower_bound is None else lower_bound)) - 1)


def select_pred_pairs(num_all: int, num_select: int, fixed_interval: bool = True,
                      first_index: int | None = None) -> torch.Tensor:
    """Consecutive pairs of the selected time indices, float32 [P, 2].

    Reference: util_learning.py:74-141 (``select_ctx_pred_pts`` with num_ctx=0)
    and :144-150.  Draws from the CPU default torch generator exactly like the
    reference (one ``randint`` for fixed intervals, one ``randperm`` otherwise).
    """
    assert num_select <= num_all
    if fixed_interval:
        interval, residual = divmod(num_all, num_select)
        if first_index is None:
            first_index = torch.randint(low=0, high=interval + residual, size=[]).item()
        assert 0 <= first_index < interval + residual
        idx = torch.arange(first_index, num_all, interval, dtype=torch.long)
    else:
        idx = torch.sort(torch.randperm(num_all)[:num_select])[0]
    pairs = torch.zeros([idx.shape[0] - 1, 2])
    pairs[:, 0] = idx[:-1]
    pairs[:, 1] = idx[1:]
    return pairs


def get_time_pairs(num_times: int, time_pairs_config: dict) -> torch.Tensor:
    """int64 [P, 2].  Reference: temporal_correlated_sampler.py:80-85."""
    return select_pred_pairs(num_all=num_times, **time_pairs_config).to(torch.long)
