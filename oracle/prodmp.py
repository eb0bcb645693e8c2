"""Oracle restatement of the ProDMP movement primitive (mp_pytorch 0.1.4).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).  **PARITY UNPINNED**: the
``mp_pytorch`` sources are not under ``/root/reference`` (dependency pinned at
``conda_env.sh:41``: ``conda-forge::mp_pytorch=0.1.4``); this file restates the
published algorithm (Li et al., "ProDMP: A Unified Perspective on Dynamic and
Probabilistic Movement Primitives", cited at ``README.md:221-233``) behind the
method surface the reference calls:

* constructor arguments: ``mprl/util/util_mp.py:11-46`` (``get_mp``)
* ``sample_trajectories`` / ``get_traj_pos`` / ``get_traj_vel``:
  ``mprl/rl/policy/temporal_correlated_policy.py:76-92``
* ``update_inputs`` / ``get_traj_pos(flat_shape=True)`` / ``get_traj_pos_cov``:
  ``mprl/rl/policy/temporal_correlated_policy.py:188-192``

Model (per DoF, scaled time s = max((t - delay)/tau, 0), x = exp(-alpha_x s)):
  tau^2 y'' = alpha (alpha/4 (g - y) - tau y') + x * phi(x)^T w
  y(s) = c1 y1(s) + c2 y2(s) + Phi_pos(s)^T [w; g],  y1 = exp(-alpha s/2), y2 = s y1
The tables (y1, y2, dy1, dy2, Phi_pos, Phi_vel) are pre-computed on a uniform
scaled-time grid of ``5 * round(tau/dt) + 1`` points over s in [0, 5]
(``pre_compute_length_factor=5``, util_mp.py:33) with the cumulative trapezoid
rule and looked up with float indices + linear interpolation
(``indexing_interpolate``, util_matrix.py:195-227).
"""
from __future__ import annotations

import math

import torch

from . import util as ou


class ProDMPTables:
    """Pre-computed basis tables of one ProDMP configuration (float64 by default)."""

    def __init__(self, *, tau, dt, num_basis, alpha, alpha_phase, basis_bandwidth_factor,
                 delay=0.0, num_basis_outside=0, pre_compute_length_factor=5,
                 dtype=torch.float64):
        self.tau, self.dt, self.delay = float(tau), float(dt), float(delay)
        self.alpha, self.alpha_phase = float(alpha), float(alpha_phase)
        self.num_basis = K = int(num_basis)
        self.factor = int(pre_compute_length_factor)
        self.dtype = dtype
        tau_t = torch.tensor(self.tau, dtype=dtype)
        self.scaled_dt = (torch.tensor(self.dt, dtype=dtype) / tau_t)
        n_pc = self.factor * int(torch.round(1 / self.scaled_dt).item()) + 1
        self.num_pc = n_pc
        s = torch.linspace(0, self.factor, n_pc, dtype=dtype)

        # normalised RBF bases in phase space (centres spread in time, mapped by the unbounded phase)
        if K > 1:
            dist = self.tau / (K - 2 * num_basis_outside - 1)
            c_t = torch.linspace(-num_basis_outside * dist + self.delay,
                                 self.tau + num_basis_outside * dist + self.delay, K, dtype=dtype)
            c_p = torch.exp(-self.alpha_phase * (c_t - self.delay) / self.tau)
            dc = torch.cat([c_p[1:] - c_p[:-1], c_p[-1:] - c_p[-2:-1]])
            bw = basis_bandwidth_factor / dc ** 2
        else:
            c_p = torch.zeros(1, dtype=dtype)
            bw = torch.full((1,), 3.0, dtype=dtype)
        self.centers_p, self.bandwidth = c_p, bw
        x = torch.exp(-self.alpha_phase * s)                    # canonical phase on the grid
        phi = torch.exp(-0.5 * bw * (x[:, None] - c_p[None, :]) ** 2)
        if K > 1:
            phi = phi / phi.sum(-1, keepdim=True)

        a2 = 0.5 * self.alpha
        y1 = torch.exp(-a2 * s)
        y2 = s * y1
        dy1 = -a2 * y1
        dy2 = -a2 * y2 + y1
        e = torch.exp(a2 * s)
        q1 = (a2 * s - 1) * e + 1
        q2 = a2 * (e - 1)
        dp1 = (s * e * x)[:, None] * phi
        dp2 = (e * x)[:, None] * phi
        ds = s[1:] - s[:-1]
        zero = torch.zeros(1, K, dtype=dtype)
        p1 = torch.cat([zero, torch.cumsum(0.5 * (dp1[1:] + dp1[:-1]) * ds[:, None], 0)])
        p2 = torch.cat([zero, torch.cumsum(0.5 * (dp2[1:] + dp2[:-1]) * ds[:, None], 0)])

        self.y1, self.y2, self.dy1, self.dy2 = y1, y2, dy1, dy2
        self.pos_basis = torch.cat([p2 * y2[:, None] - p1 * y1[:, None],
                                    (q2 * y2 - q1 * y1)[:, None]], -1)      # [N_pc, K+1]
        self.vel_basis = torch.cat([p2 * dy2[:, None] - p1 * dy1[:, None],
                                    (q2 * dy2 - q1 * dy1)[:, None]], -1)
        # auto scale: 1 / max over the whole pre-compute range of each position basis
        self.auto_scale = 1.0 / self.pos_basis.max(dim=0)[0]

    def to(self, dtype):
        out = object.__new__(ProDMPTables)
        out.__dict__.update(self.__dict__)
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                setattr(out, k, v.to(dtype))
        out.dtype = dtype
        return out

    # ---- lookups -------------------------------------------------------
    def indices(self, times: torch.Tensor) -> torch.Tensor:
        s = torch.clip((times - self.delay) / self.tau, min=0)
        if s.numel() and s.max() > self.factor:
            raise RuntimeError("Time is beyond the pre-computation range.")
        return s / self.scaled_dt.to(times.dtype)

    def lookup(self, times: torch.Tensor):
        idx = self.indices(times)
        f = lambda tab: ou.indexing_interpolate(tab, idx)
        return f(self.y1), f(self.y2), f(self.dy1), f(self.dy2), f(self.pos_basis), f(self.vel_basis)


class ProDMP:
    """Stateful ProDMP with the calling convention the reference policy uses."""

    def __init__(self, *, num_dof, tau, dt, num_basis, alpha, alpha_phase, basis_bandwidth_factor,
                 delay=0.0, num_basis_outside=0, auto_scale_basis=True, weights_scale=1.0,
                 goal_scale=1.0, relative_goal=False, relative_goal_scaled=False,
                 dtype=torch.float64, table_dtype=None, **_ignored):
        self.num_dof = int(num_dof)
        self.num_basis = int(num_basis)
        self.num_basis_g = self.num_basis + 1
        self.relative_goal = bool(relative_goal)
        # ambiguity switch: False = "physical" g_eff = scale*g + y0
        self.relative_goal_scaled = bool(relative_goal_scaled)
        self.dtype = dtype
        tabs = ProDMPTables(tau=tau, dt=dt, num_basis=num_basis, alpha=alpha, alpha_phase=alpha_phase,
                            basis_bandwidth_factor=basis_bandwidth_factor, delay=delay,
                            num_basis_outside=num_basis_outside,
                            dtype=table_dtype or torch.float64)
        self.tables = tabs.to(dtype)
        scale = self.tables.auto_scale.clone() if auto_scale_basis else torch.ones(self.num_basis_g, dtype=dtype)
        scale[:-1] *= weights_scale
        scale[-1] *= goal_scale
        self.weights_goal_scale = scale
        self.tau = float(tau)
        self.reset()

    # ---- input cache (mp_pytorch keeps inputs between calls) -------------
    def reset(self):
        self.times = self.params = self.params_L = None
        self.init_time = self.init_pos = self.init_vel = None
        self._clear()

    def _clear(self):
        self._H = None

    def update_inputs(self, times=None, params=None, params_L=None,
                      init_time=None, init_pos=None, init_vel=None):
        for name, val in (("times", times), ("params", params), ("params_L", params_L),
                          ("init_time", init_time), ("init_pos", init_pos), ("init_vel", init_vel)):
            if val is not None:
                setattr(self, name, val)
                self._clear()

    # ---- core ------------------------------------------------------------
    def _terms(self):
        """xi_1..4 [*a, T], scaled H_pos / H_vel [*a, T, K1] (App. A.5)."""
        if self._H is None:
            tb = self.tables
            y1, y2, dy1, dy2, pb, vb = tb.lookup(self.times)
            y1b, y2b, dy1b, dy2b, pbb, vbb = (v.squeeze(self.init_time.ndim) for v in
                                              tb.lookup(self.init_time[..., None]))
            det = y1b * dy2b - y2b * dy1b
            u = lambda a: (a / det)[..., None]
            xi1 = u(dy2b) * y1 - u(dy1b) * y2
            xi2 = u(y1b) * y2 - u(y2b) * y1
            xi3 = u(dy2b) * dy1 - u(dy1b) * dy2
            xi4 = u(y1b) * dy2 - u(y2b) * dy1
            Hp = pb - xi1[..., None] * pbb[..., None, :] - xi2[..., None] * vbb[..., None, :]
            Hv = vb - xi3[..., None] * pbb[..., None, :] - xi4[..., None] * vbb[..., None, :]
            sc = self.weights_goal_scale
            self._H = (xi1, xi2, xi3, xi4, Hp * sc, Hv * sc)
        return self._H

    def _theta(self):
        """[*a, D, K1] parameters; relative goal shifts the goal by the initial position."""
        th = self.params.reshape(*self.params.shape[:-1], self.num_dof, self.num_basis_g)
        if self.relative_goal:
            th = th.clone()
            shift = self.init_pos if self.relative_goal_scaled else self.init_pos / self.weights_goal_scale[-1]
            th[..., -1] = th[..., -1] + shift
        return th

    def get_traj_pos(self, times=None, params=None, init_time=None, init_pos=None, init_vel=None,
                     flat_shape=False):
        self.update_inputs(times, params, None, init_time, init_pos, init_vel)
        xi1, xi2, _, _, Hp, _ = self._terms()
        v0 = self.init_vel * self.tau
        pos = (xi1[..., None, :] * self.init_pos[..., :, None] + xi2[..., None, :] * v0[..., :, None]
               + torch.einsum('...tk,...dk->...dt', Hp, self._theta()))          # [*a, D, T]
        return pos.reshape(*pos.shape[:-2], -1) if flat_shape else pos.transpose(-1, -2)

    def get_traj_vel(self, times=None, params=None, init_time=None, init_pos=None, init_vel=None,
                     flat_shape=False):
        self.update_inputs(times, params, None, init_time, init_pos, init_vel)
        _, _, xi3, xi4, _, Hv = self._terms()
        v0 = self.init_vel * self.tau
        vel = (xi3[..., None, :] * self.init_pos[..., :, None] + xi4[..., None, :] * v0[..., :, None]
               + torch.einsum('...tk,...dk->...dt', Hv, self._theta())) / self.tau
        return vel.reshape(*vel.shape[:-2], -1) if flat_shape else vel.transpose(-1, -2)

    def basis_multi_dof(self):
        """H_multi [*a, D*T, D*K1]: block diagonal, DoF-major / time-minor rows."""
        Hp = self._terms()[4]
        T, K1, D = Hp.shape[-2], self.num_basis_g, self.num_dof
        H = Hp.new_zeros(*Hp.shape[:-2], D * T, D * K1)
        for d in range(D):
            H[..., d * T:(d + 1) * T, d * K1:(d + 1) * K1] = Hp
        return H

    def get_traj_pos_cov(self, times=None, params_L=None, init_time=None, init_pos=None, init_vel=None,
                         reg: float = 1e-4, return_reg=False, reg_override=None):
        """H (L L^T) H^T + reg * max(diag over the WHOLE batch) * I   (App. A.5).
        ``reg_override`` (tests only): use this regulariser term instead -- lets a large batch be evaluated in
        chunks with the batch-global term computed once."""
        self.update_inputs(times, None, params_L, init_time, init_pos, init_vel)
        H = self.basis_multi_dof()
        Sigma = torch.einsum('...ij,...kj->...ik', self.params_L, self.params_L)
        cov = torch.einsum('...ik,...kl,...jl->...ij', H, Sigma, H)
        reg_term = torch.max(torch.einsum('...ii->...i', cov)).item() * reg
        if reg_override is not None:
            reg_term = float(reg_override)
        cov = cov + torch.eye(cov.shape[-1], dtype=cov.dtype) * reg_term
        return (cov, reg_term) if return_reg else cov

    def sample_trajectories(self, times=None, params=None, params_L=None, init_time=None,
                            init_pos=None, init_vel=None, num_smp=1, flat_shape=False, eps=None):
        """theta ~ N(params, L L^T) (``eps`` [num_smp, *a, Dp] may be injected), then pos/vel.

        Returns tensors with a sample axis after the batch axes: [*a, num_smp, T, D].
        """
        old = (self.times, self.params, self.params_L, self.init_time, self.init_pos, self.init_vel)
        na = params.ndim - 1
        if eps is None:
            th = torch.distributions.MultivariateNormal(loc=params, scale_tril=params_L,
                                                        validate_args=False).rsample([num_smp])
        else:
            th = params + torch.einsum('...ij,s...j->s...i', params_L, eps)
        th = torch.movedim(th, 0, na)
        ex = lambda v: ou.add_expand_dim(v, [na], [num_smp])
        self.reset()
        self.update_inputs(ex(times), th, None, ex(init_time), ex(init_pos), ex(init_vel))
        pos = self.get_traj_pos(flat_shape=flat_shape)
        vel = self.get_traj_vel(flat_shape=flat_shape)
        self.reset()
        self.update_inputs(*old)
        return pos, vel


def get_mp(**kwargs) -> ProDMP:
    """Mirror of ``mprl.util.util_mp.get_mp`` (util_mp.py:11-46) for the oracle."""
    assert kwargs["type"] == "prodmp"
    a = dict(kwargs["args"])
    dtype = a.pop("dtype", torch.float64)
    if isinstance(dtype, str):
        dtype = getattr(torch, dtype.replace("torch.", ""))
    a.pop("device", None)
    return ProDMP(dtype=dtype, **a)


LOG_2PI = math.log(2 * math.pi)
