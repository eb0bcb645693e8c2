"""Oracle restatement of the Gaussian policies over ProDMP parameters.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).  Follows
``mprl/rl/policy/abstract_policy.py:166-197`` (vector <-> Cholesky),
``mprl/rl/policy/black_box_policy.py:30-224`` (Gaussian helpers through
``torch.distributions.MultivariateNormal``) and
``mprl/rl/policy/temporal_correlated_policy.py:34-203`` (trajectory synthesis
and TCE's segment-wise likelihood).  The mean network is passed in as a callable
(plain ``torch.nn`` MLP in the tests); network construction is out of scope.
"""
from __future__ import annotations

import torch
from torch.distributions import MultivariateNormal

from . import util as ou
from .prodmp import ProDMP, get_mp


class BlackBoxPolicy:
    def __init__(self, dim_out: int, mean_net=None, cov_vector: torch.Tensor | None = None,
                 variance_net=None, std_only=False, contextual=False, min_std=1e-2,
                 dtype=torch.float64):
        self.dim_out = dim_out
        self.mean_net = mean_net
        self.variance_net = variance_net
        self.std_only = std_only
        self.contextual_cov = contextual
        self.min_std = float(min_std)
        self.dtype = dtype
        if cov_vector is None and not contextual:
            n = dim_out if std_only else dim_out + dim_out * (dim_out - 1) // 2
            cov_vector = torch.zeros(n, dtype=dtype)
            # abstract_policy.py:113-116: inverse softplus of 1 with the DEFAULT bound 1e-2
            cov_vector[:dim_out] += ou.reverse_from_softplus_space(torch.ones(dim_out, dtype=dtype), None)
        self.cov_vector = cov_vector

    # duck-typed surface the projection layers use (abstract_policy.py:312-322)
    @property
    def contextual_std(self):
        return self.contextual_cov

    @property
    def is_diag(self):
        return self.std_only

    def vector_to_cholesky(self, cov_val):
        diag = ou.to_softplus_space(cov_val[..., :self.dim_out], self.min_std)
        off = None if self.std_only else cov_val[..., self.dim_out:]
        return ou.build_lower_matrix(diag, off)

    def cholesky_to_vector(self, L):
        d, off = ou.reverse_build_matrix(L, not self.std_only)
        d = ou.reverse_from_softplus_space(d, self.min_std)
        return d if self.std_only else torch.cat([d, off], -1)

    def policy(self, obs):
        mean = self.mean_net(obs)
        if self.contextual_cov:
            L = self.vector_to_cholesky(self.variance_net(obs))
        else:
            L = self.vector_to_cholesky(ou.add_expand_dim(self.cov_vector, [0], [obs.shape[0]]))
        return mean, L

    def sample(self, require_grad, params_mean, params_L, use_mean=False, eps=None):
        if use_mean:
            smp = params_mean
        elif eps is not None:
            smp = params_mean + torch.einsum('...ij,...j->...i', params_L, eps)
        else:
            smp = MultivariateNormal(params_mean, scale_tril=params_L, validate_args=False).rsample([])
        return smp if require_grad else smp.detach()

    def log_prob(self, smp_params, params_mean, params_L, **kwargs):
        return MultivariateNormal(params_mean, scale_tril=params_L, validate_args=False).log_prob(smp_params)

    def entropy(self, params):
        return MultivariateNormal(params[0], scale_tril=params[1], validate_args=False).entropy()

    def covariance(self, L):
        return torch.einsum('...ij,...kj->...ik', L, L)

    def log_determinant(self, L):
        return 2 * L.diagonal(dim1=-2, dim2=-1).log().sum(-1)

    def precision(self, L):
        eye = torch.eye(L.shape[-1], dtype=L.dtype)
        return torch.cholesky_solve(eye, L, upper=False)

    def maha(self, x, y, L):
        diff = (x - y)[..., None]
        return torch.linalg.solve_triangular(L, diff, upper=False).pow(2).sum([-2, -1])


class TemporalCorrelatedPolicy(BlackBoxPolicy):
    def __init__(self, dim_out: int, mp: ProDMP | dict, **kw):
        super().__init__(dim_out, **kw)
        self.mp = get_mp(**mp) if isinstance(mp, dict) else mp
        self.num_dof = self.mp.num_dof

    def sample(self, require_grad, params_mean, params_L, times, init_time, init_pos, init_vel,
               use_mean=False, eps=None):
        """[*a, T, 2*D] trajectory (pos | vel).  ``eps`` [*a, Dp] injects the normal draw."""
        if not use_mean:
            pos, vel = self.mp.sample_trajectories(times=times, params=params_mean, params_L=params_L,
                                                   init_time=init_time, init_pos=init_pos, init_vel=init_vel,
                                                   num_smp=1, flat_shape=False,
                                                   eps=None if eps is None else eps[None])
            pos, vel = pos.squeeze(-3), vel.squeeze(-3)
        else:
            pos = self.mp.get_traj_pos(times=times, params=params_mean, init_time=init_time,
                                       init_pos=init_pos, init_vel=init_vel, flat_shape=False)
            vel = self.mp.get_traj_vel()
        if not require_grad:
            pos, vel = pos.detach(), vel.detach()
        return torch.cat([pos, vel], -1)

    def log_prob(self, smp_traj, params_mean, params_L, times, init_time, init_pos, init_vel,
                 return_parts=False, **kwargs):
        """Segment-wise likelihood [*a, P] (temporal_correlated_policy.py:104-203)."""
        pairs = kwargs["pred_pairs"]
        P = pairs.shape[0]
        mean_e = ou.add_expand_dim(params_mean, [-2], [P])
        L_e = ou.add_expand_dim(params_L, [-3], [P])
        time_pairs = times[:, pairs]
        it = ou.add_expand_dim(init_time, [-1], [P])
        ip = ou.add_expand_dim(init_pos, [-2], [P])
        iv = ou.add_expand_dim(init_vel, [-2], [P])
        x = smp_traj[..., pairs, :self.num_dof]                 # [*a, P, 2, D]
        x = x.transpose(-1, -2).reshape(*x.shape[:-2], -1)      # dof-major: d0@ti, d0@tj, d1@ti, ...
        self.mp.update_inputs(times=time_pairs, params=mean_e, params_L=L_e,
                              init_time=it, init_pos=ip, init_vel=iv)
        traj_mean = self.mp.get_traj_pos(flat_shape=True)
        traj_cov, reg = self.mp.get_traj_pos_cov(return_reg=True, reg_override=kwargs.get("reg_override"))
        mvn = MultivariateNormal(loc=traj_mean, covariance_matrix=traj_cov, validate_args=False)
        lp = mvn.log_prob(x)
        return (lp, traj_mean, traj_cov, reg) if return_parts else lp
