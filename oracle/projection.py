"""Oracle restatement of the differentiable trust-region projection layers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).  **PARITY UNPINNED**: the
classes the reference instantiates (``mprl/rl/projection/__init__.py:2-13,19-40``)
live in ``BruceGeLi/trust-region-layers`` (branch ``TCE_ICLR24``, commit
``9f33038b...``; ``conda_env.sh:57``, ``README.md:120-123``) and, for the KL
covariance projection, in the C++ ``cpp_projection`` package
(``conda_env.sh:34``) -- neither is present under ``/root/reference``.  This file
restates the published layers (Otto et al., "Differentiable Trust Region Layers
for Deep Reinforcement Learning", ICLR 2021) behind the call surface used at
``mprl/rl/agent/temporal_correlated_agent.py:439-441,530-533,561-567,641-686``.

All layers are written with plain differentiable torch ops, so torch autograd of
this file is the gradient oracle for the hand-written CUDA backward kernels.
The KL covariance projection solves the dual exactly (scalar root find on the
generalised eigenvalues) instead of NLopt L-BFGS.
"""
from __future__ import annotations

import math

import torch


# --------------------------------------------------------------------------
# distances (projection_utils of the dependency)
# --------------------------------------------------------------------------
def batched_trace(x):
    return x.diagonal(dim1=-2, dim2=-1).sum(-1)


def mean_distance(policy, mean, mean_other, chol_other=None, scale_prec=False):
    if scale_prec:
        return policy.maha(mean, mean_other, chol_other)
    return ((mean_other - mean) ** 2).sum(-1)


def gaussian_kl(policy, p, q):
    """KL(p || q) split in (mean part, covariance part), each [B]."""
    mean, chol = p
    mean_o, chol_o = q
    k = mean.shape[-1]
    maha_part = 0.5 * policy.maha(mean, mean_o, chol_o)
    trace_part = batched_trace(policy.precision(chol_o) @ policy.covariance(chol))
    cov_part = 0.5 * (trace_part - k + policy.log_determinant(chol_o) - policy.log_determinant(chol))
    return maha_part, cov_part


def gaussian_kl_details(policy, p, q):
    """(mean, cov, shape, volume) parts, cov = shape + volume (assumed split)."""
    mean, chol = p
    mean_o, chol_o = q
    k = mean.shape[-1]
    maha_part = 0.5 * policy.maha(mean, mean_o, chol_o)
    trace_part = batched_trace(policy.precision(chol_o) @ policy.covariance(chol))
    shape_part = 0.5 * (trace_part - k)
    volume_part = 0.5 * (policy.log_determinant(chol_o) - policy.log_determinant(chol))
    return maha_part, shape_part + volume_part, shape_part, volume_part


def gaussian_frobenius(policy, p, q, scale_prec=False, return_cov=False):
    mean, chol = p
    mean_o, chol_o = q
    mean_part = mean_distance(policy, mean, mean_o, chol_o, scale_prec)
    cov_o, cov = policy.covariance(chol_o), policy.covariance(chol)
    diff = cov_o - cov
    cov_part = batched_trace(diff @ diff)
    return (mean_part, cov_part, cov, cov_o) if return_cov else (mean_part, cov_part)


def gaussian_wasserstein_commutative(policy, p, q, scale_prec=False):
    mean, sqrt = p
    mean_o, sqrt_o = q
    mean_part = mean_distance(policy, mean, mean_o, sqrt_o, scale_prec)
    cov = policy.covariance(sqrt)
    if scale_prec:
        eye = torch.eye(mean.shape[-1], dtype=sqrt.dtype)
        inv_o = torch.linalg.solve(sqrt_o, eye.expand_as(sqrt_o))
        c = inv_o @ cov @ inv_o
        cov_part = batched_trace(eye + c - 2 * inv_o @ sqrt)
    else:
        cov_part = batched_trace(policy.covariance(sqrt_o) + cov - 2 * sqrt_o @ sqrt)
    return mean_part, cov_part


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def mean_projection(mean, old_mean, maha, eps):
    """Closed-form interpolation of the mean onto the bound (App. B.2)."""
    mask = maha > eps
    if not mask.any():
        return mean
    omega = torch.ones_like(maha)
    omega[mask] = torch.sqrt(maha[mask] / eps) - 1.0
    omega = torch.max(-omega, omega)[..., None]
    m = (mean + omega * old_mean) / (1 + omega + 1e-16)
    return torch.where(mask[..., None], m, mean)


def entropy_inequality_projection(policy, p, beta):
    mean, chol = p
    k = mean.shape[-1]
    ent = policy.entropy(p)
    mask = ent < beta
    if (~mask).all():
        return p
    alpha = torch.ones_like(ent)
    alpha[mask] = torch.exp((beta[mask] - ent[mask]) / k)
    return mean, torch.where(mask[..., None, None], chol * alpha[..., None, None], chol)


def entropy_equality_projection(policy, p, beta):
    mean, chol = p
    k = mean.shape[-1]
    alpha = torch.exp((beta - policy.entropy(p)) / k)
    return mean, chol * alpha[..., None, None]


def get_entropy_schedule(kind, total_train_steps, dim):
    if kind == "linear":
        return lambda init, target, temp, step: step * (target * dim - init) / total_train_steps + init
    if kind == "exp":
        return lambda init, target, temp, step: dim * target + (init - dim * target) * temp ** (
            10 * step / total_train_steps)
    return lambda init, target, temp, step: torch.as_tensor(-math.inf)


# --------------------------------------------------------------------------
# KL covariance projection (cpp_projection / ITPAL restated; exact dual solve)
# --------------------------------------------------------------------------
def kl_cov_eta_function(lam, eta):
    """KL_cov(Sigma(eta) || Sigma_old) from the generalised eigenvalues lam [*, k]."""
    r = (lam + eta[..., None]) / (1 + eta[..., None])
    return 0.5 * (1 / r - 1 + torch.log(r)).sum(-1)


def kl_cov_solve_eta(lam, eps, iters=200):
    """Root of KL_cov(eta) = eps, eta >= 0 (monotone decreasing) -- bracketing + bisection."""
    lam = lam.detach()
    lo = torch.zeros(lam.shape[:-1], dtype=lam.dtype)
    hi = torch.ones_like(lo)
    for _ in range(200):
        too_small = kl_cov_eta_function(lam, hi) > eps
        if not too_small.any():
            break
        hi = torch.where(too_small, hi * 2, hi)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        big = kl_cov_eta_function(lam, mid) > eps
        lo = torch.where(big, mid, lo)
        hi = torch.where(big, hi, mid)
    return 0.5 * (lo + hi)


def kl_cov_projection(chol, chol_old, eps_cov):
    """Sigma_proj [B,k,k] and active mask; differentiable w.r.t. ``chol`` (implicit eta*).

    W = L~^-1 L_old, N = W^T W = Q diag(lam) Q^T,
    Sigma(eta) = L_old Q diag((1+eta)/(lam+eta)) Q^T L_old^T  (SURVEY App. B.4).
    """
    W = torch.linalg.solve_triangular(chol, chol_old, upper=False)
    N = W.transpose(-1, -2) @ W
    lam, Q = torch.linalg.eigh(N)
    zero = torch.zeros(lam.shape[:-1], dtype=lam.dtype)
    kl0 = kl_cov_eta_function(lam, zero)
    active = kl0 > eps_cov
    eta = torch.zeros_like(kl0)
    if active.any():
        eta_star = kl_cov_solve_eta(lam, eps_cov)
        # one differentiable Newton step re-attaches the implicit gradient d eta*/d lam
        e = eta_star.clone().requires_grad_(True)
        with torch.enable_grad():
            f = kl_cov_eta_function(lam.detach(), e)
            dfde, = torch.autograd.grad(f.sum(), e)
        eta_impl = eta_star - (kl_cov_eta_function(lam, eta_star) - eps_cov) / dfde.detach()
        eta = torch.where(active, eta_impl, eta)
    d = (1 + eta[..., None]) / (lam + eta[..., None])
    M = chol_old @ Q
    cov_proj = (M * d[..., None, :]) @ M.transpose(-1, -2)
    cov = chol @ chol.transpose(-1, -2)
    return torch.where(active[..., None, None], cov_proj, cov), active


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
class BaseProjectionLayer:
    def __init__(self, proj_type="", mean_bound=0.03, cov_bound=1e-3, trust_region_coeff=0.0,
                 scale_prec=True, entropy_schedule=None, action_dim=None, total_train_steps=None,
                 target_entropy=0.0, temperature=0.5, entropy_eq=False, entropy_first=False,
                 do_regression=False, cpu=True, dtype=torch.float32, **_ignored):
        self.proj_type = proj_type
        self.mean_bound = torch.as_tensor(mean_bound, dtype=dtype)
        self.cov_bound = torch.as_tensor(cov_bound, dtype=dtype)
        self.trust_region_coeff = trust_region_coeff
        self.scale_prec = scale_prec
        assert (action_dim and total_train_steps) if entropy_schedule else True
        self.entropy_proj = entropy_equality_projection if entropy_eq else entropy_inequality_projection
        self.entropy_schedule = get_entropy_schedule(entropy_schedule, total_train_steps, action_dim)
        self.target_entropy = torch.as_tensor(target_entropy, dtype=dtype)
        self.entropy_first = entropy_first
        self.temperature = temperature
        self._initial_entropy = None
        self.dtype = dtype

    @property
    def initial_entropy(self):
        return self._initial_entropy

    @initial_entropy.setter
    def initial_entropy(self, value):
        if self._initial_entropy is None:          # write once
            self._initial_entropy = value

    def __call__(self, policy, p, q, step, **kwargs):
        beta = self.entropy_schedule(self.initial_entropy, self.target_entropy, self.temperature, step)
        beta = beta * p[0].new_ones(p[0].shape[0])
        if self.entropy_first:
            p = self.entropy_proj(policy, p, beta)
        proj = self._trust_region_projection(policy, p, q, self.mean_bound, self.cov_bound, **kwargs)
        return proj if self.entropy_first else self.entropy_proj(policy, proj, beta)

    def _trust_region_projection(self, policy, p, q, eps, eps_cov, **kwargs):
        return p

    def trust_region_value(self, policy, p, q):
        return gaussian_kl(policy, p, q)

    def get_trust_region_loss(self, policy, p, proj_p, set_variance=None):
        """coeff * mean(mean_diff [+ cov_diff]) against the DETACHED projection.

        The covariance term is included when the covariance is learned by
        gradient: contextual std, or (fork, assumed -- SURVEY App. B.5) an
        explicit ``set_variance=False``.
        """
        target = (proj_p[0].detach(), proj_p[1].detach())
        mean_diff, cov_diff = self.trust_region_value(policy, p, target)
        with_cov = policy.contextual_std or (set_variance is not None and not set_variance)
        return (mean_diff + cov_diff if with_cov else mean_diff).mean() * self.trust_region_coeff

    def compute_metrics(self, policy, p, q, step=None):
        with torch.no_grad():
            ent = policy.entropy(p)
            mean_kl, cov_kl = gaussian_kl(policy, p, q)
            mean_diff, cov_diff = self.trust_region_value(policy, p, q)
            kl, con = mean_kl + cov_kl, mean_diff + cov_diff
            return {"kl": kl.mean(), "constraint": con.mean(), "mean_constraint": mean_diff.mean(),
                    "cov_constraint": cov_diff.mean(), "entropy": ent.mean(),
                    "entropy_diff": (policy.entropy(q) - ent).mean(), "kl_max": kl.max(),
                    "constraint_max": con.max(), "mean_constraint_max": mean_diff.max(),
                    "cov_constraint_max": cov_diff.max(), "entropy_max": ent.max()}


class KLProjectionLayer(BaseProjectionLayer):
    def _trust_region_projection(self, policy, p, q, eps, eps_cov, **kwargs):
        mean, chol = p
        old_mean, old_chol = q
        mean_part, _ = gaussian_kl(policy, p, q)
        proj_mean = mean_projection(mean, old_mean, mean_part, eps)
        if not policy.contextual_std:              # one shared covariance: project the first only
            chol, old_chol = chol[:1], old_chol[:1]
        if policy.is_diag:
            raise NotImplementedError("diagonal KL projection is outside the TCE configs")
        cov_proj, _ = kl_cov_projection(chol, old_chol, eps_cov)
        proj_chol = torch.linalg.cholesky(cov_proj)
        if not policy.contextual_std:
            proj_chol = proj_chol.expand(mean.shape[0], -1, -1)
        return proj_mean, proj_chol


class FrobeniusProjectionLayer(BaseProjectionLayer):
    def _trust_region_projection(self, policy, p, q, eps, eps_cov, **kwargs):
        mean, chol = p
        old_mean, _ = q
        mean_part, cov_part, cov, cov_old = gaussian_frobenius(policy, p, q, self.scale_prec, True)
        proj_mean = mean_projection(mean, old_mean, mean_part, eps)
        mask = cov_part > eps_cov
        if not mask.any():
            return proj_mean, chol
        eta = torch.ones_like(cov_part)
        eta[mask] = torch.sqrt(cov_part[mask] / eps_cov) - 1.0
        eta = torch.max(-eta, eta)[..., None, None]
        new_cov = (cov + eta * cov_old) / (1.0 + eta + 1e-16)
        # inactive rows have eta = 1 (an SPD average); their factor is discarded by ``where``
        return proj_mean, torch.where(mask[..., None, None], torch.linalg.cholesky(new_cov), chol)

    def trust_region_value(self, policy, p, q):
        return gaussian_frobenius(policy, p, q, self.scale_prec)

    def get_trust_region_loss(self, policy, p, proj_p, set_variance=None):
        target = (proj_p[0].detach(), proj_p[1].detach())
        mean_diff, _ = self.trust_region_value(policy, p, target)
        with_cov = policy.contextual_std or (set_variance is not None and not set_variance)
        if with_cov:
            mean_diff = mean_diff + (p[1] - target[1]).pow(2).sum([-1, -2])
        return mean_diff.mean() * self.trust_region_coeff


class WassersteinProjectionLayer(BaseProjectionLayer):
    def _trust_region_projection(self, policy, p, q, eps, eps_cov, **kwargs):
        mean, sqrt = p
        old_mean, old_sqrt = q
        mean_part, cov_part = gaussian_wasserstein_commutative(policy, p, q, self.scale_prec)
        proj_mean = mean_projection(mean, old_mean, mean_part, eps)
        mask = cov_part > eps_cov
        if not mask.any():
            return proj_mean, sqrt
        eta = torch.ones_like(cov_part)
        eta[mask] = torch.sqrt(cov_part[mask] / eps_cov) - 1.0
        eta = torch.max(-eta, eta)[..., None, None]
        new_sqrt = (sqrt + eta * old_sqrt) / (1.0 + eta + 1e-16)
        return proj_mean, torch.where(mask[..., None, None], new_sqrt, sqrt)

    def trust_region_value(self, policy, p, q):
        return gaussian_wasserstein_commutative(policy, p, q, self.scale_prec)


def projection_factory(typ: str, **kwargs):
    """Mirror of ``mprl/rl/projection/__init__.py:19-40`` for the oracle."""
    kwargs = dict(kwargs)
    dtype = kwargs.get("dtype", torch.float32)
    if isinstance(dtype, str):
        dtype = getattr(torch, dtype.replace("torch.", ""))
    kwargs["dtype"] = dtype
    kwargs["cpu"] = True
    kwargs.pop("device", None)
    return {"BaseProjectionLayer": BaseProjectionLayer, "KLProjectionLayer": KLProjectionLayer,
            "FrobeniusProjectionLayer": FrobeniusProjectionLayer,
            "WassersteinProjectionLayer": WassersteinProjectionLayer}[typ](**kwargs)
