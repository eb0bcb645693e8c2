"""Mathematical invariants that pin the UNPINNED parts of the oracle."""
import math

import numpy as np
import pytest
import scipy.optimize
import scipy.stats
import torch

from oracle import policy as opol
from oracle import projection as oproj
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from oracle.prodmp import ProDMP

torch.set_default_dtype(torch.float32)


def make_policy(name, contextual=True):
    cfg = MP_CONFIGS[name]
    Dp = cfg["num_dof"] * (cfg["num_basis"] + 1)
    return opol.TemporalCorrelatedPolicy(Dp, mp=dict(type="prodmp", args=dict(cfg, dtype=torch.float64)),
                                         contextual=contextual, min_std=1e-4)


@pytest.mark.parametrize("name", list(MP_CONFIGS))
def test_prodmp_initial_conditions(name):
    mp = ProDMP(**MP_CONFIGS[name])
    D, Dp = mp.num_dof, mp.num_dof * mp.num_basis_g
    g = torch.Generator().manual_seed(0)
    B = 5
    theta = torch.randn(B, Dp, generator=g, dtype=torch.float64)
    y0 = torch.randn(B, D, generator=g, dtype=torch.float64)
    v0 = torch.randn(B, D, generator=g, dtype=torch.float64)
    delay = MP_CONFIGS[name].get("delay", 0.0)
    t0 = delay + torch.tensor([0.0, 0.1, 0.25, 0.4, 0.0], dtype=torch.float64)
    # evaluate exactly at the initial time (and a bit later)
    times = torch.stack([t0, t0 + 0.05], -1)
    pos = mp.get_traj_pos(times, theta, t0, y0, v0)
    vel = mp.get_traj_vel()
    torch.testing.assert_close(pos[:, 0], y0, rtol=0, atol=1e-10)
    torch.testing.assert_close(vel[:, 0], v0, rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", list(MP_CONFIGS))
def test_prodmp_velocity_is_derivative(name):
    cfg = MP_CONFIGS[name]
    mp = ProDMP(**cfg)
    D, Dp = mp.num_dof, mp.num_dof * mp.num_basis_g
    g = torch.Generator().manual_seed(1)
    theta = torch.randn(1, Dp, generator=g, dtype=torch.float64)
    y0 = torch.randn(1, D, generator=g, dtype=torch.float64)
    v0 = torch.randn(1, D, generator=g, dtype=torch.float64)
    t0 = torch.full((1,), cfg.get("delay", 0.0), dtype=torch.float64)
    h = cfg["dt"]          # grid points: the lerp is exact there, central differences are O(h^2)
    times = (t0 + h * torch.arange(1, NUM_TIMES[name] + 1, dtype=torch.float64))[None]
    pos = mp.get_traj_pos(times, theta, t0, y0, v0)[0]
    vel = mp.get_traj_vel()[0]
    fd = (pos[2:] - pos[:-2]) / (2 * h)
    scale = vel.abs().max()
    assert (fd - vel[1:-1]).abs().max() / scale < 3e-2   # O(h^2) of the stiff start + table mis-registration (App. A.3)


@pytest.mark.parametrize("name", list(MP_CONFIGS))
def test_prodmp_matches_ode(name):
    """RK4 on tau^2 y'' = alpha(alpha/4 (g - y) - tau y') + x phi(x)^T w  (App. A.1)."""
    cfg = MP_CONFIGS[name]
    mp = ProDMP(**cfg)
    tb = mp.tables
    K = mp.num_basis
    g = torch.Generator().manual_seed(2)
    theta = torch.randn(1, mp.num_dof * mp.num_basis_g, generator=g, dtype=torch.float64)
    y0 = torch.randn(1, mp.num_dof, generator=g, dtype=torch.float64)
    v0 = torch.randn(1, mp.num_dof, generator=g, dtype=torch.float64)
    delay, tau, alpha = cfg.get("delay", 0.0), cfg["tau"], float(cfg["alpha"])
    t0 = torch.full((1,), delay, dtype=torch.float64)
    T = NUM_TIMES[name]
    times = (t0 + cfg["dt"] * torch.arange(1, T + 1, dtype=torch.float64))[None]
    pos = mp.get_traj_pos(times, theta, t0, y0, v0)[0].numpy()            # [T, D]
    th = mp._theta()[0].numpy() * mp.weights_goal_scale.numpy()           # physical weights / goal
    w, goal = th[:, :K], th[:, K]
    c_p, bw = tb.centers_p.numpy(), tb.bandwidth.numpy()

    def force(s):
        x = math.exp(-cfg["alpha_phase"] * s)
        phi = np.exp(-0.5 * bw * (x - c_p) ** 2)
        phi = phi / phi.sum()
        return x * (w @ phi)

    def rhs(s, z):                       # z = [y, y'] in scaled time
        y, dy = z
        return np.stack([dy, alpha * (alpha / 4 * (goal - y) - dy) + force(s)])

    n_sub = 40
    # quirk: the lookup index is s / scaled_dt while the table grid step is
    # factor / (N_pc - 1); they differ when tau/dt is not an integer (table tennis: 93.75 -> 94),
    # i.e. the primitive runs in a slightly warped scaled time.
    warp = (tb.factor / (tb.num_pc - 1)) / tb.scaled_dt.item()
    h = cfg["dt"] / tau / n_sub * warp
    z = np.stack([y0[0].numpy(), v0[0].numpy() * tau])
    s, out = 0.0, []
    for _ in range(T):
        for _ in range(n_sub):
            k1 = rhs(s, z)
            k2 = rhs(s + h / 2, z + h / 2 * k1)
            k3 = rhs(s + h / 2, z + h / 2 * k2)
            k4 = rhs(s + h, z + h * k3)
            z = z + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            s += h
        out.append(z[0].copy())
    out = np.stack(out)
    assert np.abs(out - pos).max() / np.abs(pos).max() < 2e-3


@pytest.mark.parametrize("name", list(MP_CONFIGS))
def test_segment_likelihood_matches_scipy(name):
    pol = make_policy(name)
    cfg = MP_CONFIGS[name]
    B, T = 3, NUM_TIMES[name]
    inp = synthetic_inputs(name, B, seed=9)
    times = ou.get_times(inp["init_time"], T, cfg["dt"])
    torch.manual_seed(1)
    pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
    args = (times, inp["init_time"], inp["init_pos"], inp["init_vel"])
    smp = pol.sample(False, inp["mean"], inp["L"], *args, eps=inp["eps"])
    lp, mu, cov, reg = pol.log_prob(smp, inp["mean"], inp["L"], *args, pred_pairs=pairs, return_parts=True)
    n = 2 * pol.num_dof
    ev = torch.linalg.eigvalsh(cov - reg * torch.eye(n, dtype=torch.float64))
    assert ev.min() > -1e-9 * ev.max()                       # H Sigma H^T is PSD
    assert torch.linalg.eigvalsh(cov).min() > 0
    D = pol.num_dof
    for b in range(B):
        for p in (0, 7, pairs.shape[0] - 1):
            x = smp[b, pairs[p], :D].T.reshape(-1).numpy()
            want = scipy.stats.multivariate_normal(mu[b, p].numpy(), cov[b, p].numpy()).logpdf(x)
            assert abs(want - lp[b, p].item()) < 1e-7 * max(1.0, abs(want))
    # the regulariser is batch global: 1e-4 * max diag
    raw = cov - reg * torch.eye(n, dtype=torch.float64)
    assert abs(reg - 1e-4 * raw.diagonal(dim1=-2, dim2=-1).max().item()) < 1e-12


def test_sample_covariance_monte_carlo():
    pol = make_policy("table_tennis")
    cfg = MP_CONFIGS["table_tennis"]
    inp = synthetic_inputs("table_tennis", 1, seed=4)
    T = NUM_TIMES["table_tennis"]
    times = ou.get_times(inp["init_time"], T, cfg["dt"])
    pairs = torch.tensor([[40, 90]])
    N = 20000
    g = torch.Generator().manual_seed(0)
    eps = torch.randn(N, pol.dim_out, generator=g, dtype=torch.float64)
    ex = lambda v: v.expand(N, *v.shape[1:])
    smp = pol.sample(False, ex(inp["mean"]), ex(inp["L"]), ex(times), ex(inp["init_time"]),
                     ex(inp["init_pos"]), ex(inp["init_vel"]), eps=eps)
    x = smp[:, pairs[0], :pol.num_dof].transpose(-1, -2).reshape(N, -1)
    _, mu, cov, reg = pol.log_prob(smp[:1], inp["mean"], inp["L"], times, inp["init_time"], inp["init_pos"],
                                   inp["init_vel"], pred_pairs=pairs, return_parts=True)
    emp = torch.cov(x.T)
    raw = cov[0, 0] - reg * torch.eye(cov.shape[-1], dtype=torch.float64)
    assert (emp - raw).abs().max() < 0.05 * raw.abs().max()
    assert (x.mean(0) - mu[0, 0]).abs().max() < 0.05 * raw.diagonal().max().sqrt()


# ---- projections ---------------------------------------------------------------------------------------
LAYERS = [("KLProjectionLayer", 0.05, 5e-4), ("FrobeniusProjectionLayer", 0.05, 5e-4),
          ("WassersteinProjectionLayer", 0.005, 2.5e-4)]


def make_layer(typ, mb, cb, Dp, schedule=None):
    return oproj.projection_factory(typ, proj_type=typ, mean_bound=mb, cov_bound=cb, trust_region_coeff=1.0,
                                    scale_prec=True, entropy_schedule=schedule, action_dim=Dp,
                                    total_train_steps=7500, target_entropy=0.0, temperature=0.7,
                                    dtype=torch.float64)


@pytest.mark.parametrize("typ,mb,cb", LAYERS)
def test_projection_satisfies_bounds(typ, mb, cb):
    pol = make_policy("box")
    inp = synthetic_inputs("box", 16, seed=3)
    layer = make_layer(typ, mb, cb, pol.dim_out)
    layer.initial_entropy = pol.entropy([inp["mean_old"], inp["L_old"]]).mean()
    if typ == "WassersteinProjectionLayer":
        # the commutative W2 closed form is a metric only for symmetric square roots; TCE feeds
        # Cholesky factors, so check the bound on diagonal factors where both agree
        inp["L"] = torch.diag_embed(inp["L"].diagonal(dim1=-2, dim2=-1))
        inp["L_old"] = torch.diag_embed(inp["L_old"].diagonal(dim1=-2, dim2=-1))
    p, q = (inp["mean"], inp["L"]), (inp["mean_old"], inp["L_old"])
    m0, c0 = layer.trust_region_value(pol, p, q)
    assert (m0 > mb).any() and (c0 > cb).any()               # the synthetic data violates the bounds
    proj = layer(pol, p, q, 100)
    m1, c1 = layer.trust_region_value(pol, proj, q)
    assert (m1 <= mb * (1 + 1e-5)).all()
    assert (c1 <= cb * (1 + 1e-5)).all()
    # inside the region -> returned unchanged
    same = layer(pol, q, q, 100)
    assert torch.equal(same[0], q[0]) and torch.allclose(same[1], q[1], rtol=0, atol=1e-12)
    assert layer.get_trust_region_loss(pol, q, same, set_variance=False).abs() < 1e-12


def test_kl_projection_matches_dual_optimum():
    pol = make_policy("table_tennis")
    inp = synthetic_inputs("table_tennis", 4, seed=8)
    eps = 5e-4
    cov_proj, active = oproj.kl_cov_projection(inp["L"], inp["L_old"], torch.tensor(eps, dtype=torch.float64))
    assert active.all()
    k = pol.dim_out
    for b in range(4):
        S_new = (inp["L"][b] @ inp["L"][b].T).numpy()
        S_old = (inp["L_old"][b] @ inp["L_old"][b].T).numpy()
        P_new, P_old = np.linalg.inv(S_new), np.linalg.inv(S_old)

        def neg_dual(eta):
            P = (P_new + eta * P_old) / (1 + eta)
            return -(-eta * eps + 0.5 * np.linalg.slogdet(S_new)[1] + 0.5 * eta * np.linalg.slogdet(S_old)[1]
                     + 0.5 * (1 + eta) * np.linalg.slogdet(P)[1])
        res = scipy.optimize.minimize_scalar(neg_dual, bounds=(0, 1e4), method="bounded",
                                             options=dict(xatol=1e-10))
        S_ref = np.linalg.inv((P_new + res.x * P_old) / (1 + res.x))
        assert np.abs(S_ref - cov_proj[b].numpy()).max() < 1e-6 * np.abs(S_ref).max()
        kl = 0.5 * (np.trace(P_old @ cov_proj[b].numpy()) - k + np.linalg.slogdet(S_old)[1]
                    - np.linalg.slogdet(cov_proj[b].numpy())[1])
        assert abs(kl - eps) < 1e-9


def test_entropy_projection_hits_bound():
    pol = make_policy("table_tennis")
    inp = synthetic_inputs("table_tennis", 6, seed=2)
    p = (inp["mean"], inp["L"])
    ent = pol.entropy(p)
    beta = ent.mean() * torch.ones(6, dtype=torch.float64)
    _, L2 = oproj.entropy_inequality_projection(pol, p, beta)
    e2 = pol.entropy((inp["mean"], L2))
    low = ent < beta
    assert low.any() and (~low).any()
    torch.testing.assert_close(e2[low], beta[low], rtol=0, atol=1e-10)
    assert torch.equal(L2[~low], inp["L"][~low])


@pytest.mark.parametrize("typ,mb,cb", LAYERS)
def test_projection_gradcheck(typ, mb, cb):
    cfg = dict(MP_CONFIGS["table_tennis"], num_dof=2, num_basis=2)
    Dp = 6
    pol = opol.TemporalCorrelatedPolicy(Dp, mp=dict(type="prodmp", args=dict(cfg, dtype=torch.float64)),
                                        contextual=True, min_std=1e-4)
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    B = 3
    mean, mean_old = rn(B, Dp), rn(B, Dp)
    mean_old = mean + 0.3 * rn(B, Dp)
    L_old = torch.tril(0.2 * rn(B, Dp, Dp), -1) + torch.diag_embed(0.5 + torch.rand(B, Dp, generator=g, dtype=torch.float64))
    L = L_old + torch.tril(0.1 * rn(B, Dp, Dp))
    L[0], mean[0] = L_old[0] + 1e-4 * torch.tril(rn(Dp, Dp)), mean_old[0] + 1e-4     # inactive branch
    layer = make_layer(typ, mb, cb, Dp, "linear")
    layer.initial_entropy = pol.entropy([mean_old, L_old]).mean()
    w_m, w_L = rn(B, Dp), torch.tril(rn(B, Dp, Dp))

    def f(m, Lv):
        pm, pL = layer(pol, (m, torch.tril(Lv)), (mean_old, L_old), 100)
        return (pm * w_m).sum() + (pL * w_L).sum()
    assert torch.autograd.gradcheck(f, (mean.requires_grad_(True), L.requires_grad_(True)), eps=1e-6, atol=1e-6,
                                    rtol=1e-5)
    # the trust-region loss treats the projection as a constant (detached target)
    with torch.no_grad():
        target = layer(pol, (mean, torch.tril(L)), (mean_old, L_old), 100)

    def tr(m, Lv):
        return layer.get_trust_region_loss(pol, (m, torch.tril(Lv)), target, set_variance=False)
    assert torch.autograd.gradcheck(tr, (mean, L), eps=1e-6, atol=1e-6, rtol=1e-5)
