"""Trust-region projections and the full policy epoch on the GPU vs the CPU oracle (fp64, same fp32 inputs).

Tolerance: projected parameters / KLs / losses <= 1e-4 absolute (north_star); gradients relative to their scale.
"""
import pytest
import torch

from oracle import agent as oa
from oracle import policy as opol
from oracle import projection as oproj
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from tce_rl_b200 import ops
    from tce_rl_b200.rl import (TemporalCorrelatedAgent, policy_factory, projection_factory)
    from tce_rl_b200.rl.agent import SegmentTimeSampler

DEV = "cuda:0"
LAYERS = [("KLProjectionLayer", 0.05, 5e-4), ("FrobeniusProjectionLayer", 0.05, 5e-4),
          ("WassersteinProjectionLayer", 0.005, 2.5e-4)]


def f64(t):
    return t.detach().double().cpu()


def layer_kwargs(typ, mb, cb, Dp, schedule="linear", scale_prec=True):
    return dict(proj_type=typ, mean_bound=mb, cov_bound=cb, trust_region_coeff=1.0, scale_prec=scale_prec,
                entropy_schedule=schedule, action_dim=Dp, total_train_steps=7500, target_entropy=0.0,
                temperature=0.7, entropy_eq=False, entropy_first=False, do_regression=False)


class FakePolicy:
    """Duck-typed policy surface the GPU layers need in these tests."""
    def __init__(self, contextual):
        self.contextual_std, self.is_diag = contextual, False

    def entropy(self, p):
        return ops.gauss_stats(p[0], p[1], p[0], p[1])[:, 4]


def case(name, B, contextual, seed=3, diag_only=False):
    inp = synthetic_inputs(name, B, seed=seed, dtype=torch.float32)
    if not contextual:
        for k in ("L", "L_old"):
            inp[k] = inp[k][:1].expand(B, -1, -1).contiguous()
    if diag_only:
        for k in ("L", "L_old"):
            inp[k] = torch.diag_embed(inp[k].diagonal(dim1=-2, dim2=-1))
    return inp


@pytest.mark.parametrize("typ,mb,cb", LAYERS)
@pytest.mark.parametrize("name,contextual", [("box", True), ("box", False), ("table_tennis", True),
                                              ("metaworld", False)])
def test_projection_forward_backward(typ, mb, cb, name, contextual):
    B = 12
    cfg = MP_CONFIGS[name]
    Dp = cfg["num_dof"] * (cfg["num_basis"] + 1)
    inp = case(name, B, contextual, diag_only=(typ == "WassersteinProjectionLayer" and name == "table_tennis"))
    # make one episode lie inside the trust region (identity branch)
    inp["mean"][0] = inp["mean_old"][0] + 1e-3
    if contextual:
        inp["L"][0] = inp["L_old"][0] * (1 + 1e-4)
    d = lambda k: inp[k].double()
    opolicy = opol.BlackBoxPolicy(Dp, contextual=contextual, min_std=1e-4)
    olayer = oproj.projection_factory(typ, dtype=torch.float64, **layer_kwargs(typ, mb, cb, Dp))
    init_ent = opolicy.entropy([d("mean_old"), d("L_old")]).mean()
    olayer.initial_entropy = init_ent
    m64, L64 = d("mean").requires_grad_(True), d("L").requires_grad_(True)
    pm, pL = olayer(opolicy, (m64, L64), (d("mean_old"), d("L_old")), 100)
    g = torch.Generator().manual_seed(0)
    wm = torch.randn(B, Dp, generator=g, dtype=torch.float64)
    wL = torch.tril(torch.randn(B, Dp, Dp, generator=g, dtype=torch.float64))
    tr = olayer.get_trust_region_loss(opolicy, (m64, L64), (pm, pL), set_variance=False)
    gm, gL = torch.autograd.grad((pm * wm).sum() + (pL * wL).sum() + tr, [m64, L64])

    layer = projection_factory(typ, device=DEV, dtype="float32", **layer_kwargs(typ, mb, cb, Dp))
    layer.initial_entropy = init_ent.float().to(DEV)
    pol = FakePolicy(contextual)
    c = lambda k: inp[k].to(DEV)
    mg, Lg = c("mean").requires_grad_(True), c("L").requires_grad_(True)
    pmg, pLg = layer(pol, (mg, Lg), (c("mean_old"), c("L_old")), 100)
    assert (f64(pmg) - pm.detach()).abs().max() <= 1e-4
    assert (f64(pLg) - pL.detach()).abs().max() <= 1e-4
    trg = layer.get_trust_region_loss(pol, (mg, Lg), (pmg, pLg), set_variance=False)
    assert abs(trg.item() - tr.item()) <= 1e-4 * max(1.0, abs(tr.item()))
    ((pmg * wm.float().to(DEV)).sum() + (pLg * wL.float().to(DEV)).sum() + trg).backward()
    assert (f64(mg.grad) - gm).abs().max() <= 2e-4 * max(1.0, gm.abs().max().item())
    gL_t, gL_g = torch.tril(gL), f64(Lg.grad)
    if not contextual:
        # one shared covariance: only the batch SUM reaches the covariance vector (the GPU layer evaluates
        # the shared covariance terms once instead of B times, so the per-copy split differs)
        gL_t, gL_g = gL_t.sum(0), gL_g.sum(0)
    assert (gL_g - gL_t).abs().max() <= 2e-4 * max(1.0, gL_t.abs().max().item())
    # both branches were exercised
    mp0, cp0 = olayer.trust_region_value(opolicy, (d("mean"), d("L")), (d("mean_old"), d("L_old")))
    assert (mp0 > mb).any() and (mp0 <= mb).any()


def test_kl_projection_constraint_satisfied_on_gpu():
    """After the GPU projection KL_cov == eps (active) -- checked with the oracle's fp64 KL."""
    B, name = 8, "box"
    Dp = 63
    inp = case(name, B, True, seed=11)
    layer = projection_factory("KLProjectionLayer", device=DEV, dtype="float32",
                               **layer_kwargs("KLProjectionLayer", 0.05, 5e-4, Dp, schedule=None))
    pol = FakePolicy(True)
    c = lambda k: inp[k].to(DEV)
    pm, pL = layer(pol, (c("mean"), c("L")), (c("mean_old"), c("L_old")), 0)
    opolicy = opol.BlackBoxPolicy(Dp, contextual=True, min_std=1e-4)
    mk, ck = oproj.gaussian_kl(opolicy, (f64(pm), f64(pL)), (inp["mean_old"].double(), inp["L_old"].double()))
    assert (mk <= 0.05 * (1 + 1e-4)).all()
    assert ((ck - 5e-4).abs() <= 2e-6).all()        # fp32 storage of proj_L limits this, not the solver


@pytest.mark.parametrize("name,typ,contextual,fast", [("box", "KLProjectionLayer", False, False),
                                                       ("box", "KLProjectionLayer", False, True),
                                                       ("box", "KLProjectionLayer", True, False),
                                                       ("metaworld", "KLProjectionLayer", False, False),
                                                       ("metaworld", "KLProjectionLayer", False, True),
                                                       ("table_tennis", "KLProjectionLayer", False, True),
                                                       ("table_tennis", "WassersteinProjectionLayer", False, False),
                                                       ("box", "FrobeniusProjectionLayer", True, False)])
def test_policy_epoch_matches_oracle(name, typ, contextual, fast):
    """Loss, logging KLs and parameter gradients of one update_policy epoch (temporal_correlated_agent.py:524-589);
    ``fast``: through the hand-scheduled shared-covariance epoch (rl/fast_epoch.py), else the generic one."""
    torch.manual_seed(0)
    B = 24
    cfg = MP_CONFIGS[name]
    D, K1, T = cfg["num_dof"], cfg["num_basis"] + 1, NUM_TIMES[name]
    Dp, obs_dim = D * K1, 10
    mb, cb = dict(LAYERS and {t: (a, b) for t, a, b in LAYERS})[typ]
    inp = case(name, B, contextual, seed=21)
    # --- GPU side: real drop-in classes ------------------------------------------------------------
    policy = policy_factory("TemporalCorrelatedPolicy", dim_in=obs_dim, dim_out=Dp,
                            mean_net_args=dict(avg_neuron=32, num_hidden=2, shape=0.0),
                            variance_net_args=dict(std_only=False, contextual=contextual, avg_neuron=32,
                                                   num_hidden=2, shape=0.0),
                            init_method="orthogonal", out_layer_gain=0.01, act_func_hidden="leaky_relu",
                            act_func_last=None, dtype="float32", device=DEV, min_std=1e-4,
                            mp=dict(type="prodmp", args=dict(cfg)))
    with torch.no_grad():                      # move the policy off its initial point
        for p in policy.parameters:
            p.add_(0.02 * torch.randn_like(p))
    layer = projection_factory(typ, device=DEV, dtype="float32", **layer_kwargs(typ, mb, cb, Dp))
    sampler = SegmentTimeSampler(cfg["dt"], T, dict(num_select=25, fixed_interval=True), device=DEV)
    torch.manual_seed(2)
    pairs = sampler.get_time_pairs()
    agent = TemporalCorrelatedAgent(policy, None, sampler, layer, dtype="float32", device=DEV, lr_policy=1e-4,
                                    lr_critic=1e-3, wd_policy=5e-5, wd_critic=5e-5, discount_factor=1.0,
                                    epochs_policy=1, epochs_critic=1, norm_advantages=True,
                                    segment_advantage="value_subtraction", set_variance=False,
                                    fused_surrogate=(name != "metaworld" or fast), fast_epoch=fast)
    if fast:
        agent.ensure_flat_grads(agent.policy_net_params)          # FlatAdam: the fast epoch's optimiser
    obs = torch.randn(B, obs_dim + 2 * D)
    c = lambda t: t.to(DEV)
    init_time = inp["init_time"]
    times = sampler.get_times(c(init_time), T)
    with torch.no_grad():
        mean_old, L_old = policy.policy(c(obs)[..., :-2 * D])
        mean_old = mean_old + 0.05 * c(torch.randn(B, Dp))
        L_old = (1.03 * L_old + 0.01 * torch.tril(c(torch.randn(B, Dp, Dp)), -1))
        if not contextual:
            L_old = L_old[:1].expand(B, -1, -1).contiguous()
        smp = policy.sample(False, mean_old, L_old, times, c(init_time), c(inp["init_pos"]), c(inp["init_vel"]),
                            eps=c(inp["eps"]))
        lp_old = policy.log_prob(smp, mean_old, L_old, times, c(init_time), c(inp["init_pos"]), c(inp["init_vel"]),
                                 pred_pairs=pairs)
        adv, ret = agent.get_advantage_return(c(inp["rewards"]), c(inp["values"]), c(inp["dones"]),
                                              c(inp["time_limit_dones"]))
        seg_adv = agent.get_segment_advantage(c(inp["rewards"]), c(inp["values"]), adv, pairs)
    dataset = dict(segment_state=c(obs), step_actions=smp, segment_log_prob_estimate=lp_old,
                   segment_params_mean=mean_old, segment_params_L=L_old, segment_advantage=seg_adv,
                   segment_init_time=c(init_time), segment_init_pos=c(inp["init_pos"]),
                   segment_init_vel=c(inp["init_vel"]))
    layer.initial_entropy = policy.entropy([mean_old, L_old]).mean()
    params0 = [p.detach().clone() for p in policy.parameters]
    metrics = agent.policy_epoch(dataset, times, pairs).cpu()
    grads = [p.grad.detach().double().cpu() for p in policy.parameters]
    assert (agent._fast is not None) == fast

    # --- oracle side: same parameters / data in fp64 -------------------------------------------------
    import copy
    mean_net = copy.deepcopy(policy.mean_net).cpu().double()
    var_net = copy.deepcopy(policy.variance_net).cpu().double() if contextual else None
    with torch.no_grad():
        for p, p0 in zip(list(mean_net.parameters()) + (list(var_net.parameters()) if contextual else []), params0):
            p.copy_(p0.double().cpu())
    cov_vec = None if contextual else params0[-1].double().cpu().requires_grad_(True)
    opolicy = opol.TemporalCorrelatedPolicy(Dp, mp=dict(type="prodmp", args=dict(cfg, dtype=torch.float64)),
                                            mean_net=mean_net, variance_net=var_net, cov_vector=cov_vec,
                                            contextual=contextual, min_std=1e-4)
    olayer = oproj.projection_factory(typ, dtype=torch.float64, **layer_kwargs(typ, mb, cb, Dp))
    olayer.initial_entropy = f64(layer.initial_entropy)
    odata = {k: (f64(v) if v.is_floating_point() else v.cpu()) for k, v in dataset.items()}
    loss, parts = oa.policy_epoch(opolicy, olayer, odata, f64(times), pairs.cpu(), 0, set_variance=False,
                                  with_metrics=True)
    oparams = list(mean_net.parameters()) + (list(var_net.parameters()) if contextual else [cov_vec])
    ograds = torch.autograd.grad(loss, oparams)
    assert abs(metrics[0].item() - parts["surrogate_loss"].item()) <= 1e-4
    assert abs(metrics[2].item() - parts["trust_region_loss"].item()) <= 1e-4
    assert abs(metrics[3].item() - loss.item()) <= 2e-4
    assert abs(metrics[4].item() - parts["entropy"].item()) <= 1e-4
    assert abs(metrics[5].item() - parts["imp_smp_ratio"].item()) <= 1e-4
    from tce_rl_b200.rl.agent import _KL_KEYS
    for i, key in enumerate(_KL_KEYS):                           # the 12 logging means of :641-686
        assert abs(metrics[7 + i].item() - parts["kl"][key].item()) <= 1e-4, key
    for g, og in zip(grads, ograds):
        assert (g - og).abs().max() <= 1e-3 * max(1e-3, og.abs().max().item())
    # seg-advantage and old log-probs that fed the epoch agree with the oracle too
    o_adv, _ = oa.get_advantage_return(inp["rewards"].double(), inp["values"].double(), inp["dones"],
                                       inp["time_limit_dones"], 1.0, 0.95)
    o_seg = oa.get_segment_advantage(inp["rewards"].double(), inp["values"].double(), o_adv, pairs.cpu(), 1.0,
                                     "value_subtraction", True)
    assert (f64(seg_adv) - o_seg).abs().max() <= 1e-4


def test_kl_projection_warm_start_is_exact():
    """Warm-starting the eigen-solve from the previous call's basis must not change the projection;
    a different old covariance must fall back to a cold start (fingerprint guard)."""
    n, Bc = 63, 3
    inp = case("box", Bc, True, seed=5)
    c = lambda k: inp[k].to(DEV)
    L, Lo = c("L"), c("L_old")
    cold = ops.proj_kl_cov(L, Lo, 5e-4, ops.kl_state(Bc, n, DEV), False)[0]
    state = ops.kl_state(Bc, n, DEV)
    ops.proj_kl_cov(L * 1.001, Lo, 5e-4, state, True)                 # previous "epoch"
    warm = ops.proj_kl_cov(L, Lo, 5e-4, state, True)[0]
    assert (warm - cold).abs().max().item() <= 2e-6
    other = ops.proj_kl_cov(L, Lo * 1.01, 5e-4, state, True)[0]       # state belongs to another L_old
    ref = ops.proj_kl_cov(L, Lo * 1.01, 5e-4, ops.kl_state(Bc, n, DEV), False)[0]
    assert (other - ref).abs().max().item() <= 2e-6
    # gradients through a warm-started forward
    Lg = L.clone().requires_grad_(True)
    W = torch.tril(torch.randn_like(L))
    (ops.proj_kl_cov(Lg, Lo, 5e-4, state, True)[0] * W).sum().backward()
    Lg2 = L.clone().requires_grad_(True)
    (ops.proj_kl_cov(Lg2, Lo, 5e-4, ops.kl_state(Bc, n, DEV), False)[0] * W).sum().backward()
    assert (Lg.grad - Lg2.grad).abs().max().item() <= 1e-4 * Lg2.grad.abs().max().item()


@pytest.mark.parametrize("name", ["box", "table_tennis"])
def test_shared_covariance_maha_matches_per_episode_kernel(name):
    """tce_tri_inverse + tce_gauss_maha_shared (one L_o for all episodes) == tce_gauss_maha on B copies == fp64
    torch, values and gradient w.r.t. the mean (tolerance 1e-9 relative: everything is fp64 on fp32 inputs)."""
    inp = synthetic_inputs(name, 37, dtype=torch.float32)
    mean, mean_o = inp["mean"].to(DEV).requires_grad_(True), inp["mean_old"].to(DEV)
    L1 = inp["L_old"][:1].to(DEV)
    n = L1.shape[-1]
    Linv = ops.tri_inverse(L1)[0]
    ref_inv = torch.linalg.inv(L1[0].double())
    assert (Linv - ref_inv).abs().max() <= 1e-10 * ref_inv.abs().max()
    assert torch.equal(Linv, torch.tril(Linv))
    w = torch.linspace(0.5, 1.5, mean.shape[0], device=DEV, dtype=torch.float64)
    got = ops.gauss_maha_shared(mean, mean_o, Linv)
    (g_got,) = torch.autograd.grad((got * w).sum(), mean)
    ref = ops.gauss_maha(mean, mean_o, L1.expand(mean.shape[0], n, n).contiguous())
    (g_ref,) = torch.autograd.grad((ref * w).sum(), mean)
    z = torch.linalg.solve_triangular(L1[0].double(), (mean.detach().double() - mean_o.double()).T, upper=False)
    exact = (z * z).sum(0)
    assert (got - exact).abs().max() <= 1e-9 * exact.abs().max()
    assert (ref - exact).abs().max() <= 1e-9 * exact.abs().max()
    assert (g_got - g_ref).abs().max() <= 2e-6 * g_ref.abs().max()     # fp32 outputs
    g_kernel = ops.gauss_maha_shared_bwd(w, mean.detach(), mean_o, Linv)  # stand-alone gradient entry point
    assert (g_kernel - g_ref).abs().max() <= 2e-6 * g_ref.abs().max()


@pytest.mark.parametrize("beta_shift,equality", [(+0.5, False), (-0.5, False), (-0.5, True)])
def test_fused_kl_entropy_equals_the_two_separate_ops(beta_shift, equality):
    """tce_proj_kl_entropy_fwd/bwd == tce_proj_kl_cov_* followed by tce_proj_entropy_* (values bit-equal up to the
    fp32 store, gradients to 1e-6): entropy bound above (active), below (inactive) and the equality variant."""
    inp = synthetic_inputs("box", 3, dtype=torch.float32)
    L0, Lo = inp["L"].to(DEV), inp["L_old"].to(DEV)
    n = L0.shape[-1]
    H = 0.5 * n * (1.0 + 1.8378770664093453) + torch.diagonal(Lo, dim1=-2, dim2=-1).double().log().sum(-1)
    beta = (H.mean() + beta_shift).reshape(1).to(torch.float64)
    w = torch.linspace(0.5, 1.5, L0.numel(), device=DEV).reshape(L0.shape)
    res = []
    for fused in (True, False):
        L = L0.clone().requires_grad_(True)
        state = ops.kl_state(L.shape[0], n, DEV)
        if fused:
            out, proj, info = ops.proj_kl_entropy(L, Lo, 5e-4, state, False, beta, equality)
        else:
            proj, info = ops.proj_kl_cov(L, Lo, 5e-4, state, False)
            out = ops.proj_entropy(proj, beta, equality)[0]
        assert int(info.abs().max()) == 0
        (out * w).sum().backward()
        res.append((out.detach(), proj.detach(), L.grad))
    (o1, p1, g1), (o2, p2, g2) = res
    assert torch.equal(p1, p2)
    assert (o1 - o2).abs().max() <= 2e-7 * o2.abs().max()
    assert (g1 - g2).abs().max() <= 2e-6 * g2.abs().max()
    # the same forward as two launches (state with Sigma first, Cholesky second; alpha from the closed-form logdet)
    state = ops.kl_state(L0.shape[0], n, DEV)
    o3, p3, info3 = ops.proj_kl_entropy(L0.clone(), Lo, 5e-4, state, False, beta, equality, True)
    assert int(info3.abs().max()) == 0 and id(state) in ops.SIGMA_READY
    assert (p3 - p1).abs().max() <= 2e-7 * p1.abs().max() and (o3 - o1).abs().max() <= 2e-7 * o1.abs().max()
    Sigma, scale = ops.kl_state_sigma(state, L0.shape[0], n)
    ref = scale[:, None, None] * Sigma
    got = o3.double() @ o3.double().transpose(-1, -2)
    assert (got - ref).abs().max() <= 1e-6 * ref.abs().max()                  # fp32 rounding of the factor


@pytest.mark.parametrize("beta_shift,equality,kl_active", [(+0.5, False, True), (-0.5, False, True), (-0.5, True, True),
                                                         (+0.5, False, False), (-0.5, False, False)])
@pytest.mark.parametrize("split", [False, True])
def test_kl_backward_in_covariance_space_equals_factor_path(beta_shift, equality, kl_active, split):
    """tce_proj_kl_bwd_sigma (gradient w.r.t. Sigma_out = alpha^2 Sigma0, as the likelihood returns it) == the factor
    path (gradient w.r.t. out_L through the Cholesky adjoint) for a loss that depends on the covariance only;
    KL step active / inactive (identity), entropy control active / inactive / equality."""
    inp = synthetic_inputs("box", 2, dtype=torch.float32)
    L0 = inp["L"].to(DEV)
    Lo = inp["L_old"].to(DEV) if kl_active else (L0 * 1.0001).contiguous()
    n = L0.shape[-1]
    H = 0.5 * n * (1.0 + 1.8378770664093453) + torch.diagonal(Lo, dim1=-2, dim2=-1).double().log().sum(-1)
    beta = (H.mean() + beta_shift).reshape(1).to(torch.float64)
    g = torch.Generator().manual_seed(1)
    W = torch.randn(L0.shape, generator=g, dtype=torch.float64)
    W = (W + W.transpose(-1, -2)).to(DEV)                                      # d loss / d Sigma_out, symmetric

    class Holder:
        pass
    res = []
    for path in ("factor", "sigma"):
        L = L0.clone().requires_grad_(True)
        state = ops.kl_state(L.shape[0], n, DEV)
        out, proj, info, sigma, scale, _inv = ops.proj_kl_entropy(L, Lo, 5e-4, state, False, beta, equality, split, Holder(),
                                                            return_sigma=True)
        active = ops.kl_state(L.shape[0], n, DEV)                              # (placeholder to keep allocator busy)
        del active
        if path == "factor":
            S = out.double() @ out.double().transpose(-1, -2)
            (S * W).sum().backward()
        else:
            assert sigma.requires_grad and sigma.shape == L.shape and sigma.dtype == torch.float64
            (sigma * W).sum().backward()               # contract: the gradient of `sigma` IS d loss / d Sigma_out
        sc = ops.kl_state_scalars(state, L.shape[0], n)
        assert bool((sc[:, 1] != 0).all()) == kl_active
        res.append(L.grad.clone())
    gf, gs = res
    assert (gf - gs).abs().max() <= 2e-5 * gf.abs().max()


def test_kl_state_generation_guard():
    """A second forward of a warm-started layer before the first one's backward must not silently use the
    overwritten state: backward raises."""
    inp = synthetic_inputs("box", 1, dtype=torch.float32)
    L0, Lo = inp["L"].to(DEV), inp["L_old"].to(DEV)
    n = L0.shape[-1]
    beta = torch.zeros(1, device=DEV, dtype=torch.float64)

    class Holder:
        pass
    h = Holder()
    state = ops.kl_state(1, n, DEV)
    La = L0.clone().requires_grad_(True)
    out_a = ops.proj_kl_entropy(La, Lo, 5e-4, state, True, beta, False, False, h)[0]
    Lb = (L0 * 1.01).requires_grad_(True)
    out_b = ops.proj_kl_entropy(Lb, Lo, 5e-4, state, True, beta, False, False, h)[0]
    out_b.sum().backward()                                                     # the latest forward: fine
    with pytest.raises(RuntimeError, match="overwrote its state"):
        out_a.sum().backward()


def test_kl_projection_large_batch_equals_small_batches():
    """More matrices than SMs take the two-CTAs-per-SM variant of the KL forward (three shared-memory buffers, 256
    threads; csrc/tce_proj.cu `compact`): projected factors, the saved state and the gradients equal those of the same
    matrices projected in chunks that take the one-CTA-per-SM variant -- cold, warm-started, and with one identity step."""
    B, name, Dp = 200, "box", 63
    inp = case(name, B, True, seed=5)
    inp["L"][3] = inp["L_old"][3] * (1 + 1e-4)                  # inside the trust region: identity branch
    c = lambda k: inp[k].to(DEV).contiguous()
    L, L_o = c("L"), c("L_old")
    beta = torch.full((1,), -1e9, device=DEV, dtype=torch.float64)          # entropy control inactive
    g = torch.tril(torch.randn(B, Dp, Dp, generator=torch.Generator().manual_seed(1))).to(DEV)

    def run(lo, hi, state, warm):
        Lr = L[lo:hi].clone().requires_grad_(True)
        out, proj, info = ops.proj_kl_entropy(Lr, L_o[lo:hi], 5e-4, state, warm, beta, False)
        assert int(info.abs().max()) == 0
        (out * g[lo:hi]).sum().backward()
        return out.detach(), Lr.grad

    big_state = ops.kl_state(B, Dp, DEV)
    chunk_states = [ops.kl_state(100, Dp, DEV) for _ in range(2)]
    for warm in (False, True, True):                            # the second / third call start from the saved eigen-basis
        out_big, grad_big = run(0, B, big_state, warm)
        outs, grads = zip(*[run(100 * k, 100 * (k + 1), chunk_states[k], warm) for k in range(2)])
        out_small, grad_small = torch.cat(outs), torch.cat(grads)
        assert (out_big - out_small).abs().max().item() <= 2e-6 * out_small.abs().max().item()
        assert (grad_big - grad_small).abs().max().item() <= 1e-5 * grad_small.abs().max().item()
        sc_big = ops.kl_state_scalars(big_state, B, Dp)
        sc_small = torch.cat([ops.kl_state_scalars(s, 100, Dp) for s in chunk_states])
        assert (sc_big[:, :3] - sc_small[:, :3]).abs().max().item() <= 1e-9 * max(1.0, sc_small[:, :3].abs().max().item())
    assert sc_big[3, 1].item() == 0.0 and sc_big[:, 1].sum().item() >= B - 20       # identity step seen, most are active
