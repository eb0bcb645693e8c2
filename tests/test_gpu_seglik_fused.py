"""Fused segment likelihood (csrc/tce_seglik_fused.cu) against the CPU oracle and against the independent staged
kernels (csrc/tce_seglik.cu): general and uniform-grid paths, shared / per-episode covariance, chained / ragged
pairs, all three gradient modes, and the BASELINE sizes (B = 1024 x P in {24, 25}, metaworld B = 4096).

Tolerances (north star): log-probs <= 1e-4 absolute; gradients <= 2e-4 of their scale (fp32 outputs)."""
import pytest
import torch

from oracle import policy as opol  # noqa: F401
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from test_gpu_kernels import f64, make_oracle_policy, setup_case

pytestmark = pytest.mark.gpu

CUDA = torch.cuda.is_available()
if CUDA:
    from tce_rl_b200 import ops

DEV = "cuda:0"


@pytest.fixture(autouse=True, params=[False, True], ids=["fused", "routed"])
def per_episode_route(request):
    """Per-episode covariance factors are routed to the staged kernels by default (ops_seglik.PER_EPISODE_STAGED:
    they are faster for that layout); every test here runs both ways so the fused kernel keeps its coverage."""
    from tce_rl_b200 import ops_seglik
    old = ops_seglik.PER_EPISODE_STAGED
    ops_seglik.PER_EPISODE_STAGED = request.param
    yield request.param
    ops_seglik.PER_EPISODE_STAGED = old


def literal_pairs(T, P):
    """The synthetic P-pair index set of SURVEY 8(d) config 2 ("25 segments": {0, 4, ..., 96, 99})."""
    step = T // (P + 1) if P + 1 <= T else 1
    idx = list(range(0, T, max(T // P, 1)))[:P] + [T - 1]
    idx = sorted(set(idx))[:P + 1]
    t = torch.tensor(idx, dtype=torch.long)
    return torch.stack([t[:-1], t[1:]], dim=1)


def device_case(name, B, pairs=None, init_time_spread=0.0, P_select=25):
    cfg, T, inp, times, pr = setup_case(name, B, init_time_spread=init_time_spread, P_select=P_select)
    pairs = pr if pairs is None else pairs
    tabs = ops.Tables(**cfg)
    c = lambda t: t.to(DEV)
    theta = ops.mvn_rsample(c(inp["mean"]), c(inp["L"]), c(inp["eps"]), 0, 0)
    smp = ops.prodmp_traj(theta, c(times), c(inp["init_time"]), c(inp["init_pos"]), c(inp["init_vel"]), tabs.handle,
                          cfg["num_dof"])
    return cfg, T, inp, times, pairs, tabs, smp


@pytest.mark.parametrize("name,B,shared,spread", [
    ("box", 1024, False, 0.0), ("box", 1024, True, 0.0), ("box", 1024, True, 0.3), ("box", 37, False, 0.3),
    ("metaworld", 256, True, 0.0), ("metaworld", 100, False, 0.2), ("table_tennis", 129, True, 0.1),
    ("table_tennis", 64, False, 0.0)])
def test_fused_equals_staged(name, B, shared, spread):
    """Same inputs through the fused and through the staged kernels: log-probs, regulariser, both gradients."""
    cfg, T, inp, times, pairs, tabs, smp = device_case(name, B, init_time_spread=spread)
    c = lambda t: t.to(DEV)
    w = torch.linspace(0.5, 1.5, B * pairs.shape[0], device=DEV).reshape(B, -1)
    res = []
    for fn in (ops.seg_logprob, ops.seg_logprob_staged):
        mean = c(inp["mean"]).requires_grad_(True)
        L0 = c(inp["L"][:1] if shared else inp["L"]).requires_grad_(True)
        L = L0.expand(B, -1, -1) if shared else L0
        lp, info, dmax = fn(smp, mean, L, c(times), c(inp["init_time"]), c(inp["init_pos"]), c(inp["init_vel"]),
                            c(pairs), tabs, return_info=True)
        (lp * w).sum().backward()
        res.append((lp.detach(), dmax.detach().reshape(-1)[0], mean.grad, L0.grad))
        assert int(info.abs().max()) == 0
    (lp0, d0, gm0, gL0), (lp1, d1, gm1, gL1) = res
    assert abs(d0.item() - d1.item()) <= 1e-6 * abs(d1.item())
    assert (lp0.double() - lp1.double()).abs().max() <= 2e-5
    assert (gm0 - gm1).abs().max() <= 2e-5 * gm1.abs().max()
    assert gL0.shape == gL1.shape
    assert (gL0 - gL1).abs().max() <= 5e-5 * gL1.abs().max()
    assert float(gL0.triu(1).abs().max()) == 0.0


@pytest.mark.parametrize("name,B,P_select", [("box", 1024, 25), ("box", 1024, "literal25"), ("metaworld", 4096, 25),
                                             ("table_tennis", 1024, 25)])
@pytest.mark.parametrize("contextual", [False, True])
def test_likelihood_at_baseline_sizes_vs_oracle(name, B, P_select, contextual):
    """BASELINE config sizes against the fp64 oracle: log-probs <= 1e-4 abs, gradients <= 2e-4 of their scale.
    Non-contextual: ONE factor (the uniform-grid path is taken automatically: init_time = 0 for all episodes)."""
    if name == "metaworld" and contextual:
        B = 1024                                           # the oracle materialises [B, P, Dp, Dp]
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    pairs = literal_pairs(T, 25) if P_select == "literal25" else None
    cfg, T, inp, times, pairs, tabs, smp = device_case(name, B, pairs=pairs)
    if P_select == "literal25":
        assert pairs.shape[0] == 25 and pairs[-1, 1] == T - 1
    pol = make_oracle_policy(name)
    d = lambda k: inp[k].double()
    args64 = (times.double(), d("init_time"), d("init_pos"), d("init_vel"))
    mean64 = d("mean").requires_grad_(True)
    L64 = (d("L") if contextual else d("L")[:1]).clone().requires_grad_(True)
    P = pairs.shape[0]
    w64 = torch.linspace(0.5, 1.5, B * P, dtype=torch.float64).reshape(B, P)
    want = torch.empty(B, P, dtype=torch.float64)
    gm = torch.zeros_like(mean64)
    gL = torch.zeros_like(L64)
    # the regulariser is batch global: evaluate it once on the whole batch, then chunk the oracle
    smp64 = smp.double().cpu()
    with torch.no_grad():
        reg = None
        for s in range(0, B, 256):
            sl = slice(s, s + 256)
            Ls = L64[sl] if contextual else L64.expand(B, -1, -1)[sl]
            _, _, cov, _ = pol.log_prob(smp64[sl], mean64[sl], Ls, *[a[sl] for a in args64], pred_pairs=pairs,
                                        return_parts=True, reg_override=1.0)    # (+1: table tennis has C = 0
            m = torch.diagonal(cov, dim1=-2, dim2=-1).max() - 1.0                # before its delay -- not PD)
            reg = m if reg is None else torch.maximum(reg, m)
        reg = float(reg) * 1e-4
    for s in range(0, B, 256):
        sl = slice(s, s + 256)
        Ls = L64[sl] if contextual else L64.expand(B, -1, -1)[sl]
        lp = pol.log_prob(smp64[sl], mean64[sl], Ls, *[a[sl] for a in args64], pred_pairs=pairs, reg_override=reg)
        want[sl] = lp.detach()
        g1, g2 = torch.autograd.grad((lp * w64[sl]).sum(), [mean64, L64])
        gm += g1
        gL += g2
    c = lambda t: t.to(DEV)
    mean = c(inp["mean"]).requires_grad_(True)
    L0 = c(inp["L"] if contextual else inp["L"][:1]).requires_grad_(True)
    L = L0 if contextual else L0.expand(B, -1, -1)
    lp, info, dmax = ops.seg_logprob(smp, mean, L, c(times), c(inp["init_time"]), c(inp["init_pos"]),
                                     c(inp["init_vel"]), c(pairs), tabs, return_info=True)
    assert abs(dmax.reshape(-1)[0].item() * 1e-4 - reg) <= 1e-6 * reg
    assert (f64(lp) - want).abs().max() <= 1e-4
    (lp * c(w64.float())).sum().backward()
    gLt = torch.tril(gL)
    assert (f64(mean.grad) - gm).abs().max() <= 2e-4 * gm.abs().max()
    assert (f64(L0.grad) - gLt).abs().max() <= 2e-4 * gLt.abs().max()


@pytest.mark.parametrize("name,B,shared,spread,ragged", [
    ("box", 1024, True, 0.0, False), ("box", 1024, False, 0.0, False), ("box", 200, True, 0.25, False),
    ("box", 65, True, 0.0, True), ("box", 33, False, 0.2, True), ("metaworld", 512, True, 0.0, False),
    ("table_tennis", 300, True, 0.0, False)])
def test_fused_surrogate_vs_oracle(name, B, shared, spread, ragged):
    """Fused likelihood + surrogate (forward produces the gradients) against oracle autograd; uniform path when the
    grid is uniform and the covariance shared, general path otherwise; ragged = non-chained, repeated pairs."""
    T = NUM_TIMES[name]
    pairs = torch.tensor([[0, 1], [0, T - 1], [5, 50], [T - 2, T - 1], [10, 11], [11, 12], [40, 80]]) if ragged else None
    cfg, T, inp, times, pairs, tabs, smp = device_case(name, B, pairs=pairs, init_time_spread=spread)
    P = pairs.shape[0]
    pol = make_oracle_policy(name)
    d = lambda k: inp[k].double()
    args64 = (times.double(), d("init_time"), d("init_pos"), d("init_vel"))
    g = torch.Generator().manual_seed(3)
    adv = torch.randn(B, P, generator=g)
    mean64 = d("mean").requires_grad_(True)
    L64 = (d("L")[:1] if shared else d("L")).clone().requires_grad_(True)
    smp64 = smp.double().cpu()
    with torch.no_grad():
        lp0 = pol.log_prob(smp64, d("mean"), (d("L")[:1].expand(B, -1, -1) if shared else d("L")), *args64,
                           pred_pairs=pairs)
    lp_old = (lp0 + 0.1 * torch.randn(B, P, generator=g).double()).float()
    lp = pol.log_prob(smp64, mean64, L64.expand(B, -1, -1) if shared else L64, *args64, pred_pairs=pairs)
    ratio = (lp - lp_old.double()).exp()
    loss64 = -(ratio * adv.double()).mean()
    gm, gL = torch.autograd.grad(loss64, [mean64, L64])
    c = lambda t: t.to(DEV)
    mean = c(inp["mean"]).requires_grad_(True)
    L0 = c(inp["L"][:1] if shared else inp["L"]).requires_grad_(True)
    L = L0.expand(B, -1, -1) if shared else L0
    loss, rmean, lpg = ops.seg_surrogate(smp, mean, L, c(times), c(inp["init_time"]), c(inp["init_pos"]),
                                         c(inp["init_vel"]), c(pairs), c(lp_old), c(adv), tabs)
    assert (f64(lpg) - lp.detach()).abs().max() <= 1e-4
    assert abs(loss.item() - loss64.item()) <= 1e-4 * max(1.0, abs(loss64.item()))
    assert abs(rmean.item() - ratio.mean().item()) <= 1e-4
    (2.5 * loss).backward()                                  # a non-unit upstream gradient
    gLt = torch.tril(gL)
    assert (f64(mean.grad) / 2.5 - gm).abs().max() <= 3e-4 * gm.abs().max()
    assert (f64(L0.grad) / 2.5 - gLt).abs().max() <= 3e-4 * gLt.abs().max()


def test_uniform_path_equals_general_path():
    """Shared covariance + uniform grid: the uniform kernels and the general fused kernel agree."""
    name, B = "box", 777
    cfg, T, inp, times, pairs, tabs, smp = device_case(name, B)
    c = lambda t: t.to(DEV)
    P = pairs.shape[0]
    adv = torch.linspace(-1, 1, B * P, device=DEV).reshape(B, P)
    out = []
    for uniform in (True, False):
        mean = c(inp["mean"]).requires_grad_(True)
        L0 = c(inp["L"][:1]).requires_grad_(True)
        lp = ops.seg_logprob(smp, mean, L0, c(times), c(inp["init_time"]), c(inp["init_pos"]), c(inp["init_vel"]),
                             c(pairs), tabs, uniform=uniform)
        lp_old = lp.detach() - 0.05
        loss, ratio, lp2 = ops.seg_surrogate(smp, mean, L0, c(times), c(inp["init_time"]), c(inp["init_pos"]),
                                             c(inp["init_vel"]), c(pairs), lp_old, adv, tabs, uniform=uniform)
        loss.backward()
        out.append((lp.detach(), lp2, loss.detach(), ratio, mean.grad, L0.grad))
    a, b = out
    assert (a[0] - b[0]).abs().max() <= 2e-5 and (a[1] - b[1]).abs().max() <= 2e-5
    assert abs(a[2].item() - b[2].item()) <= 1e-6 and abs(a[3].item() - b[3].item()) <= 1e-6
    assert (a[4] - b[4]).abs().max() <= 2e-5 * b[4].abs().max()
    assert (a[5] - b[5]).abs().max() <= 1e-4 * b[5].abs().max()


def test_wrong_chained_claim_is_refused():
    """tce_seglik_fused with chained = 1 on pairs that are not a chain must not compute anything (info[0] = -7)."""
    name, B = "box", 8
    pairs = torch.tensor([[0, 5], [10, 20], [30, 31]])
    cfg, T, inp, times, pairs, tabs, smp = device_case(name, B, pairs=pairs)
    c = lambda t: t.to(DEV)
    out = ops.seglik(smp, c(inp["mean"]), c(inp["L"]), None, None, c(times), c(inp["init_time"]), c(inp["init_pos"]),
                     c(inp["init_vel"]), c(pairs), tabs.handle, 1e-4, 0, None, None, None, True, False, False)
    assert int(out[1].reshape(-1)[0]) == -7
