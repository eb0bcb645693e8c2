"""The stateful ``mp_pytorch.mp.ProDMP`` surface (``tce_rl_b200.mp.ProDMP``) against the CPU oracle, called the way the
reference calls it (mprl/rl/policy/temporal_correlated_policy.py:76-92 sample_trajectories / get_traj_pos / get_traj_vel,
:152-192 update_inputs -> get_traj_pos(flat_shape=True) / get_traj_pos_cov), the out-of-range error of the pre-computed
tables, and the third segment-advantage mode (temporal_correlated_agent.py:288-319)."""
import pytest
import torch

from oracle import agent as oa
from oracle import prodmp as oprodmp
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from tce_rl_b200 import mp as gmp, ops

DEV = "cuda:0"


def f64(t):
    return t.detach().double().cpu()


def _case(name, B):
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    inp = synthetic_inputs(name, B, dtype=torch.float32)
    times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float()
    shim = gmp.get_mp(type="prodmp", args=dict(cfg))
    ref = oprodmp.get_mp(type="prodmp", args=dict(cfg, dtype=torch.float64))
    return cfg, T, inp, times, shim, ref


@pytest.mark.parametrize("name", ["box", "table_tennis", "metaworld"])
def test_update_inputs_then_getters(name):
    B = 5
    cfg, T, inp, times, shim, ref = _case(name, B)
    c = lambda t: t.to(DEV)
    d = lambda t: t.double()
    shim.update_inputs(times=c(times), params=c(inp["mean"]), params_L=None, init_time=c(inp["init_time"]),
                       init_pos=c(inp["init_pos"]), init_vel=c(inp["init_vel"]))
    ref.update_inputs(times=d(times), params=d(inp["mean"]), params_L=None, init_time=d(inp["init_time"]),
                      init_pos=d(inp["init_pos"]), init_vel=d(inp["init_vel"]))
    for flat in (False, True):
        for getter in ("get_traj_pos", "get_traj_vel"):
            got, want = getattr(shim, getter)(flat_shape=flat), getattr(ref, getter)(flat_shape=flat)
            assert tuple(got.shape) == tuple(want.shape), (getter, flat)
            assert (f64(got) - want).abs().max() <= 1e-5 * want.abs().max(), (getter, flat)
    # the getters accept new inputs like mp_pytorch (update + evaluate); the cached trajectory is invalidated
    got = shim.get_traj_pos(params=c(inp["mean"] * 0.5), flat_shape=True)
    want = ref.get_traj_pos(params=d(inp["mean"] * 0.5), flat_shape=True)
    assert (f64(got) - want).abs().max() <= 1e-5 * want.abs().max()


@pytest.mark.parametrize("name", ["box", "table_tennis"])
def test_pos_cov_of_time_pairs_like_the_reference_calls_it(name):
    """temporal_correlated_policy.py:152-192: everything is expanded to [B, P, ...] and the times are the pairs."""
    B = 4
    cfg, T, inp, times, shim, ref = _case(name, B)
    torch.manual_seed(0)
    pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
    P = pairs.shape[0]
    Dp = inp["mean"].shape[-1]
    ex = lambda v: ou.add_expand_dim(v, [1], [P])
    tp = times[:, pairs]                                                    # [B, P, 2]
    c = lambda t: t.to(DEV).contiguous()
    d = lambda t: t.double()
    args = dict(times=tp, params=ex(inp["mean"]), params_L=ex(inp["L"]), init_time=ex(inp["init_time"]),
                init_pos=ex(inp["init_pos"]), init_vel=ex(inp["init_vel"]))
    shim.update_inputs(**{k: c(v) for k, v in args.items()})
    ref.update_inputs(**{k: d(v) for k, v in args.items()})
    mean_g, mean_w = shim.get_traj_pos(flat_shape=True), ref.get_traj_pos(flat_shape=True)
    assert tuple(mean_g.shape) == (B, P, 2 * cfg["num_dof"])
    assert (f64(mean_g) - mean_w).abs().max() <= 1e-5 * mean_w.abs().max()
    cov_g, cov_w = shim.get_traj_pos_cov(), ref.get_traj_pos_cov()
    assert tuple(cov_g.shape) == tuple(cov_w.shape) == (B, P, 2 * cfg["num_dof"], 2 * cfg["num_dof"])
    assert (f64(cov_g) - cov_w).abs().max() <= 2e-5 * cov_w.abs().max()
    # SPD with the batch-global regulariser: torch's MVN accepts it
    torch.linalg.cholesky(f64(cov_g))


def test_sample_trajectories_with_injected_noise_and_state_restore():
    B, S = 3, 2
    cfg, T, inp, times, shim, ref = _case("box", B)
    c = lambda t: t.to(DEV)
    d = lambda t: t.double()
    eps = torch.randn(S, B, inp["mean"].shape[-1], generator=torch.Generator().manual_seed(3))
    keep = c(times[:, :7].contiguous())
    shim.update_inputs(times=keep, params=c(inp["mean"]), init_time=c(inp["init_time"]), init_pos=c(inp["init_pos"]),
                       init_vel=c(inp["init_vel"]))
    pos_g, vel_g = shim.sample_trajectories(times=c(times), params=c(inp["mean"]), params_L=c(inp["L"]),
                                            init_time=c(inp["init_time"]), init_pos=c(inp["init_pos"]),
                                            init_vel=c(inp["init_vel"]), num_smp=S, eps=c(eps))
    pos_w, vel_w = ref.sample_trajectories(times=d(times), params=d(inp["mean"]), params_L=d(inp["L"]),
                                           init_time=d(inp["init_time"]), init_pos=d(inp["init_pos"]),
                                           init_vel=d(inp["init_vel"]), num_smp=S, eps=d(eps))
    assert tuple(pos_g.shape) == tuple(pos_w.shape) == (B, S, T, cfg["num_dof"])
    assert (f64(pos_g) - pos_w).abs().max() <= 1e-5 * pos_w.abs().max()
    assert (f64(vel_g) - vel_w).abs().max() <= 1e-5 * vel_w.abs().max()
    assert shim.times is keep                                   # the previous inputs are restored (mp_pytorch behaviour)
    # without injected noise: Philox draws, reproducible for a fixed seed, different across samples
    a1, _ = shim.sample_trajectories(times=c(times), params=c(inp["mean"]), params_L=c(inp["L"]),
                                     init_time=c(inp["init_time"]), init_pos=c(inp["init_pos"]),
                                     init_vel=c(inp["init_vel"]), num_smp=2, seed=11)
    a2, _ = shim.sample_trajectories(times=c(times), params=c(inp["mean"]), params_L=c(inp["L"]),
                                     init_time=c(inp["init_time"]), init_pos=c(inp["init_pos"]),
                                     init_vel=c(inp["init_vel"]), num_smp=2, seed=11)
    assert torch.equal(a1, a2) and not torch.equal(a1[:, 0], a1[:, 1])


def test_time_beyond_the_precomputed_range_raises_like_mp_pytorch():
    cfg, T, inp, times, shim, ref = _case("box", 2)
    late = times.clone()
    late[1, -1] = cfg["tau"] * 5 + 0.5                          # factor = 5 periods (util_mp.py:33)
    with pytest.raises(RuntimeError, match="pre-computation range"):
        ref.update_inputs(times=late.double(), params=inp["mean"].double(), init_time=inp["init_time"].double(),
                          init_pos=inp["init_pos"].double(), init_vel=inp["init_vel"].double())
        ref.get_traj_pos()
    with pytest.raises(RuntimeError, match="pre-computation range"):
        shim.update_inputs(times=late.to(DEV))
    shim.strict_range = False                                   # lazy: the flag is read when the caller asks
    shim.update_inputs(times=late.to(DEV))
    shim.update_inputs(times=times.to(DEV))                     # a later in-range call does not clear the flag
    with pytest.raises(RuntimeError, match="pre-computation range"):
        shim.check_range()
    shim.check_range()                                          # cleared by the read
    # exactly at the end of the range is allowed
    edge = times.clone()
    edge[0, -1] = cfg["tau"] * 5
    shim.strict_range = True
    shim.update_inputs(times=edge.to(DEV))


@pytest.mark.parametrize("B,T,gamma", [(7, 100, 1.0), (33, 60, 0.98)])
def test_segment_advantage_accumulated_rewards_mode(B, T, gamma):
    """Mode 2 of get_segment_advantage (temporal_correlated_agent.py:288-319) through the agent mirror."""
    from tce_rl_b200.rl.agent import TemporalCorrelatedAgent

    g = torch.Generator().manual_seed(9)
    rewards, values = torch.randn(B, T, generator=g), torch.randn(B, T + 1, generator=g)
    torch.manual_seed(2)
    pairs = ou.get_time_pairs(T, dict(num_select=12, fixed_interval=True))
    adv = torch.randn(B, T, generator=g)
    want = oa.get_segment_advantage(rewards.double(), values.double(), adv.double(), pairs, gamma,
                                    "accumulated_rewards", False)
    agent = TemporalCorrelatedAgent.__new__(TemporalCorrelatedAgent)
    agent.segment_advantage, agent.norm_advantages, agent.clip_advantages = "accumulated_rewards", False, False
    agent._gamma, agent.process_group = gamma, None
    c = lambda t: t.to(DEV)
    got = agent.get_segment_advantage(c(rewards), c(values), c(adv), c(pairs))
    assert tuple(got.shape) == tuple(want.shape)
    assert (f64(got) - want).abs().max() <= 1e-5 * max(1.0, want.abs().max().item())
    raw = ops.segment_advantage(2, c(rewards), c(values), c(adv), c(pairs), gamma, False)
    assert (f64(raw - raw.mean(dim=0)) - want).abs().max() <= 1e-5 * max(1.0, want.abs().max().item())
