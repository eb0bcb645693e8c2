"""BlackBoxAgent (the BBRL baseline, mprl/rl/agent/black_box_agent.py) on the GPU kernels: one policy epoch against the
oracle restatement (loss terms, logging KLs, parameter gradients), then the agent API (process_dataset, update_critic,
update_policy with the balance check and projection.compute_metrics)."""
import copy

import pytest
import torch

from oracle import agent as oa
from oracle import policy as opol
from oracle import projection as oproj
from oracle.gen_golden import synthetic_inputs

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from tce_rl_b200.rl import agent_factory, critic_factory, policy_factory, projection_factory
    from tce_rl_b200.rl.agent import _KL_KEYS

DEV = "cuda:0"
PROJ_KW = dict(proj_type="kl", mean_bound=0.05, cov_bound=5e-4, trust_region_coeff=10.0, scale_prec=True,
               entropy_schedule="linear", total_train_steps=100, target_entropy=0.0, temperature=0.7,
               entropy_eq=False, entropy_first=False, do_regression=False)


def f64(t):
    return t.detach().double().cpu()


def build(B=48, Dp=63, obs_dim=9, typ="KLProjectionLayer", **agent_kw):
    torch.manual_seed(0)
    policy = policy_factory("BlackBoxPolicy", dim_in=obs_dim, dim_out=Dp,
                            mean_net_args=dict(avg_neuron=32, num_hidden=2, shape=0.0),
                            variance_net_args=dict(std_only=False, contextual=False), init_method="orthogonal",
                            out_layer_gain=0.01, act_func_hidden="leaky_relu", act_func_last=None, dtype="float32",
                            device=DEV, min_std=1e-4)
    critic = critic_factory("ValueFunction", dim_in=obs_dim, dim_out=1, hidden=dict(avg_neuron=32, num_hidden=2, shape=0.0),
                            init_method="orthogonal", out_layer_gain=1, act_func_hidden="leaky_relu",
                            act_func_last=None, dtype="float32", device=DEV)
    layer = projection_factory(typ, device=DEV, dtype="float32", action_dim=Dp, **PROJ_KW)
    kw = dict(lr_policy=3e-4, lr_critic=1e-3, wd_policy=5e-5, wd_critic=5e-5, epochs_policy=3, epochs_critic=2,
              num_minibatchs=2, norm_advantages=True, set_variance=False, balance_check=10)
    kw.update(agent_kw)
    agent = agent_factory("BlackBoxAgent", policy=policy, critic=critic, sampler=None, projection=layer, dtype="float32",
                          device=DEV, **kw)
    inp = synthetic_inputs("box", B, seed=4, dtype=torch.float32)
    g = torch.Generator().manual_seed(6)
    obs = torch.randn(B, obs_dim, generator=g).to(DEV)
    with torch.no_grad():
        mean_old, L_old = policy.policy(obs)
        mean_old = (mean_old + 0.05 * torch.randn(B, Dp, generator=g).to(DEV)).contiguous()
        L_old = (1.03 * L_old[:1] + 0.01 * torch.tril(torch.randn(Dp, Dp, generator=g), -1).to(DEV)).expand(B, -1, -1) \
            .contiguous()
        actions = policy.sample(False, mean_old, L_old, eps=inp["eps"].to(DEV))
        lp_old = policy.log_prob(actions, mean_old, L_old)
        for p in policy.parameters:
            p.add_(0.02 * torch.randn(p.shape, generator=g).to(DEV))
    dataset = dict(segment_state=obs, segment_action=actions, segment_log_prob=lp_old, segment_params_mean=mean_old,
                   segment_params_L=L_old, segment_reward=torch.randn(B, generator=g).to(DEV),
                   segment_value=torch.randn(B, generator=g).to(DEV))
    return agent, dataset


def test_bbrl_epoch_matches_oracle():
    agent, dataset = build()
    policy, layer = agent.policy, agent.projection
    dataset = agent.process_dataset(dataset)
    odata = {k: f64(v) for k, v in dataset.items()}
    assert (f64(dataset["segment_advantage"]) - oa.bbrl_process_dataset(odata)).abs().max() <= 1e-5
    layer.initial_entropy = policy.entropy([dataset["segment_params_mean"], dataset["segment_params_L"]]).mean()
    params0 = [p.detach().clone() for p in policy.parameters]
    metrics = agent.policy_epoch(dataset, None, None).cpu()
    grads = [p.grad.detach().double().cpu() for p in policy.parameters]
    # oracle
    mean_net = copy.deepcopy(policy.mean_net).cpu().double()
    with torch.no_grad():
        for p, p0 in zip(mean_net.parameters(), params0):
            p.copy_(p0.double().cpu())
    cov_vec = params0[-1].double().cpu().requires_grad_(True)
    opolicy = opol.BlackBoxPolicy(63, mean_net=mean_net, cov_vector=cov_vec, contextual=False, min_std=1e-4)
    olayer = oproj.projection_factory("KLProjectionLayer", dtype=torch.float64, action_dim=63, **PROJ_KW)
    olayer.initial_entropy = f64(layer.initial_entropy)
    loss, parts = oa.policy_epoch_bbrl(opolicy, olayer, odata, 0, set_variance=False, with_metrics=True)
    ograds = torch.autograd.grad(loss, list(mean_net.parameters()) + [cov_vec])
    assert abs(metrics[0].item() - parts["surrogate_loss"].item()) <= 1e-4
    assert abs(metrics[2].item() - parts["trust_region_loss"].item()) <= 1e-4 * max(1.0, abs(parts["trust_region_loss"].item()))
    assert abs(metrics[4].item() - parts["entropy"].item()) <= 1e-4
    for i, key in enumerate(_KL_KEYS):
        assert abs(metrics[7 + i].item() - parts["kl"][key].item()) <= 1e-4, key
    for g, og in zip(grads, ograds):
        assert (g - og).abs().max() <= 1e-3 * max(1e-3, og.abs().max().item())


@pytest.mark.parametrize("use_graph", [False, True])
def test_bbrl_agent_api(use_graph):
    agent, dataset = build(use_cuda_graph=use_graph)
    agent.num_iterations = 1                               # 1 % balance_check == 1: the balance check runs
    dataset = agent.process_dataset(dataset)
    c0 = [p.detach().clone() for p in agent.critic.parameters]
    out_c = agent.update_critic(dataset)
    assert out_c["critic_loss_mean"] > 0 and any((p - q).abs().max() > 0 for p, q in zip(agent.critic.parameters, c0))
    p0 = [p.detach().clone() for p in agent.policy.parameters]
    out = agent.update_policy(dataset)
    assert any((p - q).abs().max() > 0 for p, q in zip(agent.policy.parameters, p0))
    for key in ("surrogate_loss_mean", "trust_region_loss_mean", "policy_grad_norm_mean", "clipped_policy_grad_norm_mean",
                "projection_kl", "projection_constraint_max", "projection_entropy", "projection_new_old_cov_diff_mean",
                "surrogate_grad_norm_mean", "trust_region_grad_norm_mean", "balance_ratio"):
        assert key in out, key
    assert out["balance_ratio"] > 0
    assert out["projection_proj_old_mean_diff_max"] <= 0.05 * (1 + 1e-3)
    agent.num_iterations = 2                               # no balance check: graph replay when requested
    out2 = agent.update_policy(dataset)
    assert "balance_ratio" not in out2 and out2["policy_loss_mean"] == out2["policy_loss_mean"]


def test_tce_balance_check_keys():
    """The TCE agent's balance check (temporal_correlated_agent.py:446-522, 601-612)."""
    from test_gpu_agent import build as build_tce
    agent, dataset = build_tce(epochs=2)
    agent.balance_check = 5
    agent.num_iterations = 6                               # 6 % 5 == 1
    dataset = agent.process_dataset(dataset)
    out = agent.update_policy(dataset)
    assert out["surrogate_grad_norm_mean"] > 0 and out["trust_region_grad_norm_mean"] >= 0
    assert abs(out["balance_ratio"] - out["surrogate_grad_norm_mean"] / out["trust_region_grad_norm_mean"]) < 1e-9
