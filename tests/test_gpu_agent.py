"""The drop-in agent API on the GPU: process_dataset -> update_critic -> update_policy (eager and CUDA graph)."""
import copy

import numpy as np
import pytest
import torch

from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from tce_rl_b200.rl import (TemporalCorrelatedAgent, critic_factory, policy_factory, projection_factory)
    from tce_rl_b200.rl.agent import SegmentTimeSampler

DEV = "cuda:0"


def build(name="box", B=64, use_graph=False, epochs=4, seed=0):
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    D, K1 = cfg["num_dof"], cfg["num_basis"] + 1
    Dp, obs_dim = D * K1, 12
    torch.manual_seed(seed)
    policy = policy_factory("TemporalCorrelatedPolicy", dim_in=obs_dim, dim_out=Dp,
                            mean_net_args=dict(avg_neuron=32, num_hidden=2, shape=0.0),
                            variance_net_args=dict(std_only=False, contextual=False), init_method="orthogonal",
                            out_layer_gain=0.01, act_func_hidden="leaky_relu", act_func_last=None, dtype="float32",
                            device=DEV, min_std=1e-4, mp=dict(type="prodmp", args=dict(cfg)))
    critic = critic_factory("ValueFunction", dim_in=obs_dim, dim_out=1, hidden=dict(avg_neuron=32, num_hidden=2, shape=0.0),
                            init_method="orthogonal", out_layer_gain=1, act_func_hidden="leaky_relu",
                            act_func_last=None, dtype="float32", device=DEV)
    proj = projection_factory("KLProjectionLayer", proj_type="kl", mean_bound=0.05, cov_bound=5e-4,
                              trust_region_coeff=1.0, scale_prec=True, entropy_schedule="linear", action_dim=Dp,
                              total_train_steps=100, target_entropy=0.0, temperature=0.7, entropy_eq=False,
                              entropy_first=False, do_regression=False, dtype="float32", device=DEV)
    sampler = SegmentTimeSampler(cfg["dt"], T, dict(num_select=25, fixed_interval=True), device=DEV)
    torch.manual_seed(1)
    pairs = sampler.get_time_pairs()
    agent = TemporalCorrelatedAgent(policy, critic, sampler, proj, dtype="float32", device=DEV, lr_policy=3e-4,
                                    lr_critic=1e-3, wd_policy=5e-5, wd_critic=5e-5, discount_factor=1.0,
                                    epochs_policy=epochs, epochs_critic=2, num_minibatchs=2, norm_advantages=True,
                                    segment_advantage="value_subtraction", set_variance=False, gae_scaling=0.95,
                                    use_cuda_graph=use_graph, schedule_lr_policy=True, schedule_lr_critic=True)
    inp = synthetic_inputs(name, B, seed=3, dtype=torch.float32)
    c = lambda t: t.to(DEV)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(B, obs_dim + 2 * D, generator=g)
    step_states = torch.randn(B, T, obs_dim + 2 * D, generator=g)
    with torch.no_grad():
        times = sampler.get_times(c(inp["init_time"]), T)
        mean_old, L_old = policy.policy(c(obs)[..., :-2 * D])
        smp = policy.sample(False, mean_old, L_old, times, c(inp["init_time"]), c(inp["init_pos"]), c(inp["init_vel"]),
                            eps=c(inp["eps"]))
        lp_old = policy.log_prob(smp, mean_old, L_old, times, c(inp["init_time"]), c(inp["init_pos"]),
                                 c(inp["init_vel"]), pred_pairs=pairs)
    dataset = dict(segment_state=c(obs), step_actions=smp, segment_log_prob_estimate=lp_old,
                   segment_params_mean=mean_old.clone(), segment_params_L=L_old.clone(),
                   segment_init_time=c(inp["init_time"]), segment_init_pos=c(inp["init_pos"]),
                   segment_init_vel=c(inp["init_vel"]), step_states=c(step_states), step_rewards=c(inp["rewards"]),
                   step_values=c(inp["values"]), step_dones=c(inp["dones"]),
                   step_time_limit_dones=c(inp["time_limit_dones"]))
    return agent, dataset


def test_full_update_eager_and_graph_agree():
    outs, params = [], []
    for use_graph in (False, True):
        agent, dataset = build(use_graph=use_graph)
        agent.num_iterations = 1
        dataset = agent.process_dataset(dataset)
        assert dataset["segment_advantage"].shape == (64, 24)
        assert abs(dataset["segment_advantage"].mean().item()) < 1e-4            # normalised
        p0 = [p.detach().clone() for p in agent.policy.parameters]
        out = agent.update_policy(dataset)
        assert all(torch.isfinite(torch.tensor(float(v))) for v in out.values())
        assert any((p - q).abs().max() > 0 for p, q in zip(agent.policy.parameters, p0))      # Adam moved them
        for key in ("surrogate_loss_mean", "trust_region_loss_mean", "policy_loss_mean", "entropy_mean",
                    "policy_grad_norm_mean", "projection_proj_old_cov_diff_mean", "projection_new_old_mean_diff_max"):
            assert key in out
        outs.append(out)
        params.append([p.detach().clone() for p in agent.policy.parameters])
    for k in outs[0]:
        assert abs(outs[0][k] - outs[1][k]) <= 1e-4 * max(1.0, abs(outs[0][k])), k
    for p, q in zip(*params):
        assert (p - q).abs().max().item() <= 1e-5


def test_update_critic_and_projection_bounds_hold():
    agent, dataset = build(epochs=6)
    agent.num_iterations = 1
    dataset = agent.process_dataset(dataset)
    c0 = [p.detach().clone() for p in agent.critic.parameters]
    out = agent.update_critic(dataset)
    assert out["critic_loss_mean"] > 0 and any((p - q).abs().max() > 0 for p, q in zip(agent.critic.parameters, c0))
    out = agent.update_policy(dataset)
    # the projected policy respects the trust region at every epoch (KL metric, logged means)
    assert out["projection_proj_old_mean_diff_max"] <= 0.05 * (1 + 1e-3)
    assert out["projection_proj_old_cov_diff_max"] <= 5e-4 * (1 + 1e-2)
    with pytest.raises(NotImplementedError):
        agent.step()                     # environment rollout is outside the B200 path


@pytest.mark.parametrize("use_graph,minibatches,clip", [(False, 3, 0.0), (True, 3, 0.5), (True, 1, 0.0), (False, 1, 0.0)])
def test_update_critic_matches_oracle(use_graph, minibatches, clip):
    """update_critic (temporal_correlated_agent.py:323-379) against the oracle restatement with the SAME numpy
    permutations: per-step losses, gradient norms and the updated critic weights."""
    import numpy as np
    from oracle import agent as oa
    agent, dataset = build(B=16, use_graph=use_graph)
    agent.epochs_critic, agent.num_minibatchs, agent.clip_grad_norm = 3, minibatches, clip
    agent.num_iterations = 1
    dataset = agent.process_dataset(dataset)
    net = copy.deepcopy(agent.critic.net).cpu().double()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-5)
    odata = {k: (v.detach().double().cpu() if v.is_floating_point() else v.cpu()) for k, v in dataset.items()
             if torch.is_tensor(v)}
    np.random.seed(123)
    losses, norms, clipped = oa.update_critic(net, opt, odata, 3, minibatches, agent.policy.num_dof, 0.0, clip)
    np.random.seed(123)
    out = agent.update_critic(dataset)
    assert abs(out["critic_loss_mean"] - np.mean(losses)) <= 1e-4 * np.mean(losses)
    assert abs(out["critic_loss_min"] - np.min(losses)) <= 1e-4 * np.mean(losses)
    assert abs(out["critic_grad_norm_max"] - np.max(norms)) <= 1e-4 * np.max(norms)
    assert abs(out["clipped_critic_grad_norm_mean"] - np.mean(clipped)) <= 1e-4 * np.mean(clipped)
    for p, q in zip(agent.critic.parameters, net.parameters()):
        assert (p.detach().double().cpu() - q.detach()).abs().max().item() <= 5e-5
    # a second call (new learning rate after a scheduler step) re-captures and keeps training
    if agent.critic_lr_scheduler:
        agent.critic_lr_scheduler.step()
    out2 = agent.update_critic(dataset)
    assert out2["critic_loss_mean"] < out["critic_loss_mean"]


def test_overlapped_epoch_equals_serial_epoch():
    """Side streams (covariance chain first, trust-region loss / logging branches, weight gradients on per-layer
    streams, early gradient clearing) only reorder independent work: metrics and updated parameters of two epochs
    must equal those of the fully serial schedule (tolerance 1e-6: the flat-buffer gradient norm sums in another
    order)."""
    res = []
    for overlap in (True, False):
        agent, dataset = build(epochs=2)
        agent.fast_epoch = False                                   # this test is about the generic epoch's streams
        agent.num_iterations = 1
        dataset = agent.process_dataset(dataset)
        agent.ensure_flat_grads(agent.policy_net_params)
        old = [dataset["segment_params_mean"], dataset["segment_params_L"]]
        agent.projection.initial_entropy = agent.policy.entropy(old).mean()       # as update_policy does
        if not overlap:
            agent.overlap_logging = False
            agent.projection.overlap = False
            agent.policy.mean_net.side_wgrad = False
        times = agent.sampler.get_times(dataset["segment_init_time"], agent.sampler.num_times)
        rows = [agent.policy_epoch(dataset, times, agent.sampler.pred_pairs) for _ in range(2)]
        torch.cuda.synchronize()
        res.append((torch.stack(rows).cpu(), [p.detach().clone().cpu() for p in agent.policy.parameters]))
    (m1, p1), (m2, p2) = res
    assert torch.isfinite(m1).all()
    assert (m1 - m2).abs().max().item() <= 1e-6 * max(1.0, m2.abs().max().item())
    for a, b in zip(p1, p2):
        assert (a - b).abs().max().item() <= 1e-6


@pytest.mark.parametrize("name,B", [("box", 64), ("box", 1024), ("metaworld", 200), ("table_tennis", 33)])
def test_fast_epoch_equals_generic_epoch(name, B):
    """The hand-scheduled shared-covariance epoch (rl/fast_epoch.py: closed-form trust-region / logging terms, mean
    chain in two kernels) against the generic autograd-driven epoch: metrics of three consecutive epochs (incl. warm-
    started projections and Adam steps in between) and the updated parameters."""
    res = []
    for fast in (True, False):
        agent, dataset = build(name=name, B=B, epochs=3)
        agent.fast_epoch = fast
        agent.num_iterations = 1
        dataset = agent.process_dataset(dataset)
        with torch.no_grad():                                     # away from old == new: both projection branches
            for p in agent.policy.mean_net.parameters():
                p.add_(0.05 * torch.randn_like(p))
            agent.policy.variance_net.variable.add_(0.02 * torch.randn_like(agent.policy.variance_net.variable))
        agent.ensure_flat_grads(agent.policy_net_params)
        old = [dataset["segment_params_mean"], dataset["segment_params_L"]]
        agent.projection.initial_entropy = agent.policy.entropy(old).mean()
        times = agent.sampler.get_times(dataset["segment_init_time"], agent.sampler.num_times)
        rows = [agent.policy_epoch(dataset, times, agent.sampler.pred_pairs).clone() for _ in range(3)]
        torch.cuda.synchronize()
        assert (agent._fast is not None) == fast
        res.append((torch.stack(rows).cpu(), [p.detach().clone().cpu() for p in agent.policy.parameters]))
    (m1, p1), (m2, p2) = res
    assert torch.isfinite(m1).all()
    assert (m2[:, 7] > 0.05).any() or (m2[:, 8] > 5e-4).any()     # the trust region was active
    err = (m1 - m2).abs() / m2.abs().clamp_min(1.0)
    assert err.max().item() <= 2e-5, err
    for a, b in zip(p1, p2):
        assert (a - b).abs().max().item() <= 2e-6


def test_save_and_load_agent_resume_training(tmp_path):
    """save_agent / load_agent (abstract_agent.py:109-174) in the reference's file layout: a resumed agent continues
    exactly like the uninterrupted one (weights, Adam moments and step counters of BOTH FlatAdam optimisers)."""
    import os
    agent, dataset = build(epochs=2)
    agent.num_iterations = 1
    dataset = agent.process_dataset(dataset)
    agent.update_critic(dataset)
    agent.update_policy(dataset)
    agent.save_agent(str(tmp_path), 1)
    files = set(os.listdir(tmp_path))
    assert {"policy_optimizer_state_1", "critic_optimizer_state_1", "ValueFunction_mlp_weights_1",
            "ValueFunction_mlp_parameters.pkl", "TemporalCorrelatedPolicy_mean_mlp_weights_1",
            "TemporalCorrelatedPolicy_variance_variable_weights_1"} <= files, files
    sd = agent.policy_optimizer.state_dict()                   # torch.optim.Adam layout
    assert set(sd) == {"state", "param_groups"} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd["state"][0]["step"]) == 2.0
    import numpy as np
    np.random.seed(5)                                          # the critic's minibatch shuffles (numpy global generator)
    agent.update_critic(dataset)
    out_a = agent.update_policy(dataset)
    want = [p.detach().clone() for p in agent.policy.parameters + agent.critic.parameters]
    # a fresh agent (different weights), resumed from the checkpoint
    agent2, _ = build(epochs=2, seed=123)
    agent2.ensure_flat_grads(agent2.policy_net_params)
    agent2._critic_flat_adam()
    agent2.load_agent(str(tmp_path), 1)
    assert agent2.num_iterations == 1
    agent2.projection.initial_entropy = agent.projection.initial_entropy
    np.random.seed(5)
    agent2.update_critic(dataset)
    out_b = agent2.update_policy(dataset)
    got = [p.detach() for p in agent2.policy.parameters + agent2.critic.parameters]
    for a, b in zip(want, got):
        assert (a - b).abs().max().item() <= 1e-6
    assert abs(out_a["policy_loss_mean"] - out_b["policy_loss_mean"]) <= 1e-5 * max(1.0, abs(out_a["policy_loss_mean"]))


def test_dataset_to_device_broadcasts_the_shared_old_factor():
    agent, dataset = build()
    host = {k: v.cpu().pin_memory() for k, v in dataset.items()}
    dev = agent.dataset_to_device(host)
    L = dev["segment_params_L"]
    assert L.shape == dataset["segment_params_L"].shape and L.stride(0) == 0 and L._tce_first.shape[0] == 1
    assert torch.equal(L[5], dataset["segment_params_L"][5])
    for k, v in dataset.items():
        if k != "segment_params_L":
            assert torch.equal(dev[k], v), k
    ptrs = {k: v.data_ptr() for k, v in dev.items()}
    host["segment_params_mean"] = (host["segment_params_mean"] + 1.0).pin_memory()
    dev2 = agent.dataset_to_device(host, out=dev)
    torch.cuda.synchronize()
    assert dev2 is dev and all(dev[k].data_ptr() == ptrs[k] for k in dev)          # static addresses
    assert torch.equal(dev["segment_params_mean"].cpu(), host["segment_params_mean"])


@pytest.mark.parametrize("max_norm,wd", [(0.0, 0.0), (0.5, 5e-5)])
def test_flat_adam_matches_torch_adam(max_norm, wd):
    """tce_grad_sumsq + tce_adam_step == clip_grad_norm_ + torch.optim.Adam.step over several steps (fp32 updates:
    1e-6 relative), including the reported gradient norm."""
    from tce_rl_b200.rl.optim import FlatAdam
    torch.manual_seed(0)
    shapes = [(33, 7), (33,), (5, 33), (5,), (2016,)]
    ref = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    flat = torch.zeros(sum(p.numel() for p in mine), device=DEV)
    off = 0
    for p in mine:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    opt_ref = torch.optim.Adam(ref, lr=3e-3, weight_decay=wd)
    opt = FlatAdam(mine, flat, lr=3e-3, weight_decay=wd)
    for it in range(6):
        opt.begin()
        gs = [torch.randn_like(p) * (10.0 if it == 2 else 0.1) for p in ref]
        for p, q, g in zip(ref, mine, gs):
            p.grad = g.clone()
            q.grad.add_(g)
        norm_ref = torch.linalg.vector_norm(torch.cat([g.reshape(-1) for g in gs]))
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt_ref.step()
        opt.step(max_norm=max_norm)
        assert abs(opt.grad_norm().item() - norm_ref.item()) <= 1e-5 * norm_ref.item()
        for p, q in zip(ref, mine):
            assert (p - q).abs().max().item() <= 2e-6 * max(1.0, p.abs().max().item()), it
    assert opt.stats[0].item() == 6.0


def test_side_stream_gradients_are_scoped_to_the_update():
    """The agent switches the mean network's side-stream weight gradients on only while its own update code runs
    (``_side_grad_scope``): outside, ``torch.autograd.grad`` / a user's own backward through the policy see stock
    autograd (every parameter gets its gradient)."""
    agent, dataset = build(epochs=1)
    net = agent.policy.mean_net
    assert net.side_wgrad is False
    dataset = agent.process_dataset(dataset)
    metrics = agent.update_policy(dataset)
    assert net.side_wgrad is False and bool(np.isfinite(metrics["policy_loss_mean"]))
    obs = dataset["segment_state"][..., :12]
    params = list(net.parameters())
    grads = torch.autograd.grad(net(obs).square().sum(), params)
    assert all(g is not None and torch.isfinite(g).all() and g.abs().sum() > 0 for g in grads)
    for p in params:                                            # .grad buffers exist (views of the flat buffer): a plain
        assert p.grad is not None                               # backward accumulates into them on the current stream
    before = [p.grad.clone() for p in params]
    net(obs).square().sum().backward()
    torch.cuda.synchronize()
    assert all(((p.grad - b) - g).abs().max() <= 1e-4 * max(1.0, g.abs().max().item()) for p, b, g in zip(params, before, grads))
