"""Oracle restatement of the TCE agent arithmetic (GAE, segment advantages, losses, one policy epoch).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).  Follows
``mprl/rl/agent/temporal_correlated_agent.py``: ``get_advantage_return`` :118-181,
``get_segment_advantage`` :183-321, ``update_policy`` epoch body :524-599,
``kl_old_new_proj`` :641-686, ``value_loss`` :688-716, ``surrogate_loss`` :718-739,
``entropy_loss`` :741-745.  Pinned against the real reference code by
``oracle/gen_golden.py`` -> ``tests/golden/agent_*.pt``.
"""
from __future__ import annotations

import torch

from .projection import gaussian_kl_details


def get_advantage_return(rewards, values, dones, time_limit_dones, discount_factor, gae_scaling,
                         use_gae=True):
    """GAE(gamma, lambda) reversed scan -> (advantages [B,T], returns [B,T])."""
    returns = torch.zeros_like(values)
    nd = torch.logical_not(dones)
    ntl = torch.logical_not(time_limit_dones)
    T = rewards.shape[1]
    discount = torch.as_tensor(discount_factor, dtype=values.dtype) * nd
    if use_gae:
        gae = 0
        for t in reversed(range(T)):
            td = rewards[..., t] + discount[..., t] * values[..., t + 1] - values[..., t]
            gae = (td + discount[..., t] * gae_scaling * gae) * ntl[..., t]
            returns[..., t] = gae + values[..., t]
    else:
        returns[..., -1] = values[..., -1]
        for t in reversed(range(T)):
            returns[..., t] = ntl[..., t] * (rewards[..., t] + discount[..., t] * returns[..., t + 1]) \
                + time_limit_dones[..., t] * values[..., t]
    returns = returns[..., :-1]
    return (returns - values[..., :-1]).clone().detach(), returns.clone().detach()


def get_segment_advantage(rewards, values, advantages, pred_pairs, discount_factor,
                          mode="value_subtraction", norm_advantages=True, clip_advantages=0.0):
    """Advantage of each (start, end) time pair [B, P]."""
    gamma = torch.as_tensor(discount_factor, dtype=rewards.dtype)
    norm = lambda a: (a - a.mean()) / (a.std() + 1e-8)
    start, end = pred_pairs[..., 0], pred_pairs[..., 1]
    T = rewards.shape[-1]
    if mode == "accumulate":
        if norm_advantages:
            advantages = norm(advantages)
        if clip_advantages > 0:
            advantages = torch.clamp(advantages, -clip_advantages, clip_advantages)
        seg = torch.stack([advantages[:, lo:hi + 1].sum(-1) for lo, hi in pred_pairs.tolist()], -1)
        return norm(seg) if norm_advantages else seg
    idx = torch.arange(T)
    disc_rewards = rewards * gamma.pow(idx)
    mask = torch.logical_and(start[:, None] <= idx, idx < end[:, None]).to(rewards.dtype)
    acc = torch.einsum('ik,jk->ij', disc_rewards, mask)
    first = gamma.pow(start)
    if mode == "value_subtraction":
        seg = acc / first + gamma.pow(end - start) * values[:, end] - values[:, start]
        return norm(seg) if norm_advantages else seg
    if mode == "accumulated_rewards":
        return (acc - acc.mean(dim=0)) / first
    raise NotImplementedError(mode)


def value_loss(values, returns, old_vs, clip_critic=0.0):
    loss = (returns - values).pow(2)
    if clip_critic > 0:
        clipped = old_vs + (values - old_vs).clamp(-clip_critic, clip_critic)
        loss = torch.max(loss, (clipped - returns).pow(2))
    return loss.mean()


def generate_minibatches(n, n_minibatches):
    """mprl/util/util_data_structure.py:378-391 -- the GLOBAL numpy generator, shuffle of arange(n), array_split."""
    import numpy as np
    idx = np.arange(n)
    np.random.shuffle(idx)
    return np.array_split(idx, n_minibatches)


def grad_norm_clip(bound, params):
    """mprl/util/util_numerical.py:244-275: (norm before, norm after clip_grad_norm_)."""
    params = [p for p in params if p.grad is not None]
    norm = torch.linalg.vector_norm(torch.stack([p.grad.norm(2) for p in params]))
    if bound > 0:
        torch.nn.utils.clip_grad_norm_(params, bound)
        clipped = torch.linalg.vector_norm(torch.stack([p.grad.norm(2) for p in params]))
    else:
        clipped = norm
    return norm.item(), clipped.item()


def update_critic(critic_net, optimizer, dataset, epochs_critic, num_minibatchs, num_dof, clip_critic=0.0,
                  clip_grad_norm=0.0):
    """TemporalCorrelatedAgent.update_critic (temporal_correlated_agent.py:323-379) -> (losses, grad norms, clipped
    grad norms), one entry per optimiser step."""
    states = dataset["step_states"].flatten(0, 1)
    old_values = dataset["step_values"][:, :-1].flatten(0, 1)
    returns = dataset["step_returns"].flatten(0, 1)
    losses, norms, clipped = [], [], []
    for _ in range(epochs_critic):
        for idx in generate_minibatches(states.shape[0], num_minibatchs):
            sel = torch.as_tensor(idx)
            values_new = critic_net(states[sel][..., :-num_dof * 2]).squeeze(-1)
            loss = value_loss(values_new, returns[sel], old_values[sel], clip_critic)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            n0, n1 = grad_norm_clip(clip_grad_norm, [p for g in optimizer.param_groups for p in g["params"]])
            optimizer.step()
            losses.append(loss.item())
            norms.append(n0)
            clipped.append(n1)
    return losses, norms, clipped


def surrogate_loss(advantages, log_prob_new, log_prob_old):
    ratio = (log_prob_new - log_prob_old).exp()
    return -(ratio * advantages).mean(), {"imp_smp_ratio": ratio.mean()}


def entropy_loss(policy, params_mean, params_L, entropy_penalty_coef):
    ent = policy.entropy([params_mean, params_L]).mean()
    return -entropy_penalty_coef * ent, {"entropy": ent}


def kl_old_new_proj(policy, new, old, proj):
    """The 12 logging scalars of temporal_correlated_agent.py:641-686 (means over the batch)."""
    out = {}
    for name, (p, q) in (("new_old", (new, old)), ("new_proj", (new, proj)), ("proj_old", (proj, old))):
        m, c, s, v = gaussian_kl_details(policy, p, q)
        out[f"{name}_mean_diff"], out[f"{name}_cov_diff"] = m.mean(), c.mean()
        out[f"{name}_shape_diff"], out[f"{name}_volume_diff"] = s.mean(), v.mean()
    return out


def policy_epoch(policy, projection, dataset, times, pred_pairs, num_iterations, *,
                 entropy_penalty_coef=0.0, set_variance=False, with_metrics=False):
    """One epoch body of ``update_policy`` up to the loss (no optimiser step).

    Returns (policy_loss, dict of intermediates).  ``dataset`` keys follow
    temporal_correlated_sampler.py:318-337.
    """
    D2 = policy.num_dof * 2
    old = (dataset["segment_params_mean"], dataset["segment_params_L"])
    if projection.initial_entropy is None:
        projection.initial_entropy = policy.entropy(list(old)).mean()
    new = policy.policy(dataset["segment_state"][..., :-D2])
    proj = projection(policy, new, old, num_iterations)
    lp_new = policy.log_prob(dataset["step_actions"], params_mean=proj[0], params_L=proj[1], times=times,
                             init_time=dataset["segment_init_time"], init_pos=dataset["segment_init_pos"],
                             init_vel=dataset["segment_init_vel"], pred_pairs=pred_pairs)
    sur, sur_stats = surrogate_loss(dataset["segment_advantage"], lp_new, dataset["segment_log_prob_estimate"])
    ent, ent_stats = entropy_loss(policy, proj[0], proj[1], entropy_penalty_coef)
    trl = projection.get_trust_region_loss(policy, new, proj, set_variance=set_variance)
    out = {"new": new, "proj": proj, "log_prob_new": lp_new, "surrogate_loss": sur, "entropy_loss": ent,
           "trust_region_loss": trl, **sur_stats, **ent_stats}
    if with_metrics:
        with torch.no_grad():
            out["kl"] = kl_old_new_proj(policy, new, old, proj)
    return sur + ent + trl, out


def bbrl_process_dataset(dataset, norm_advantages=True, clip_advantages=0.0):
    """BlackBoxAgent.process_dataset (black_box_agent.py:90-103)."""
    adv = dataset["segment_reward"] - dataset["segment_value"]
    if norm_advantages:
        std = adv.std() if len(adv) != 1 else 1.0
        adv = (adv - adv.mean()) / (std + 1e-8)
    if clip_advantages > 0:
        adv = torch.clamp(adv, -clip_advantages, clip_advantages)
    return adv


def policy_epoch_bbrl(policy, projection, dataset, num_iterations, *, entropy_penalty_coef=0.0, set_variance=False,
                      with_metrics=False):
    """One epoch body of BlackBoxAgent.update_policy (black_box_agent.py:285-334) up to the loss."""
    old = (dataset["segment_params_mean"], dataset["segment_params_L"])
    if projection.initial_entropy is None:
        projection.initial_entropy = policy.entropy(list(old)).mean()
    new = policy.policy(dataset["segment_state"])
    proj = projection(policy, new, old, num_iterations)
    lp_new = policy.log_prob(dataset["segment_action"], params_mean=proj[0], params_L=proj[1])
    sur, sur_stats = surrogate_loss(dataset["segment_advantage"], lp_new, dataset["segment_log_prob"])
    ent, ent_stats = entropy_loss(policy, proj[0], proj[1], entropy_penalty_coef)
    trl = projection.get_trust_region_loss(policy, new, proj, set_variance=set_variance)
    out = {"new": new, "proj": proj, "log_prob_new": lp_new, "surrogate_loss": sur, "entropy_loss": ent,
           "trust_region_loss": trl, **sur_stats, **ent_stats}
    if with_metrics:
        with torch.no_grad():
            out["kl"] = kl_old_new_proj(policy, new, old, proj)
            out["projection_metrics"] = projection.compute_metrics(policy, new, proj, num_iterations)
    return sur + ent + trl, out
