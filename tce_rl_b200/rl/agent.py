"""TCE agent (policy-update side) with the reference's method surface.

Reference: mprl/rl/agent/abstract_agent.py:12-107 (optimisers, LR schedulers),
mprl/rl/agent/temporal_correlated_agent.py: process_dataset :102-116, get_advantage_return :118-181,
get_segment_advantage :183-321, update_critic :323-379, update_policy :381-639, kl_old_new_proj :641-686,
value_loss :688-716, surrogate_loss :718-739, entropy_loss :741-745.

Differences that do not change results: all per-epoch logging scalars are accumulated in ONE device
buffer and read back once per update (the reference issues >= 20 host synchronisations per epoch, SURVEY
3.3); the epoch can be captured in a CUDA graph (``use_cuda_graph=True``); at >1 GPU the gradients are
all-reduced (SUM of per-rank means scaled by 1/world) in one flat NCCL call per optimiser step.
Environment rollout (``sampler.run``) is out of scope: ``step()`` needs a sampler that provides it.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import LinearLR

from .. import ops, util
from .._lib import TceError
from .projection import KLProjectionLayer, gaussian_kl_details

_KL_KEYS = ["new_old_mean_diff", "new_old_cov_diff", "new_old_shape_diff", "new_old_volume_diff",
            "new_proj_mean_diff", "new_proj_cov_diff", "new_proj_shape_diff", "new_proj_volume_diff",
            "proj_old_mean_diff", "proj_old_cov_diff", "proj_old_shape_diff", "proj_old_volume_diff"]
_LOSS_KEYS = ["surrogate_loss", "entropy_loss", "trust_region_loss", "policy_loss", "entropy", "imp_smp_ratio",
              "policy_grad_norm"]


class SegmentTimeSampler:
    """The two sampler methods the update path needs (temporal_correlated_sampler.py:64-85).
    Time-pair selection draws from the HOST torch generator exactly like the reference."""

    def __init__(self, dt: float, num_times: int, time_pairs_config: dict, device="cuda", dtype=torch.float32):
        self.dt, self.num_times = float(dt), int(num_times)
        self.time_pairs_config = dict(time_pairs_config)
        self.device, self.dtype = torch.device(device), dtype
        self.pred_pairs = None

    def get_times(self, init_time, num_times):
        return util.tensor_linspace(start=init_time + self.dt, end=init_time + num_times * self.dt,
                                    steps=num_times).T.contiguous()

    def get_time_pairs(self):
        pairs = util.select_pred_pairs(num_all=self.num_times, **self.time_pairs_config)
        self.pred_pairs = pairs.to(torch.long).to(self.device)
        return self.pred_pairs


def _stats(values, name):
    a = np.asarray(values, dtype=np.float64)
    return {f"{name}_mean": a.mean(), f"{name}_max": a.max(), f"{name}_min": a.min(), f"{name}_std": a.std(),
            f"{name}_median": np.median(a)}


class TemporalCorrelatedAgent:
    def __init__(self, policy, critic, sampler, projection, dtype=torch.float32, device="cuda", **kwargs):
        self.policy, self.critic, self.sampler, self.projection = policy, critic, sampler, projection
        self.dtype, self.device = util.parse_dtype_device(dtype, device)
        self.lr_policy, self.lr_critic = float(kwargs["lr_policy"]), float(kwargs["lr_critic"])
        self.wd_policy, self.wd_critic = float(kwargs["wd_policy"]), float(kwargs["wd_critic"])
        self.schedule_lr_policy = kwargs.get("schedule_lr_policy", False)
        self.schedule_lr_critic = kwargs.get("schedule_lr_critic", False)
        self.total_iterations = kwargs.get("total_iterations", 10000)
        self.discount_factor = torch.tensor(float(kwargs["discount_factor"]), dtype=self.dtype, device=self.device)
        self._gamma = float(kwargs["discount_factor"])
        self.epochs_policy, self.epochs_critic = kwargs["epochs_policy"], kwargs["epochs_critic"]
        self.clip_critic = float(kwargs.get("clip_critic", 0.0))
        self.clip_grad_norm = float(kwargs.get("clip_grad_norm", 0.0))
        self.num_minibatchs = kwargs.get("num_minibatchs", 10)
        self.norm_advantages = kwargs.get("norm_advantages", False)
        self.clip_advantages = kwargs.get("clip_advantages", False)
        self.entropy_penalty_coef = float(kwargs.get("entropy_penalty_coef", 0.0))
        self.use_gae = kwargs.get("use_gae", True)
        self.gae_scaling = float(kwargs.get("gae_scaling", 0.95))
        self.segment_advantage = kwargs.get("segment_advantage", "accumulate")
        self.set_variance = kwargs.get("set_variance", False)
        self.balance_check = kwargs.get("balance_check", 10)
        self.evaluation_interval = kwargs.get("evaluation_interval", 1)
        self.use_cuda_graph = bool(kwargs.get("use_cuda_graph", False))
        self.fused_surrogate = bool(kwargs.get("fused_surrogate", True))
        self.overlap_logging = bool(kwargs.get("overlap_logging", True))
        self.use_flat_adam = bool(kwargs.get("flat_adam", True))
        # hand-scheduled epoch for the shipped configuration (shared covariance + KL projection), rl/fast_epoch.py;
        # False = always the generic autograd-driven epoch below (any policy / projection layer)
        self.fast_epoch = bool(kwargs.get("fast_epoch", True))
        # data parallel: exchange the flat gradient with the fused NVLink peer-memory kernel (rl/p2p.py) instead of NCCL
        self.p2p_allreduce = bool(kwargs.get("p2p_allreduce", True))
        self._p2p = None
        self._fast = None
        self._scope_depth = 0                              # _side_grad_scope (re-entrant)
        self._log_stream = None
        self._log_stream2 = None
        self._tr_stream = None
        self._flat_grad = None
        self._flat_adam_tried = False
        self.process_group = kwargs.get("process_group", None)      # torch.distributed group (None = single GPU)
        if self.process_group is not None:
            # the batch-global quantities of the path are global over the SAME group as the gradients: the
            # likelihood regulariser (MAX over all episodes, mp_pytorch get_traj_pos_cov) and the advantage
            # normalisation statistics (temporal_correlated_agent.py:281-284)
            ops.set_regulariser_group(self.process_group)
            ops.set_stats_group(self.process_group)
        self.policy_net_params = policy.parameters
        self.critic_net_params = critic.parameters if critic is not None else []
        capt = dict(capturable=True, fused=True) if self.device.type == "cuda" else {}   # one multi-tensor kernel
        self.policy_optimizer = torch.optim.Adam(self.policy_net_params, lr=self.lr_policy,
                                                 weight_decay=self.wd_policy, **capt)
        self.critic_optimizer = (torch.optim.Adam(self.critic_net_params, lr=self.lr_critic,
                                                  weight_decay=self.wd_critic, **capt)
                                 if self.critic_net_params else None)
        mk = lambda opt: LinearLR(opt, start_factor=1, end_factor=0.01, total_iters=self.total_iterations)
        self.policy_lr_scheduler = mk(self.policy_optimizer) if self.schedule_lr_policy else None
        self.critic_lr_scheduler = (mk(self.critic_optimizer)
                                    if self.schedule_lr_critic and self.critic_optimizer else None)
        self.num_iterations = 0
        self.num_global_steps = 0
        self._graph = None

    # ---- distributed helpers ---------------------------------------------------------------------------
    @property
    def world_size(self):
        return dist.get_world_size(self._group()) if self._distributed else 1

    @property
    def _distributed(self):
        return self.process_group is not None and dist.is_available() and dist.is_initialized()

    def _group(self):
        return None if self.process_group is True else self.process_group

    def _allreduce_grads(self, params):
        """One flat all-reduce(SUM) of the gradients; losses are local means, so divide by the world size."""
        if not self._distributed:
            return
        flat = self._flat_grad
        if (getattr(self, "_p2p", None) is not None and flat is not None and params is self.policy_net_params
                and getattr(self.policy_optimizer, "reducer", None) is self._p2p):
            return                 # exchanged inside FlatAdam.step (fused with the gradient norm, rl/p2p.py)
        if flat is not None and self._flat_grad_ok(params):        # gradients already live in one buffer
            if dist.get_backend(self._group()) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self._group())
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self._group())
                flat.div_(self.world_size)
            return
        grads = [p.grad for p in params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self._group())
        flat.div_(self.world_size)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _flat_grad_ok(self, params):
        off, base = 0, self._flat_grad.data_ptr()
        for p in params:
            if p.grad is None or p.grad.data_ptr() != base + off * self._flat_grad.element_size():
                return False
            off += p.numel()
        return off == self._flat_grad.numel()

    def ensure_flat_grads(self, params):
        """Give every parameter a ``.grad`` that is a view into ONE flat buffer (autograd accumulates in place):
        the data-parallel all-reduce then needs no gather / scatter kernels around it."""
        params = list(params)
        if self._flat_grad is not None and self._flat_grad_ok(params):
            return
        numel = sum(p.numel() for p in params)
        self._p2p = None
        if self._distributed and self.p2p_allreduce and self.use_flat_adam and params[0].is_cuda:
            from . import p2p
            if p2p.available(self._group()):
                self._p2p = p2p.P2PGradBuffer(numel, params[0].device, self._group())
        flat = self._p2p.buffer if self._p2p is not None else torch.zeros(numel, dtype=params[0].dtype,
                                                                          device=params[0].device)
        off = 0
        for p in params:
            view = flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()
        self._flat_grad = flat
        if params and params[0] is self.policy_net_params[0] and self.use_flat_adam:
            self._use_flat_adam()
            if self._p2p is not None and hasattr(self.policy_optimizer, "reducer"):
                self.policy_optimizer.reducer = self._p2p

    def _global_mean(self, x):
        """Mean over the global batch (equal shard sizes)."""
        m = x.mean()
        if self._distributed:
            dist.all_reduce(m, op=dist.ReduceOp.SUM, group=self._group())
            m = m / self.world_size
        return m

    # ---- dataset processing -----------------------------------------------------------------------------
    def dataset_to_device(self, host_dataset, out=None, non_blocking=True):
        """Move a (pinned) host dataset in the reference's layout (sampler output / ``process_dataset``) to the
        device.  ``out``: a dict returned by an earlier call -- its buffers are overwritten in place (static
        addresses: a captured epoch graph can be replayed on the new data).

        A non-contextual policy stores ONE old covariance factor B times in ``segment_params_L`` [B, n, n]
        (black_box_policy.py:50-53 repeats the parameter): only the first is transferred and the device tensor
        is a stride-0 broadcast of it -- 16 KB instead of 16 MB per 1024 box-pushing episodes."""
        dev = self.device
        shared_L = not self.policy.contextual_cov
        res = out if out is not None else {}
        for k, v in host_dataset.items():
            if not torch.is_tensor(v):
                res[k] = v
                continue
            if k == "segment_params_L" and shared_L and v.dim() == 3:
                first = getattr(res.get(k), "_tce_first", None)
                if first is None:
                    first = torch.empty((1,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
                    res[k] = first.expand(v.shape[0], -1, -1)
                    res[k]._tce_first = first
                first.copy_(v[:1], non_blocking=non_blocking)
            elif k in res and torch.is_tensor(res[k]) and res[k].shape == v.shape:
                res[k].copy_(v, non_blocking=non_blocking)
            else:
                res[k] = v.to(dev, non_blocking=non_blocking)
        return res

    def process_dataset(self, dataset):
        adv, ret = self.get_advantage_return(dataset["step_rewards"], dataset["step_values"], dataset["step_dones"],
                                             dataset["step_time_limit_dones"])
        dataset["step_advantages"], dataset["step_returns"] = adv, ret
        dataset["segment_advantage"] = self.get_segment_advantage(dataset["step_rewards"], dataset["step_values"],
                                                                  adv, self.sampler.pred_pairs)
        return dataset

    def get_advantage_return(self, rewards, values, dones, time_limit_dones):
        """GAE(gamma, lambda) in one launch (``tce::gae``)."""
        return ops.gae(rewards, values, dones, time_limit_dones, self._gamma, self.gae_scaling, bool(self.use_gae))

    def get_segment_advantage(self, rewards, values, advantages, pred_pairs, **kwargs):
        norm = bool(self.norm_advantages)
        if self.segment_advantage == "accumulate":
            if norm:
                advantages = ops.normalize(advantages)
            if self.clip_advantages > 0:
                advantages = torch.clamp(advantages, -self.clip_advantages, self.clip_advantages)
            return ops.segment_advantage(0, rewards, values, advantages.contiguous(), pred_pairs, self._gamma, norm)
        if self.segment_advantage == "value_subtraction":
            return ops.segment_advantage(1, rewards, values, advantages, pred_pairs, self._gamma, norm)
        if self.segment_advantage == "accumulated_rewards":
            acc = ops.segment_advantage(2, rewards, values, advantages, pred_pairs, self._gamma, False)
            mean = acc.mean(dim=0)
            if self._distributed:
                dist.all_reduce(mean, op=dist.ReduceOp.SUM, group=self._group())
                mean = mean / self.world_size
            return acc - mean        # gamma^start cancels: (acc - mean) / gamma^start with acc already divided
        raise NotImplementedError

    # ---- losses ---------------------------------------------------------------------------------------------
    def value_loss(self, values, returns, old_vs):
        vf_loss = (returns - values).pow(2)
        if self.clip_critic > 0:
            vs_clipped = old_vs + (values - old_vs).clamp(-self.clip_critic, self.clip_critic)
            vf_loss = torch.max(vf_loss, (vs_clipped - returns).pow(2))
        return vf_loss.mean()

    @staticmethod
    def surrogate_loss(advantages, log_prob_new, log_prob_old):
        ratio = (log_prob_new - log_prob_old).exp()
        return -(ratio * advantages).mean(), {"imp_smp_ratio": ratio.mean()}

    def entropy_loss(self, params_mean, params_L):
        entropy = self.policy.entropy([params_mean, params_L]).mean()
        return -self.entropy_penalty_coef * entropy, {"entropy": entropy}

    def _entropy_term(self, proj):
        if self.entropy_penalty_coef != 0.0:
            return self.entropy_loss(proj[0], proj[1])
        with torch.no_grad():                             # coefficient 0 in every config: logging value only
            return self.entropy_loss(proj[0], proj[1])

    def kl_old_new_proj(self, new, old, proj):
        """The 12 logging means of temporal_correlated_agent.py:641-686 as one device vector.  Parts that the
        projection / trust-region loss of this epoch already evaluated (``projection.cache``) are reused."""
        cache = getattr(self.projection, "cache", {})
        kl_metric = isinstance(self.projection, KLProjectionLayer)
        with torch.no_grad():
            mp = cache.get("new_old_mean") if kl_metric else None
            linv = getattr(self.projection, "_old_linv", None) if kl_metric else None
            no_cov, po_cov = cache.get("new_old_cov"), cache.get("proj_old_cov")
            if kl_metric and no_cov is not None and po_cov is not None and mp is not None:
                # per-episode covariances projected by the fused KL kernel: every covariance term is a closed form of
                # the eigen-systems in the projection's state -- only the two missing mean terms are evaluated
                parts = [mp, *no_cov]
                parts += list(cache["new_proj"]) if "new_proj" in cache else list(
                    gaussian_kl_details(self.policy, new, proj))
                parts += [0.5 * ops.gauss_maha(proj[0], old[0], old[1]), *po_cov]
                return torch.stack([x.expand(new[0].shape[0]) for x in parts]).mean(dim=1)
            second = None
            if new[0].is_cuda and self.overlap_logging:     # the two decompositions are independent chains (each
                cur = torch.cuda.current_stream()            # has a single-CTA kernel): run them side by side
                if self._log_stream2 is None:
                    self._log_stream2 = torch.cuda.Stream(device=new[0].device)
                second = self._log_stream2
                second.wait_stream(cur)
                with torch.cuda.stream(second):
                    last = list(gaussian_kl_details(self.policy, proj, old, q_linv=linv))
                    for x in last:
                        x.record_stream(cur)
            parts = list(gaussian_kl_details(self.policy, new, old, mean_part=mp, q_linv=linv))
            parts += list(cache["new_proj"]) if "new_proj" in cache else list(
                gaussian_kl_details(self.policy, new, proj))
            if second is not None:
                cur.wait_stream(second)
            else:
                last = list(gaussian_kl_details(self.policy, proj, old, q_linv=linv))
            parts += last
            return torch.stack([x.expand(new[0].shape[0]) for x in parts]).mean(dim=1)    # one reduction

    # ---- critic ---------------------------------------------------------------------------------------------
    def _flat_buffer(self, params):
        """Give every parameter a ``.grad`` view into ONE new flat buffer (see ``ensure_flat_grads``)."""
        flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            view = flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()
        return flat

    def _critic_flat_adam(self):
        """Flat gradient buffer + ``FlatAdam`` for the critic (same maths as torch Adam + clip_grad_norm_)."""
        from .optim import FlatAdam
        if getattr(self, "_critic_flat", None) is not None or isinstance(self.critic_optimizer, FlatAdam):
            return isinstance(self.critic_optimizer, FlatAdam)
        params = self.critic_net_params
        if (self.device.type != "cuda" or not self.use_flat_adam or self.critic_optimizer.state
                or len(params) > 32):
            self._critic_flat = False
            return False
        self._critic_flat = self._flat_buffer(params)
        g = self.critic_optimizer.param_groups[0]
        opt = FlatAdam(params, self._critic_flat, lr=g["lr"], betas=g["betas"], eps=g["eps"],
                       weight_decay=g["weight_decay"])
        if "initial_lr" in g:
            opt.param_groups[0]["initial_lr"] = g["initial_lr"]
        if self.critic_lr_scheduler is not None:
            sched = LinearLR(opt, start_factor=1, end_factor=0.01, total_iters=self.total_iterations)
            sched.load_state_dict(self.critic_lr_scheduler.state_dict())
            opt.param_groups[0]["lr"] = g["lr"]
            self.critic_lr_scheduler = sched
        self.critic_optimizer = opt
        return True

    def _critic_step(self, states, returns, old_values, sel, row):
        """One optimiser step of update_critic on the minibatch ``sel`` (None: the whole batch in its own order);
        writes {loss, grad norm, clipped grad norm} into ``row`` [3] on the device."""
        D2 = self.policy.num_dof * 2
        flat = isinstance(getattr(self, "_critic_flat", None), torch.Tensor)
        keep_tail = getattr(self, "_critic_keep_tail", False)        # BBRL: the critic sees the whole state
        if sel is None:
            st, rt, ov = states, returns, old_values
        else:
            st, rt, ov = states.index_select(0, sel), returns.index_select(0, sel), old_values.index_select(0, sel)
        values_new = self.critic.critic(st if keep_tail else st[..., :-D2]).squeeze(-1)
        loss = self.value_loss(values_new, rt, ov)
        if flat:
            self.critic_optimizer.begin()
        else:
            self.critic_optimizer.zero_grad(set_to_none=True)
        loss.backward()
        util.join_side_grads()
        if self._distributed:
            if flat:
                buf = self._critic_flat
                if dist.get_backend(self._group()) == "nccl":
                    dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self._group())
                else:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self._group())
                    buf.div_(self.world_size)
            else:
                self._allreduce_grads(self.critic_net_params)
        if flat:
            self.critic_optimizer.step(max_norm=float(self.clip_grad_norm))
            norm = self.critic_optimizer.grad_norm()
        else:
            norm = self._grad_norm_clip(self.critic_net_params)
            self.critic_optimizer.step()
        clipped = torch.clamp(norm, max=self.clip_grad_norm) if self.clip_grad_norm > 0 else norm
        row[0].copy_(loss.detach())
        row[1].copy_(norm.detach())
        row[2].copy_(clipped.detach())

    def update_critic(self, dataset):
        """temporal_correlated_agent.py:323-379: ``epochs_critic`` passes over the flattened step states in
        ``num_minibatchs`` shuffled minibatches (numpy GLOBAL generator, util_data_structure.py:378-391), value loss,
        gradient-norm clipping, Adam.  All permutations of the update are drawn up front (same generator sequence as
        the reference: nothing else consumes numpy's generator in between) and uploaded once; the per-step losses and
        norms stay on the device and are read back once.  With ``use_cuda_graph`` the minibatch step (gather, MLP
        forward/backward, all-reduce, norm + Adam in two launches) is captured once per minibatch size and replayed.
        ``num_minibatchs == 1`` (every shipped config): the full-batch mean does not depend on the order, so the
        gather is skipped."""
        states = dataset["step_states"].flatten(0, 1)
        old_values = dataset["step_values"][:, :-1].flatten(0, 1).contiguous()
        returns = dataset["step_returns"].flatten(0, 1)
        n, E, M = states.shape[0], int(self.epochs_critic), int(self.num_minibatchs)
        self._critic_flat_adam()
        splits = []
        for _ in range(E):
            idx = np.arange(n)
            np.random.shuffle(idx)                                   # reference RNG call sequence
            splits.append(np.array_split(idx, M))
        gather = M > 1
        if gather:
            perm = torch.as_tensor(np.concatenate([np.concatenate(sp) for sp in splits])).to(states.device)
        out = torch.zeros(E * M, 3, device=states.device, dtype=torch.float64)
        graphs = getattr(self, "_critic_graphs", None)
        key = (states.data_ptr(), returns.data_ptr(), old_values.data_ptr(), n, M,
               float(self.critic_optimizer.param_groups[0]["lr"]))   # the learning rate is baked into a captured step
        if self.use_cuda_graph and states.is_cuda and (graphs is None or graphs["key"] != key):
            graphs = self._critic_graphs = {"key": key, "by_size": {}}
        k = 0
        off = 0
        for e in range(E):
            for m in range(M):
                size = len(splits[e][m])
                sel = perm[off:off + size] if gather else None
                off += size
                if self.use_cuda_graph and states.is_cuda:
                    ent = graphs["by_size"].get(size)
                    if ent is None:
                        sel_static = torch.empty(size, device=states.device, dtype=torch.long) if gather else None
                        row_static = torch.zeros(3, device=states.device, dtype=torch.float64)
                        if gather:
                            sel_static.copy_(sel)
                        side = torch.cuda.Stream()
                        side.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(side):                # warm-up outside the capture (lazy inits)
                            self._critic_step(states, returns, old_values, sel_static, row_static)
                        torch.cuda.current_stream().wait_stream(side)
                        out[k].copy_(row_static)
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._critic_step(states, returns, old_values, sel_static, row_static)
                        graphs["by_size"][size] = (g, sel_static, row_static)
                        k += 1
                        continue
                    g, sel_static, row_static = ent
                    if gather:
                        sel_static.copy_(sel)
                    g.replay()
                    out[k].copy_(row_static)
                else:
                    self._critic_step(states, returns, old_values, sel, out[k])
                k += 1
        res = out.cpu().numpy()                                      # the ONE synchronisation of the update
        return {**_stats(res[:, 0], "critic_loss"), **_stats(res[:, 1], "critic_grad_norm"),
                **_stats(res[:, 2], "clipped_critic_grad_norm")}

    def _zero_policy_grads(self):
        from .optim import FlatAdam
        if isinstance(self.policy_optimizer, FlatAdam):
            self.policy_optimizer.begin()               # gradients + norm accumulator
        else:
            self._flat_grad.zero_()

    def _use_flat_adam(self):
        """Swap the policy optimiser for ``FlatAdam`` once the flat gradient buffer exists (CUDA only; same maths,
        the learning-rate scheduler is re-attached to it with its progress preserved)."""
        from .optim import FlatAdam
        if (self._flat_adam_tried or self.device.type != "cuda" or self._flat_grad is None
                or not self._flat_grad_ok(self.policy_net_params)):
            return
        self._flat_adam_tried = True
        old = self.policy_optimizer
        if old.state:                                   # already stepped with torch's Adam: keep it
            return
        g = old.param_groups[0]
        try:
            opt = FlatAdam(self.policy_net_params, self._flat_grad, lr=g["lr"], betas=g["betas"], eps=g["eps"],
                           weight_decay=g["weight_decay"])
        except TceError:                                # e.g. more than 32 parameter tensors: torch's Adam stays
            return
        if "initial_lr" in g:
            opt.param_groups[0]["initial_lr"] = g["initial_lr"]
        if self.policy_lr_scheduler is not None:
            sched = LinearLR(opt, start_factor=1, end_factor=0.01, total_iters=self.total_iterations)
            sched.load_state_dict(self.policy_lr_scheduler.state_dict())
            opt.param_groups[0]["lr"] = g["lr"]
            self.policy_lr_scheduler = sched
        self.policy_optimizer = opt

    def _grad_norm_clip(self, params):
        """util_numerical.py:244-275 without the per-parameter .item(): norm on the device."""
        if self._flat_grad is not None and self._flat_grad_ok(params):     # one reduction over the flat buffer
            norm = torch.linalg.vector_norm(self._flat_grad)
            if self.clip_grad_norm > 0:                                     # clip_grad_norm_: coef = max / (norm + 1e-6)
                self._flat_grad.mul_(torch.clamp(self.clip_grad_norm / (norm + 1e-6), max=1.0))
            return norm
        norm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm([p.grad for p in params])))
        if self.clip_grad_norm > 0:
            torch.nn.utils.clip_grad_norm_(params, self.clip_grad_norm)
        return norm

    # ---- policy -----------------------------------------------------------------------------------------------
    _segment_wise = True       # TCE: segment-wise trajectory likelihood; BlackBoxAgent: likelihood of the parameters

    def _policy_obs(self, dataset):
        """Network input: the observation without the desired position / velocity tail (:352, :527)."""
        return dataset["segment_state"][..., :-self.policy.num_dof * 2]

    def _log_probs(self, dataset, proj, times, pred_pairs):
        """-> (log-prob of the data under the projected policy, log-prob recorded at sampling time)."""
        lp = self.policy.log_prob(dataset["step_actions"], params_mean=proj[0], params_L=proj[1], times=times,
                                  init_time=dataset["segment_init_time"], init_pos=dataset["segment_init_pos"],
                                  init_vel=dataset["segment_init_vel"], pred_pairs=pred_pairs)
        return lp, dataset["segment_log_prob_estimate"]

    def _flat_grad_norm(self):
        return torch.linalg.vector_norm(self._flat_grad).double()

    def balance_norms(self, dataset, times, pred_pairs):
        """The balance check of temporal_correlated_agent.py:446-522 / black_box_agent.py:221-283: the gradient norm of
        the surrogate loss alone and of the trust-region loss alone (two extra forward + backward passes, no optimiser
        step).  -> device tensor [2] = (surrogate_grad_norm, trust_region_grad_norm)."""
        with self._side_grad_scope():
            return self._balance_norms(dataset, times, pred_pairs)

    def _balance_norms(self, dataset, times, pred_pairs):
        old = (dataset["segment_params_mean"], dataset["segment_params_L"])
        obs = self._policy_obs(dataset)
        norms = []
        for which in ("surrogate", "trust_region"):
            new = self.policy.policy(obs)
            proj = self.projection(self.policy, new, old, self.num_iterations)
            if which == "surrogate":
                lp, lp_old = self._log_probs(dataset, proj, times, pred_pairs)
                loss, _ = self.surrogate_loss(dataset["segment_advantage"], lp, lp_old)
            else:
                loss = self.projection.get_trust_region_loss(self.policy, new, proj, set_variance=self.set_variance)
            self._zero_policy_grads()
            loss.backward()
            util.join_side_grads()
            if self._p2p is not None and getattr(self.policy_optimizer, "reducer", None) is self._p2p:
                st = torch.zeros(3, device=self._flat_grad.device, dtype=torch.float64)
                self._p2p.allreduce_sumsq(st)                 # global gradient: averaged over the ranks
                norms.append(st[1].sqrt())
            else:
                self._allreduce_grads(self.policy_net_params)
                norms.append(self._flat_grad_norm())
        return torch.stack(norms)

    @contextlib.contextmanager
    def _side_grad_scope(self):
        """While the agent's own update code runs: weight / bias gradients of the mean network on side streams
        (``MLP.side_wgrad``, joined by ``util.join_side_grads`` after every backward in here) and torch's
        accumulate-grad stream-mismatch warning off (the covariance chain deliberately runs forward and backward on a
        side stream).  Both are restored on exit, so user code around the agent (``torch.autograd.grad``, hooks, own
        losses) sees stock autograd behaviour."""
        net = getattr(self.policy, "mean_net", None)
        use = bool(self.overlap_logging and net is not None and hasattr(net, "side_wgrad"))
        prev = net.side_wgrad if use else None
        warn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if use:
            net.side_wgrad = True
        if warn is not None and self._scope_depth == 0:
            warn(False)
        self._scope_depth += 1
        try:
            yield
        finally:
            self._scope_depth -= 1
            if use:
                net.side_wgrad = prev
            if warn is not None and self._scope_depth == 0:
                warn(True)

    def policy_epoch(self, dataset, times, pred_pairs):
        """One epoch body of ``update_policy`` (temporal_correlated_agent.py:524-589): returns the metrics
        vector [7 + 12] (``_LOSS_KEYS`` then ``_KL_KEYS``) living on the device."""
        with self._side_grad_scope():
            return self._policy_epoch(dataset, times, pred_pairs)

    def _policy_epoch(self, dataset, times, pred_pairs):
        if self.fast_epoch:
            from .fast_epoch import SharedCovKLEpoch
            if SharedCovKLEpoch.applicable(self, dataset):
                if self._fast is None:
                    self._fast = SharedCovKLEpoch(self)
                return self._fast.run(dataset, times, pred_pairs)
        old = (dataset["segment_params_mean"], dataset["segment_params_L"])
        obs = self._policy_obs(dataset)
        # gradients are cleared early (after the covariance chain is launched, while this stream has slack), not in
        # front of backward()
        zeroed_early = self._flat_grad is not None and self._flat_grad_ok(self.policy_net_params)
        pre = None
        if not self.policy.contextual_cov and hasattr(self.policy, "shared_params_L"):
            # shared covariance: start its projection (side stream) before the mean net -- same values as
            # policy.policy(obs) followed by the projection, see BaseProjectionLayer.start_cov_projection
            params_L = self.policy.shared_params_L(obs.shape[0])
            pre = self.projection.start_cov_projection(self.policy, params_L, old[1], self.num_iterations)
            if zeroed_early:
                self._zero_policy_grads()
            new = (self.policy.mean_net(obs), params_L)
        else:
            if zeroed_early:
                self._zero_policy_grads()
            new = self.policy.policy(obs)
        proj = self.projection(self.policy, new, old, self.num_iterations, cov_projected=pre,
                               defer_factor=bool(self.fused_surrogate and self._segment_wise
                                                 and hasattr(self.policy, "segment_surrogate")))
        cov_pending = getattr(self.projection, "_cov_pending", False)
        # trust-region loss: small (partly single-CTA) kernels that only need `new` and `proj` -- a parallel
        # branch next to the segment likelihood, forward and (autograd replays the streams) backward
        tr_stream = None
        # entropy coefficient 0 (every config): the entropy term is a logging value, policy_loss + (-0.0) is
        # policy_loss -- evaluate it with the other logging values next to the backward
        defer_entropy = self.entropy_penalty_coef == 0.0 and self.overlap_logging and proj[0].is_cuda
        proj_ready = None
        if self.overlap_logging and proj[0].is_cuda:
            if self._tr_stream is None:
                self._tr_stream = torch.cuda.Stream(device=proj[0].device)
            tr_stream = self._tr_stream
            proj_ready = torch.cuda.Event()                # everything the trust-region branch reads exists here
            proj_ready.record()
        if self.fused_surrogate and self._segment_wise and hasattr(self.policy, "segment_surrogate"):
            surrogate, ratio, _ = self.policy.segment_surrogate(
                dataset["step_actions"], proj[0], proj[1], times, dataset["segment_init_time"],
                dataset["segment_init_pos"], dataset["segment_init_vel"], pred_pairs,
                dataset["segment_log_prob_estimate"], dataset["segment_advantage"])
            sur_stats = {"imp_smp_ratio": ratio}
            if cov_pending:        # stage 1 ran on Sigma alone; everything after this reads the projected factor
                self.projection.join_covariance()
        else:
            if cov_pending:
                self.projection.join_covariance()
            log_prob_new, log_prob_old = self._log_probs(dataset, proj, times, pred_pairs)
            surrogate, sur_stats = self.surrogate_loss(dataset["segment_advantage"], log_prob_new, log_prob_old)
        # the trust-region loss is back-propagated exactly once with a unit seed next to the projection when the two
        # loss terms are separate autograd roots (below): its covariance gradient may then be folded into the
        # projection's backward kernel (no kernels of its own)
        fold = dict(fold=True) if (defer_entropy and tr_stream is not None
                                   and isinstance(self.projection, KLProjectionLayer)) else {}
        if tr_stream is None:
            ent_loss, ent_stats = self._entropy_term(proj)
            tr_loss = self.projection.get_trust_region_loss(self.policy, new, proj, set_variance=self.set_variance)
        else:
            # Built AFTER the likelihood in program order (autograd runs nodes in reverse creation order: this
            # branch's backward -- a long single-CTA kernel -- is then queued before, not behind, the likelihood's
            # backward stage) but ordered on the device only after `proj_ready`, i.e. next to the likelihood.
            cur = torch.cuda.current_stream()
            tr_stream.wait_event(proj_ready)
            if cov_pending:
                self.projection.join_covariance(tr_stream)
            with torch.cuda.stream(tr_stream):
                tr_loss = self.projection.get_trust_region_loss(self.policy, new, proj, set_variance=self.set_variance,
                                                                **fold)
                tr_loss.record_stream(cur)
                if not defer_entropy:
                    ent_loss, ent_stats = self._entropy_term(proj)
                    ent_loss.record_stream(cur)
                    ent_stats["entropy"].record_stream(cur)
        # With the entropy term deferred the total loss is surrogate + tr_loss: the two terms are handed to autograd
        # as separate roots seeded with a cached 1 (no add / fill / cast kernels in front of the likelihood's
        # backward stage, and the main stream need not wait for the trust-region branch in the forward); the SUM
        # is only a logging value and is formed on the logging branch.
        split_roots = defer_entropy and tr_stream is not None
        if not split_roots:
            if tr_stream is not None:
                torch.cuda.current_stream().wait_stream(tr_stream)
            policy_loss = surrogate + tr_loss if defer_entropy else surrogate + ent_loss + tr_loss
        # logging-only KL decomposition: a parallel branch (side stream) next to backward + Adam
        main, side = None, None
        if self.overlap_logging and surrogate.is_cuda:
            main = torch.cuda.current_stream()
            if self._log_stream is None:
                self._log_stream = torch.cuda.Stream(device=surrogate.device)
            side = self._log_stream
            side.wait_stream(main)
            if split_roots:
                side.wait_stream(tr_stream)
            with torch.cuda.stream(side):
                kl = self.kl_old_new_proj(new, old, proj)
                kl.record_stream(main)
                if defer_entropy:
                    ent_loss, ent_stats = self._entropy_term(proj)
                    ent_loss.record_stream(main)
                    ent_stats["entropy"].record_stream(main)
                if split_roots:
                    policy_loss = surrogate.detach() + tr_loss.detach()
                    policy_loss.record_stream(main)
                # everything of the metrics vector but the gradient norm is known here: assemble it on this branch
                early = torch.stack([surrogate.detach(), ent_loss.detach().to(surrogate.dtype), tr_loss.detach(),
                                     policy_loss.detach(), ent_stats["entropy"].detach().to(surrogate.dtype),
                                     sur_stats["imp_smp_ratio"].detach()]).double()
                kl = kl.double()
                early.record_stream(main)
                kl.record_stream(main)
        else:
            kl = self.kl_old_new_proj(new, old, proj)
        if not zeroed_early:
            self.policy_optimizer.zero_grad(set_to_none=False)
        if split_roots:
            torch.autograd.backward([surrogate, tr_loss], [ops.unit_seed(surrogate.device, surrogate.dtype),
                                                           ops.unit_seed(tr_loss.device, tr_loss.dtype)])
        else:
            policy_loss.backward()
        util.join_side_grads()                             # weight gradients of the mean net (side streams)
        self._allreduce_grads(self.policy_net_params)
        if hasattr(self.policy_optimizer, "grad_norm"):     # FlatAdam: norm, clipping and update in two launches
            self.policy_optimizer.step(max_norm=float(self.clip_grad_norm))
            grad_norm = None
        else:
            grad_norm = self._grad_norm_clip(self.policy_net_params)
            self.policy_optimizer.step()
        if side is not None:
            main.wait_stream(side)
        if grad_norm is None:
            grad_norm = self.policy_optimizer.grad_norm()
        if side is not None:                                  # two launches after the optimiser step: sqrt/cast + cat
            return torch.cat([early, grad_norm.detach().double().reshape(1), kl])
        head = torch.stack([surrogate.detach(), ent_loss.detach().to(surrogate.dtype), tr_loss.detach(),
                            policy_loss.detach(), ent_stats["entropy"].detach().to(surrogate.dtype),
                            sur_stats["imp_smp_ratio"].detach(), grad_norm.detach()])
        return torch.cat([head.double(), kl.double()])

    def _update_inputs(self, dataset):
        init_time = dataset["segment_init_time"]
        return init_time, self.sampler.get_times(init_time, self.sampler.num_times), self.sampler.pred_pairs

    def update_policy(self, dataset):
        init_time, times, pred_pairs = self._update_inputs(dataset)
        old = (dataset["segment_params_mean"], dataset["segment_params_L"])
        if self.projection.initial_entropy is None:
            self.projection.initial_entropy = self._global_mean(self.policy.entropy(list(old)))
        self.ensure_flat_grads(self.policy_net_params)
        if self._distributed and times is not None:
            ops.sync_uniform(init_time, times)       # same time grid on all ranks: no per-epoch all-reduce(MAX)
        check_balance = isinstance(self.balance_check, int) and self.balance_check > 0 \
            and self.num_iterations % self.balance_check == 1
        rows, balance_rows = [], []
        if self.use_cuda_graph and not check_balance:   # NCCL all-reduces are captured with the epoch
            metrics = self._graphed_epochs(dataset, times, pred_pairs, rows)
        else:
            for _ in range(self.epochs_policy):
                if check_balance:
                    balance_rows.append(self.balance_norms(dataset, times, pred_pairs))
                rows.append(self.policy_epoch(dataset, times, pred_pairs))
            metrics = torch.stack(rows)
        if balance_rows:
            metrics = torch.cat([metrics, torch.stack(balance_rows).to(metrics.dtype)], dim=1)
        metrics = metrics.cpu().numpy()                      # the ONE synchronisation of the update
        if self._p2p is not None and float(self.policy_optimizer.stats[2].item()) != 0.0:
            raise RuntimeError("data-parallel gradient exchange: a peer did not arrive (tce_p2p_allreduce_sumsq timed out)")
        if not np.isfinite(metrics[:, :4]).all():
            raise Exception("NAN loss detected")           # temporal_correlated_agent.py:569-577
        out = {}
        for i, k in enumerate(_LOSS_KEYS):
            out.update(_stats(metrics[:, i], k))
        for i, k in enumerate(_KL_KEYS):
            out.update(_stats(metrics[:, len(_LOSS_KEYS) + i], "projection_" + k))
        gn = metrics[:, _LOSS_KEYS.index("policy_grad_norm")]
        out.update(_stats(np.minimum(gn, self.clip_grad_norm) if self.clip_grad_norm > 0 else gn,
                          "clipped_policy_grad_norm"))
        if balance_rows:                                     # temporal_correlated_agent.py:601-612
            nb = len(_LOSS_KEYS) + len(_KL_KEYS)
            out.update(_stats(metrics[:, nb], "surrogate_grad_norm"))
            out.update(_stats(metrics[:, nb + 1], "trust_region_grad_norm"))
            out["balance_ratio"] = out["surrogate_grad_norm_mean"] / out["trust_region_grad_norm_mean"]
        out.update(self._after_update_metrics(dataset, old))
        if self.set_variance and not self.policy.contextual_cov:
            with torch.no_grad():
                new = self.policy.policy(self._policy_obs(dataset))
                proj = self.projection(self.policy, new, old, self.num_iterations)
            self.policy.set_cov_variable(proj[1][0].detach())
        return out

    def _after_update_metrics(self, dataset, old):
        return {}

    def _graphed_epochs(self, dataset, times, pred_pairs, rows):
        """Capture one epoch (forward, backward, Adam) in a CUDA graph and replay it ``epochs_policy`` times."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                        # warm-up outside the capture (lazy inits)
            rows.append(self.policy_epoch(dataset, times, pred_pairs).clone())
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_metrics = self.policy_epoch(dataset, times, pred_pairs)
        for _ in range(self.epochs_policy - 1):
            graph.replay()
            rows.append(static_metrics.clone())
        self._graph = graph
        return torch.stack(rows)

    # ---- checkpoints (abstract_agent.py:109-174) ---------------------------------------------------------------
    def save_agent(self, log_dir: str, epoch: int):
        """Policy / critic weights and both optimiser states in the reference's files
        (``*_parameters.pkl``, ``*_weights_<epoch>``, ``{policy,critic}_optimizer_state_<epoch>``)."""
        from .. import rollout
        self.policy.save_weights(log_dir, epoch)
        if self.critic is not None:
            self.critic.save_weights(log_dir, epoch)
        for name, opt in (("policy_optimizer", self.policy_optimizer), ("critic_optimizer", self.critic_optimizer)):
            if opt is not None:
                with open(rollout.get_training_state_save_path(log_dir, name, epoch), "wb") as f:
                    torch.save(opt.state_dict(), f)

    def load_agent(self, log_dir: str, epoch: int):
        """Inverse of ``save_agent``: weights are copied INTO the existing parameters (the flat gradient buffers and
        ``FlatAdam`` keep referring to them), the Adam moments and step counters are restored, the LR schedulers are
        re-created and ``num_iterations = epoch`` as in the reference."""
        from .. import rollout
        self.policy.load_weights(log_dir, epoch)
        if self.critic is not None:
            self.critic.load_weights(log_dir, epoch)
        for name, opt in (("policy_optimizer", self.policy_optimizer), ("critic_optimizer", self.critic_optimizer)):
            if opt is not None:
                sd = torch.load(rollout.get_training_state_save_path(log_dir, name, epoch), map_location=self.device,
                                weights_only=False)
                opt.load_state_dict(sd)
        mk = lambda opt: LinearLR(opt, start_factor=1, end_factor=0.01, total_iters=self.total_iterations)
        self.policy_lr_scheduler = mk(self.policy_optimizer) if self.schedule_lr_policy else None
        self.critic_lr_scheduler = (mk(self.critic_optimizer)
                                    if self.schedule_lr_critic and self.critic_optimizer else None)
        self.num_iterations = epoch

    def step(self):
        if not hasattr(self.sampler, "run"):
            raise NotImplementedError("environment rollout is outside the B200 hot path: provide a sampler with "
                                      "run() (temporal_correlated_sampler.py:91-344) or call update_* directly")
        self.num_iterations += 1
        dataset, n = self.sampler.run(training=True, policy=self.policy, critic=self.critic)
        self.num_global_steps += n
        dataset = self.process_dataset(dataset)
        out = {**self.update_critic(dataset)}
        if self.critic_lr_scheduler:
            self.critic_lr_scheduler.step()
        out.update(self.update_policy(dataset))
        if self.policy_lr_scheduler:
            self.policy_lr_scheduler.step()
        return out


class BlackBoxAgent(TemporalCorrelatedAgent):
    """The BBRL baseline agent on the same kernels (mprl/rl/agent/black_box_agent.py): one Gaussian over the whole
    ProDMP parameter vector, episode-level advantage ``segment_reward - segment_value`` (:90-103), critic on the
    episode's initial state (:105-158), ``update_policy`` with ``BlackBoxPolicy.log_prob`` (63-dim MVN through
    ``tce_gauss_maha``), the same projections / trust-region loss / logging and ``projection.compute_metrics``
    (:159-389).  Dataset keys: segment_state, segment_action, segment_log_prob, segment_params_mean, segment_params_L,
    segment_reward, segment_value."""
    _segment_wise = False

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("discount_factor", 1.0)
        super().__init__(*args, **kwargs)
        self.fast_epoch = False                        # the hand-scheduled epoch is the TCE likelihood's

    def process_dataset(self, dataset):
        adv = dataset["segment_reward"] - dataset["segment_value"]
        if self.norm_advantages:
            if adv.numel() == 1 and not self._distributed:          # black_box_agent.py:95: std := 1 for one episode
                adv = (adv - adv.mean()) / (1.0 + 1e-8)
            else:
                adv = ops.normalize(adv.contiguous())               # global mean / unbiased std (all ranks)
        if self.clip_advantages > 0:
            adv = torch.clamp(adv, -self.clip_advantages, self.clip_advantages)
        dataset["segment_advantage"] = adv
        return dataset

    def _policy_obs(self, dataset):
        return dataset["segment_state"]

    def _log_probs(self, dataset, proj, times, pred_pairs):
        return (self.policy.log_prob(dataset["segment_action"], params_mean=proj[0], params_L=proj[1]),
                dataset["segment_log_prob"])

    def _update_inputs(self, dataset):
        return None, None, None

    def update_critic(self, dataset):
        flat = dict(step_states=dataset["segment_state"][:, None], step_returns=dataset["segment_reward"][:, None],
                    step_values=torch.stack([dataset["segment_value"], dataset["segment_value"]], dim=1))
        self._critic_keep_tail = True
        return super().update_critic(flat)

    def _after_update_metrics(self, dataset, old):
        """projection.compute_metrics(policy, new, proj, step) after the last epoch (black_box_agent.py:359-363)."""
        with torch.no_grad():
            new = self.policy.policy(self._policy_obs(dataset))
            proj = self.projection(self.policy, new, old, self.num_iterations)
            met = self.projection.compute_metrics(self.policy, new, proj, self.num_iterations)
            keys = sorted(met)
            vals = torch.stack([met[k].double() for k in keys]).cpu().numpy()
        return {"projection_" + k: float(v) for k, v in zip(keys, vals)}


def agent_factory(typ: str, **kwargs):
    """mprl/rl/agent/__init__.py:8-19."""
    agents = {"TemporalCorrelatedAgent": TemporalCorrelatedAgent, "BlackBoxAgent": BlackBoxAgent}
    if typ not in agents:
        raise NotImplementedError(f"{typ}: not on the B200 path (SURVEY section 8(f))")
    return agents[typ](**kwargs)
