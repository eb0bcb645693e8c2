"""One TCE policy epoch for the configuration every shipped config uses -- non-contextual full covariance, KL
trust-region projection with the entropy control last (or none), entropy coefficient 0 -- as ONE explicit sequence of
kernels: forward and backward of everything but the mean network are issued here by hand, autograd only sees the MLP.

Reference: the epoch body of ``TemporalCorrelatedAgent.update_policy``
(mprl/rl/agent/temporal_correlated_agent.py:524-599): policy -> projection -> log_prob -> surrogate_loss -> entropy_loss
-> get_trust_region_loss -> kl_old_new_proj (logging) -> backward -> grad clip -> Adam.  Same values as the generic path
of ``rl/agent.py`` (``tests/test_gpu_agent.py::test_fast_epoch_equals_generic_epoch`` and the oracle parity tests run
through it), ~35 GPU launches instead of ~92:

  covariance stream : head -> KL projection + entropy control (state: eigen-system, Sigma_out, closed-form logging
                      scalars) .................................... -> KL backward in covariance space -> head backward
  main stream       : mean net -> L_old^-1 -> mean chain forward -> [Sigma ready] segment likelihood + surrogate
                      (forward produces d loss / d proj_mean and d loss / d Sigma_out) -> mean chain backward (+ trust-
                      region mean gradient) -> mean net backward -> [join] all-reduce -> Adam -> metrics

The trust-region loss needs no kernels of its own: its covariance term and gradient come in closed form from the
projection's eigen-system (``tce_proj_kl_bwd_sigma(tr_coeff)``), its mean term uses
Sigma_out^-1 = (Sigma~^-1 + eta Sigma_old^-1) / (alpha^2 (1 + eta)) inside ``tce_epoch_mean_bwd``; the twelve logging KL
parts are closed forms of the same eigenvalues (``csrc/tce_proj.cu: save_tr_value``).
"""
from __future__ import annotations

import os

import torch

from .. import _lib, ops, ops_seglik, util
from .projection import KLProjectionLayer, _first


EARLY_EXCHANGE = os.environ.get("TCE_P2P_EARLY", "1") != "0"    # data parallel: mean-net slice of the gradient goes ahead


def _p(t):
    return None if t is None else t.data_ptr()


class SharedCovKLEpoch:
    def __init__(self, agent):
        self.agent = agent
        self._cov_stream = None
        self._tr_stream = None

    # ---- when the hand-scheduled epoch applies ------------------------------------------------------------------
    @staticmethod
    def applicable(agent, dataset) -> bool:
        pol, proj = agent.policy, agent.projection
        return (type(proj) is KLProjectionLayer and not pol.contextual_cov and not pol.is_diag
                and hasattr(pol, "mp") and hasattr(pol, "shared_params_L") and not proj.entropy_first
                and agent.entropy_penalty_coef == 0.0 and agent.fused_surrogate
                and hasattr(agent.policy_optimizer, "grad_norm") and dataset["segment_params_mean"].is_cuda
                and dataset["segment_params_mean"].shape[0] > 0
                and dataset["segment_params_mean"].shape[-1] <= 64)

    def _stream(self, device):
        if self._cov_stream is None:
            self._cov_stream = torch.cuda.Stream(device=device, priority=-1)   # the covariance chain is the critical path
        return self._cov_stream

    def run(self, dataset, times, pred_pairs):
        """-> metrics [19] fp64 on the device (``_LOSS_KEYS`` then ``_KL_KEYS`` of rl/agent.py)."""
        ag = self.agent
        pol, proj = ag.policy, ag.projection
        D2 = pol.num_dof * 2
        obs = dataset["segment_state"][..., :-D2]
        mean_old = dataset["segment_params_mean"]
        L_old1 = _first(dataset["segment_params_L"]).contiguous()               # [1, n, n]
        B, n = mean_old.shape
        dev = mean_old.device
        f32, f64 = torch.float32, torch.float64
        main = torch.cuda.current_stream()
        cov = self._stream(dev)
        step = ag.num_iterations
        with_cov = bool(proj._with_cov(pol, ag.set_variance))
        coeff = float(proj.trust_region_coeff)
        params = ag.policy_net_params
        vec = pol.variance_net.variable

        acc = torch.zeros(4, device=dev, dtype=f64)
        # ---- covariance chain, forward (side stream; needs neither the observations nor the mean net) ----------------
        cov.wait_stream(main)
        with torch.cuda.stream(cov):
            L_new = torch.empty(1, n, n, device=dev, dtype=f32)         # built from the vector inside the KL kernel
            state = proj._state_for(L_new)
            beta = proj._entropy_bound(step, dev)
            scratch = torch.empty(2, n, n, device=dev, dtype=f32)                # factor outputs of an identity step
            info = torch.empty(1, device=dev, dtype=torch.int32)
            _lib.call("tce_proj_kl_entropy_fwd_sigma_vec", _p(vec), 0, float(pol.min_std), _p(L_new), _p(L_old1),
                      float(proj.cov_bound), _p(beta), 0, int(proj.entropy_eq), _p(scratch[0]), _p(scratch[1]), _p(state),
                      _p(info), int(proj.warm_start), 1, n, cov.cuda_stream)
            sigma_ready = torch.cuda.Event()
            sigma_ready.record(cov)
            # the gradient-independent part of the backward (K = L~^-T U~) while the likelihood is busy
            _lib.call("tce_proj_kl_bwd_prep", _p(state), 1, n, cov.cuda_stream)
            for t in (L_new, scratch, info, acc):
                t.record_stream(cov)
        ag._zero_policy_grads()
        # ---- mean chain, forward ---------------------------------------------------------------------------------------
        mean = pol.mean_net(obs)
        mean_d = mean.detach()
        Linv_old = torch.empty(n, n, device=dev, dtype=f64)
        _lib.call("tce_tri_inverse", _p(L_old1), n * n, _p(Linv_old), 1, n, main.cuda_stream)
        proj_mean = torch.empty(B, n, device=dev, dtype=f32)
        maha_old = torch.empty(B, device=dev, dtype=f64)
        u_old = torch.empty(B, n, device=dev, dtype=f32)
        _lib.call("tce_epoch_mean_fwd", _p(mean_d), _p(mean_old), _p(Linv_old), float(proj.mean_bound), _p(proj_mean),
                  _p(maha_old), _p(u_old), _p(acc), B, n, main.cuda_stream)
        # ---- segment likelihood + surrogate (forward + backward in one pass, gradient in covariance space) ----------
        mean_fwd_done = torch.cuda.Event()
        mean_fwd_done.record(main)
        main.wait_event(sigma_ready)
        nn_ = n * n
        sigma0 = state[3 * nn_:4 * nn_]
        sc = state[4 * nn_ + n:]
        # trust-region mean term (value + gradient): needs the projection's state only -> beside the likelihood
        if self._tr_stream is None:
            self._tr_stream = torch.cuda.Stream(device=dev)
        tr_stream = self._tr_stream
        tr_grad = torch.empty(B, n, device=dev, dtype=f32)
        chained, uniform = ops_seglik._facts(pred_pairs, dataset["segment_init_time"], times, True, None)
        logp, linfo, lacc, g_pm, _, g_S = ops_seglik.seglik(
            dataset["step_actions"], proj_mean, None, sigma0, sc[6:7], times, dataset["segment_init_time"],
            dataset["segment_init_pos"], dataset["segment_init_vel"], pred_pairs, pol.mp.tables.handle, 1e-4, 2, None,
            dataset["segment_log_prob_estimate"], dataset["segment_advantage"], chained, uniform, False, True)
        lik_done = torch.cuda.Event()
        lik_done.record(main)
        # ---- covariance chain, backward --------------------------------------------------------------------------------
        with torch.cuda.stream(cov):
            cov.wait_event(lik_done)
            _lib.call("tce_proj_kl_bwd_sigma_k_vec", _p(L_new), _p(vec), 0, _p(g_S), _p(state), 1,
                      coeff if with_cov else 0.0, _p(vec.grad), 1, n, cov.cuda_stream)
            cov_done = torch.cuda.Event()
            cov_done.record(cov)
            g_S.record_stream(cov)
        # ---- mean chain, backward ----------------------------------------------------------------------------------------
        tr_stream.wait_event(mean_fwd_done)          # maha_old, u_old
        tr_stream.wait_event(sigma_ready)            # the projection's state (L~^-1, eta, alpha)
        with torch.cuda.stream(tr_stream):
            _lib.call("tce_epoch_tr_mean", _p(mean_d), _p(mean_old), _p(maha_old), _p(u_old),
                      _p(state[2 * nn_:3 * nn_]), _p(sc), float(proj.mean_bound), coeff, _p(tr_grad), _p(acc), B, n,
                      tr_stream.cuda_stream)
            for t in (mean_d, maha_old, u_old, tr_grad, acc):
                t.record_stream(tr_stream)
        g_mean = torch.empty(B, n, device=dev, dtype=f32)
        main.wait_stream(tr_stream)
        _lib.call("tce_epoch_mean_combine", _p(g_pm), _p(mean_d), _p(mean_old), _p(maha_old), _p(u_old), _p(tr_grad),
                  float(proj.mean_bound), _p(g_mean), B, n, main.cuda_stream)
        torch.autograd.backward([mean], [g_mean])
        util.join_side_grads()
        # data parallel: the mean network's slice of the flat gradient is final -- exchange it under the covariance
        # chain's backward; only the covariance vector's slice trails the chain (FlatAdam.step)
        if getattr(ag.policy_optimizer, "reducer", None) is not None and EARLY_EXCHANGE:
            ag.policy_optimizer.exchange_early(sum(p.numel() for p in params[:-1]))
        main.wait_event(cov_done)
        ag._allreduce_grads(params)
        metrics = torch.empty(19, device=dev, dtype=f64)
        metrics_done = torch.cuda.Event()

        def metrics_beside_adam():          # needs the gradient norm, not the update: runs next to the Adam kernel
            norm_ready = torch.cuda.Event()
            norm_ready.record(main)
            tr_stream.wait_event(norm_ready)
            _lib.call("tce_epoch_metrics", _p(acc), _p(lacc[1:]), _p(sc), _p(ag.policy_optimizer.stats), B, coeff,
                      int(with_cov), float(ag.entropy_penalty_coef), _p(metrics), tr_stream.cuda_stream)
            metrics_done.record(tr_stream)
            for t in (acc, lacc, metrics):
                t.record_stream(tr_stream)

        ag.policy_optimizer.step(max_norm=float(ag.clip_grad_norm), after_norm=metrics_beside_adam)
        main.wait_event(metrics_done)
        self.last = dict(logp=logp, info=linfo, proj_mean=proj_mean, state=state, mean=mean_d)
        return metrics
