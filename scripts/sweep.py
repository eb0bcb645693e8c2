"""BASELINE.json configs 3-5 on one GPU: segment-likelihood sweep (roofline report) and the policy epoch on the
metaworld / table-tennis shapes.  One JSON line per point (CUDA events, L2 flushed between iterations)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tce_rl_b200 import ops, _lib
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from oracle import util as ou

dev = "cuda:0"
flush = torch.empty(192 * 1024 * 1024, device=dev, dtype=torch.int32)
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e-3


def likelihood_point(name, B, P):
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    D, K1 = cfg["num_dof"], cfg["num_basis"] + 1
    Dp, n = D * K1, 2 * D
    inp = synthetic_inputs(name, min(B, 4096), dtype=torch.float32)
    rep = (B + inp["mean"].shape[0] - 1) // inp["mean"].shape[0]
    g = {k: v.repeat(rep, *([1] * (v.dim() - 1)))[:B].contiguous().to(dev) for k, v in inp.items()}
    times = ou.get_times(g["init_time"].double().cpu(), T, cfg["dt"]).float().to(dev)
    idx = torch.arange(0, T, T // (P + 1))[:P + 1]
    pairs = torch.stack([idx[:-1], idx[1:]], 1).to(dev)
    tabs = ops.Tables(**cfg)
    theta = ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0)
    traj = ops.prodmp_traj(theta, times, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, D)
    work = ops._work(tabs.handle, B, P, dev); adj = torch.empty_like(work)
    dmax = torch.zeros(1, device=dev, dtype=torch.float64)
    logp = torch.empty(B, P, device=dev); info = torch.empty(B, P, device=dev, dtype=torch.int32)
    glp = torch.ones(B, P, device=dev); gm = torch.empty_like(g["mean"]); gL = torch.empty_like(g["L"])
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()

    def fwd():
        _lib.call("tce_seglik_gram", tabs.handle, p(traj), p(g["mean"]), p(g["L"]), Dp * Dp, p(times), p(g["init_time"]),
                  p(g["init_pos"]), p(g["init_vel"]), p(pairs), p(work), p(dmax), B, T, P, st)
        _lib.call("tce_seglik_chol", tabs.handle, p(work), None, p(dmax), 1e-4, None, None, None, 0.0, None, p(logp),
                  p(info), B, P, st)

    def fwd_bwd():
        _lib.call("tce_seglik_gram", tabs.handle, p(traj), p(g["mean"]), p(g["L"]), Dp * Dp, p(times), p(g["init_time"]),
                  p(g["init_pos"]), p(g["init_vel"]), p(pairs), p(work), p(dmax), B, T, P, st)
        _lib.call("tce_seglik_chol", tabs.handle, p(work), p(adj), p(dmax), 1e-4, p(glp), None, None, 0.0, None, p(logp),
                  p(info), B, P, st)
        _lib.call("tce_seglik_bwd", tabs.handle, p(adj), p(g["L"]), Dp * Dp, p(times), p(g["init_time"]), p(pairs), None,
                  p(gm), p(gL), B, T, P, st)

    tf, tfb = timed(fwd), timed(fwd_bwd)
    tri = Dp * (Dp + 1) // 2
    fwd_bytes = 4 * (tri + Dp + (P + 1) * D + (1 + 2 * D) + P) * B
    bwd_bytes = fwd_bytes + 4 * (P + Dp + tri) * B
    mac = Dp * (Dp + 1) + K1 * (4 * sum(d * (d + 1) // 2 for d in range(D)) + 3 * D * (D + 1) // 2) + n ** 3 // 6 + n * n // 2 + n * K1
    flops_fwd = 2 * mac * P * B
    assert int(info.abs().max()) == 0
    return {"kind": "segment_likelihood", "shape": name, "B": B, "P": P, "fwd_us": round(tf * 1e6, 1),
            "fwd_bwd_us": round(tfb * 1e6, 1), "seg_logprobs_per_s_fwd": round(B * P / tf),
            "seg_logprobs_per_s_fwd_bwd": round(B * P / tfb),
            "hbm_frac_fwd": round(fwd_bytes / tf / 1e9 / peaks["hbm_gbs"], 4),
            "hbm_frac_fwd_bwd": round((fwd_bytes + bwd_bytes) / tfb / 1e9 / peaks["hbm_gbs"], 4),
            "alg_tflops_fwd": round(flops_fwd / tf / 1e12, 2), "alg_tflops_fwd_bwd": round(3 * flops_fwd / tfb / 1e12, 2)}


def epoch_point(name, B, typ, mean_bound, cov_bound):
    import bench
    from tce_rl_b200.rl import TemporalCorrelatedAgent, policy_factory, projection_factory
    from tce_rl_b200.rl.agent import SegmentTimeSampler
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    D, K1 = cfg["num_dof"], cfg["num_basis"] + 1
    Dp = D * K1
    torch.manual_seed(0)
    policy = policy_factory("TemporalCorrelatedPolicy", dim_in=20, dim_out=Dp, dtype="float32", device=dev,
                            mp=dict(type="prodmp", args=dict(cfg)), **bench.POLICY)
    proj = projection_factory(typ, device=dev, dtype="float32", action_dim=Dp,
                              **dict(bench.PROJ, mean_bound=mean_bound, cov_bound=cov_bound))
    sampler = SegmentTimeSampler(cfg["dt"], T, dict(num_select=25, fixed_interval=True), device=dev)
    torch.manual_seed(0)
    pairs = sampler.get_time_pairs()
    agent = TemporalCorrelatedAgent(policy, None, sampler, proj, dtype="float32", device=dev, **bench.AGENT)
    inp = synthetic_inputs(name, B, dtype=torch.float32)
    c = lambda t: t.to(dev)
    obs = torch.randn(B, 20 + 2 * D)
    with torch.no_grad():
        times = sampler.get_times(c(inp["init_time"]), T)
        mean0, L0 = policy.policy(c(obs)[..., :-2 * D])
        mean_old = mean0 + 0.05 * c(torch.randn(B, Dp))
        L_old = (1.05 * L0[:1] + torch.tril(0.01 * c(torch.randn(Dp, Dp)), -1)).expand(B, -1, -1).contiguous()
        smp = policy.sample(False, mean_old, L_old, times, c(inp["init_time"]), c(inp["init_pos"]), c(inp["init_vel"]),
                            eps=c(inp["eps"]))
        lp_old = policy.log_prob(smp, mean_old, L_old, times, c(inp["init_time"]), c(inp["init_pos"]),
                                 c(inp["init_vel"]), pred_pairs=pairs)
        adv, _ = agent.get_advantage_return(c(inp["rewards"]), c(inp["values"]), c(inp["dones"]),
                                            c(inp["time_limit_dones"]))
        seg = agent.get_segment_advantage(c(inp["rewards"]), c(inp["values"]), adv, pairs)
        for q in policy.mean_net.parameters():
            q.add_(0.05 * torch.randn_like(q))
        policy.variance_net.variable.add_(0.02 * torch.randn_like(policy.variance_net.variable))
    ds = dict(segment_state=c(obs), step_actions=smp, segment_log_prob_estimate=lp_old, segment_params_mean=mean_old,
              segment_params_L=L_old, segment_advantage=seg, segment_init_time=c(inp["init_time"]),
              segment_init_pos=c(inp["init_pos"]), segment_init_vel=c(inp["init_vel"]))
    proj.initial_entropy = policy.entropy([mean_old, L_old]).mean()
    agent.num_iterations = 100
    agent.ensure_flat_grads(agent.policy_net_params)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            agent.policy_epoch(ds, times, pairs)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        m = agent.policy_epoch(ds, times, pairs)
    t = timed(gr.replay, reps=20, warm=5)
    assert torch.isfinite(m).all()
    return {"kind": "policy_epoch", "shape": name, "B": B, "P": int(pairs.shape[0]), "projection": typ,
            "us_per_epoch": round(t * 1e6, 1), "episodes_per_s": round(B / t)}


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "sweep"):
        for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
            for P in (10, 25, 49):
                print(json.dumps(likelihood_point("box", B, P)), flush=True)
    if what in ("all", "epochs"):
        print(json.dumps(epoch_point("box", 1024, "KLProjectionLayer", 0.05, 5e-4)), flush=True)
        print(json.dumps(epoch_point("metaworld", 4096, "KLProjectionLayer", 0.005, 5e-4)), flush=True)
        print(json.dumps(epoch_point("table_tennis", 1024, "WassersteinProjectionLayer", 0.005, 2.5e-4)), flush=True)
        print(json.dumps(epoch_point("table_tennis", 8192, "WassersteinProjectionLayer", 0.005, 2.5e-4)), flush=True)
