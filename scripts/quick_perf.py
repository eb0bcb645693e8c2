"""Quick kernel timings on one GPU (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tce_rl_b200 import ops, _lib
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from oracle import util as ou

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

name = sys.argv[1] if len(sys.argv) > 1 else "box"
for B in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1024", "16384"])]:
    cfg = MP_CONFIGS[name]; T = NUM_TIMES[name]
    inp = synthetic_inputs(name, B, dtype=torch.float32)
    times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float()
    torch.manual_seed(0)
    pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
    c = lambda t: t.to(dev)
    tabs = ops.Tables(**cfg)
    g = {k: c(v) for k, v in inp.items()}
    times_g, pairs_g = c(times), c(pairs)
    theta = ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0)
    traj = ops.prodmp_traj(theta, times_g, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, cfg["num_dof"])
    P = pairs.shape[0]
    work = ops._work(tabs.handle, B, P, dev); adj = torch.empty_like(work)
    dmax = torch.zeros(1, device=dev, dtype=torch.float64)
    logp = torch.empty(B, P, device=dev); info = torch.empty(B, P, device=dev, dtype=torch.int32)
    glp = torch.ones(B, P, device=dev)
    gm = torch.empty_like(g["mean"]); gL = torch.empty_like(g["L"])
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()
    Dp = g["mean"].shape[1]
    res = {"B": B, "name": name}
    res["traj_us"] = timeit(lambda: ops.prodmp_traj(theta, times_g, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, cfg["num_dof"]))
    res["rsample_us"] = timeit(lambda: ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0))
    res["gram_us"] = timeit(lambda: _lib.call("tce_seglik_gram", tabs.handle, p(traj), p(g["mean"]), p(g["L"]), Dp*Dp, p(times_g), p(g["init_time"]), p(g["init_pos"]), p(g["init_vel"]), p(pairs_g), p(work), p(dmax), B, T, P, st))
    res["chol_fwd_us"] = timeit(lambda: _lib.call("tce_seglik_chol", tabs.handle, p(work), None, p(dmax), 1e-4, None, None, None, 0.0, None, p(logp), p(info), B, P, st))
    res["chol_grad_us"] = timeit(lambda: _lib.call("tce_seglik_chol", tabs.handle, p(work), p(adj), p(dmax), 1e-4, p(glp), None, None, 0.0, None, p(logp), p(info), B, P, st))
    res["bwd_us"] = timeit(lambda: _lib.call("tce_seglik_bwd", tabs.handle, p(adj), p(g["L"]), Dp*Dp, p(times_g), p(g["init_time"]), p(pairs_g), None, p(gm), p(gL), B, T, P, st))
    res["gauss_stats_us"] = timeit(lambda: ops.gauss_stats(g["mean"], g["L"], g["mean_old"], g["L_old"]))
    Linv = ops.tri_inverse(g["L_old"][:1].contiguous())[0]
    gm64 = torch.ones(B, device=dev, dtype=torch.float64)
    res["tri_inverse_us"] = timeit(lambda: ops.tri_inverse(g["L_old"][:1].contiguous()))
    res["maha_us"] = timeit(lambda: ops.gauss_maha(g["mean"], g["mean_old"], g["L_old"]))
    res["maha_bwd_us"] = timeit(lambda: ops.gauss_maha_bwd(gm64, g["mean"], g["mean_old"], g["L_old"]))
    res["maha_shared_us"] = timeit(lambda: ops.gauss_maha_shared(g["mean"], g["mean_old"], Linv))
    res["maha_shared_bwd_us"] = timeit(lambda: ops.gauss_maha_shared_bwd(gm64, g["mean"], g["mean_old"], Linv))
    res["gae_us"] = timeit(lambda: ops.gae(g["rewards"], g["values"], g["dones"], g["time_limit_dones"], 1.0, 0.95, True))
    res["segadv_us"] = timeit(lambda: ops.segment_advantage(1, g["rewards"], g["values"], g["rewards"], pairs_g, 1.0, True))
    res["launch_floor_us"] = timeit(lambda: _lib.call("tce_normalize_by_stats", p(logp), p(dmax.new_ones(3)), 1, st))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}))
