"""Graph-replayed policy-epoch time under different switches (where does the step time go?)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tce_rl_b200.rl import projection_factory

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(192 * 1024 * 1024, device=dev, dtype=torch.int32)


def measure(tag, mutate):
    agent, dataset, times, pairs = bench.build_gpu_workload(dev, 0, 1)
    mutate(agent)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            agent.policy_epoch(dataset, times, pairs)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        agent.policy_epoch(dataset, times, pairs)
    for _ in range(5):
        g.replay()
    tot = 0.0
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print(f"{tag:40s} {tot / 20 * 1000:8.1f} us/step", flush=True)
    if tag == "baseline":                      # phase clocks of the last KL forward of the replayed epoch
        import ctypes
        from tce_rl_b200 import _lib
        buf = (ctypes.c_longlong * 32)()
        _lib.call("tce_debug_kl_phase_cycles", buf)
        st = list(buf)
        names = ["load", "trsm W", "jacobi", "eta solve", "gemm M + save", "scale", "gemm Sigma", "chol", "store"]
        print("    KL fwd in the replayed epoch: sweeps", st[15], " ".join(f"{n}={st[i+1]-st[i]}" for i, n in enumerate(names)),
              "total", st[9] - st[0], "cycles", flush=True)


def swap_proj(typ, **kw):
    def f(agent):
        args = dict(bench.PROJ, action_dim=bench.DP, **kw)
        p = projection_factory(typ, device=dev, dtype="float32", **args)
        p.initial_entropy = agent.projection.initial_entropy
        agent.projection = p
    return f


measure("baseline", lambda a: None)
measure("no side streams", lambda a: (setattr(a.projection, "overlap", False), setattr(a, "overlap_logging", False),
                                      setattr(a.policy.mean_net, "side_wgrad", False)))
measure("no KL warm start", lambda a: setattr(a.projection, "warm_start", False))
measure("unfused surrogate", lambda a: setattr(a, "fused_surrogate", False))
measure("BaseProjectionLayer (no cov projection)", swap_proj("BaseProjectionLayer"))
measure("W2 projection", swap_proj("WassersteinProjectionLayer"))
measure("Frobenius projection", swap_proj("FrobeniusProjectionLayer"))
measure("KL, no entropy schedule", swap_proj("KLProjectionLayer", entropy_schedule=None))
