"""Likelihood kernel timings on one GPU (CUDA events on the launching stream, L2 flushed between iterations):
fused path (diagmax / fused / reduce, uniform prep / main / finish) against the staged kernels it replaces.
usage: python scripts/perf_seglik.py [name] [B,B,...] [P_select]   -> one JSON line per batch size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from tce_rl_b200 import _lib, ops

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
    return round(ts[len(ts) // 2], 2)


name = sys.argv[1] if len(sys.argv) > 1 else "box"
Bs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1024", "16384"])]
psel = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for B in Bs:
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    inp = synthetic_inputs(name, B, dtype=torch.float32)
    times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float()
    torch.manual_seed(0)
    pairs = ou.get_time_pairs(T, dict(num_select=psel, fixed_interval=True))
    c = lambda t: t.to(dev)
    tabs = ops.Tables(**cfg)
    g = {k: c(v) for k, v in inp.items()}
    tg, pg = c(times), c(pairs)
    theta = ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0)
    traj = ops.prodmp_traj(theta, tg, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, cfg["num_dof"])
    P, Dp = pairs.shape[0], g["mean"].shape[1]
    L1 = g["L"][:1].contiguous()
    lp_old = ops.seg_logprob(traj, g["mean"], g["L"], tg, g["init_time"], g["init_pos"], g["init_vel"], pg, tabs) - 0.05
    adv = torch.randn(B, P, device=dev)
    glp = torch.ones(B, P, device=dev)
    args = (tg, g["init_time"], g["init_pos"], g["init_vel"], pg)
    H = tabs.handle
    sl = lambda L, mode, uniform, **kw: ops.seglik(traj, g["mean"], L, None, None, *args, H, 1e-4, mode,
                                                   glp if mode == 1 else None, lp_old if mode == 2 else None,
                                                   adv if mode == 2 else None, True, uniform, mode != 0)
    res = {"name": name, "B": B, "P": P}
    # whole calls (all launches of the op)
    res["ctx_fwd_us"] = timeit(lambda: sl(g["L"], 0, False))
    res["ctx_fwd_bwd_us"] = timeit(lambda: sl(g["L"], 2, False))
    res["shared_fwd_us"] = timeit(lambda: sl(L1, 0, False))
    res["shared_fwd_bwd_us"] = timeit(lambda: sl(L1, 2, False))
    res["uniform_fwd_us"] = timeit(lambda: sl(L1, 0, True))
    res["uniform_fwd_bwd_us"] = timeit(lambda: sl(L1, 2, True))
    # single kernels through the ABI
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()
    dmax = torch.zeros(1, device=dev, dtype=torch.float64)
    stats = torch.zeros(2, device=dev, dtype=torch.float64)
    logp = torch.empty(B, P, device=dev)
    info = torch.empty(B, P, device=dev, dtype=torch.int32)
    gm = torch.empty(B, Dp, device=dev)
    gL = torch.empty(B, Dp, Dp, device=dev)
    from tce_rl_b200 import ops_seglik
    cfgf = ops_seglik.fused_config(H, B, P, True)
    part = torch.empty(cfgf["part_floats"], device=dev)
    pre = torch.empty(B * cfgf["pre_doubles"], device=dev, dtype=torch.float64)
    res["E"], res["grid"] = cfgf["E"], cfgf["grid"]
    for tag, L, ldb in (("ctx", g["L"], Dp * Dp), ("shared", L1, 0)):
        res[f"{tag}_prepass_us"] = timeit(lambda: _lib.call(
            "tce_seglik_prepass", H, p(traj), p(g["mean"]), p(L), ldb, None, None, p(tg), p(g["init_time"]),
            p(g["init_pos"]), p(g["init_vel"]), p(pg), p(pre), p(dmax), 1, B, T, P, st))
        for mode in (0, 2):
            res[f"{tag}_fused_mode{mode}_us"] = timeit(lambda: _lib.call(
                "tce_seglik_fused", H, p(pre), p(L), ldb, None, None, p(pg), p(dmax), 1e-4, mode, None,
                p(lp_old) if mode else None, p(adv) if mode else None, 1.0 / (B * P), p(stats), p(logp), p(info),
                p(gm) if mode else None, p(gL) if (mode and ldb) else None, p(part) if (mode and not ldb) else None, 1,
                B, P, st))
    gL1 = torch.empty(1, Dp, Dp, device=dev)

    def red():
        part[-4:].zero_()                               # re-arm the ticket (the fused kernel does it in a real call)
        _lib.call("tce_seglik_dsigma_reduce", H, p(part), cfgf["grid"], p(L1), None, p(gL1), None, st)
    res["dsigma_reduce_plus_fill_us"] = timeit(red)
    ws = torch.empty(cfgf["ws_doubles"], device=dev, dtype=torch.float64)
    apart = torch.empty(cfgf["apart_doubles"], device=dev, dtype=torch.float64)
    res["uni_prep_us"] = timeit(lambda: _lib.call("tce_seglik_uniform_prep", H, p(L1), None, None, p(tg),
                                                  p(g["init_time"]), p(pg), p(ws), p(dmax), 1e-4, 3, P, st))
    res["uni_main_us"] = timeit(lambda: _lib.call("tce_seglik_uniform_main", H, p(ws), p(traj), p(g["mean"]),
                                                  p(g["init_pos"]), p(g["init_vel"]), p(pg), 2, None, p(lp_old), p(adv),
                                                  1.0 / (B * P), p(stats), p(logp), p(info), p(gm), p(apart), B, T, P,
                                                  st))
    res["uni_finish_us"] = timeit(lambda: _lib.call("tce_seglik_uniform_finish", H, p(ws), p(apart), cfgf["uni_parts"],
                                                    p(L1), None, p(gL1), None, P, st))
    # staged kernels (round 1)
    work = ops._work(H, B, P, dev)
    res["staged_gram_us"] = timeit(lambda: _lib.call("tce_seglik_gram", H, p(traj), p(g["mean"]), p(g["L"]), Dp * Dp,
                                                     p(tg), p(g["init_time"]), p(g["init_pos"]), p(g["init_vel"]),
                                                     p(pg), p(work), p(dmax), B, T, P, st))
    adj = torch.empty_like(work)
    res["staged_chol_us"] = timeit(lambda: _lib.call("tce_seglik_chol", H, p(work), p(adj), p(dmax), 1e-4, p(glp), None,
                                                     None, 0.0, None, p(logp), p(info), B, P, st))
    res["staged_bwd_us"] = timeit(lambda: _lib.call("tce_seglik_bwd", H, p(adj), p(g["L"]), Dp * Dp, p(tg),
                                                    p(g["init_time"]), p(pg), None, p(gm), p(gL), B, T, P, st))
    # algorithmic FLOPs (SURVEY 8(d) convention): fwd = 14.7 kFLOP / segment for the box shape
    print(json.dumps(res))
