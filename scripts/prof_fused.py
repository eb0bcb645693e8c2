"""One launch of each fused-likelihood kernel (for ncu): python scripts/prof_fused.py [B] [ctx|shared|uniform ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from tce_rl_b200 import ops

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
which = sys.argv[2:] or ["ctx", "shared", "uniform"]
name = "box"
cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
inp = synthetic_inputs(name, B, dtype=torch.float32)
times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float()
torch.manual_seed(0)
pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
c = lambda t: t.to(dev)
tabs = ops.Tables(**cfg)
g = {k: c(v) for k, v in inp.items()}
tg, pg = c(times), c(pairs)
theta = ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0)
traj = ops.prodmp_traj(theta, tg, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, cfg["num_dof"])
P = pairs.shape[0]
L1 = g["L"][:1].contiguous()
sigma = (L1[0].double() @ L1[0].double().T).contiguous()
lp_old = ops.seg_logprob(traj, g["mean"], L1, tg, g["init_time"], g["init_pos"], g["init_vel"], pg, tabs) - 0.05
adv = torch.randn(B, P, device=dev)
args = (tg, g["init_time"], g["init_pos"], g["init_vel"], pg)
for rep in range(2):
    if "ctx" in which:
        ops.seglik(traj, g["mean"], g["L"], None, None, *args, tabs.handle, 1e-4, 2, None, lp_old, adv, True, False, True)
    if "shared" in which:
        ops.seglik(traj, g["mean"], L1, sigma, None, *args, tabs.handle, 1e-4, 2, None, lp_old, adv, True, False, True)
    if "uniform" in which:
        ops.seglik(traj, g["mean"], L1, sigma, None, *args, tabs.handle, 1e-4, 2, None, lp_old, adv, True, True, True)
    torch.cuda.synchronize()
print("done")
