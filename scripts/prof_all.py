"""Launch every data-parallel kernel once (after warm-up) on a large batch -- for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tce_rl_b200 import ops
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs
from oracle import util as ou
dev = "cuda:0"
name, B = "box", int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
inp = synthetic_inputs(name, B, dtype=torch.float32)
times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float().to(dev)
torch.manual_seed(0)
pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True)).to(dev)
g = {k: v.to(dev) for k, v in inp.items()}
tabs = ops.Tables(**cfg)
for _ in range(3):
    theta = ops.mvn_rsample(g["mean"], g["L"], g["eps"], 0, 0)
    traj = ops.prodmp_traj(theta, times, g["init_time"], g["init_pos"], g["init_vel"], tabs.handle, cfg["num_dof"])
    adv, ret = ops.gae(g["rewards"], g["values"], g["dones"], g["time_limit_dones"], 1.0, 0.95, True)
    seg = ops.segment_advantage(1, g["rewards"], g["values"], adv, pairs, 1.0, True)
    mh = ops.gauss_maha(g["mean"], g["mean_old"], g["L_old"])
    L = ops.policy_head(torch.zeros(2016, device=dev), B, 63, 1e-4)
torch.cuda.synchronize()
print("ok")
