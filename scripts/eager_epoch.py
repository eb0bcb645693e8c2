"""Three eager policy epochs (for ncu -k filters on single kernels with the real inputs of the epoch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
agent, dataset, times, pairs = bench.build_gpu_workload(dev, 0, 1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    agent.policy_epoch(dataset, times, pairs)
torch.cuda.synchronize()
print("ok")
