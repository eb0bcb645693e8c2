"""Phase clocks of the KL forward kernel INSIDE the graph-replayed headline epoch (SM cycles of block 0) plus the
Jacobi diagnostics (sweeps, largest cosine^2 met per sweep)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tce_rl_b200 import _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
contextual = len(sys.argv) > 1 and sys.argv[1] == "contextual"        # per-episode covariance factors: 1024 projections per step
agent, dataset, times, pairs = bench.build_gpu_workload(dev, 0, 1, contextual=contextual)
step_fn, metrics, *_ = bench.capture_epoch(agent, dataset, times, pairs)
for _ in range(10):
    step_fn()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
_lib.call("tce_debug_kl_phase_cycles", buf)
st = list(buf)
names = ["load", "tri_inverse+W", "jacobi", "eta solve", "gemm M + save", "scale", "gemm Sigma", "chol", "store"]
print("sweeps", st[15], "max cos^2 per sweep", [v * 1e-12 for v in st[12:15]])
print(" ".join(f"{n}={st[i+1]-st[i]}" for i, n in enumerate(names)), "total", st[9] - st[0])
print("tri_inverse", st[10] - st[1], "zero+gemm W", st[11] - st[10], "save Li", st[2] - st[11])
bn = ["load Sbar", "load M", "gemm Sbar M", "gemm M^T(.)", "scalars + Nt", "load U", "gemm U Nt", "gemm (.)U^T", "load Li", "gemm Li^T(.)", "store"]
print("KL bwd (covariance space):", " ".join(f"{n}={st[17+i]-st[16+i]}" for i, n in enumerate(bn)), "total", st[27] - st[16])
