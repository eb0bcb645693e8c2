import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tce_rl_b200 import ops, _lib
from oracle.gen_golden import synthetic_inputs
inp = synthetic_inputs("box", 4, dtype=torch.float32)
L, Lo = inp["L"][:1].cuda(), inp["L_old"][:1].cuda()
import sys as _s
warm = len(_s.argv) > 1 and _s.argv[1] == "warm"
state = ops.kl_state(1, 63, "cuda")
for i in range(3):
    out = ops.proj_kl_cov(L + (1e-4 * i if warm else 0.0) * torch.tril(torch.ones_like(L)), Lo, 5e-4, state, warm)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
_lib.call("tce_debug_kl_phase_cycles", buf)
st = list(buf)[:10]
names = ["load", "trsm W", "jacobi", "eta solve", "save", "load+gemm M", "gemm Sigma", "chol", "store"]
print("state tail {eta, active, kl0, fingerprint, alpha, ent_active}", state[-8:-2].tolist(), "sweeps", buf[15])
for i, n in enumerate(names):
    print(f"{n:14s} {st[i+1]-st[i]:9d} cycles")
print("total", st[9]-st[0])
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.proj_kl_cov(L, Lo, 5e-4, state, warm)
b.record(); torch.cuda.synchronize()
print("us per call (incl. python)", a.elapsed_time(b) * 100)
if any(buf[10:15]):
    nm = ["dot+STS", "barrier", "combine+F2F", "rotation", "apply+shfl(loop)"]
    tot = sum(buf[10:15])
    for i, k in enumerate(nm):
        print(f"  jacobi/{k:18s} {buf[10+i]:9d} cycles (last sweep)  {buf[10+i]/max(tot,1):.2f}")
