import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tce_rl_b200 import _lib
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
agent, dataset, times, pairs = bench.build_gpu_workload(dev, 0, 1)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): agent.policy_epoch(dataset, times, pairs)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): agent.policy_epoch(dataset, times, pairs)
for _ in range(5): g.replay()
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * (8 * 160 * 4))()
lib.tce_debug_maha(buf)
import numpy as np
a = np.array(buf, dtype=np.uint64).reshape(8, 160, 4).astype(np.int64)
t0 = a[:, :128, 0][a[:, :128, 0] > 0].min()
for s in range(8):
    x = a[s, :128]
    if x[:, 0].max() == 0: continue
    st, en = x[:, 0] - t0, x[:, 1] - t0
    dur = en - st
    worst = np.argsort(-dur)[:4]
    print(f"slot {s} bwd={x[0,3]} kernel start {st.min()/1e3:.1f} us end {en.max()/1e3:.1f} us  cta dur median {np.median(dur)/1e3:.1f} max {dur.max()/1e3:.1f} us; late starters: {(st > st.min() + 3000).sum()}  worst ctas {[(int(w), int(x[w,2]), round(dur[w]/1e3,1), round((st[w]-st.min())/1e3,1)) for w in worst]}")
