"""Kernel timeline (stream, start, duration) of ONE graph-replayed policy epoch, from the torch profiler (CUPTI)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
dev = torch.device("cuda", local)
torch.cuda.set_device(local)
if world > 1:           # under torchrun: the data-parallel epoch (peer-memory gradient exchange), one timeline per rank
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
agent, dataset, times, pairs = bench.build_gpu_workload(dev, rank, world)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        agent.policy_epoch(dataset, times, pairs)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
hi = torch.cuda.Stream(priority=-1) if os.environ.get("TCE_HIGH_PRIO") else None
with torch.cuda.graph(g, stream=hi):
    agent.policy_epoch(dataset, times, pairs)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.txt"
if world > 1:
    out = out.replace(".txt", f"_rank{rank}.txt")
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
tmp = f"/tmp/trace{rank}.json"
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# keep the last replay: everything after the previous epoch's last kernel (metrics assembly / policy head as markers)
# (the metrics kernel runs beside the Adam kernel, which ends the epoch: cut after whichever of the two comes last)
ends = [i for i, e in enumerate(ev) if "adam_kernel" in e["name"]]
ends = [max(i, *[j for j in range(i - 2, min(i + 3, len(ev))) if "epoch_metrics_kernel" in ev[j]["name"]] or [i]) for i in ends]
if len(ends) >= 2:
    ev = ev[ends[-2] + 1:ends[-1] + 1]
else:
    cut = max(i for i, e in enumerate(ev) if "head_fwd_kernel" in e["name"])
    ev = ev[cut:]
t0 = ev[0]["ts"]
with open(out, "w") as f:
    f.write(f"# {len(ev)} GPU activities, span {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us\n# start_us dur_us stream name\n")
    for e in ev:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream', '?'):<4} {e['name'][:90]}\n")
print(open(out).read()[:300])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
