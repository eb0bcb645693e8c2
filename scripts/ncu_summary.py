"""Summarise an `ncu --page raw --csv` export: one line per captured launch with duration, DRAM traffic, pipe
utilisation, occupancy, bank conflicts -> profiles/r02_ncu_summary.txt, and per-kernel DRAM bytes (last instance)
-> profiles/r02_ncu_traffic.json (read by bench.py; every entry carries the hash of ITS kernel's sources at capture time)."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except ValueError:
        return float("nan")
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
KEYS = [("us", "gpu__time_duration.sum", None), ("dram_rd_MB", "dram__bytes_read.sum", None), ("dram_wr_MB", "dram__bytes_write.sum", None),
        ("dram%", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", 1), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1),
        ("waves", "launch__waves_per_multiprocessor", 1), ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1)]
unit_of = lambda k: units[col[k]]
KEYS = [(k, c, (TIME if k == "us" else BYTES)[unit_of(c)] if sc is None else sc) for k, c, sc in KEYS]
lines = ["# ncu --set full --clock-control none (one launch per line, capture order); times are cold-cache and serialised",
         "# us = gpu__time_duration.sum, dram_* = dram__bytes_{read,write}.sum, *% = pct_of_peak_sustained_elapsed (warps%: of active)",
         f"{'kernel':44s}" + "".join(f"{k:>12s}" for k, _, _ in KEYS)]
last = {}
for r in data:
    name = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::", "", r[col["Kernel Name"]]).split("(")[0][:44]
    vals = {k: f(r, c) * s for k, c, s in KEYS}
    lines.append(f"{name:44s}" + "".join(f"{vals[k]:12.2f}" if vals[k] == vals[k] else f"{'n/a':>12s}" for k, _, _ in KEYS))
    last[(name, int(vals["grid"]))] = vals
txt = "\n".join(lines)
out_txt = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_ncu_summary.txt")
open(out_txt, "w").write(txt + "\n")
import bench
abi = {"proj_kl_cov_fwd_kernel": ["tce_proj_kl_entropy_fwd_sigma", "tce_proj_kl_cov_fwd", "tce_proj_kl_entropy_fwd",
                                  "tce_proj_kl_entropy_fwd_sigma_vec"],
       "proj_kl_cov_bwd_sigma_kernel": ["tce_proj_kl_bwd_sigma", "tce_proj_kl_bwd_sigma_k", "tce_proj_kl_bwd_sigma_k_vec"],
       "traj_uniform_kernel<9>": ["tce_prodmp_traj_fwd_uniform"], "maha_kernel": ["tce_gauss_maha"],
       "rsample_kernel": ["tce_mvn_rsample"], "uniform_main_kernel<7, 9, 8>": ["tce_seglik_uniform_main"],
       "uniform_prep_kernel<7, 9, true>": ["tce_seglik_uniform_prep"], "uniform_finish_kernel<7, 9>": ["tce_seglik_uniform_finish"],
       "epoch_mean_fwd_kernel": ["tce_epoch_mean_fwd"], "epoch_tr_mean_kernel": ["tce_epoch_tr_mean"],
       "proj_kl_cov_bwd_kernel": ["tce_proj_kl_cov_bwd", "tce_proj_kl_entropy_bwd", "tce_proj_kl_entropy_bwd_tr"],
       "traj_fwd_warp_kernel<9, 8>": ["tce_prodmp_traj_fwd"]}
kern = {}
for (name, grid), vals in last.items():
    for a in abi.get(name, []):
        kern.setdefault(a, {})[f"grid={grid}"] = {"kernel": name, "dram_bytes": int((vals["dram_rd_MB"] + vals["dram_wr_MB"]) * 1e6),
                                                  "us_under_ncu": round(vals["us"], 2), "src_hash": bench.src_hash_of(a)}
json.dump({"source": os.path.relpath(out_txt, ROOT) + " (ncu --set full, last captured instance of each kernel; "
           "keyed by launch grid: grid=1 = the ONE matrix of the shared-covariance epoch, grid=1024 = per-episode covariances)",
           "csrc_hash": bench.csrc_hash(), "kernels": kern},
          open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(txt)
