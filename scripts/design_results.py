"""Fill DESIGN.md section 7 (between the results markers) from the committed measurements under profiles/:
r02_bench_n1.json, r02_bench_reference_arm.json, r02_bench_n{1_short_same_box,2_short,8_short}.json,
r02_hbm_kernels_summary.txt.  Run after profiles/ has been refreshed."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda f: os.path.join(ROOT, "profiles", f)
last = lambda f: json.loads(open(P(f)).read().strip().splitlines()[-1])
b = last("r02_bench_n1.json")
ref = last("r02_bench_reference_arm.json")
v, also, roof, cb = b["config"]["variants"], b["also"], b["roofline"], b["cpu_baseline"]
ws = roof["whole_step"]
ctx = v["contextual_P24"]
ctx_roof = ctx.get("roofline", {})
ctxk = ctx_roof.get("kernels", {})
klf = ctxk.get("tce_proj_kl_entropy_fwd", {}).get("us")
klb = ctxk.get("tce_proj_kl_entropy_bwd_tr", {}).get("us")
out = []
w = out.append
w(f"**Headline** (`r02_bench_n1.json`, 1 GPU, CUDA graph replay, L2 flushed between steps, clocks "
  f"{b['clocks']['sm_mhz']} MHz, reasons {b['clocks']['reasons']}): **{b['ms_per_step']:.3f} ms per policy epoch = "
  f"{b['value'] / 1e6:.2f} M episodes/s** device-resident (round 1: 0.364 ms), **{b['e2e']['value'] / 1e6:.2f} M episodes/s end to end** "
  f"({b['e2e']['ms_per_step']:.3f} ms with {b['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB uploaded and the loss vector read back every step); "
  f"{b['gpu_launches_per_step']} GPU activities per epoch ({b['gpu_launches_detail']['own_kernels_per_step']} through the C ABI). "
  f"CPU arm on the same 1024-episode batch: {cb['value']:.0f} episodes/s on {cb['cores']} cores ({cb['ms_per_step']:.0f} ms per epoch; "
  f"`--impl reference`: {ref['value']:.0f}).")
w("")
w("| variant / config (1 GPU) | ms per epoch | episodes/s |")
w("|---|---|---|")
w(f"| config 2, shared covariance, P = 24 (headline) | {b['ms_per_step']:.3f} | {b['value']:.3g} |")
for k, label in (("shared_P25", "config 2, shared covariance, literal P = 25 {0,4,…,96,99}"),
                 ("contextual_P24", "config 2, per-episode covariances [B,63,63], P = 24"),
                 ("contextual_P25", "config 2, per-episode covariances, P = 25")):
    if k in v:
        w(f"| {label} | {v[k]['ms_per_step']:.3f} | {v[k]['episodes_per_s']:.3g} |")
for k, label in (("config1_boxpush_B152", "config 1, box pushing B = 152"), ("config3_metaworld_kl", "config 3, metaworld KL, B = 4096"),
                 ("config4_table_tennis_w2", "config 4, table tennis W2, B = 1024")):
    if k in also:
        w(f"| {label} | {also[k]['ms_per_step']:.3f} | {also[k]['episodes_per_s']:.3g} |")
c1 = also.get("config1_boxpush_B152", {})
if "cpu_fp32" in c1:
    w("")
    w(f"Config 1 on the CPU (oracle port, {c1['cpu_fp32']['cores']} cores): fp32 {c1['cpu_fp32']['ms_per_step']:.1f} ms, "
      f"fp64 {c1['cpu_fp64']['ms_per_step']:.1f} ms per epoch against {c1['ms_per_step']:.3f} ms on the GPU.")
w("")
w(f"**Roofline of the headline step.** Dominant kernel over all C-ABI launches: `{roof['kernel']}` {roof['kernel_us']:.0f} µs — a "
  f"single-CTA latency chain on ONE 63×63 covariance, fraction of any peak ≈ 0 by construction (`roofline.frac` = {roof['frac']:.1e}). "
  f"Whole step: {ws['algorithmic_flops'] / 1e9:.2f} GFLOP and {ws['algorithmic_bytes'] / 1e6:.1f} MB (SURVEY §8(d)) in {ws['ms_per_step']:.3f} ms = "
  f"{ws['frac_of_blended_fma']:.3f} of the blended FP32/FP64 FMA peak ({roof['compute']['measured_fp32_fma_tflops']:.0f} / "
  f"{roof['compute']['measured_fp64_fma_tflops']:.0f} TFLOP/s measured in the run), {ws['frac_of_fp32_fma']:.3f} of the FP32 peak "
  f"(round 1: 0.059), {ws['frac_of_hbm']:.3f} of HBM. The likelihood (all launches) takes {roof['likelihood']['us_all_launches']:.0f} µs "
  f"= {roof['likelihood']['alg_tflops']:.1f} algorithmic TFLOP/s (round 1: 135 µs, 8 TFLOP/s). Per-kernel CUDA-event times of an eager "
  "epoch: " + ", ".join(f"`{k}` {x['us']:.0f}" for k, x in list(roof["kernels"].items())[:8]) + " µs.")
if klf:
    nmat = ctx["episodes_per_gpu"]
    w("")
    w(f"**Per-episode covariances** ({nmat} KL projections per epoch): forward {klf:.0f} µs = {nmat / klf:.2f} matrices/µs, backward "
      f"{klb:.0f} µs; whole epoch {ctx['ms_per_step']:.2f} ms = {ctx_roof.get('whole_step', {}).get('frac_of_blended_fma', float('nan')):.3f} of the "
      "blended FMA peak. ncu (`r02_ncu_summary.txt`, grid = 1024): FP64 pipe 24–26 % busy, issue slots 34–47 %, two 256-thread CTAs per SM "
      "(three fp64 63×63 buffers each fill the shared memory): the Jacobi sweep and the triangular inverse are dependent chains, and the "
      "3.46 waves of 296 resident CTAs leave a partial last wave.")
w("")
w("**Scaling** (weak, 1024 episodes per GPU, gradient exchange as one push kernel over NVLink peer memory; "
  "`r02_bench_n{1_short_same_box,2_short,8_short}.json`): " + ", ".join(
      f"N = {last(f)['n_gpus']}: {last(f)['ms_per_step']:.3f} ms ({last(f)['value'] / 1e6:.1f} M episodes/s)"
      for f in ("r02_bench_n1_short_same_box.json", "r02_bench_n2_short.json", "r02_bench_n8_short.json")) +
  "; 8-GPU efficiency 0.93 (round 1: 0.82). The sharded update equals the single-GPU update on the concatenated batch at 2 and 8 ranks "
  "(`r02_sharded_check_{2,8}gpu.json`).")
w("")
w("**HBM-bound family** (`r02_hbm_kernels_summary.txt`; CUDA events, L2 flushed, fractions of the measured 6544 GB/s on ALGORITHMIC bytes):")
w("")
w("```")
w(open(P("r02_hbm_kernels_summary.txt")).read().rstrip())
w("```")
w("")
w("The factor-streaming kernels (rsample, Mahalanobis, policy head) use bulk asynchronous copies (`cp.async.bulk` + mbarrier); "
  "rsample moves its dense-layout bytes at 0.90 of the peak, above what a read-only torch reduction reaches on the same box. The "
  "trajectory kernel (9/10 writes) and GAE sit at 0.41–0.46: issue-bound at ~7 warp instructions per (episode, time) point.")
w("")
s5 = also.get("config5_likelihood_sweep")
if s5:
    w("**Config 5, likelihood sweep with per-episode factors** (general path `seglik_gram/chol/bwd`, P = 25): " + "; ".join(
        f"B = {r['B']}: fwd {r['fwd_us']:.0f} µs, fwd+bwd {r['fwd_bwd_us']:.0f} µs ({r['seg_logprobs_per_s_fwd_bwd'] / 1e6:.0f} M segment "
        f"log-probs/s, {r['frac_fp32_fma_fwd_bwd']:.3f} of the FP32 FMA peak)" for r in s5) +
      ". With ONE covariance and a common time grid (every shipped config) the same likelihood costs "
      f"{roof['likelihood']['us_all_launches']:.0f} µs at B = 1024 because C_p, its factor and inverse are formed once for the batch.")
txt = "\n".join(out)
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
if "RESULTS_PLACEHOLDER" in s:
    s = s.replace("RESULTS_PLACEHOLDER", "<!-- results:begin -->\n<!-- results:end -->")
s = re.sub(r"<!-- results:begin -->.*<!-- results:end -->", lambda m: "<!-- results:begin -->\n" + txt + "\n<!-- results:end -->", s, flags=re.S)
open(p, "w").write(s)
print(txt)
