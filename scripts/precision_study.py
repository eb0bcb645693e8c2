"""Precision evidence behind "no tensor cores for the segment likelihood" (BASELINE north star: tensor cores only if
3xTF32 meets the 1e-4 log-prob tolerance).

The segment likelihood of BASELINE config 2 (box pushing, B = 1024 x 24 segments, per-episode covariance factors) is
evaluated in four arithmetics against the fp64 oracle on the SAME fp32 inputs:

  (i)   fp32 everywhere: Sigma = L L^T, C = H Sigma H^T, Cholesky, solves and the log-prob in float32 (what a plain
        fp32 / cuBLAS-SGEMM formulation does);
  (ii)  emulated 3xTF32 for the two contractions Sigma = L L^T and C = H Sigma H^T (operands split into TF32 big + small
        parts, three products, fp32 accumulation -- the standard error-compensated tensor-core scheme), the rest fp32;
  (iii) fp32 contractions but fp64 factorisation / quadratic forms (isolates where the error comes from);
  (iv)  the shipped GPU path (``ops.seg_logprob``: fp32 inputs, Sigma on the FP32 pipe, quadratic forms / Cholesky /
        adjoint on the FP64 pipe) -- only when a CUDA device is present.

Output: max |delta logp| and max relative gradient error per variant -> profiles/r02_precision.txt.
3xTF32 reproduces AT BEST the fp32-accumulate product, so (ii) can only match (i); if (i) misses the tolerance, a
tensor-core formulation does too.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import policy as opol
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS, NUM_TIMES, synthetic_inputs

LOG_2PI = 1.8378770664093453


def tf32_round(x):
    """Round fp32 to TF32 (10 explicit mantissa bits, round to nearest even) -- what the tensor core sees."""
    xd = x.detach().contiguous()
    i = xd.view(torch.int32)
    lsb = (i >> 13) & 1
    i = (i + 0xFFF + lsb) & ~0x1FFF
    return x + (i.view(torch.float32) - xd)          # straight-through: autograd sees the identity


def mm_3xtf32(a, b):
    """a @ b with both operands split into TF32 big + small parts: big*big + big*small + small*big, fp32 accumulate."""
    ab, bb = tf32_round(a), tf32_round(b)
    asm, bsm = tf32_round(a - ab), tf32_round(b - bb)
    return ab @ bb + (ab @ bsm + asm @ bb)


def variants_logp(H, c, x, L, reg_rel, mode):
    """H [B,P,n,Dp] fp64 (exact basis), c [B,P,n] offset, x [B,P,n] data, L [B,Dp,Dp] fp32 -> logp [B,P] fp64 + grads."""
    B, P, n, Dp = H.shape
    L = L.clone().requires_grad_(True)
    H32, x32, c32 = H.float(), x.float(), c.float()
    if mode in ("fp32", "fp32_contract_fp64_factor"):
        Sigma = L @ L.transpose(-1, -2)
        A = H32 @ Sigma[:, None]
        C = A @ H32.transpose(-1, -2)
    elif mode == "3xtf32":
        Sigma = mm_3xtf32(L, L.transpose(-1, -2))
        A = mm_3xtf32(H32, Sigma[:, None].expand(B, P, Dp, Dp))
        C = mm_3xtf32(A, H32.transpose(-1, -2).contiguous())
    else:
        raise ValueError(mode)
    if mode == "fp32_contract_fp64_factor":
        C = C.double()
        r = (x - c)
    else:
        r = (x32 - c32)
    eye = torch.eye(n, dtype=C.dtype)
    reg = reg_rel * torch.diagonal(C.detach(), dim1=-2, dim2=-1).max()
    S = torch.linalg.cholesky(C + reg * eye)
    z = torch.linalg.solve_triangular(S, r.unsqueeze(-1), upper=False).squeeze(-1)
    lp = -0.5 * (n * LOG_2PI + (z * z).sum(-1)) - torch.diagonal(S, dim1=-2, dim2=-1).log().sum(-1)
    w = torch.linspace(0.5, 1.5, B * P, dtype=lp.dtype).reshape(B, P)
    (gL,) = torch.autograd.grad((lp * w).sum(), [L])
    return lp.detach().double(), torch.tril(gL.detach().double())


def main(B=1024, name="box"):
    torch.manual_seed(0)
    cfg, T = MP_CONFIGS[name], NUM_TIMES[name]
    D, K1 = cfg["num_dof"], cfg["num_basis"] + 1
    Dp, n = D * K1, 2 * D
    inp = synthetic_inputs(name, B, dtype=torch.float32)
    times = ou.get_times(inp["init_time"].double(), T, cfg["dt"]).float()
    torch.manual_seed(0)
    pairs = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
    P = pairs.shape[0]
    pol = opol.TemporalCorrelatedPolicy(Dp, mp=dict(type="prodmp", args=dict(cfg, dtype=torch.float64)), contextual=True,
                                        min_std=1e-4)
    d = lambda k: inp[k].double()
    a64 = (times.double(), d("init_time"), d("init_pos"), d("init_vel"))
    smp = pol.sample(False, d("mean"), d("L"), *a64, eps=d("eps")).float().double()
    # ---- fp64 oracle (reference values and the exact linear map mu = H theta + c) ----------------------------------
    L64 = d("L").clone().requires_grad_(True)
    want = torch.empty(B, P, dtype=torch.float64)
    gL_want = torch.zeros(B, Dp, Dp, dtype=torch.float64)
    w = torch.linspace(0.5, 1.5, B * P, dtype=torch.float64).reshape(B, P)
    reg = 0.0
    for s in range(0, B, 256):                                     # batch-global regulariser first
        sl = slice(s, s + 256)
        with torch.no_grad():
            _, _, cov, _ = pol.log_prob(smp[sl], d("mean")[sl], d("L")[sl], *[a[sl] for a in a64], pred_pairs=pairs,
                                        return_parts=True, reg_override=0.0)
        reg = max(reg, float(torch.diagonal(cov, dim1=-2, dim2=-1).max()) * 1e-4)
    Hs, cs, xs = [], [], []
    for s in range(0, B, 256):
        sl = slice(s, s + 256)
        lp = pol.log_prob(smp[sl], d("mean")[sl], L64[sl], *[a[sl] for a in a64], pred_pairs=pairs, reg_override=reg)
        want[sl] = lp.detach()
        (g,) = torch.autograd.grad((lp * w[sl]).sum(), [L64])
        gL_want += g
        with torch.no_grad():
            Hm = pol.mp.basis_multi_dof()                          # [b, P, n, Dp] of the last update_inputs
            mu = pol.mp.get_traj_pos(flat_shape=True)
            Hs.append(Hm.clone())
            cs.append(mu - torch.einsum('bpnk,bk->bpn', Hm, d("mean")[sl]))
            x = smp[sl][..., pairs, :D]
            xs.append(x.transpose(-1, -2).reshape(*x.shape[:-2], -1))
    H, c0, x = torch.cat(Hs), torch.cat(cs), torch.cat(xs)
    c = c0 + torch.einsum('bpnk,bk->bpn', H, d("mean"))
    gL_want = torch.tril(gL_want)
    lines = [f"# segment likelihood, {name} shape, B = {B}, P = {P}, n = {n}, Dp = {Dp}; reference = fp64 oracle on the same "
             f"fp32 inputs; tolerance (north star): |delta logp| <= 1e-4",
             f"# regulariser reg = {reg:.6e}; cond(C) ~ {float(torch.linalg.cond(H[0, 0] @ (d('L')[0] @ d('L')[0].T) @ H[0, 0].T + reg * torch.eye(n, dtype=torch.float64))):.2e}",
             f"{'variant':46s} {'max|dlogp|':>12s} {'max rel dgrad_L':>16s}  verdict"]

    def report(tag, lp, gL):
        e1 = (lp - want).abs().max().item()
        e2 = ((gL - gL_want).abs().max() / gL_want.abs().max()).item()
        lines.append(f"{tag:46s} {e1:12.3e} {e2:16.3e}  {'meets 1e-4' if e1 <= 1e-4 else 'MISSES 1e-4'}")

    for mode, tag in (("fp32", "(i)   fp32 everywhere"),
                      ("3xtf32", "(ii)  3xTF32 contractions, fp32 rest"),
                      ("fp32_contract_fp64_factor", "(iii) fp32 contractions, fp64 factor/solve")):
        lp, gL = variants_logp(H, c, x, inp["L"], 1e-4, mode)
        report(tag, lp, gL)
    if torch.cuda.is_available():
        from tce_rl_b200 import ops
        dev = "cuda:0"
        tabs = ops.Tables(**cfg)
        cu = lambda t: t.to(dev)
        Lg = cu(inp["L"]).requires_grad_(True)
        lp = ops.seg_logprob(cu(smp.float()), cu(inp["mean"]), Lg, cu(times), cu(inp["init_time"]), cu(inp["init_pos"]),
                             cu(inp["init_vel"]), cu(pairs), tabs)
        (lp * cu(w.float())).sum().backward()
        report("(iv)  shipped CUDA path (fp32 io, fp64 forms)", lp.detach().double().cpu(), Lg.grad.double().cpu())
    lines.append("# (ii) can at best equal (i): 3xTF32 recovers the fp32 product, the accumulation stays fp32.  The error of (i)")
    lines.append("# comes from cancellation in C = H Sigma H^T (neighbouring time points are almost perfectly correlated; the")
    lines.append("# regulariser is 1e-4 of the largest variance) and from the fp32 Cholesky; (iii) shows that rounding the")
    lines.append("# CONTRACTIONS to fp32 is already too much once the factorisation is exact.")
    out = "\n".join(lines)
    print(out)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                "profiles", "r02_precision.txt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(out + "\n")


if __name__ == "__main__":
    main(B=int(os.environ.get("PRECISION_B", "1024")))
