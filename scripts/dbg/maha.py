import sys, torch
from tce_rl_b200 import ops
n, B, shared, bwd = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = torch.Generator().manual_seed(1)
mean, mean_o = torch.randn(B, n, generator=g), torch.randn(B, n, generator=g)
L = torch.tril(0.1 * torch.randn(B, n, n, generator=g), -1) + torch.diag_embed(0.5 + torch.rand(B, n, generator=g))
if shared: L = L[:1].clone()
md, od, Ld = (t.cuda().requires_grad_(bool(bwd)) for t in (mean, mean_o, L))
got = ops.gauss_maha(md, od, Ld)
torch.cuda.synchronize()
z = torch.linalg.solve_triangular(L.double().expand(B, n, n), (mean - mean_o).double()[..., None], upper=False)
print(sys.argv[1:], "fwd err", (got.cpu() - z.square().sum((1, 2))).abs().max().item(), flush=True)
if bwd:
    gs = torch.autograd.grad(got.sum(), [md, od, Ld])
    torch.cuda.synchronize()
    print("bwd ok", [float(x.abs().max()) for x in gs], flush=True)
