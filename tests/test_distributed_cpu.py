"""Host-side multi-GPU logic on CPU: world_size-2 ``gloo`` processes.

What is exercised: the flat gradient all-reduce of the agent (per-rank local-mean losses -> global-mean
gradient), the global mean used for ``initial_entropy``, the SUM-reduction of the advantage-normalisation
statistics and the MAX-reduction of the likelihood regulariser seed.  The kernels themselves need a GPU
and are covered by the ``gpu`` tests; here only the collective plumbing runs (on CPU tensors).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tce_rl_b200 import ops
        from tce_rl_b200.rl.agent import TemporalCorrelatedAgent

        class _Pol:
            def __init__(self):
                torch.manual_seed(0)
                self.net = torch.nn.Linear(4, 3)
                self.parameters = list(self.net.parameters())
                self.num_dof = 1

        pol = _Pol()
        agent = TemporalCorrelatedAgent(pol, None, None, None, dtype="float32", device="cpu", lr_policy=1e-3,
                                        lr_critic=1e-3, wd_policy=0.0, wd_critic=0.0, discount_factor=1.0,
                                        epochs_policy=1, epochs_critic=1, process_group=True)
        # every rank owns a shard of a global batch; local-mean loss per rank
        g = torch.Generator().manual_seed(123)
        X, Y = torch.randn(8, 4, generator=g), torch.randn(8, 3, generator=g)
        xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
        loss = ((pol.net(xs) - ys) ** 2).mean()
        loss.backward()
        agent._allreduce_grads(pol.parameters)
        # single-process reference: global-mean loss
        torch.manual_seed(0)
        ref = torch.nn.Linear(4, 3)
        ((ref(X) - Y) ** 2).mean().backward()
        err = max((p.grad - q.grad).abs().max().item() for p, q in zip(pol.parameters, ref.parameters()))
        gm = agent._global_mean(torch.full((4,), float(rank + 1)))
        # statistics SUM and regulariser MAX
        ops.set_stats_group(True)
        ops.set_regulariser_group(True)
        stats = torch.tensor([4.0, xs.sum().item(), (xs ** 2).sum().item()], dtype=torch.float64)
        ops._reduce_stats(stats)
        dmax = torch.tensor([float(rank) + 0.5], dtype=torch.float64)
        ops._reduce_diag_max(dmax)
        ok_stats = abs(stats[1].item() - X.sum().item()) < 1e-5 and stats[0].item() == 8.0
        out[rank] = (err, gm.item(), ok_stats, dmax.item())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_gradient_and_scalar_collectives():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, gm, ok_stats, dmax = out[rank]
        assert err < 1e-6          # all-reduced local-mean gradients == global-mean gradient
        assert abs(gm - 1.5) < 1e-6
        assert ok_stats
        assert dmax == 1.5         # max over ranks {0.5, 1.5}
