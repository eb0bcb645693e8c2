"""2-rank check (run under torchrun, NCCL): the data-parallel policy update -- episodes sharded over the ranks, the
likelihood regulariser MAX-reduced (or provably rank-independent), advantage statistics SUM-reduced, gradients
AVG-reduced -- reproduces the single-GPU update on the concatenated batch: segment advantages, per-epoch losses and
the updated parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_sharded_epoch.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from tce_rl_b200 import ops
from tce_rl_b200.rl import TemporalCorrelatedAgent, policy_factory, projection_factory
from tce_rl_b200.rl.agent import SegmentTimeSampler


def make_agent(device, group, graph, spread):
    cfg, T, d, k1, dp = bench.shape_dims("box")
    torch.manual_seed(0)
    policy = policy_factory("TemporalCorrelatedPolicy", dim_in=bench.OBS_DIM, dim_out=dp, dtype="float32",
                            device=device, mp=dict(type="prodmp", args=dict(cfg)), **bench.POLICY)
    proj = projection_factory("KLProjectionLayer", device=device, dtype="float32", action_dim=dp, **bench.PROJ)
    sampler = SegmentTimeSampler(cfg["dt"], T, dict(num_select=25, fixed_interval=True), device=device)
    torch.manual_seed(0)
    sampler.get_time_pairs()
    agent = TemporalCorrelatedAgent(policy, None, sampler, proj, dtype="float32", device=device, process_group=group,
                                    use_cuda_graph=graph, **dict(bench.AGENT, epochs_policy=4))
    return agent


def full_dataset(agent, device, B, spread):
    """The rollout data of ALL ranks (same seed everywhere)."""
    cfg, T, d, k1, dp = bench.shape_dims("box")
    host = bench.synthetic_host_data(B, seed=99)
    c = lambda t: t.to(device)
    pol, sampler = agent.policy, agent.sampler
    init_time = c(host["init_time"])
    if spread:
        init_time = init_time + spread * torch.rand(B, generator=torch.Generator().manual_seed(5)).to(device)
    with torch.no_grad():
        times = sampler.get_times(init_time, T)
        mean0, L0 = pol.policy(c(host["obs"])[..., :-2 * d])
        mean_old = mean0 + c(host["mean_noise"])
        L_old = (1.05 * L0[:1] + c(host["L_noise"])).expand(B, -1, -1).contiguous()
        smp = pol.sample(False, mean_old, L_old, times, init_time, c(host["init_pos"]), c(host["init_vel"]),
                         eps=c(host["eps"]))
        lp_old = pol.log_prob(smp, mean_old, L_old, times, init_time, c(host["init_pos"]), c(host["init_vel"]),
                              pred_pairs=sampler.pred_pairs)
    ds = dict(segment_state=c(host["obs"]), step_actions=smp, segment_log_prob_estimate=lp_old,
              segment_params_mean=mean_old, segment_params_L=L_old, segment_init_time=init_time,
              segment_init_pos=c(host["init_pos"]), segment_init_vel=c(host["init_vel"]), step_rewards=c(host["rewards"]),
              step_values=c(host["values"]), step_dones=c(host["dones"]),
              step_time_limit_dones=c(host["time_limit_dones"]))
    return ds


def perturb(agent):
    torch.manual_seed(7)
    with torch.no_grad():
        for p in agent.policy.mean_net.parameters():
            p.add_(0.05 * torch.randn_like(p))
        agent.policy.variance_net.variable.add_(0.02 * torch.randn_like(agent.policy.variance_net.variable))


def run(agent, ds):
    agent.num_iterations = 100
    ds = agent.process_dataset(dict(ds))
    out = agent.update_policy(ds)
    return ds["segment_advantage"], out, [p.detach().clone() for p in agent.policy.parameters]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    Bs = 256
    B = Bs * world
    results = []
    for spread, graph in ((0.0, False), (0.0, True), (0.3, False)):
        # ---- single GPU on the concatenated batch (no process group anywhere) -----------------------------------------
        ops.set_regulariser_group(None)
        ops.set_stats_group(None)
        ref = make_agent(device, None, graph, spread)
        ds_full = full_dataset(ref, device, B, spread)
        perturb(ref)
        adv_ref, out_ref, par_ref = run(ref, ds_full)
        # ---- sharded ----------------------------------------------------------------------------------------------------
        agent = make_agent(device, True, graph, spread)
        perturb(agent)
        sl = slice(rank * Bs, (rank + 1) * Bs)
        ds = {k: (v[sl].contiguous() if torch.is_tensor(v) else v) for k, v in ds_full.items()}
        adv, out, par = run(agent, ds)
        err_adv = (adv - adv_ref[sl]).abs().max().item()
        err_par = max((a - b).abs().max().item() for a, b in zip(par, par_ref))
        # rank-local logging means average to the global ones
        keys = ["surrogate_loss_mean", "trust_region_loss_mean", "policy_grad_norm_mean", "imp_smp_ratio_mean",
                "projection_new_old_mean_diff_mean", "projection_proj_old_cov_diff_mean"]
        loc = torch.tensor([out[k] for k in keys], device=device, dtype=torch.float64)
        dist.all_reduce(loc, op=dist.ReduceOp.SUM)
        loc /= world
        refv = torch.tensor([out_ref[k] for k in keys], dtype=torch.float64)
        err_met = ((loc.cpu() - refv).abs() / refv.abs().clamp_min(1.0))
        moved = max((a - b).abs().max().item() for a, b in zip(par_ref, [p for p in make_agent(device, None, False, 0).policy.parameters]))
        errs = torch.tensor([err_adv, err_par, err_met.max().item()], device=device, dtype=torch.float64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        results.append(dict(init_time_spread=spread, cuda_graph=graph, world=world, episodes_per_rank=Bs,
                            max_err_segment_advantage=errs[0].item(), max_err_updated_parameters=errs[1].item(),
                            max_rel_err_logging_means=errs[2].item(), parameters_moved_by=moved,
                            uniform_everywhere=bool(__import__("tce_rl_b200.ops_seglik", fromlist=["x"])._GLOBAL_UNIFORM),
                            grad_exchange="p2p_fused" if agent._p2p is not None else "nccl", epochs=4))
    ok = all(r["max_err_segment_advantage"] <= 1e-5 and r["max_err_updated_parameters"] <= 1e-6
             and r["max_rel_err_logging_means"] <= 1e-5 for r in results)
    if rank == 0:
        line = json.dumps(dict(check="sharded policy update == single-GPU update on the concatenated batch", ok=ok,
                               cases=results))
        print(line)
        if len(sys.argv) > 1:
            open(sys.argv[1], "w").write(line + "\n")
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
