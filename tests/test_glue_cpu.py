"""Rollout-side glue (tce_rl_b200/rollout.py) and the oracle's agent pieces against fixtures produced by the REAL reference
code (``oracle/gen_golden.py: gen_ref_glue`` -> tests/golden/ref_glue.pt): RunningMeanStd, make_mdp_reward, checkpoint paths
and files, numpy minibatch shuffles, BlackBoxAgent.process_dataset and TemporalCorrelatedAgent.update_critic."""
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from oracle import agent as oa
from tce_rl_b200 import rollout, util


@pytest.fixture(scope="module")
def glue(golden):
    return golden("ref_glue.pt")


def test_running_mean_std_matches_reference(glue):
    rec = glue["rms"]
    rms = rollout.RunningMeanStd(name="obs", shape=(5,), dtype="torch.float64", device="cpu")
    for b, want in zip(rec["batches"][:2], rec["hist"][:2]):
        rms.update(b)
        assert torch.equal(rms.mean, want["mean"]) and torch.equal(rms.var, want["var"]) and rms.count == want["count"]
    other = rollout.RunningMeanStd(shape=(5,), dtype="torch.float64", device="cpu")
    other.update(rec["batches"][2])
    cp = rms.copy()
    rms.combine(other)
    want = rec["hist"][2]
    assert torch.equal(rms.mean, want["mean"]) and torch.equal(rms.var, want["var"]) and rms.count == want["count"]
    assert not torch.equal(cp.mean, rms.mean)                      # copy() is detached from the original
    x = rec["batches"][0]
    assert torch.equal(rollout.apply_normalization(x, rms), (x - rms.mean) / torch.sqrt(rms.var + 1e-8))


def test_running_mean_std_save_load(glue, tmp_path):
    rms = rollout.RunningMeanStd(name="obs", shape=(5,), dtype="torch.float64", device="cpu")
    rms.update(glue["rms"]["batches"][1])
    rms.save(str(tmp_path), 4)
    assert os.path.exists(tmp_path / "obs_state_4")                 # get_training_state_save_path layout
    new = rollout.RunningMeanStd(name="obs", shape=(5,), dtype="torch.float64", device="cpu")
    new.load(str(tmp_path), 4)
    assert torch.equal(new.mean, rms.mean) and torch.equal(new.var, rms.var) and new.count == rms.count


def test_make_mdp_reward_matches_reference(glue):
    rec = glue["mdp"]
    infos = [{"hit_ball": rec["event"][e].tolist(), "has_left_floor": rec["event"][e].flip(0).tolist()}
             for e in range(rec["event"].shape[0])]
    got = rollout.make_mdp_reward("TableTennis4D-v0", rec["rewards"].clone(), infos)
    assert torch.equal(got, rec["table_tennis"])
    assert torch.equal(rollout.make_mdp_reward("TableTennis4D-v0", rec["rewards"].clone(), rec["event"]), rec["table_tennis"])
    assert torch.equal(rollout.make_mdp_reward("HopperJumpSparse", rec["rewards"].clone(), infos), rec["hopper"])
    assert torch.equal(rollout.make_mdp_reward("BoxPushingDense", rec["rewards"].clone(), infos), rec["other"])
    assert not torch.equal(rec["table_tennis"], rec["rewards"])   # the fixture exercises the event branch


def test_checkpoint_paths_match_reference(glue):
    p = glue["paths"]
    assert rollout.get_nn_save_paths("/log", "policy_mean_mlp", 12) == tuple(p["nn"])
    assert rollout.get_nn_save_paths("/log", "x", None) == tuple(p["nn_none"])
    assert rollout.get_training_state_save_path("/log", "policy_optimizer", 3) == p["state"]
    assert rollout.get_training_state_save_path("/log", "obs", None) == p["state_none"]


def test_reference_checkpoint_loads_and_our_checkpoint_has_the_reference_layout(glue, tmp_path):
    rec = glue["ckpt"]
    d = str(tmp_path)
    # a checkpoint as the REFERENCE writes it (structure pickle + state_dict keyed "<name>_mlp.<i>.weight")
    structure = dict(rec["structure"], dtype=torch.float64, device=torch.device("cpu"))
    with open(os.path.join(d, "critic_net_mlp_parameters.pkl"), "wb") as f:
        pkl.dump(structure, f)
    torch.save(rec["weights"], os.path.join(d, "critic_net_mlp_weights_7"))
    mlp = util.MLP(name="critic_net", dim_in=4, dim_out=1, hidden_layers=[8, 6], init_method="orthogonal",
                   out_layer_gain=1.0, act_func_hidden="leaky_relu", act_func_last=None, dtype=torch.float64, device="cpu")
    rollout.load_mlp(mlp, d, 7)
    assert torch.equal(mlp(rec["x"]).detach(), rec["y"])
    # ... and ours, written next to it, has the same files / keys / structure entries
    d2 = str(tmp_path / "ours")
    os.makedirs(d2)
    rollout.save_mlp(mlp, d2, 7)
    var = util.TrainableVariable("cov", torch.arange(5, dtype=torch.float64))
    rollout.save_variable(var, d2, 7)
    assert sorted(os.listdir(d2)) == rec["files"]
    w = torch.load(os.path.join(d2, "critic_net_mlp_weights_7"), weights_only=False)
    assert list(w.keys()) == list(rec["weights"].keys()) and all(torch.equal(w[k], rec["weights"][k]) for k in w)
    with open(os.path.join(d2, "critic_net_mlp_parameters.pkl"), "rb") as f:
        s = pkl.load(f)
    assert {k: (str(v) if k in ("dtype", "device") else v) for k, v in s.items()} == rec["structure"]
    with open(os.path.join(d2, "cov_variable_parameters.pkl"), "rb") as f:
        vs = pkl.load(f)
    assert vs["variable_name"] == rec["var_structure"]["variable_name"] and tuple(vs["variable_shape"]) == (5,)
    var2 = util.TrainableVariable("cov", torch.zeros(5, dtype=torch.float64))
    handle = var2.variable
    rollout.load_variable(var2, d2, 7)
    assert var2.variable is handle and torch.equal(var2.variable.detach(), rec["var_saved"])
    wrong = util.MLP(name="critic_net", dim_in=4, dim_out=1, hidden_layers=[8, 5], init_method="orthogonal",
                     out_layer_gain=1.0, act_func_hidden="leaky_relu", act_func_last=None, dtype=torch.float64, device="cpu")
    with pytest.raises(AssertionError):
        rollout.load_mlp(wrong, d, 7)


def test_minibatch_shuffle_is_the_reference_sequence(glue):
    np.random.seed(7)
    for want in glue["minibatches"]:
        got = oa.generate_minibatches(23, 4)
        assert len(got) == len(want) and all(np.array_equal(g, w.numpy()) for g, w in zip(got, want))


def test_bbrl_process_dataset_matches_reference(glue):
    for n in (9, 1):
        rec = glue[f"bbrl_process_{n}"]
        got = oa.bbrl_process_dataset(rec["inputs"], norm_advantages=True, clip_advantages=1.5)
        assert torch.equal(got, rec["advantage"])


@pytest.mark.parametrize("clip_critic,clip_norm", [(0.0, 0.0), (0.2, 0.5)])
def test_oracle_update_critic_matches_reference(glue, clip_critic, clip_norm):
    """The oracle restatement that the GPU update_critic is tested against, pinned to the real
    TemporalCorrelatedAgent.update_critic (losses, gradient norms, final weights)."""
    inp, rec = glue["update_critic_inputs"], glue[f"update_critic_{clip_critic}_{clip_norm}"]
    net = util.MLP(name="ValueFunction", dim_in=6, dim_out=1, hidden_layers=[16, 16], init_method="orthogonal",
                   out_layer_gain=1.0, act_func_hidden="leaky_relu", act_func_last=None, dtype=torch.float64, device="cpu")
    net.load_state_dict({k.replace("ValueFunction_mlp.", "layers."): v for k, v in inp["w0"].items()})
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=5e-5)
    np.random.seed(99)
    losses, norms, clipped = oa.update_critic(net, opt, inp["dataset"], 3, 4, inp["num_dof"], clip_critic, clip_norm)
    st = rec["stats"]
    assert np.isclose(np.mean(losses), st["critic_loss_mean"], rtol=1e-12) and np.isclose(np.max(losses), st["critic_loss_max"], rtol=1e-12)
    assert np.isclose(np.mean(norms), st["critic_grad_norm_mean"], rtol=1e-10)
    assert np.isclose(np.mean(clipped), st["clipped_critic_grad_norm_mean"], rtol=1e-6)
    for k, v in rec["weights"].items():
        assert torch.allclose(net.state_dict()[k.replace("ValueFunction_mlp.", "layers.")], v, rtol=0, atol=1e-12), k


def test_assemble_dataset_keys_and_shapes():
    E, T, D, Dp, P, obs = 3, 10, 2, 6, 4, 5
    mk = lambda: dict(step_actions=torch.randn(E, T, 2 * D), segment_log_prob_estimate=torch.randn(E, P),
                      step_states=torch.randn(E, T + 1, obs), step_rewards=torch.randn(E, T),
                      episode_init_state=torch.randn(E, obs), episode_reward=torch.randn(E),
                      step_dones=torch.zeros(E, T, dtype=torch.bool), step_values=torch.randn(E, T + 1),
                      init_time=torch.zeros(E), init_pos=torch.randn(E, D), init_vel=torch.randn(E, D),
                      params_mean=torch.randn(E, Dp), params_L=torch.randn(E, Dp, Dp), success=torch.ones(E))
    parts = [mk(), mk()]
    ds = rollout.assemble_dataset(parts, task_specified_metrics=["success"])
    want = {"step_actions", "segment_log_prob_estimate", "step_states", "step_rewards", "segment_state", "segment_reward",
            "episode_reward", "step_dones", "step_values", "segment_init_time", "segment_init_pos", "segment_init_vel",
            "step_time_limit_dones", "segment_params_mean", "segment_params_L", "success"}   # sampler.run :318-337
    assert set(ds) == want
    assert ds["step_states"].shape == (2 * E, T, obs) and ds["step_values"].shape == (2 * E, T + 1)
    assert torch.equal(ds["segment_reward"], ds["step_rewards"].sum(-1)) and not ds["step_time_limit_dones"].any()
    assert torch.equal(ds["step_states"][:E], parts[0]["step_states"][:, :-1])
