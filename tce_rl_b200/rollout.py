"""Rollout-side glue of the update path: what turns the sampler's per-environment lists into the
device-resident dataset ``update_critic`` / ``update_policy`` consume, and the checkpoint file layout.

Reference: ``RunningMeanStd`` (mprl/util/util_numerical.py:278-350), ``apply_normalization``
(mprl/rl/sampler/temporal_correlated_sampler.py:87-89), ``make_mdp_reward`` (mprl/util/util_experiment.py:261-328), the
result dictionary of ``TemporalCorrelatedSampler.run`` (temporal_correlated_sampler.py:318-337), checkpoint paths
(mprl/util/util_file.py:280-317) and ``save_agent`` / ``load_agent`` (mprl/rl/agent/abstract_agent.py:109-174).
Environment stepping itself (MuJoCo / fancy_gym / stable-baselines3) stays on the CPU and is out of scope.

Everything here is elementwise / reduction work on tensors that already live on the device (a few launches per rollout,
not per epoch), expressed with torch tensor ops; no host synchronisation except where the reference has one.
"""
from __future__ import annotations

import os
import pickle as pkl
from typing import Optional, Tuple, Union

import torch

from . import util


# ---- checkpoint paths (util_file.py:280-317) -------------------------------------------------------------------------
def get_nn_save_paths(log_dir: str, nn_name: str, epoch: Optional[int]) -> Tuple[str, str]:
    s_path = os.path.join(log_dir, nn_name + "_parameters.pkl")
    w_path = os.path.join(log_dir, nn_name + "_weights")
    if epoch is not None:
        w_path = w_path + "_{:d}".format(epoch)
    return s_path, w_path


def get_training_state_save_path(log_dir: str, name: str, epoch: Optional[int]) -> str:
    o_path = os.path.join(log_dir, name + "_state")
    if epoch is not None:
        o_path = o_path + "_{:d}".format(epoch)
    return o_path


# ---- running observation statistics (util_numerical.py:278-350) ------------------------------------------------------
class RunningMeanStd:
    """Running mean / variance of a data stream (parallel-variance update), state on the device."""

    def __init__(self, name: str = "", epsilon: float = 1e-4, shape: Tuple[int, ...] = (),
                 dtype: str = "torch.float32", device: str = "cuda"):
        self.name = "running_mean_std" if name == "" else name
        self.shape = shape
        self.dtype, self.device = util.parse_dtype_device(dtype, device)
        self.mean = torch.zeros(shape, dtype=self.dtype, device=self.device)
        self.var = torch.ones(shape, dtype=self.dtype, device=self.device)
        self.count = epsilon

    def copy(self) -> "RunningMeanStd":
        new = RunningMeanStd(shape=self.mean.shape, dtype=self.dtype, device=self.device)
        new.mean, new.var, new.count = self.mean.clone(), self.var.clone(), float(self.count)
        return new

    def combine(self, other: "RunningMeanStd") -> None:
        self.update_from_moments(other.mean, other.var, other.count)

    def update(self, arr: torch.Tensor) -> None:
        self.update_from_moments(torch.mean(arr, dim=0), torch.var(arr, dim=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count: Union[int, float]) -> None:
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_2 = self.var * self.count + batch_var * batch_count \
            + torch.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), batch_count + self.count

    def save(self, log_dir: str, epoch: int):
        with open(get_training_state_save_path(log_dir, self.name, epoch), "wb") as f:
            torch.save({"mean": self.mean, "var": self.var, "count": self.count}, f)

    def load(self, log_dir: str, epoch: int):
        d = torch.load(get_training_state_save_path(log_dir, self.name, epoch), map_location=self.device,
                       weights_only=False)
        self.mean, self.var, self.count = d["mean"], d["var"], d["count"]


def apply_normalization(raw: torch.Tensor, rms: RunningMeanStd) -> torch.Tensor:
    return (raw - rms.mean) / torch.sqrt(rms.var + 1e-8)


# ---- non-MDP -> MDP rewards (util_experiment.py:261-328) -------------------------------------------------------------
def get_item_from_dicts(dicts, key, to_value=lambda x: x):
    """mprl/util/util_data_structure.py: collect ``d[key]`` over a list of dicts."""
    return [to_value(d[key]) for d in dicts]


def make_mdp_reward(task_id: str, step_rewards: torch.Tensor, step_infos, dtype=None, device=None) -> torch.Tensor:
    """Sum the rewards from the key event on (table-tennis ball hit, hopper leaving the floor) into the step of the event
    and zero everything after it; other tasks: unchanged.  ``step_infos``: the list of per-environment info dicts of
    the reference, or directly the [num_env, num_times] boolean event tensor.  Modifies ``step_rewards`` in place and
    returns it, like the reference."""
    if "TableTennis" in task_id:
        key = "hit_ball"
    elif "HopperJump" in task_id:
        key = "has_left_floor"
    else:
        return step_rewards
    device = step_rewards.device if device is None else device
    event = step_infos if torch.is_tensor(step_infos) else torch.as_tensor(
        __import__("numpy").asarray(get_item_from_dicts(step_infos, key)), device=device)
    event_index = torch.where(event.to(device), 1.0, 0.0).to(step_rewards.dtype)
    first = torch.argmax(event_index, dim=-1)
    after = (step_rewards * event_index).sum(dim=-1)
    happened = first > 0
    # no boolean-mask indexing (a device->host sync in torch): one scatter and one masked fill
    cols = first.unsqueeze(1)
    cur = step_rewards.gather(1, cols)
    step_rewards.scatter_(1, cols, torch.where(happened.unsqueeze(1), after.unsqueeze(1), cur))
    mask = torch.arange(step_rewards.size(1), device=device).unsqueeze(0) > cols
    step_rewards.masked_fill_(torch.logical_and(happened.unsqueeze(1), mask), 0)
    return step_rewards


# ---- dataset assembly (temporal_correlated_sampler.py:232-337) -------------------------------------------------------
def assemble_dataset(rollouts, task_specified_metrics=None) -> dict:
    """The result dictionary of ``TemporalCorrelatedSampler.run`` from per-iteration pieces that already live on the
    device.  ``rollouts``: list of dicts with keys step_actions [E,T,2D], segment_log_prob_estimate [E,P],
    step_states [E,T+1,obs] (initial state included, normalised), step_rewards [E,T], episode_init_state [E,obs],
    episode_reward [E], step_dones [E,T] bool, step_values [E,T+1], init_time [E], init_pos [E,D], init_vel [E,D],
    params_mean [E,Dp], params_L [E,Dp,Dp] (+ task metrics)."""
    cat = lambda k: torch.cat([r[k] for r in rollouts], dim=0)
    res = dict()
    res["step_actions"] = cat("step_actions")
    res["segment_log_prob_estimate"] = cat("segment_log_prob_estimate")
    res["step_states"] = cat("step_states")[:, :-1]
    res["step_rewards"] = cat("step_rewards")
    res["segment_state"] = cat("episode_init_state")
    res["segment_reward"] = res["step_rewards"].sum(dim=-1)
    res["episode_reward"] = cat("episode_reward")
    res["step_dones"] = cat("step_dones")
    res["step_values"] = cat("step_values")
    res["segment_init_time"] = cat("init_time")
    res["segment_init_pos"] = cat("init_pos")
    res["segment_init_vel"] = cat("init_vel")
    res["step_time_limit_dones"] = torch.zeros_like(res["step_dones"], dtype=torch.bool)
    res["segment_params_mean"] = cat("params_mean")
    res["segment_params_L"] = cat("params_L")
    for metric in task_specified_metrics or ():
        res[metric] = cat(metric)
    return res


# ---- network / optimiser checkpoints in the reference's file layout -------------------------------------------------
def save_mlp(mlp: "util.MLP", log_dir: str, epoch: int) -> None:
    """``MLP.save`` (mprl/util/util_nn.py:164-193): ``<name>_parameters.pkl`` (structure) + ``<name>_weights_<epoch>``
    (state_dict whose keys carry the reference's module name ``<name>.<i>.weight``)."""
    s_path, w_path = get_nn_save_paths(log_dir, mlp.mlp_name, epoch)
    lay = list(mlp.layers)
    with open(s_path, "wb") as f:
        pkl.dump({"dim_in": lay[0].in_features, "dim_out": lay[-1].out_features,
                  "hidden_layers": [l.out_features for l in lay[:-1]], "act_func_hidden_type": mlp.act_hidden_name,
                  "act_func_last_type": mlp.act_last_name, "dtype": lay[0].weight.dtype,
                  "device": lay[0].weight.device}, f)
    sd = {k.replace("layers.", mlp.mlp_name + ".", 1): v for k, v in mlp.state_dict().items()}
    with open(w_path, "wb") as f:
        torch.save(sd, f)


def load_mlp(mlp: "util.MLP", log_dir: str, epoch: int) -> None:
    s_path, w_path = get_nn_save_paths(log_dir, mlp.mlp_name, epoch)
    lay = list(mlp.layers)
    with open(s_path, "rb") as f:
        p = pkl.load(f)
    assert (lay[0].in_features == p["dim_in"] and lay[-1].out_features == p["dim_out"]
            and [l.out_features for l in lay[:-1]] == list(p["hidden_layers"])
            and mlp.act_hidden_name == p["act_func_hidden_type"] and mlp.act_last_name == p["act_func_last_type"]), \
        "NN structure parameters do not match"
    sd = torch.load(w_path, map_location=lay[0].weight.device, weights_only=False)
    mlp.load_state_dict({k.replace(mlp.mlp_name + ".", "layers.", 1): v for k, v in sd.items()})


def save_variable(var: "util.TrainableVariable", log_dir: str, epoch: int) -> None:
    """``TrainableVariable.save`` (util_nn.py:465-492)."""
    name = var.name + "_variable"
    s_path, w_path = get_nn_save_paths(log_dir, name, epoch)
    with open(s_path, "wb") as f:
        pkl.dump({"variable_name": name, "variable_shape": var.variable.shape, "dtype": var.variable.dtype,
                  "device": var.variable.device}, f)
    with open(w_path, "wb") as f:
        torch.save(var.variable, f)


def load_variable(var: "util.TrainableVariable", log_dir: str, epoch: int) -> None:
    name = var.name + "_variable"
    s_path, w_path = get_nn_save_paths(log_dir, name, epoch)
    with open(s_path, "rb") as f:
        p = pkl.load(f)
    assert name == p["variable_name"] and tuple(var.variable.shape) == tuple(p["variable_shape"]), \
        f"Variable {name}'s parameters do not match"
    loaded = torch.load(w_path, map_location=var.variable.device, weights_only=False)
    with torch.no_grad():                       # keep the Parameter object (optimiser / flat gradient views refer to it)
        var.variable.copy_(loaded.detach() if torch.is_tensor(loaded) else loaded)


def save_net(net, log_dir, epoch):
    (save_mlp if isinstance(net, util.MLP) else save_variable)(net, log_dir, epoch)


def load_net(net, log_dir, epoch):
    (load_mlp if isinstance(net, util.MLP) else load_variable)(net, log_dir, epoch)
