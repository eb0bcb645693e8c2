"""Value-function critic (mprl/rl/critic/abstract_critic.py:8-120, value_function_critic.py:4-17).
A plain MLP on cuBLAS through torch: outside the four hand-written subsystems."""
from __future__ import annotations

from .. import util


class ValueFunction:
    def __init__(self, dim_in, dim_out, hidden, init_method, out_layer_gain, act_func_hidden, act_func_last,
                 dtype="torch.float32", device="cuda", **kwargs):
        self.dtype, self.device = util.parse_dtype_device(dtype, device)
        self.net = util.MLP(name="ValueFunction", dim_in=dim_in, dim_out=dim_out,
                            hidden_layers=util.mlp_arch_3_params(**hidden), init_method=init_method,
                            out_layer_gain=out_layer_gain, act_func_hidden=act_func_hidden,
                            act_func_last=act_func_last, dtype=self.dtype, device=self.device)

    @property
    def parameters(self):
        return list(self.net.parameters())

    def critic(self, state):
        return self.net(state)

    def save_weights(self, log_dir: str, epoch: int):
        """abstract_critic.py:83-105, files in the reference's layout."""
        from .. import rollout
        rollout.save_net(self.net, log_dir, epoch)

    def load_weights(self, log_dir: str, epoch: int):
        from .. import rollout
        rollout.load_net(self.net, log_dir, epoch)


def critic_factory(typ: str, **kwargs):
    return {"ValueFunction": ValueFunction}[typ](**kwargs)
