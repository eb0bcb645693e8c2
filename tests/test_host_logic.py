"""Host-side mirrors of the reference interface (no GPU): ``tce_rl_b200.util`` helpers against the fixtures
produced by the REAL reference code, the bit-exact time-pair sampling, factories and their error behaviour."""
import pytest
import torch

from tce_rl_b200 import util
from tce_rl_b200.rl.agent import SegmentTimeSampler


def test_util_matches_reference_fixtures(golden):
    g = golden("ref_util.pt")
    r = g["build_lower_matrix"]
    assert torch.equal(util.build_lower_matrix(r["diag"], r["off"]), r["L"])
    d, off = util.reverse_build_matrix(r["L"], True)
    assert torch.equal(d, g["reverse_build_matrix"]["diag"]) and torch.equal(off, g["reverse_build_matrix"]["off"])
    r = g["add_expand_dim"]
    assert torch.equal(util.add_expand_dim(r["x"], [1, 3, 5], [2, 3, 5]), r["a"])
    assert torch.equal(util.add_expand_dim(r["x"], [1, -3, -1], [2, 3, 5]), r["b"])
    assert torch.equal(util.add_expand_dim(r["x"], [-2], [7]), r["c"])
    assert util.add_expand_dim(r["x"], [-2], [7]).stride(-2) == 0            # a view, like the reference
    r = g["tensor_linspace"]
    assert torch.equal(util.tensor_linspace(0, r["end"].clone(), 11), r["out"])
    r = g["softplus"]
    assert torch.equal(util.to_softplus_space(r["x"], None), r["none"])
    assert torch.equal(util.to_softplus_space(r["x"], 2.0), r["two"])
    assert torch.equal(util.reverse_from_softplus_space(r["none"], None), r["inv"])


def test_time_pair_sampling_is_bit_exact(golden):
    """Index sampling runs on the host torch generator exactly like util_learning.py:74-150."""
    for (T, seed), want in golden("ref_util.pt")["select_pred_pairs"].items():
        torch.manual_seed(seed)
        if T == "random":
            got = util.select_pred_pairs(num_all=100, num_select=10, fixed_interval=False).to(torch.long)
        else:
            s = SegmentTimeSampler(0.02, T, dict(num_select=25, fixed_interval=True), device="cpu")
            got = s.get_time_pairs()
            assert s.pred_pairs is got
        assert got.dtype == torch.long and torch.equal(got, want)


def test_get_times_matches_reference(golden):
    r = golden("ref_util.pt")["get_times"]
    s = SegmentTimeSampler(r["dt"], r["T"], dict(num_select=25, fixed_interval=True), device="cpu", dtype=torch.float64)
    assert torch.equal(s.get_times(r["init_time"], r["T"]), r["out"])


def test_select_ctx_pred_pts_branches():
    torch.manual_seed(0)
    ctx, pred = util.select_ctx_pred_pts(num_ctx=3, num_all=50, num_select=10, fixed_interval=True, first_index=2,
                                         ctx_before_pred=True)
    assert ctx.tolist() == [2, 7, 12] and pred.tolist() == list(range(17, 50, 5))
    with pytest.raises(AssertionError):
        util.select_ctx_pred_pts(num_ctx=0, num_all=10, num_select=11)
    with pytest.raises(AssertionError):
        util.select_ctx_pred_pts(num_ctx=0, num_all=100, num_select=25, fixed_interval=True, first_index=4)


def test_mlp_arch_and_mlp():
    assert util.mlp_arch_3_params(128, 2, 0.0) == [128, 128]
    assert util.mlp_arch_3_params(100, 3, -1.0) == [200, 100, 1]        # contracting, last layer clamps to 1
    assert util.mlp_arch_3_params(100, 3, 1.0) == [1, 100, 200]
    torch.manual_seed(0)
    net = util.MLP("m", 5, 3, [16, 16], "orthogonal", 0.01, "leaky_relu", None, dtype=torch.float64)
    assert [tuple(l.weight.shape) for l in net.layers] == [(16, 5), (16, 16), (3, 16)]
    assert all(float(l.bias.detach().abs().max()) == 0.0 for l in net.layers)
    w = net.layers[0].weight                       # orthogonal init with gain sqrt(2): columns orthogonal
    assert torch.allclose(w.T @ w, 2.0 * torch.eye(5, dtype=torch.float64), atol=1e-10)
    assert net.layers[-1].weight.norm() < 0.05      # out_layer_gain = 0.01
    assert net(torch.zeros(4, 5, dtype=torch.float64)).shape == (4, 3)
    with pytest.raises(ValueError):
        util.MLP("m", 5, 3, [16], "uniform", 1.0, "tanh", None)


def test_parse_dtype_device():
    assert util.parse_dtype_device("float32", "cpu") == (torch.float32, torch.device("cpu"))
    assert util.parse_dtype_device("torch.float64", "cuda")[0] == torch.float64
    with pytest.raises(NotImplementedError):
        util.parse_dtype_device("float16", "cpu")


def test_factories_refuse_what_is_out_of_scope():
    from tce_rl_b200.rl import agent_factory, projection_factory
    with pytest.raises(NotImplementedError):
        projection_factory("PAPIProjection", device="cuda", dtype="float32")
    with pytest.raises(NotImplementedError):
        projection_factory("KLProjectionLayer", device="cpu", dtype="float32")          # no CPU path
    with pytest.raises(NotImplementedError):
        agent_factory("SomeOtherAgent")
    from tce_rl_b200.rl import BlackBoxAgent, TemporalCorrelatedAgent          # both agents of the reference are built
    assert issubclass(BlackBoxAgent, TemporalCorrelatedAgent)
    layer = projection_factory("KLProjectionLayer", device="cuda", dtype="float32", mean_bound=0.05, cov_bound=5e-4,
                               entropy_schedule="linear", action_dim=63, total_train_steps=7500)
    layer.initial_entropy = torch.tensor(3.0)
    layer.initial_entropy = torch.tensor(5.0)                                           # write once
    assert float(layer.initial_entropy) == 3.0
    beta0 = layer.entropy_schedule(layer.initial_entropy, layer.target_entropy, layer.temperature, 0)
    beta_end = layer.entropy_schedule(layer.initial_entropy, layer.target_entropy, layer.temperature, 7500)
    assert float(beta0) == 3.0 and abs(float(beta_end)) < 1e-6


def test_ops_reject_cpu_tensors():
    from tce_rl_b200 import ops
    from tce_rl_b200._lib import TceError
    x = torch.zeros(2, 4)
    with pytest.raises(TceError):
        ops.gae(x, torch.zeros(2, 5), torch.zeros(2, 4, dtype=torch.bool), torch.zeros(2, 4, dtype=torch.bool), 1.0,
                0.95, True)
    with pytest.raises(TceError):
        ops.gauss_maha(torch.zeros(2, 3), torch.zeros(2, 3), torch.eye(3).expand(2, 3, 3))


def test_shared_kl_loss_node_matches_the_plain_formula():
    """``_SharedKLLoss`` (one autograd node, constant gradients) == coeff * (mean 1/2 maha + shape + volume) written with
    plain torch ops: values, detached pieces and gradients, with the cached unit seed and with a general seed."""
    from tce_rl_b200 import ops
    from tce_rl_b200.rl.projection import _SharedKLLoss
    torch.manual_seed(0)
    B, k, coeff = 17, 6, 0.8
    for with_cov in (True, False):
        for seed_scale in (None, 2.5):
            maha = torch.rand(B, dtype=torch.float64).requires_grad_(True)
            st = torch.rand(1, 5, dtype=torch.float64).requires_grad_(True)
            loss, mean_diff, cov_diff, shape, volume = _SharedKLLoss.apply(maha, st, coeff, with_cov, k, torch.float32)
            m2, s2 = maha.detach().clone().requires_grad_(True), st.detach().clone().requires_grad_(True)
            sh, vo = 0.5 * (s2[:, 1] - k), 0.5 * (s2[:, 3] - s2[:, 2])
            ref = ((0.5 * m2 + (sh + vo if with_cov else 0.0)).mean() * coeff).to(torch.float32)
            assert loss.dtype == torch.float32 and abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
            assert torch.allclose(mean_diff, 0.5 * m2.detach()) and torch.allclose(cov_diff, (sh + vo).detach())
            assert torch.allclose(shape, sh.detach()) and torch.allclose(volume, vo.detach())
            if seed_scale is None:
                loss.backward(ops.unit_seed(loss.device, loss.dtype))
                ref.backward()
            else:
                loss.backward(torch.tensor(seed_scale))
                ref.backward(torch.tensor(seed_scale))
            assert torch.allclose(maha.grad, m2.grad, rtol=1e-6, atol=0)
            if with_cov:
                assert torch.allclose(st.grad, s2.grad, rtol=1e-6, atol=1e-12)
            else:
                assert float(st.grad.abs().max()) == 0.0 and s2.grad is None or float(s2.grad.abs().max()) == 0.0


def test_stats_to_float_node():
    from tce_rl_b200 import ops
    stats = torch.tensor([0.25, 1.5], dtype=torch.float64, requires_grad=True)
    loss, ratio = ops._StatsToFloat.apply(stats)
    assert loss.dtype == torch.float32 and float(loss.detach()) == 0.25 and float(ratio.detach()) == 1.5
    loss.backward(ops.unit_seed(loss.device, loss.dtype))
    assert stats.grad.tolist() == [1.0, 0.0]
    stats.grad = None
    loss2, _ = ops._StatsToFloat.apply(stats)
    (3.0 * loss2).backward()
    assert stats.grad.tolist() == [3.0, 0.0]


def test_flat_adam_and_loader_guard_rails():
    from tce_rl_b200._lib import TceError
    from tce_rl_b200.rl.optim import FlatAdam
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(TceError):
        FlatAdam([p], torch.zeros(4))                       # CPU buffers: there is no CPU path
    from tce_rl_b200.rl import projection as pj
    L1 = torch.eye(3).reshape(1, 3, 3)
    e = pj._expand_first(L1, 5)
    assert e.shape == (5, 3, 3) and e.stride(0) == 0 and pj._first(e) is L1
    dense = torch.eye(3).expand(5, 3, 3).contiguous()
    assert pj._first(dense).shape == (1, 3, 3)
