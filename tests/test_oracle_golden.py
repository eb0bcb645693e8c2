"""Oracle vs the reference's own known answers and vs fixtures produced by the REAL reference code
(``oracle/gen_golden.py``, run in the build container)."""
import math

import pytest
import torch

from oracle import agent as oa
from oracle import policy as opol
from oracle import util as ou
from oracle.gen_golden import MP_CONFIGS


# ---- known answers quoted in the reference's tests ------------------------------------------------
def test_known_answer_build_lower_matrix():
    # mprl/test/util_test/util_matrix_test.py:7-20 : diag 0.5, off-diag 1..15 row-major strictly lower
    L = ou.build_lower_matrix(torch.ones(6) * 0.5, torch.arange(1, 16, dtype=torch.float32))
    assert torch.equal(L.diagonal(), torch.full((6,), 0.5))
    assert L[1, 0] == 1 and L[2, 0] == 2 and L[2, 1] == 3 and L[5, 4] == 15 and L[0, 1] == 0
    d, off = ou.reverse_build_matrix(L, True)
    assert torch.equal(off, torch.arange(1, 16, dtype=torch.float32))


def test_known_answer_softplus():
    # mprl/test/util_test/util_numerical_test.py:19-36
    z = torch.zeros(1, dtype=torch.float64)
    assert abs(ou.to_softplus_space(z, None).item() - 0.7031) < 1e-4
    assert abs(ou.to_softplus_space(z, 2.0).item() - 2.6931) < 1e-4
    assert abs(ou.reverse_from_softplus_space(ou.to_softplus_space(z, None), None).item()) < 1e-12


def test_known_answer_linspace_and_lerp():
    # util_matrix_test.py:89-105
    out = ou.tensor_linspace(0, torch.arange(0, 11, dtype=torch.float64), 11)
    assert out.shape == (11, 11) and torch.allclose(out[-1], torch.arange(0, 11, dtype=torch.float64))
    data = torch.arange(10, dtype=torch.float64)[:, None].expand(10, 2)
    idx = torch.tensor([[0.5, 1.5, 2.5, 3.5, 4.5]] * 3, dtype=torch.float64)
    r = ou.indexing_interpolate(data, idx)
    assert r.shape == (3, 5, 2) and torch.allclose(r[0, :, 0], idx[0])


def test_known_answer_first_index_mt19937():
    # CPU randint(0, r) == mt19937(seed) first u32 % r
    first_u32 = {0: 2357136044, 1: 1791095845, 2: 1872583848}
    for T, hi in ((100, 4), (500, 20), (350, 14)):
        for s, u in first_u32.items():
            torch.manual_seed(s)
            pairs = ou.select_pred_pairs(T, 25, True)
            assert int(pairs[0, 0]) == u % hi
            assert pairs.shape[0] == 24


# ---- fixtures produced by the real reference ---------------------------------------------------------
def test_ref_util(golden):
    g = golden("ref_util.pt")
    r = g["build_lower_matrix"]
    assert torch.equal(ou.build_lower_matrix(r["diag"], r["off"]), r["L"])
    d, off = ou.reverse_build_matrix(r["L"], True)
    assert torch.equal(d, g["reverse_build_matrix"]["diag"]) and torch.equal(off, g["reverse_build_matrix"]["off"])
    r = g["add_expand_dim"]
    assert torch.equal(ou.add_expand_dim(r["x"], [1, 3, 5], [2, 3, 5]), r["a"])
    assert torch.equal(ou.add_expand_dim(r["x"], [1, -3, -1], [2, 3, 5]), r["b"])
    assert torch.equal(ou.add_expand_dim(r["x"], [-2], [7]), r["c"])
    r = g["tensor_linspace"]
    assert torch.equal(ou.tensor_linspace(0, r["end"].clone(), 11), r["out"])
    for key in ("indexing_interpolate", "indexing_interpolate_edge"):
        r = g[key]
        assert torch.equal(ou.indexing_interpolate(r["data"], r["idx"]), r["out"])
    r = g["softplus"]
    assert torch.equal(ou.to_softplus_space(r["x"], None), r["none"])
    assert torch.equal(ou.to_softplus_space(r["x"], 2.0), r["two"])
    assert torch.equal(ou.reverse_from_softplus_space(r["none"], None), r["inv"])
    r = g["get_times"]
    assert torch.equal(ou.get_times(r["init_time"], r["T"], r["dt"]), r["out"])


def test_ref_select_pred_pairs_bit_exact(golden):
    for (T, seed), want in golden("ref_util.pt")["select_pred_pairs"].items():
        torch.manual_seed(seed)
        if T == "random":
            got = ou.get_time_pairs(100, dict(num_select=10, fixed_interval=False))
        else:
            got = ou.get_time_pairs(T, dict(num_select=25, fixed_interval=True))
        assert got.dtype == torch.long and torch.equal(got, want)


@pytest.mark.parametrize("tag", ["g1", "g099"])
def test_ref_agent(golden, tag):
    r = golden("ref_agent.pt")[tag]
    for use_gae in (True, False):
        adv, ret = oa.get_advantage_return(r["rewards"], r["values"], r["dones"], r["time_limit_dones"],
                                           r["gamma"], r["lam"], use_gae)
        assert torch.equal(adv, r[f"adv_gae{int(use_gae)}"]) and torch.equal(ret, r[f"ret_gae{int(use_gae)}"])
    for mode in ("accumulate", "value_subtraction", "accumulated_rewards"):
        for norm in (True, False):
            got = oa.get_segment_advantage(r["rewards"], r["values"], r["adv_gae1"], r["pred_pairs"],
                                           r["gamma"], mode, norm)
            torch.testing.assert_close(got, r[f"seg_{mode}_norm{int(norm)}"], rtol=1e-13, atol=1e-13)
    sur, _ = oa.surrogate_loss(r["seg_value_subtraction_norm1"], r["lp_new"], r["lp_old"])
    assert torch.equal(sur, r["surrogate"])
    for clip in (0.0, 0.2):
        got = oa.value_loss(r["values_new"], r["ret_gae1"], r["values"][:, :-1], clip)
        assert torch.equal(got, r[f"value_loss_clip{clip}"])


@pytest.mark.parametrize("name", ["box", "table_tennis"])
def test_ref_policy(golden, name):
    r = golden("ref_policy.pt")[name]
    cfg = MP_CONFIGS[name]
    Dp = cfg["num_dof"] * (cfg["num_basis"] + 1)
    layers, st = [], r["mean_net_state"]
    keys = sorted({k.rsplit(".", 1)[0] for k in st})
    for i, k in enumerate(keys):
        lin = torch.nn.Linear(st[k + ".weight"].shape[1], st[k + ".weight"].shape[0], dtype=torch.float64)
        lin.load_state_dict({"weight": st[k + ".weight"], "bias": st[k + ".bias"]})
        layers += [lin] + ([torch.nn.LeakyReLU()] if i < len(keys) - 1 else [])
    pol = opol.TemporalCorrelatedPolicy(Dp, mp=dict(type="prodmp", args=dict(cfg, dtype=torch.float64)),
                                        mean_net=torch.nn.Sequential(*layers), cov_vector=r["cov_vector"],
                                        contextual=False, min_std=1e-4)
    with torch.no_grad():
        mean, L = pol.policy(r["obs"])
    torch.testing.assert_close(mean, r["mean"], rtol=1e-12, atol=1e-14)
    assert torch.equal(L, r["L"])
    inp = r["inputs"]
    assert torch.equal(pol.entropy([mean, L]), r["entropy"])
    assert torch.equal(pol.covariance(L), r["covariance"])
    assert torch.equal(pol.log_determinant(L), r["log_determinant"])
    assert torch.equal(pol.precision(L), r["precision"])
    assert torch.equal(pol.maha(inp["mean"], inp["mean_old"], inp["L_old"]), r["maha"])
    assert torch.equal(opol.BlackBoxPolicy.log_prob(pol, inp["mean_old"], inp["mean"], inp["L"]), r["bb_log_prob"])
    args = (r["times"], r["init_time"], inp["init_pos"], inp["init_vel"])
    assert torch.equal(pol.sample(False, inp["mean"], inp["L"], *args, use_mean=True), r["traj_mean"])
    lp = pol.log_prob(r["smp_traj"], inp["mean"], inp["L"], *args, pred_pairs=r["pred_pairs"])
    assert torch.equal(lp, r["log_prob"])
    # initial covariance vector quirk (abstract_policy.py:113-116): diag = softplus(v0) + min_std = 0.99 + 1e-4
    fresh = opol.BlackBoxPolicy(Dp, contextual=False, min_std=1e-4)
    L0 = fresh.vector_to_cholesky(fresh.cov_vector)
    assert abs(L0[0, 0].item() - 0.9901) < 1e-12
    assert abs(fresh.entropy([torch.zeros(Dp, dtype=torch.float64), L0]).item()
               - (0.5 * Dp * (1 + math.log(2 * math.pi)) + Dp * math.log(0.9901))) < 1e-9
