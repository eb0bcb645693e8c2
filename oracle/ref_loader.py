"""Import the REAL reference package (``/root/reference/mprl``) in the build container.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container
(never on the GPU box), so this module is used solely by ``oracle/gen_golden.py``
and by CPU tests that skip when the path is absent.  Third-party modules the
reference imports but that are not installed (cw2, natsort, matplotlib,
git_repos_tracker, fancy_gym, gymnasium, stable_baselines3, wandb) are replaced
by ``MagicMock``; ``mp_pytorch`` and ``trust_region_projections`` are replaced by
thin shims over the oracle restatements so that the reference's own
``get_mp`` / ``projection_factory`` / policy / agent code runs unmodified.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("TCE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mprl"))


def _shim_mp_pytorch():
    from . import prodmp as op

    class ExpDecayPhaseGenerator:
        def __init__(self, tau, delay=0.0, alpha_phase=3.0, **kw):
            self.tau, self.delay, self.alpha_phase = tau, delay, alpha_phase

    class ProDMPBasisGenerator:
        def __init__(self, phase_generator, num_basis, basis_bandwidth_factor, num_basis_outside=0,
                     dt=0.01, alpha=25, pre_compute_length_factor=5, dtype=None, device=None):
            self.pg, self.dtype = phase_generator, dtype
            self.kw = dict(num_basis=num_basis, basis_bandwidth_factor=basis_bandwidth_factor,
                           num_basis_outside=num_basis_outside, dt=dt, alpha=alpha)
            assert pre_compute_length_factor == 5

    class ProDMP(op.ProDMP):
        def __init__(self, basis_gn, num_dof, auto_scale_basis=False, weights_scale=1, goal_scale=1,
                     disable_weights=False, disable_goal=False, relative_goal=False, dtype=None,
                     device=None, **kw):
            assert not disable_weights and not disable_goal
            pg = basis_gn.pg
            super().__init__(num_dof=num_dof, tau=pg.tau, delay=pg.delay, alpha_phase=pg.alpha_phase,
                             auto_scale_basis=auto_scale_basis, weights_scale=weights_scale,
                             goal_scale=goal_scale, relative_goal=relative_goal, dtype=dtype,
                             **basis_gn.kw)

    root = types.ModuleType("mp_pytorch")
    for sub, objs in (("basis_gn", {"ProDMPBasisGenerator": ProDMPBasisGenerator}),
                      ("mp", {"ProDMP": ProDMP}),
                      ("phase_gn", {"ExpDecayPhaseGenerator": ExpDecayPhaseGenerator})):
        m = types.ModuleType(f"mp_pytorch.{sub}")
        m.__dict__.update(objs)
        setattr(root, sub, m)
        sys.modules[f"mp_pytorch.{sub}"] = m
    sys.modules["mp_pytorch"] = root


def _shim_trust_region():
    from . import projection as oj
    root = types.ModuleType("trust_region_projections")
    utils = types.ModuleType("trust_region_projections.utils")
    pu = types.ModuleType("trust_region_projections.utils.projection_utils")
    pu.gaussian_kl_details = oj.gaussian_kl_details
    pu.gaussian_kl = oj.gaussian_kl
    projs = types.ModuleType("trust_region_projections.projections")
    sys.modules.update({"trust_region_projections": root, "trust_region_projections.utils": utils,
                        "trust_region_projections.utils.projection_utils": pu,
                        "trust_region_projections.projections": projs})
    table = {"base_projection_layer": ("BaseProjectionLayer", oj.BaseProjectionLayer),
             "frob_projection_layer": ("FrobeniusProjectionLayer", oj.FrobeniusProjectionLayer),
             "kl_projection_layer": ("KLProjectionLayer", oj.KLProjectionLayer),
             "papi_projection": ("PAPIProjection", MagicMock()),
             "w2_projection_layer": ("WassersteinProjectionLayer", oj.WassersteinProjectionLayer),
             "w2_projection_layer_non_com": ("WassersteinProjectionLayerNonCommuting", MagicMock())}
    for mod, (name, obj) in table.items():
        m = types.ModuleType(f"trust_region_projections.projections.{mod}")
        setattr(m, name, obj)
        sys.modules[m.__name__] = m


_MOCKED = ["cw2", "cw2.cw_data", "cw2.cw_data.cw_wandb_logger", "cw2.cw_data.cw_logging", "cw2.experiment",
           "cw2.cluster_work", "cw2.cw_error", "natsort", "matplotlib", "matplotlib.pyplot",
           "matplotlib.animation", "git_repos_tracker", "git_repos_tracker.tracker", "fancy_gym",
           "gymnasium", "stable_baselines3", "stable_baselines3.common",
           "stable_baselines3.common.vec_env", "wandb"]


def load():
    """Return the imported reference ``mprl`` package (raises if unavailable)."""
    if not available():
        raise ImportError(f"reference not present at {REF_ROOT}")
    if "mprl" in sys.modules:
        return sys.modules["mprl"]
    for name in _MOCKED:
        sys.modules.setdefault(name, MagicMock())
    _shim_mp_pytorch()
    _shim_trust_region()
    sys.path.insert(0, REF_ROOT)
    try:
        import mprl  # noqa: F401
        import mprl.util  # noqa: F401
        import mprl.rl.policy  # noqa: F401
        import mprl.rl.agent  # noqa: F401
        import mprl.rl.projection  # noqa: F401
    finally:
        sys.path.remove(REF_ROOT)
    return sys.modules["mprl"]
