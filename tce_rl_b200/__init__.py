"""tce_rl_b200 -- B200-native (sm_100a) kernels for TCE's episodic policy-update path.

Layout: ``csrc/`` hand-written CUDA + the C ABI (``include/tce_b200.h``), ``_lib`` ctypes binding,
``ops`` torch custom ops, ``mp`` / ``rl`` host-side mirrors of the reference's ProDMP, policy,
projection and agent interfaces (same names, argument meaning and error behaviour).
"""
__version__ = "0.1.0"
