"""CPU oracle for the TCE episodic policy-update path.  TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU / numpy restatement of the algorithm that the
reference (BruceGeLi/TCE_RL) runs through ``mprl.rl`` + ``mp_pytorch`` +
``trust_region_projections``.  It exists so that the CUDA path in
``tce_rl_b200`` can be checked; it is **never** imported by the product
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it.

Parity status
-------------
* PINNED (checked against the reference's own known-answer tests and against
  the reference's real code imported in the build container, fixtures under
  ``tests/golden``): ``build_lower_matrix`` / ``reverse_build_matrix`` layout,
  ``add_expand_dim``, ``tensor_linspace`` / ``get_times``,
  ``indexing_interpolate``, softplus-space transforms, ``select_pred_pairs``
  (bit exact), GAE, segment advantages, surrogate / value losses, the
  ``BlackBoxPolicy`` Gaussian helpers and the tensor plumbing of
  ``TemporalCorrelatedPolicy.log_prob`` / ``sample``.
* PARITY UNPINNED: ``oracle.prodmp`` (mp_pytorch 0.1.4, conda-forge) and
  ``oracle.projection`` (BruceGeLi/trust-region-layers@TCE_ICLR24 + the C++
  ``cpp_projection`` package) restate third-party dependencies whose sources
  are absent from ``/root/reference``; the reference holds no test or golden
  vector for them.  They follow the published algorithms (ProDMP paper, Otto et
  al. ICLR'21) and are validated through mathematical invariants
  (``tests/test_oracle_invariants.py``).
"""
