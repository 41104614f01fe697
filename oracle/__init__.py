"""CPU oracle for the VCGPCM ELBO hot path.  **TEST INFRASTRUCTURE ONLY.**

This package restates, on numpy / torch-CPU float64, the algorithm of the reference
(wesselb/cgpcm, ``src/core``) for the one path this repository accelerates: the interdomain
expected-kernel ("Psi") statistics, the saturated VCGPCM evidence lower bound and its gradient.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker or the timed CPU baseline.  The product
(``cgpcm_b200``) never imports it and has no CPU fallback.

Pinning status
------------------------------
* ``oracle.expq`` is pinned by the reference's only known-answer tests
  (``src/core/exponentiated_quadratic_test.py:24-42``), carried in ``tests/test_oracle_golden.py``.
* ``oracle.bvn`` restates the published Genz (2004) BVND algorithm that the un-vendored, un-pinned
  external dependency ``wesselb/bvn-cdf`` wraps (``src/core/tf_util.py:9-13``, called at
  ``src/core/exponentiated_quadratic.py:552``; ``doc/paper.tex:341``).  It is pinned element-wise
  against scipy's bivariate normal CDF and mpmath quadrature in ``tests/test_oracle_bvn.py``.
* The reference holds **no** golden value for any Psi matrix, the ELBO or a gradient, and it cannot
  run here (Python 2 + TensorFlow 1.x + bvn-cdf): for those quantities **parity is unpinned** against
  the reference itself; the oracle is instead tied to the reference's algorithm by evaluating the
  reference's own integrands (``src/core/cgpcm.py:111-121``) through ``oracle.expq`` and to the exact
  mathematics by quadrature / finite differences (``tests/test_oracle_*.py``).
"""
