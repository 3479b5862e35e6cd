"""CPU oracle for the GPdoemd design-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpdoemd_b200/`` may import this package; it is used
by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the *checker* and the *CPU baseline*, never as the product path.

The reference (cog-imperial/GPdoemd) is pure Python on top of GPy.  GPy is an un-vendored,
unpinned third-party dependency (``requirements.txt:6``, ``setup.py:25,31``; era 2018-19, i.e. GPy
1.9.x + paramz >= 0.9.0) that is not installed in this image and cannot be fetched (no network).
The arithmetic GPy performs for the reference's call sites is therefore *restated* here from GPy's
published algorithms (``oracle/gpy_kernels.py``, ``oracle/gp.py``), while the reference's own
NumPy-only modules (``design_criteria.py``, ``marginal/*``, ``utils.py``, ``transform.py``) are
restated in ``oracle/criteria.py``, ``oracle/marginal.py``, ``oracle/routing.py`` and checked
against the reference modules themselves wherever ``/root/reference`` is importable
(``tests/test_oracle_vs_reference.py``, ``tests/golden/make_golden.py``).

Parity status:
* criteria, Taylor marginals, routing, transforms: PINNED against the reference's own code run in
  the build container (golden vectors under ``tests/golden/``) and the reference's golden J vectors.
* GP predictive mean / variance / p-derivatives: **parity unpinned against real GPy** (GPy absent
  here and on the GPU box).  Pinned instead by internal consistency: central finite differences of
  the oracle's own mean/variance, extended-precision (80-bit) recomputation, sigma^2 via L vs via
  K^-1, and interpolation at the training inputs (``tests/test_oracle_gp.py``); and by an INDEPENDENT
  third-party implementation of the same mathematics: scikit-learn's anisotropic RBF / Matern kernels and
  exact-GP ``predict`` reproduce the oracle's kernel matrices, means and variances
  (``tests/test_oracle_vs_sklearn.py``).  GPy-specific conventions remain unverified.
* discrimination criteria on the observed set (``oracle/discrimination.py``): PINNED against the
  reference's own module (``tests/golden/obs_golden.npz``).
"""
