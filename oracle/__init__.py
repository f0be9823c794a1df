"""CPU oracle for the PoissonGPLVMJump1D EM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported only from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
(``poor_man_gplvm_b200``) never imports this package and fails loudly when its
CUDA library is missing.

PARITY UNPINNED: the reference (``/root/reference``, pure Python on JAX) ships no
golden vectors, no passing tests and cannot be executed in this image (no jax /
optax wheels, no network).  The restatement in ``ref_numpy.py`` is therefore
validated by (i) mathematical invariants, (ii) agreement with the independent
linear-space fp64 derivation in ``linear_ref.py`` and (iii) recovery of planted
structure on synthetic data; see ``tests/test_oracle_*.py``.
"""
