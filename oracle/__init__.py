"""CPU oracle for the PoissonGPLVMJump1D EM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported only from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
(``poor_man_gplvm_b200``) never imports this package and fails loudly when its
CUDA library is missing.

How the oracle is pinned.  The reference (``/root/reference``, pure Python on
JAX) ships no golden vectors and no passing tests, and JAX itself cannot be
installed in this image (no wheel, no network).  The pin is therefore the
reference's *own source files*, executed unmodified in the build container with
``oracle/jaxshim`` standing in for the ``jax`` / ``optax`` import names on torch
CPU tensors (``oracle/ref_loader.py``, ``tests/golden/make_golden.py``): their
outputs are committed as ``tests/golden/*.npz`` and

  * ``ref_numpy.py`` (the NumPy restatement, reference operation order) agrees
    with them to 1e-10 in fp64 and to fp32 rounding in fp32
    (``tests/test_oracle_golden.py``);
  * the CUDA path is compared with the same fixtures directly
    (``tests/test_gpu_golden.py``).

What this pins: everything written in the reference's source (control flow,
index conventions, the ``prior[t+1]``/``post[t]`` pairing, masks, chunk loop,
Adam loop and stopping rule).  What it cannot pin: XLA's floating-point
evaluation order, the LAPACK/cuSOLVER SVD sign convention behind
``generate_basis`` (fixtures carry the basis) and JAX's PRNG bit streams
(fixtures carry every random input).  ``linear_ref.py`` is an independent
linear-space derivation used as a second check (``tests/test_oracle.py``).
"""
