"""CPU oracle for the DaXBench simulator step (TEST INFRASTRUCTURE ONLY).

This package is a torch-CPU restatement of the reference's JAX arithmetic for
the hot path named in BASELINE.json (MLS-MPM substep/step, rigid primitives,
safe SVD, mass-spring cloth step) with torch autograd standing in for
``jax.grad``.  Each function cites the reference file:line it follows
(paths relative to /root/reference/DaXBench/daxbench/).

PARITY STATUS: **parity unpinned by reference tests** -- the reference ships no
tests for this path and JAX cannot be installed in this image, so the
reference cannot be executed directly.  What pins this oracle instead:
  * tests/golden/* fixtures produced by running the UNMODIFIED reference
    sources under ``oracle/jaxshim`` (a minimal numpy/torch stand-in for the
    handful of jax APIs the path uses) -- see oracle/gen_golden.py;
  * the reference's own artefacts (goal.npy lattices, expert-demo primitive
    kinematics) -- see tests/test_oracle_fixtures.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
legs may import this package.  The product (unidom_b200/) never does.
"""
