"""jax._src.lax.control_flow stand-in (the reference imports fori_loop from here: mpm_simulator.py:10)."""
from ...lax import fori_loop, scan, cond  # noqa: F401
