"""optax stand-in: only global_norm (optax/_src/linear_algebra.py: sqrt(sum(sum(x**2))) over leaves)."""
import jax.numpy as jnp
from jax import tree_util


def global_norm(updates):
    leaves = tree_util.tree_leaves(updates)
    return jnp.sqrt(sum(jnp.sum(x.astype(jnp.float32) ** 2 if hasattr(x, "astype") else jnp.array(x) ** 2) for x in leaves))
