"""jax.tree_util stand-in: pytrees over tuple / list / dict / NamedTuple / None; everything else is a leaf."""


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree):
    leaves = []

    def rec(x):
        if x is None:
            return ("none",)
        if _is_namedtuple(x):
            return ("nt", type(x), [rec(c) for c in x])
        if isinstance(x, tuple):
            return ("tuple", [rec(c) for c in x])
        if isinstance(x, list):
            return ("list", [rec(c) for c in x])
        if isinstance(x, dict):
            keys = sorted(x.keys())
            return ("dict", keys, [rec(x[k]) for k in keys])
        leaves.append(x)
        return ("leaf",)
    treedef = rec(tree)
    return leaves, treedef


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind = d[0]
        if kind == "none":
            return None
        if kind == "leaf":
            return next(it)
        if kind == "nt":
            return d[1](*[rec(c) for c in d[2]])
        if kind == "tuple":
            return tuple(rec(c) for c in d[1])
        if kind == "list":
            return [rec(c) for c in d[1]]
        if kind == "dict":
            return {k: rec(c) for k, c in zip(d[1], d[2])}
        raise ValueError(kind)
    return rec(treedef)


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_map(f, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def tree_reduce(f, tree, initializer=None):
    import functools
    leaves = tree_leaves(tree)
    if initializer is None:
        return functools.reduce(f, leaves)
    return functools.reduce(f, leaves, initializer)


tree_multimap = tree_map
