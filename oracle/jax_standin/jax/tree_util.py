"""jax.tree_util stand-in: registration decorators are no-ops; tree_map over tuples / lists / dicts / registered classes."""
_REGISTRY = {}


def register_pytree_node_class(cls):
    _REGISTRY[cls] = (lambda x: x.tree_flatten(), lambda aux, ch: cls.tree_unflatten(aux, ch))
    return cls


def register_pytree_node(cls, flatten, unflatten):
    _REGISTRY[cls] = (flatten, unflatten)


def tree_flatten(tree):
    leaves = []

    def rec(x):
        if type(x) in _REGISTRY:
            children, aux = _REGISTRY[type(x)][0](x)
            return ("node", type(x), aux, [rec(c) for c in children])
        if isinstance(x, (tuple, list)):
            return ("seq", type(x), None, [rec(c) for c in x])
        if isinstance(x, dict):
            return ("dict", None, list(x.keys()), [rec(x[k]) for k in x])
        if x is None:
            return ("none", None, None, [])
        leaves.append(x)
        return ("leaf", None, None, [])
    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind, typ, aux, ch = d
        if kind == "leaf":
            return next(it)
        if kind == "none":
            return None
        if kind == "seq":
            return typ(rec(c) for c in ch)
        if kind == "dict":
            return {k: rec(c) for k, c in zip(aux, ch)}
        return _REGISTRY[typ][1](aux, [rec(c) for c in ch])
    return rec(treedef)


def tree_map(fn, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])


tree_leaves = lambda tree: tree_flatten(tree)[0]


def tree_reduce(fn, tree, initializer=None):
    import functools
    leaves = tree_leaves(tree)
    return functools.reduce(fn, leaves) if initializer is None else functools.reduce(fn, leaves, initializer)
