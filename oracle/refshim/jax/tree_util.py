"""jax.tree_util stand-in: Partial and tree_map over dict / list / tuple containers."""
import functools


class Partial(functools.partial):
    pass


def tree_map(f, tree):
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    if isinstance(tree, list):
        return [tree_map(f, v) for v in tree]
    if isinstance(tree, tuple):
        return tuple(tree_map(f, v) for v in tree)
    return f(tree)
