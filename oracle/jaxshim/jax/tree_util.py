"""jax.tree_util stand-in: pytree registration is a no-op without tracing."""


def register_pytree_node_class(cls):
    return cls
