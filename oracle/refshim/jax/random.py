"""jax.random is only used for jittered clouds (noise_key): its stream cannot be reproduced without JAX."""


def _no(*a, **k):
    raise NotImplementedError("jax.random is not reproduced by the stand-in; use noise_key=None")


split = uniform = permutation = PRNGKey = _no
