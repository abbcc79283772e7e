"""Stand-in for opt-einsum 3.4.0 (absent here, no network). TEST INFRASTRUCTURE ONLY.

The reference always calls ``contract(expr, *ops, optimize=[pairwise path], backend=...)`` with an
explicit pairwise path, so dispatching to ``numpy.einsum`` with that same path performs the same
pairwise contractions in the same order.
"""
import numpy as np


def contract(expr, *operands, optimize=None, backend="numpy", **_kw):
    if optimize is None or optimize is True or optimize is False:
        return np.einsum(expr, *operands)
    return np.einsum(expr, *operands, optimize=["einsum_path", *[tuple(p) for p in optimize]])
