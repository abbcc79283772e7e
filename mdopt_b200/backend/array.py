"""Array-backend shim (mirror of ``mdopt/backend/array.py:11-115``), fixed to the single sm_100a path.

The reference switches between NumPy and CuPy through ``MDOPT_BACKEND``; here there is exactly one
backend -- our CUDA kernels, with PyTorch for device memory and streams -- so ``GPU`` is always True and
there is nothing to dispatch on.
"""
from __future__ import annotations

import contextlib

import numpy as np

GPU = True


def _torch():
    from mdopt_b200 import _lib

    return _lib.require_gpu()[1]


def to_device(x):
    torch = _torch()
    return torch.as_tensor(np.asarray(x), device="cuda")


def to_host(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


@contextlib.contextmanager
def stream():
    torch = _torch()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        yield st
    st.synchronize()


def synchronize():
    _torch().cuda.synchronize()


def svd(a, full_matrices=False):
    from mdopt_b200.utils.utils import svd as _svd

    u, s, vh, _ = _svd(to_host(a), cut=-1.0, chi_max=int(1e9))
    return u, s, vh
