"""Array-backend shim (mirror of ``mdopt/backend/array.py:11-115``), fixed to the single sm_100a path.

The reference switches between NumPy and CuPy through ``MDOPT_BACKEND``; here there is exactly one backend -- our
CUDA kernels, with PyTorch for device memory and streams -- so ``GPU`` is always True and there is nothing to
dispatch on.  The surface is the reference's: ``backend_name / is_cuda_backend`` (:36-43), ``to_device / to_host /
stream / synchronize`` (:46-85), ``einsum / svd / asfortran`` (:91-101) and module-level attribute forwarding
(:108-115), which here resolves ``xp.<name>`` against the device array library (``torch``) the way the reference
resolves it against ``cupy``; ``xp.linalg.svd`` is the batched Jacobi kernel of this repo.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np

GPU = True


def backend_name() -> str:
    """The active backend: always the CUDA path (the reference answers 'cupy' or 'numpy', array.py:36-38)."""
    return "mdopt_b200"


def is_cuda_backend() -> bool:
    return True


def _torch():
    from mdopt_b200 import _lib

    return _lib.require_gpu()[1]


def to_device(x):
    """Move / ensure the array is on the device (array.py:49-51)."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        return x.cuda()
    return torch.as_tensor(np.asarray(x), device="cuda")


def to_host(x):
    """Move / ensure the array is on the host as NumPy (array.py:53-55)."""
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


@contextlib.contextmanager
def stream():
    """A non-blocking CUDA stream context (array.py:57-60); synchronised on exit."""
    torch = _torch()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        yield st
    st.synchronize()


def synchronize():
    _torch().cuda.synchronize()


def asarray(x, dtype=None):
    t = to_device(x)
    return t if dtype is None else t.to(dtype)


def einsum(expr, *args, **kw):
    """``xp.einsum`` on device arrays (array.py:91-92)."""
    torch = _torch()
    return torch.einsum(expr, *[to_device(a) for a in args])


def svd(x, full_matrices=False):
    """A consistent SVD surface (array.py:95-97): reduced SVD by this repo's batched Jacobi kernel, host arrays out
    like the reference's ``svd`` returns after ``_to_numpy`` (utils.py:114-116)."""
    if full_matrices:
        raise NotImplementedError("only the reduced SVD is provided (the reference never asks for the full one)")
    from mdopt_b200.utils.utils import svd as _svd

    u, s, vh, _ = _svd(to_host(x), cut=-1.0, chi_max=int(1e9))
    return u, s, vh


def asfortran(a):
    """Column-major copy (array.py:100-101).  Device tensors carry strides, so this is a transposed-contiguous view."""
    t = to_device(a)
    return t.t().contiguous().t() if t.dim() == 2 else t.contiguous()


linalg = types.SimpleNamespace(svd=svd)


def __getattr__(name):
    """Module-level attribute forwarding (array.py:108-115): ``xp.zeros``, ``xp.tensordot`` ... resolve against the
    device array library."""
    import torch

    if hasattr(torch, name):
        return getattr(torch, name)
    raise AttributeError(
        f"module '{__name__}' has no attribute '{name}'. "
        f"The current backend ('{backend_name()}') also does not have this attribute. "
        "This may be due to a typo or an API change."
    )
