"""Loader for ``libmdopt_b200.so`` (the sm_100a kernels behind the C ABI in ``include/mdopt_b200.h``).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every compute entry point
raises.  Device memory, streams and multi-GPU plumbing come from PyTorch; the kernels are ours.
"""
from __future__ import annotations

import ctypes as C
import os

from mdopt_b200 import _abi

_LIB_NAME = "libmdopt_b200.so"
_lib = None


class MdoptCudaError(RuntimeError):
    pass


def lib_path() -> str:
    """In-tree library; MDOPT_B200_LIB points at another build of the same ABI (kernel experiments)."""
    override = os.environ.get("MDOPT_B200_LIB")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def _try_build() -> None:
    """Build the library in-tree when it is missing and nvcc is available (still no CPU fallback)."""
    import importlib.util
    import shutil

    if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
        return
    entry = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "__graft_entry__.py")
    if not os.path.exists(entry):
        return
    spec = importlib.util.spec_from_file_location("_mdopt_graft_entry", entry)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()


def load() -> C.CDLL:
    """Load and bind the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            _try_build()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  mdopt_b200 has no CPU fallback."
            )
        _lib = _abi.bind(C.CDLL(path))
    return _lib


def require_gpu():
    """Return (lib, torch) or raise: the product path never runs without the CUDA kernels."""
    import torch

    lib = load()
    if not torch.cuda.is_available():
        raise MdoptCudaError("mdopt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return lib, torch


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    names = {_abi.E_BADARG: "bad argument", _abi.E_CAPACITY: "exceeds compiled capacity", _abi.E_NOGPU: "no GPU"}
    if rc < 0:
        raise MdoptCudaError(f"{what}: {names.get(rc, rc)}")
    raise MdoptCudaError(f"{what}: CUDA error {rc}")
