"""``svd`` / ``split_two_site_tensor`` with the reference's truncation semantics, on the GPU.

Drop-in for ``mdopt/utils/utils.py:29-138`` and ``:290-383`` (svd strategy).  One CTA per matrix runs a
one-sided Jacobi SVD in shared memory (``mdopt_svd_batch_device``).  Results come back as host NumPy
arrays, like the reference's (utils.py:114-116).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from mdopt_b200 import _abi, _lib


def svd_batch(mats: np.ndarray, cut: float = 1e-12, chi_max: int = int(1e4), renormalise: bool = False):
    """Batched truncated SVD of ``mats[b]`` (all the same shape).  Returns (u, s, vh, kept, trunc_err, status)."""
    lib, torch = _lib.require_gpu()
    mats = np.ascontiguousarray(mats, dtype=np.float64)
    if mats.ndim != 3:
        raise ValueError("svd_batch expects an array of shape (batch, m, n)")
    if not np.isfinite(mats).all():
        raise RuntimeError("All SVD methods failed. Last error: array must not contain infs or NaNs")
    batch, m, n = mats.shape
    transposed = n > m  # orthogonalise the shorter side: fewer Jacobi players, same singular values
    work = np.ascontiguousarray(np.swapaxes(mats, 1, 2)) if transposed else mats
    wm, wn = work.shape[1], work.shape[2]
    cap = _abi.chi_capacity((max(wm, wn) + 1) // 2)
    if cap == 0:
        raise _lib.MdoptCudaError(f"matrix {m}x{n} exceeds the shared-memory SVD capacity (128x128)")
    k = min(m, n)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_a = torch.from_numpy(work).to(dev)
    d_u = torch.empty((batch, wm, k), dtype=torch.float64, device=dev)
    d_s = torch.empty((batch, k), dtype=torch.float64, device=dev)
    d_vh = torch.empty((batch, k, wn), dtype=torch.float64, device=dev)
    d_kept = torch.empty(batch, dtype=torch.int32, device=dev)
    d_trunc = torch.empty(batch, dtype=torch.float64, device=dev)
    d_status = torch.empty(batch, dtype=torch.int32, device=dev)
    chi = int(min(int(chi_max), k))
    rc = lib.mdopt_svd_batch_device(cap, batch, wm, wn, C.c_void_p(d_a.data_ptr()), chi, float(cut),
                                    int(bool(renormalise)), C.c_void_p(d_u.data_ptr()), C.c_void_p(d_s.data_ptr()),
                                    C.c_void_p(d_vh.data_ptr()), C.c_void_p(d_kept.data_ptr()),
                                    C.c_void_p(d_trunc.data_ptr()), C.c_void_p(d_status.data_ptr()),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mdopt_svd_batch_device")
    u, s, vh = d_u.cpu().numpy(), d_s.cpu().numpy(), d_vh.cpu().numpy()
    if transposed:
        u, vh = np.ascontiguousarray(np.swapaxes(vh, 1, 2)), np.ascontiguousarray(np.swapaxes(u, 1, 2))
    return u, s, vh, d_kept.cpu().numpy(), d_trunc.cpu().numpy(), d_status.cpu().numpy()


def svd(mat: np.ndarray, cut: float = float(1e-12), chi_max: int = int(1e4), renormalise: bool = False,
        return_truncation_error: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Optional[float]]:
    """Truncated SVD keeping ``min(chi_max, #(s > cut))`` values (reference utils.py:29-138)."""
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError(f"A valid matrix must have 2 dimensions while the one given has {mat.ndim}.")
    if np.iscomplexobj(mat) or max(mat.shape) > 128:
        # complex or beyond the shared-memory Jacobi kernel (128 x 128): cuSOLVER through the library tier
        from mdopt_b200 import generic

        torch = _lib.require_gpu()[1]
        dmat = torch.as_tensor(np.ascontiguousarray(mat), device=generic._DEVICE)
        u, s, vh, err = generic.svd_device(dmat, cut=cut, chi_max=chi_max, renormalise=renormalise)
        out = (u.cpu().numpy(), s.cpu().numpy().astype(float), vh.cpu().numpy())
        return out + ((err,) if return_truncation_error else (None,))
    u, s, vh, kept, trunc, _ = svd_batch(mat[None], cut=cut, chi_max=chi_max, renormalise=renormalise)
    k = int(kept[0])
    out = (np.ascontiguousarray(u[0][:, :k]), np.ascontiguousarray(s[0][:k]), np.ascontiguousarray(vh[0][:k, :]))
    return out + ((float(trunc[0]),) if return_truncation_error else (None,))


def qr(mat: np.ndarray, cut: float = float(1e-12), chi_max: int = int(1e4), renormalise: bool = False,
       return_truncation_error: bool = False):
    """Pivoted QR with truncation by |diag R| (reference utils.py:141-212), on the device through the library tier."""
    mat = np.asarray(mat)
    if len(mat.shape) != 2:
        raise ValueError(f"Input matrix must have 2 dimensions, got {len(mat.shape)}.")
    from mdopt_b200 import generic

    torch = _lib.require_gpu()[1]
    dmat = torch.as_tensor(np.ascontiguousarray(mat), device=generic._DEVICE)
    q, r, err = generic.qr_device(dmat, cut=cut, chi_max=int(min(chi_max, 1 << 30)), renormalise=renormalise,
                                  return_truncation_error=return_truncation_error)
    return q.cpu().numpy(), r.cpu().numpy(), err


def split_two_site_tensor(tensor: np.ndarray, chi_max: int = int(1e4), cut: float = float(1e-12),
                          renormalise: bool = False, strategy: str = "svd",
                          return_truncation_error: bool = False) -> Tuple:
    """Split ``(i, j, k, l)`` into ``U(i,j,m), s(m), V(m,k,l)`` (``"svd"``) or ``Q(i,j,m), R(m,k,l)`` (``"qr"``)
    (reference utils.py:290-383)."""
    tensor = np.asarray(tensor)
    if tensor.ndim != 4:
        raise ValueError(f"A valid two-site tensor must have 4 legs while the one given has {tensor.ndim}.")
    if strategy not in ["svd", "qr"]:
        raise ValueError("The strategy must be either `svd` or `qr`.")
    i, j, k, l = tensor.shape
    if strategy == "qr":   # pivoted QR (utils.py:141-212, :370-383) through the library tier on the device
        from mdopt_b200 import generic

        torch = _lib.require_gpu()[1]
        dmat = torch.as_tensor(np.ascontiguousarray(tensor.reshape(i * j, k * l)), device=generic._DEVICE)
        q, r, err = generic.qr_device(dmat, cut=cut, chi_max=int(min(chi_max, 1 << 30)), renormalise=renormalise,
                                      return_truncation_error=return_truncation_error)
        chi_cut = min(int(min(chi_max, 1 << 30)), q.shape[1])
        q_l = q.cpu().numpy().reshape(i, j, chi_cut)
        r_r = r.cpu().numpy().reshape(chi_cut, k, l)
        if return_truncation_error:
            return q_l, r_r, err, None
        return q_l, r_r, None, None
    u, s, vh, err = svd(tensor.reshape(i * j, k * l), cut=cut, chi_max=chi_max, renormalise=renormalise,
                        return_truncation_error=return_truncation_error)
    u = u.reshape(i, j, s.size)
    vh = vh.reshape(s.size, k, l)
    if return_truncation_error:
        return u, s, vh, err
    return u, s, vh
