"""Generic device path for inputs OUTSIDE the specialised engines' domain: complex dtypes, physical dimension != 2,
MPO bond > 2, bonds beyond the compiled capacities, arbitrary-site marginals.

The hand-written kernels of this repo (``csrc/``) cover the decode hot path: real FP64, physical dimension 2, MPO
bond <= 2 (``mps_core.cuh``, ``mono_core.cuh``).  The reference's API accepts more than that (its own tests contract
random complex MPOs of bond 3..9, ``tests/contractor/test_contractor.py:152-208``; the random-circuit example is
complex128 at chi = 1024, ``examples/random_circuit/mps-rand-circ.py:57-79``).  For those inputs the same algorithm
-- the reference's zip-up, contraction by contraction (``mdopt/contractor/contractor.py:248-378``), two-site centre
moves (``mdopt/mps/canonical.py:345-420``), truncation rule ``k = min(chi_max, #(s > cut))``
(``mdopt/utils/utils.py:119``) and marginal (``canonical.py:906-991``) -- runs on the GPU through LIBRARY kernels:
``torch.einsum`` / ``tensordot`` (cuBLAS) and ``torch.linalg.svd`` (cuSOLVER).  It is not a CPU fallback (it raises
without a CUDA device) and it is not the product's hot path; DESIGN.md lists it as the library tier.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


_DEVICE = "cuda"   # the library tier has no CPU variant: _torch() raises without a CUDA device


def _torch():
    from mdopt_b200 import _lib

    return _lib.require_gpu()[1]


def needs_generic(tensors: Sequence[np.ndarray], mpo: Optional[Sequence[np.ndarray]] = None, max_bond: int = 256) -> bool:
    """True when the specialised engine cannot take this input (see module docstring)."""
    for t in tensors:
        t = np.asarray(t)
        if np.iscomplexobj(t) or t.ndim != 3 or t.shape[1] != 2 or max(t.shape[0], t.shape[2]) > max_bond:
            return True
    for w in mpo or ():
        w = np.asarray(w)
        if np.iscomplexobj(w) or w.ndim != 4 or w.shape[2] != 2 or w.shape[3] != 2 or max(w.shape[0], w.shape[1]) > 2:
            return True
    return False


def _dev(a, dtype):
    torch = _torch()
    return torch.as_tensor(np.ascontiguousarray(a), device=_DEVICE).to(dtype)


def _dtype_for(arrays):
    torch = _torch()
    return torch.complex128 if any(np.iscomplexobj(np.asarray(a)) for a in arrays) else torch.float64


def svd_device(mat, cut: float = 1e-12, chi_max: int = int(1e4), renormalise: bool = False):
    """Truncated SVD of a device matrix with the reference's rule (utils.py:119-129); cuSOLVER does the SVD."""
    torch = _torch()
    if not bool(torch.isfinite(mat.abs()).all()):
        raise RuntimeError("All SVD methods failed. Last error: array must not contain infs or NaNs")
    u, s, vh = torch.linalg.svd(mat, full_matrices=False)
    k = min(int(chi_max), int((s > cut).sum().item()))
    err = float((s[k:] ** 2).sum().item())
    u, s, vh = u[:, :k], s[:k], vh[:k, :]
    if renormalise and k > 0:
        nrm = float(torch.linalg.norm(s).item())
        if nrm > 0:
            s = s / nrm
    return u, s, vh, err


def qr_pivoted_device(a):
    """Householder QR with column pivoting on the device (what ``scipy.linalg.qr(pivoting=True, mode="economic")``
    gives the reference at utils.py:188-190: at every step the remaining column of largest norm is eliminated next).
    Returns (Q (m, k), R (k, n) in pivoted column order, pivots) with k = min(m, n)."""
    torch = _torch()
    m, n = a.shape
    k = min(m, n)
    r = a.clone()
    q = torch.eye(m, dtype=a.dtype, device=a.device)
    piv = torch.arange(n, device=a.device)
    for j in range(k):
        norms = (r[j:, j:].abs() ** 2).sum(dim=0)
        p = j + int(torch.argmax(norms).item())
        if p != j:
            r[:, [j, p]] = r[:, [p, j]]
            piv[[j, p]] = piv[[p, j]]
        x = r[j:, j]
        normx = torch.linalg.norm(x)
        if float(normx.item()) == 0.0:
            break
        alpha = x[0]
        phase = alpha / alpha.abs() if float(alpha.abs().item()) > 0 else torch.ones((), dtype=a.dtype, device=a.device)
        v = x.clone()
        v[0] = v[0] + phase * normx          # beta = -phase * |x| (LAPACK's sign choice: no cancellation)
        v = v / torch.linalg.norm(v)
        r[j:, j:] = r[j:, j:] - 2.0 * torch.outer(v, v.conj() @ r[j:, j:])
        q[:, j:] = q[:, j:] - 2.0 * torch.outer(q[:, j:] @ v, v.conj())
    return q[:, :k], torch.triu(r[:k, :]), piv


def qr_device(mat, cut: float = 1e-12, chi_max: int = int(1e4), renormalise: bool = False,
              return_truncation_error: bool = False):
    """``mdopt.utils.utils.qr`` (utils.py:141-212): pivoted QR truncated to ``min(chi_max, #(|diag R| > cut))`` rows,
    pivoting undone on R, optional renormalisation by the norm of the kept diagonal, truncation error ||A - QR||."""
    torch = _torch()
    q, r, piv = qr_pivoted_device(mat)
    diag = torch.diagonal(r).abs()
    rank = min(int(chi_max), int((diag > cut).sum().item()))
    inv = torch.argsort(piv)
    r = r[:rank, :][:, inv]
    q = q[:, :rank]
    if renormalise and rank > 0:
        nrm = float(torch.linalg.norm(torch.diagonal(r[:rank, :rank])).item())
        if nrm > 0:
            r = r / nrm
    err = float(torch.linalg.norm(mat - q @ r).item()) if return_truncation_error else None
    return q, r, err


def split_device(theta, chi_max=int(1e4), cut=1e-12, renormalise=False):
    i, j, k, l = theta.shape
    u, s, vh, err = svd_device(theta.reshape(i * j, k * l), cut=cut, chi_max=chi_max, renormalise=renormalise)
    return u.reshape(i, j, s.numel()), s, vh.reshape(s.numel(), k, l), err


def _reverse(ts):
    return [t.permute(2, 1, 0).contiguous() for t in reversed(ts)]


def move_centre_device(ts: List, centre: int, final_pos: int, chi_max: int):
    """canonical.py:345-420: one two-site split per bond with chi = the MPS's own chi_max and cut = 1e-12; moving
    left runs on the reversed chain (:390-393, 413-415)."""
    n = len(ts)
    if centre == final_pos:
        return ts, centre
    rev = centre > final_pos
    if rev:
        ts, begin, final = _reverse(ts), n - 1 - centre, n - 1 - final_pos
    else:
        ts, begin, final = list(ts), centre, final_pos
    torch = _torch()
    for i in range(begin, final):
        theta = torch.tensordot(ts[i], ts[i + 1], dims=([2], [0]))
        u, s, v, _ = split_device(theta, chi_max=chi_max, cut=1e-12, renormalise=False)
        ts[i] = u
        ts[i + 1] = v * s.to(v.dtype)[:, None, None]
    if rev:
        ts = _reverse(ts)
    return ts, final_pos


def mps_mpo_contract_generic(tensors: Sequence[np.ndarray], centre: int, mpo: Sequence[np.ndarray], start_site: int,
                             renormalise: bool, chi_max: int, cut: float, mps_chi_max: int = int(1e4)
                             ) -> Tuple[List[np.ndarray], int]:
    """The reference's zip-up (contractor.py:248-378) on the device; returns host tensors and the new centre."""
    torch = _torch()
    dt = _dtype_for(list(tensors) + list(mpo))
    ts = [_dev(t, dt) for t in tensors]
    ws = [_dev(w, dt) for w in mpo]
    ts, _ = move_centre_device(ts, centre, start_site, mps_chi_max)            # :253-254
    c = start_site
    theta = torch.einsum("ijk,klm,nojp,oqlr->iprqm", ts[c], ts[c + 1], ws[0], ws[1])   # :276-291
    theta = theta.reshape(theta.shape[0], theta.shape[1], theta.shape[2], -1)
    for i in range(len(ws) - 2):                                               # :294-342
        u, s, b, _ = split_device(theta, chi_max=chi_max, cut=cut, renormalise=renormalise)
        ts[c] = u
        c += 1
        carried = (b * s.to(b.dtype)[:, None, None]).reshape(s.numel(), ws[i + 1].shape[3], ws[i + 1].shape[1],
                                                           ts[c + 1].shape[0])
        theta = torch.einsum("ijkl,lmn,komp->ijpon", carried, ts[c + 1], ws[i + 2])
        theta = theta.reshape(theta.shape[0], theta.shape[1], theta.shape[2], -1)
    u, s, b, _ = split_device(theta, chi_max=chi_max, cut=cut, renormalise=renormalise)   # :345-360
    ts[c] = u
    ts[c + 1] = b * s.to(b.dtype)[:, None, None]
    if renormalise:   # :363-369 normalises tensors[orth_centre_index] = the left neighbour of the new centre (quirk kept)
        nrm = float(torch.linalg.norm(ts[c]).item())
        if nrm != 0:
            ts[c] = ts[c] / nrm
    if _DEVICE == "cuda":
        torch.cuda.synchronize()                                               # :371-372
    return [t.cpu().numpy() for t in ts], c + 1


def move_orth_centre_generic(tensors, centre, final_pos, chi_max):
    dt = _dtype_for(tensors)
    ts, c = move_centre_device([_dev(t, dt) for t in tensors], centre, final_pos, chi_max)
    return [t.cpu().numpy() for t in ts], c


def marginal_generic(tensors: Sequence[np.ndarray], sites: Sequence[int], renormalise: bool) -> List[np.ndarray]:
    """canonical.py:956-989 for any set of sites: trace with (1, .., 1)/sqrt(d), absorb into the right neighbour (the
    left one for the last site), renormalise the last tensor after each step if asked (unguarded, like the
    reference: a zero norm gives NaN)."""
    torch = _torch()
    dt = _dtype_for(tensors)
    t = [_dev(x, dt) for x in tensors]
    for site in sorted(sites, reverse=True):
        d = t[site].shape[1]
        vec = torch.ones(d, dtype=dt, device=_DEVICE) / np.sqrt(d)
        traced = torch.tensordot(t[site], vec, dims=([1], [0]))
        if site < len(t) - 1:
            merged = torch.tensordot(traced, t[site + 1], dims=([1], [0]))
            del t[site + 1]
            del t[site]
            t.insert(site, merged)
        elif site > 0:
            t[site - 1] = torch.tensordot(t[site - 1], traced, dims=([2], [0]))
            del t[site]
        else:
            t[site] = traced.reshape(traced.shape + (1,))
        if renormalise:
            t[-1] = t[-1] / torch.linalg.norm(t[-1])
    return [x.cpu().numpy() for x in t]


def compress_bond_generic(tensors, centre, bond, chi_max, cut, renormalise, mps_chi_max, strategy="svd",
                          want_error=True):
    """canonical.py:684-835: centre to `bond`, then
    "svd": split the two-site tensor with the user's chi_max / cut, leave U.S on `bond` (:741-760);
    "qr": pivoted-QR split, Q on `bond`, R on `bond + 1` (:762-775);
    "svd_advanced": SVD of each tensor, SVD of the small core, sqrt(s) on both sides (:777-833).
    Returns (host tensors, centre, truncation error)."""
    torch = _torch()
    dt = _dtype_for(tensors)
    ts, _ = move_centre_device([_dev(t, dt) for t in tensors], centre, bond, mps_chi_max)
    left, right = ts[bond], ts[bond + 1]
    if strategy == "svd":
        theta = torch.tensordot(left, right, dims=([2], [0]))
        u, s, v, err = split_device(theta, chi_max=chi_max, cut=cut, renormalise=renormalise)
        ts[bond] = u * s.to(u.dtype)[None, None, :]
        ts[bond + 1] = v
    elif strategy == "qr":
        theta = torch.tensordot(left, right, dims=([2], [0]))
        i, j, k, l = theta.shape
        q, r, err = qr_device(theta.reshape(i * j, k * l), cut=cut, chi_max=chi_max, renormalise=renormalise,
                              return_truncation_error=want_error)
        ts[bond] = q.reshape(i, j, q.shape[1])
        ts[bond + 1] = r.reshape(r.shape[0], k, l)
    elif strategy == "svd_advanced":
        cl, pl, _ = left.shape
        _, pr, cr = right.shape
        u0, sl, u1, _ = svd_device(left.reshape(cl * pl, -1), cut=cut, chi_max=chi_max, renormalise=renormalise)
        v0, sr, v1, _ = svd_device(right.reshape(-1, pr * cr), cut=cut, chi_max=chi_max, renormalise=renormalise)
        core = (u1 * sl.to(u1.dtype)[:, None]) @ (v0 * sr.to(v0.dtype)[None, :])
        un, sn, vn, err = svd_device(core, cut=cut, chi_max=chi_max, renormalise=renormalise)
        root = torch.sqrt(sn).to(un.dtype)
        ts[bond] = (u0 @ (un * root[None, :])).reshape(cl, pl, sn.numel())
        ts[bond + 1] = ((root[:, None] * vn) @ v1).reshape(sn.numel(), pr, cr)
    else:
        raise ValueError(f"Unsupported compression strategy: {strategy}")
    return [t.cpu().numpy() for t in ts], bond, err
