"""Canonical-tie-break SVD hook.  TEST INFRASTRUCTURE ONLY (see oracle/mps_oracle.py header).

Why it exists.  With ``chi_max`` active, many truncating SVDs of a decode cut *inside an exactly degenerate
multiplet* (SURVEY.md section 7): which vectors of the multiplet survive is decided by LAPACK rounding, and the
reference's own ``gesdd`` and ``gesvd`` drivers disagree by up to 8.5e-2 in the final marginals.  "Equal to the
reference within 1e-10" is therefore only a well-posed statement once the tie-break is fixed.  This module states
the rule the CUDA engine uses, as a drop-in for ``numpy.linalg.svd(mat, full_matrices=False)`` -- the call the
reference makes at ``mdopt/utils/utils.py:80`` through its backend shim (``xp.linalg.svd``).  Installing it there
(``with installed(): ...``) leaves every line of the reference unmodified and makes truncating decodes
reproducible to rounding.

The rule.  In a decode with product-state inputs and 0/1 constraint tensors every two-site matrix ``theta`` the
reference hands to ``svd`` has the form ``theta = Q diag(sigma) T`` with ``T`` row-orthonormal and ``Q`` a
"row-singleton" isometry: the rows of ``theta`` fall into groups of mutually parallel rows, and rows of different
groups are orthogonal (each group is one trellis state of the parity constraints).  Then the
singular triplets can be read off without any rotation:

  * ``sigma_g = || theta[rows of g, :] ||_F``, ``U[r, g] = +-||theta[r, :]|| / sigma_g``, ``Vh = diag(1/sigma) U^T theta``;
  * groups with ``sigma_g <= 4 eps * max sigma`` are numerically zero (s = 0, never kept);
  * ORDER: sigma descending, where values that agree to ``TIE_REL`` relative (chained over the sorted list) form
    one cluster, and inside a cluster the group whose first row comes first goes first.

The order is robust against rounding (genuinely distinct values differ by far more than ``TIE_REL``, rounding noise
is far smaller), so an implementation with a different summation order reproduces it; the first-row key is
intrinsic to ``theta`` (no access to the column labels of the carried matrix is needed).  Matrices without the
structure fall through to LAPACK unchanged.
"""
from __future__ import annotations

import contextlib

import numpy as np

EPS = 2.220446049250313e-16
TIE_REL = 1e-10          # relative tolerance (on sigma) below which two singular values count as tied
PARALLEL_TOL = 1e-9      # |cos| of two rows must be within this of 0 or 1
STATS = {"structured": 0, "lapack": 0}

_LAPACK_SVD = np.linalg.svd


def canonical_order(sig2, first_row):
    """Order of the groups: sigma^2 descending in chained clusters of relative width 2*TIE_REL, first row ascending
    inside a cluster.  Shared with the tests of the CUDA engine's ordering."""
    sig2 = np.asarray(sig2, dtype=float)
    first_row = np.asarray(first_row)
    idx = np.lexsort((first_row, -sig2))           # sigma^2 descending, then first row
    out = []
    start = 0
    n = len(idx)
    for i in range(1, n + 1):
        if i == n or sig2[idx[i]] < sig2[idx[i - 1]] * (1.0 - 2.0 * TIE_REL):
            cluster = idx[start:i]
            out.extend(sorted(cluster, key=lambda g: first_row[g]))
            start = i
    return np.asarray(out, dtype=int)


def row_groups(mat):
    """Groups of mutually parallel rows, or None when the matrix does not have the row-singleton structure."""
    rho2 = np.einsum("ij,ij->i", mat, mat)
    live = np.nonzero(rho2 > 0.0)[0]
    if live.size == 0:
        return [], rho2, None
    sub = mat[live]
    gram = sub @ sub.T
    rho = np.sqrt(rho2[live])
    cosine = gram / np.outer(rho, rho)
    a = np.abs(cosine)
    if np.any((a > PARALLEL_TOL) & (a < 1.0 - PARALLEL_TOL)):
        return None, rho2, None
    same = a >= 1.0 - PARALLEL_TOL
    groups, seen = [], np.zeros(live.size, dtype=bool)
    signs = np.ones(mat.shape[0])
    for i in range(live.size):
        if seen[i]:
            continue
        members = np.nonzero(same[i])[0]
        # parallelism must be an equivalence relation
        if np.any(seen[members]) or not np.all(same[np.ix_(members, members)]):
            return None, rho2, None
        seen[members] = True
        signs[live[members]] = np.sign(cosine[i, members])
        groups.append(live[members])
    return groups, rho2, signs


def structured_svd(mat, full_matrices=False, **kw):
    """``numpy.linalg.svd`` replacement (reduced form) with the canonical tie-break; LAPACK for unstructured input."""
    mat = np.asarray(mat)
    if full_matrices or kw or mat.ndim != 2 or np.iscomplexobj(mat) or not np.all(np.isfinite(mat)):
        STATS["lapack"] += 1
        return _LAPACK_SVD(mat, full_matrices=full_matrices, **kw)
    m, n = mat.shape
    groups, rho2, signs = row_groups(mat)
    if groups is None or len(groups) > min(m, n):
        STATS["lapack"] += 1
        return _LAPACK_SVD(mat, full_matrices=False)
    STATS["structured"] += 1
    kdim = min(m, n)
    u = np.zeros((m, kdim))
    s = np.zeros(kdim)
    vh = np.zeros((kdim, n))
    if not groups:
        return u, s, vh
    sig2 = np.array([sum(rho2[r] for r in g) for g in groups])     # row order, like the engine's column sums
    first = np.array([g[0] for g in groups])
    floor2 = sig2.max() * (4.0 * EPS) ** 2
    order = canonical_order(sig2, first)
    pos = 0
    for gi in order:
        if not sig2[gi] > floor2:
            continue
        g = groups[gi]
        sg = np.sqrt(sig2[gi])
        s[pos] = sg
        u[g, pos] = signs[g] * np.sqrt(rho2[g]) * (1.0 / sg)
        vh[pos] = (u[g, pos] @ mat[g]) * (1.0 / sg)
        pos += 1
    return u, s, vh


@contextlib.contextmanager
def installed():
    """Install the hook where the reference (and the oracle port) pick up their SVD: ``numpy.linalg.svd``."""
    prev = np.linalg.svd
    np.linalg.svd = structured_svd
    try:
        yield STATS
    finally:
        np.linalg.svd = prev
