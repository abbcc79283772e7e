"""Logical-constraint MPO tensors and constraint strings (host side).

Mirror of the reference's ``mdopt/optimiser/utils.py`` for the decode path: the six 0/1 site tensors with
legs ``(vL, vR, pU, pD)`` (reference :23-43), ``ConstraintString`` (:141-213) and ``apply_constraints``
(:216-359).  The tensors are plain numpy constants; applying them to an MPS runs on the GPU.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def _table(shape, rule) -> np.ndarray:
    out = np.zeros(shape, dtype=float)
    for idx in np.ndindex(*shape):
        out[idx] = 1.0 if rule(*idx) else 0.0
    return out


# (vL, vR, pU, pD); pU is contracted with the MPS physical leg.
IDENTITY = _table((1, 1, 2, 2), lambda l, r, u, d: u == d)
COPY_LEFT = _table((1, 2, 2, 2), lambda l, r, u, d: r == d)                      # ignores pU, emits pD = vR
XOR_BULK = _table((2, 2, 2, 2), lambda l, r, u, d: r == (l ^ u) and u == d)       # vR = vL xor p
XOR_LEFT = _table((1, 2, 2, 2), lambda l, r, u, d: r == u and u == d)             # vR = p
XOR_RIGHT = _table((2, 1, 2, 2), lambda l, r, u, d: l == u and u == d)            # keep vL == p
SWAP = _table((2, 2, 2, 2), lambda l, r, u, d: l == r and u == d)                 # pass both through


def _front_profile(spans, order, n_cols):
    """(largest front, summed front) of applying the strings in `order`: the front after each string is the number
    of sites covered by an applied string that still have unapplied strings crossing them."""
    cover, ahead = np.zeros(n_cols, dtype=int), np.zeros(n_cols, dtype=int)
    for lo, hi in spans:
        ahead[lo:hi + 1] += 1
    worst = total = 0
    for r in order:
        lo, hi = spans[r]
        cover[lo:hi + 1] += 1
        ahead[lo:hi + 1] -= 1
        front = int(np.count_nonzero((cover > 0) & (ahead > 0)))
        worst, total = max(worst, front), total + front
    return worst, total


def front_minimising_order(location_matrix: np.ndarray) -> List[int]:
    """A matrix-front minimiser for the "Optimised" strategy (what ``matrex.msro`` is described as; the real routine
    is unavailable -- no source, wheel or recorded order in the reference tree).  The *front* after applying a
    string is the number of MPS sites covered by an applied string that still have unapplied strings crossing them,
    which is what drives the intermediate bond dimensions.  Candidates: the (first site, last site) sweep, the
    (last site, first site) sweep, and a greedy order that always applies the string leaving the smallest front; the
    candidate with the smallest (largest front, summed front) wins.  Select it with ``order="front"``
    (``orders={"xc": "front"}`` in the decoders); the default stays the (first site, last site) order the reference
    fixtures were generated with.  Untruncated results do not depend on the order."""
    m = np.asarray(location_matrix) != 0
    n_rows, n_cols = m.shape
    spans = []
    for r in range(n_rows):
        nz = np.nonzero(m[r])[0]
        spans.append((int(nz[0]), int(nz[-1])) if nz.size else (0, 0))
    remaining = set(range(n_rows))
    cover, ahead = np.zeros(n_cols, dtype=int), np.zeros(n_cols, dtype=int)
    for lo, hi in spans:
        ahead[lo:hi + 1] += 1
    greedy: List[int] = []
    while remaining:
        best, best_key = None, None
        for r in remaining:
            lo, hi = spans[r]
            cov, ah = cover.copy(), ahead.copy()
            cov[lo:hi + 1] += 1
            ah[lo:hi + 1] -= 1
            key = (int(np.count_nonzero((cov > 0) & (ah > 0))), lo, hi, r)
            if best_key is None or key < best_key:
                best, best_key = r, key
        lo, hi = spans[best]
        cover[lo:hi + 1] += 1
        ahead[lo:hi + 1] -= 1
        remaining.remove(best)
        greedy.append(best)
    first_last = sorted(range(n_rows), key=lambda r: (spans[r][0], spans[r][1], r))
    last_first = sorted(range(n_rows), key=lambda r: (spans[r][1], spans[r][0], r))
    return min((first_last, last_first, greedy), key=lambda o: _front_profile(spans, o, n_cols))


def msro(location_matrix: np.ndarray) -> List[int]:
    """Default row order for the "Optimised" strategy, standing in for ``matrex.msro`` (optimiser/utils.py:270-279):
    strings sorted by (first site, last site), which keeps the active front moving left to right.  The order is an
    explicit input shared with the oracle (``order=`` / ``orders=``); ``front_minimising_order`` is the alternative."""
    m = np.asarray(location_matrix)
    keys = []
    for r in range(m.shape[0]):
        nz = np.nonzero(m[r])[0]
        keys.append((int(nz[0]) if nz.size else 0, int(nz[-1]) if nz.size else 0, r))
    return [k[2] for k in sorted(keys)]


class ConstraintString:
    """A string of logical constraints laid over a contiguous span of sites (reference :141-213)."""

    def __init__(self, constraints: Sequence[np.ndarray], sites: Sequence[Sequence[int]]) -> None:
        self.constraints = [np.asarray(c) for c in constraints]
        self.sites = [list(group) for group in sites]
        if not self.constraints:
            raise ValueError("Empty list of constraints passed.")
        if not self.sites:
            raise ValueError("Empty list of sites passed.")
        if len(self.sites) != len(self.constraints):
            raise ValueError(
                f"We have {len(self.constraints)} constraint(s) in the constraints list, "
                f"but {len(self.sites)} constraint(s) assumed by the sites list."
            )
        flat = self.flat()
        if len(set(flat)) != len(flat):
            raise ValueError("Non-unique sites encountered in the list.")
        ordered = sorted(flat)
        if ordered != list(range(ordered[0], ordered[-1] + 1)):
            raise ValueError("The string should not have breaks in it.")

    def __getitem__(self, index: int) -> tuple:
        return self.sites[index], self.constraints[index]

    def flat(self, sort: bool = False) -> List[int]:
        out = [s for group in self.sites for s in group]
        return sorted(out) if sort else out

    def span(self) -> int:
        flat = self.flat()
        return max(flat) - min(flat) + 1

    def start(self) -> int:
        return min(self.flat())

    def mpo(self) -> List[np.ndarray]:
        base = self.start()
        out: List[np.ndarray] = [np.array(None, dtype=float)] * self.span()
        for tensor, group in zip(self.constraints, self.sites):
            for site in group:
                out[site - base] = tensor
        return out

    def tensor_slots(self) -> List[int]:
        """Index into ``constraints`` for every site of the span (what the device program stores)."""
        base = self.start()
        out = [-1] * self.span()
        for g, group in enumerate(self.sites):
            for site in group:
                out[site - base] = g
        return out


def apply_constraints(mps, strings, logical_tensors, chi_max=int(1e4), cut=float(1e-17), renormalise=True,
                      strategy="Naive", silent=False, dense=False, return_entropies_and_bond_dims=False,
                      order=None, mode="svd"):
    """Apply constraint strings to an MPS on the GPU (reference :216-359).

    Returns a new ``CanonicalMPS``; the input is not modified.  ``dense`` and
    ``return_entropies_and_bond_dims`` are diagnostics of the reference that are not on the decode path
    and are not provided by the GPU engine.
    """
    if dense or return_entropies_and_bond_dims:
        raise NotImplementedError("dense / entropy diagnostics are outside the GPU decode path")
    from mdopt_b200.engine import run_constraints  # deferred: needs the CUDA library

    return run_constraints(mps, strings, logical_tensors, chi_max=chi_max, cut=cut, renormalise=renormalise,
                           strategy=strategy, order=order, mode=mode)
