"""``CanonicalMPS``: host container with GPU-backed operations.

Mirror of the part of ``mdopt/mps/canonical.py`` the decode path touches: constructor and attributes
(:87-140), ``copy`` (:142), ``reverse`` (:150), ``dense`` (:241), ``move_orth_centre`` (:345),
``compress_bond`` / ``compress`` (:684-904, "svd" strategy), ``marginal`` (:906).  Tensors stay host NumPy arrays (the reference returns host arrays even on its GPU
backend, contractor.py:375-376); ``move_orth_centre`` and the MPO contraction run on the GPU engine.
"""
from __future__ import annotations

from copy import deepcopy
from functools import reduce
from typing import List, Optional

import numpy as np


class CanonicalMPS:
    def __init__(self, tensors: List[np.ndarray], orth_centre: Optional[int] = None,
                 tolerance: float = float(1e-12), chi_max: int = int(1e4)) -> None:
        self.tensors = tensors
        self.num_sites = len(tensors)
        self.num_bonds = self.num_sites - 1
        self.orth_centre = orth_centre
        self.dtype = tensors[0].dtype
        self.tolerance = tolerance
        self.chi_max = chi_max
        if orth_centre and orth_centre not in range(self.num_sites):
            raise ValueError(
                f"The orthogonality centre index must reside anywhere from site 0 "
                f"to {self.num_sites-1}, the one given is at position {orth_centre}."
            )
        for i, tensor in enumerate(tensors):
            if tensor.ndim != 3:
                raise ValueError(f"A valid MPS tensor must have 3 legs while the one at site {i} has {tensor.ndim}.")

    @property
    def bond_dimensions(self) -> List[int]:
        return [t.shape[-1] for t in self.tensors[:-1]]

    @property
    def phys_dimensions(self) -> List[int]:
        return [t.shape[1] for t in self.tensors]

    def __len__(self) -> int:
        return self.num_sites

    def copy(self) -> "CanonicalMPS":
        return CanonicalMPS(deepcopy(self.tensors), self.orth_centre, self.tolerance, self.chi_max)

    def reverse(self) -> "CanonicalMPS":
        rev = [np.transpose(t) for t in reversed(self.tensors)]
        if self.orth_centre:  # 0 is falsy in the reference too (canonical.py:155)
            return CanonicalMPS(rev, self.num_sites - 1 - self.orth_centre, self.tolerance, self.chi_max)
        return CanonicalMPS(rev, None, self.tolerance, self.chi_max)

    def dense(self, flatten: bool = True, renormalise: bool = False, norm=2) -> np.ndarray:
        """Dense vector of a *small* MPS (host; used only for <= 12 logical sites in the reference)."""
        out = reduce(lambda a, b: np.tensordot(a, b, (-1, 0)), self.tensors).squeeze()
        if flatten:
            out = out.flatten()
        if renormalise:
            n = float(np.linalg.norm(out, ord=norm))
            if n > 1e-17:
                out = out / n
        return out

    def resolve_orth_centre(self, tol: float = 1e-12) -> int:
        """Orthogonality-centre bootstrap for an MPS whose ``orth_centre`` is None
        (mdopt/mps/utils.py:38-124 + mdopt/optimiser/utils.py:299-322)."""
        n = self.num_sites
        left, right, centres = [], [], []
        for i, t in enumerate(self.tensors):
            gl = np.einsum("ijk,ijl->kl", t, np.conjugate(t))
            gr = np.einsum("ijk,ljk->il", t, np.conjugate(t))
            fl = bool(np.isclose(np.linalg.norm(gl - np.eye(gl.shape[0])), 0, atol=tol))
            fr = bool(np.isclose(np.linalg.norm(gr - np.eye(gr.shape[0])), 0, atol=tol))
            left.append(fl)
            right.append(fr)
            if not fl and not fr:
                centres.append(i)
        neg = [not f for f in left]
        if left in ([True] + [False] * (n - 1), [False] * n) and right == neg:
            centres.append(0)
        if left in ([True] * (n - 1) + [False], [True] * n) and right == neg:
            centres.append(n - 1)
        if all(left) and all(right):
            centres.append(0)
        oc = centres[0] if len(centres) == 1 else None
        if all(left) and all(right):
            oc = 0
        elif left == [True] + [False] * (n - 1):
            oc = n - 1 if right == neg else oc
        elif left == [True] * (n - 1) + [False]:
            oc = 0 if right == neg else oc
        elif all(right):
            oc = 0
        elif all(left):
            oc = n - 1
        if len(centres) > 1:
            oc = centres[0]
        if oc is None:
            raise ValueError("The orthogonality centre value is set to None.")
        return oc

    def move_orth_centre(self, final_pos: int, return_singular_values: bool = False, renormalise: bool = True,
                         mode: str = "svd"):
        """Move the orthogonality centre on the GPU (reference canonical.py:345-420).

        ``renormalise`` of the per-bond spectra is not applied (the decode path always passes False).
        """
        if final_pos not in range(self.num_sites):
            raise ValueError(
                "The final position of the orthogonality centre should be"
                f"from 0 to {self.num_sites-1}, given {final_pos}."
            )
        if return_singular_values:
            raise NotImplementedError("return_singular_values is not provided by the GPU engine")
        centre = self.orth_centre if self.orth_centre is not None else self.resolve_orth_centre()
        if centre == final_pos:
            return self
        from mdopt_b200 import engine, generic, schedule

        if generic.needs_generic(self.tensors):   # complex / phys dim != 2 / large bonds: library tier (generic.py)
            tensors, c = generic.move_orth_centre_generic(self.tensors, centre, final_pos, self.chi_max)
            return CanonicalMPS(tensors, c, self.tolerance, self.chi_max)

        ops = np.array([schedule.OP_MOVE, final_pos, schedule.OP_END], dtype=np.int32)
        table = schedule.TensorTable()
        from mdopt_b200.optimiser.utils import IDENTITY

        table.add(IDENTITY)
        wtab, wdim = table.arrays()
        res = engine.run_program_on_mps(self.tensors, centre, ops, wtab, wdim, self.chi_max, 1e-12, False, mode=mode)
        return CanonicalMPS(res["tensors"], res["centre"], self.tolerance, self.chi_max)

    def _run_compress(self, bonds, chi_max, cut, renormalise, start_centre):
        """OP_COMPRESS for each bond in order on the GPU; returns (new MPS, truncation error per bond)."""
        from mdopt_b200 import engine, schedule
        from mdopt_b200.optimiser.utils import IDENTITY

        table = schedule.TensorTable()
        table.add(IDENTITY)
        wtab, wdim = table.arrays()
        ops = []
        for bond in bonds:
            ops.extend([schedule.OP_COMPRESS, int(bond), int(bool(renormalise))])
        ops.append(schedule.OP_END)
        chi = int(min(int(chi_max), 1 << 30))
        res = engine.run_program_on_mps(self.tensors, start_centre, np.array(ops, dtype=np.int32), wtab, wdim, chi,
                                        float(cut), False, mode="svd", trace_max=4 * self.num_sites + 8)
        errors = []
        for row in res["trace"]:
            rows, cols, kept, is_move = int(row[0]), int(row[1]), int(row[2]), int(row[3])
            if rows == 0 and cols == 0:
                break
            if not is_move:  # the compressing split itself; centre moves are traced with is_move = 1
                sv = row[4:4 + min(rows, cols)]
                errors.append(float(np.sum(sv[kept:] ** 2)))
        out = CanonicalMPS(res["tensors"], res["centre"], self.tolerance, self.chi_max)
        return out, errors

    def compress_bond(self, bond: int, chi_max: int = int(1e4), cut: float = float(1e-17),
                      renormalise: bool = False, strategy: str = "svd", return_truncation_error: bool = False):
        """Truncate one bond on the GPU (reference canonical.py:684-784): centre to ``bond``, SVD of the two-site
        tensor with ``k = min(chi_max, #(s > cut))``, singular values absorbed into site ``bond``.  Only the
        ``"svd"`` strategy runs on the engine."""
        if bond not in range(self.num_bonds):
            raise ValueError(f"Bond given {bond}, with the number of bonds in the MPS {self.num_bonds}.")
        if strategy not in ["svd", "qr", "svd_advanced"]:
            raise ValueError(f"Unsupported compression strategy: {strategy}")
        centre = self.orth_centre if self.orth_centre is not None else self.resolve_orth_centre()
        from mdopt_b200 import generic

        if strategy != "svd" or generic.needs_generic(self.tensors):
            # "qr" / "svd_advanced" (canonical.py:762-833), complex or large bonds: library tier, same steps
            tensors, c, err = generic.compress_bond_generic(self.tensors, centre, bond, chi_max, cut, renormalise,
                                                            self.chi_max, strategy, return_truncation_error)
            return (CanonicalMPS(tensors, c, self.tolerance, self.chi_max), (err if return_truncation_error else None))
        out, errors = self._run_compress([bond], chi_max, cut, renormalise, centre)
        return out, (errors[0] if return_truncation_error else None)

    def compress(self, chi_max: int = int(1e4), cut: float = float(1e-17), renormalise: bool = False,
                 strategy: str = "svd", return_truncation_errors: bool = False):
        """``compress_bond`` for every bond (reference canonical.py:836-904): the centre first goes to the nearer
        border; from the last site the chain is processed mirrored and the errors are reported in bond order."""
        if strategy not in ["svd", "qr", "svd_advanced"]:
            raise ValueError(f"Unsupported compression strategy: {strategy}")
        centre = self.orth_centre if self.orth_centre is not None else self.resolve_orth_centre()
        from_last = centre != 0 and (centre == self.num_sites - 1 or centre > int(self.num_bonds / 2))
        work = self
        if from_last:  # :888-889 -- work on the mirrored chain
            work = CanonicalMPS([np.ascontiguousarray(np.transpose(t)) for t in reversed(self.tensors)],
                                self.num_sites - 1 - centre, self.tolerance, self.chi_max)
            centre = work.orth_centre
        from mdopt_b200 import generic

        if strategy != "svd" or generic.needs_generic(work.tensors):
            out, errors = work, []
            for bond in range(self.num_bonds):
                out, err = out.compress_bond(bond, chi_max=chi_max, cut=cut, renormalise=renormalise,
                                             strategy=strategy, return_truncation_error=True)
                errors.append(err)
        else:
            out, errors = work._run_compress(list(range(self.num_bonds)), chi_max, cut, renormalise, centre)
        if from_last:
            out = CanonicalMPS([np.ascontiguousarray(np.transpose(t)) for t in reversed(out.tensors)],
                               self.num_sites - 1 - out.orth_centre, self.tolerance, self.chi_max)
            errors = errors[::-1]
        return out, (errors if return_truncation_errors else [None])

    def marginal(self, sites_to_marginalise: List[int], renormalise: bool = False) -> "CanonicalMPS":
        """Trace sites out with (1,..,1)/sqrt(d); MUTATES and returns self (reference canonical.py:906-991).

        The engine kernel implements the case the decoder uses -- a contiguous tail ``k..N-1`` traced into the leading
        ``k`` sites (decoding.py:1348-1356); any other site set runs the same chain on the device through the library
        tier (``mdopt_b200/generic.py``).
        """
        if not sites_to_marginalise:
            return self
        if set(sites_to_marginalise) - set(range(self.num_sites)):
            raise ValueError("The list of sites to marginalise must be a subset of the list of all sites.")
        sites = sorted(set(sites_to_marginalise))
        k = sites[0]
        from mdopt_b200 import generic

        if sites != list(range(k, self.num_sites)) or not 1 <= k <= 12 or generic.needs_generic(self.tensors):
            # any other site set (or dtype / dimensions outside the engine): the same chain through the library tier
            self.tensors[:] = generic.marginal_generic(self.tensors, sites, renormalise)
            self.num_sites = len(self.tensors)
            self.num_bonds = self.num_sites - 1
            self.orth_centre = self.num_sites - 1
            return self
        from mdopt_b200 import engine, schedule
        from mdopt_b200.optimiser.utils import IDENTITY

        table = schedule.TensorTable()
        table.add(IDENTITY)
        wtab, wdim = table.arrays()
        ops = np.array([schedule.OP_MARGINAL, schedule.OP_END], dtype=np.int32)
        res = engine.run_program_on_mps(self.tensors, self.orth_centre or 0, ops, wtab, wdim, self.chi_max, 1e-12,
                                        renormalise, n_logical=k, want_dist=True)
        self.tensors[:] = res["tensors"][:k]
        self.num_sites = k
        self.num_bonds = k - 1
        self.orth_centre = k - 1
        return self
