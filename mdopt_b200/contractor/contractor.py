"""MPS-MPO contractor on the GPU (mirror of ``mdopt/contractor/contractor.py``)."""
from __future__ import annotations

from typing import List, Union

import numpy as np

from mdopt_b200.mps.canonical import CanonicalMPS
from mdopt_b200.mps.explicit import ExplicitMPS


def apply_one_site_operator(tensor: np.ndarray, operator: np.ndarray) -> np.ndarray:
    """``t'[i,l,k] = sum_j t[i,j,k] op[j,l]`` (reference contractor.py:16-63).

    A (vL,2,vR) x (2,2) relabelling with no reduction over bonds; the batched decode path applies the bias
    inside the decode kernel, this helper only serves API parity on single tensors.
    """
    tensor, operator = np.asarray(tensor), np.asarray(operator)
    if tensor.ndim != 3:
        raise ValueError(f"A valid MPS tensor must have 3 legs while the one given has {tensor.ndim}.")
    if operator.ndim != 2:
        raise ValueError(f"A valid one-site operator must have 2 legs while the one given has {operator.ndim}.")
    return np.einsum("ijk,jl->ilk", tensor, operator)


def mps_mpo_contract(mps: Union[ExplicitMPS, CanonicalMPS], mpo: List[np.ndarray], start_site: int = 0,
                     renormalise: bool = False, chi_max: int = int(1e4), cut: float = float(1e-17),
                     inplace: bool = False, mode: str = "svd") -> CanonicalMPS:
    """Apply an MPO to a canonical MPS, left to right, truncating every bond to ``chi_max``
    (reference contractor.py:177-378).  Runs as one CTA of the GPU engine; ``mode="svd"`` follows the
    reference's SVD truncation rule on every split, ``mode="fast"`` uses rank-revealing QR where no
    truncation is needed (user ``cut`` must then be negligible)."""
    from mdopt_b200 import engine, schedule

    if isinstance(mps, ExplicitMPS):
        work = mps.mixed_canonical(start_site)
    else:
        work = mps if inplace else mps.copy()
    for i, tensor in enumerate(mpo):
        if np.ndim(tensor) != 4:
            raise ValueError(f"A valid MPO tensor must have 4 legs while tensor {i} has {np.ndim(tensor)}.")
    if start_site + len(mpo) > len(work):
        raise ValueError(
            "The length of the MPO should correspond to the number of sites where it is applied, "
            f"given MPO of length {len(mpo)} and the starting site {start_site}, "
            f"while the MPS has length {len(work)}."
        )
    if len(mpo) < 2:
        raise ValueError("the MPO must cover at least two sites")
    centre = work.orth_centre if work.orth_centre is not None else work.resolve_orth_centre()
    from mdopt_b200 import generic

    if generic.needs_generic(work.tensors, mpo):
        # complex dtype, physical dimension != 2, MPO bond > 2 or bonds beyond the compiled capacities: the same
        # zip-up through the library tier (cuBLAS / cuSOLVER on the device), see mdopt_b200/generic.py
        tensors, new_centre = generic.mps_mpo_contract_generic(work.tensors, centre, mpo, start_site, renormalise,
                                                               chi_max, cut, work.chi_max)
        out = CanonicalMPS(tensors, new_centre, work.tolerance, work.chi_max)
        if inplace and isinstance(mps, CanonicalMPS):
            mps.tensors, mps.orth_centre = out.tensors, out.orth_centre
            return mps
        return out
    table = schedule.TensorTable()
    ids = [table.add(np.asarray(t)) for t in mpo]
    ops = np.array([schedule.OP_SWEEP, start_site, len(mpo), 0, int(bool(renormalise))] + ids + [schedule.OP_END],
                   dtype=np.int32)
    wtab, wdim = table.arrays()
    res = engine.run_program_on_mps(work.tensors, centre, ops, wtab, wdim, chi_max, cut, False, mode=mode)
    tensors = res["tensors"]
    if renormalise:  # reference quirk: the *left neighbour* of the new centre is normalised (:363-369)
        idx = start_site + len(mpo) - 2
        nrm = float(np.linalg.norm(tensors[idx]))
        if nrm != 0:
            tensors[idx] = tensors[idx] / nrm
    out = CanonicalMPS(tensors, res["centre"], work.tolerance, work.chi_max)
    if inplace and isinstance(mps, CanonicalMPS):
        mps.tensors, mps.orth_centre = out.tensors, out.orth_centre
        return mps
    return out
