"""Product-state constructors (mirror of ``mdopt/mps/utils.py:360-522``)."""
from __future__ import annotations

from typing import Union

import numpy as np

from mdopt_b200.mps.canonical import CanonicalMPS
from mdopt_b200.mps.explicit import ExplicitMPS


def _product(tensors, form, tolerance):
    singular_values = [[1.0] for _ in range(len(tensors) + 1)]
    if form == "Right-canonical":
        mps = ExplicitMPS(tensors, singular_values, tolerance=tolerance).right_canonical()
        mps.orth_centre = None
        return mps
    if form == "Left-canonical":
        mps = ExplicitMPS(tensors, singular_values, tolerance=tolerance).left_canonical()
        mps.orth_centre = None
        return mps
    if form == "Mixed-canonical":
        raise ValueError("Mixed-canonical form is not defined for a product state.")
    return ExplicitMPS(tensors, singular_values, tolerance=tolerance)


def create_simple_product_state(num_sites: int, which: str = "0", phys_dim: int = 2, form: str = "Right-canonical",
                                tolerance: float = float(1e-12)) -> Union[ExplicitMPS, CanonicalMPS]:
    tensor = np.zeros((phys_dim,))
    if which == "0":
        tensor[0] = 1.0
    if which == "1":
        tensor[-1] = 1.0
    if which == "+":
        tensor[:] = 1 / np.sqrt(phys_dim)
    return _product([tensor.reshape((1, phys_dim, 1)).copy() for _ in range(num_sites)], form, tolerance)


def create_custom_product_state(string: str, phys_dim: int = 2, form: str = "Right-canonical",
                                tolerance: float = float(1e-12)) -> Union[ExplicitMPS, CanonicalMPS]:
    tensors = []
    for symbol in string:
        tensor = np.zeros((phys_dim,))
        if symbol == "0":
            tensor[0] = 1.0
        elif symbol == "1":
            tensor[-1] = 1.0
        elif symbol == "+":
            tensor[:] = 1 / np.sqrt(phys_dim)
        else:
            raise ValueError(f"The string argument accepts only 0, 1 or + as elements, given {symbol}.")
        tensors.append(tensor.reshape((1, phys_dim, 1)))
    return _product(tensors, form, tolerance)


def find_orth_centre(mps: CanonicalMPS, return_orth_flags: bool = False, tolerance: float = 1e-12):
    if isinstance(mps, ExplicitMPS):
        raise ValueError("Orthogonality centre is undefined for an Explicit MPS instance.")
    return [mps.resolve_orth_centre(tolerance)]
