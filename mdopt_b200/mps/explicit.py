"""``ExplicitMPS`` (Vidal Gamma/Lambda form): the constructor route for product states.

Mirror of the slice of ``mdopt/mps/explicit.py`` the decode path needs: the container (:36-105) and
``right_canonical`` / ``left_canonical`` / ``mixed_canonical`` (:312-381), which absorb the singular values
into the site tensors and hand back a ``CanonicalMPS``.
"""
from __future__ import annotations

from copy import deepcopy
from typing import List

import numpy as np

from mdopt_b200.mps.canonical import CanonicalMPS


class ExplicitMPS:
    def __init__(self, tensors: List[np.ndarray], singular_values: List[list], tolerance: float = float(1e-12),
                 chi_max: int = int(1e4)) -> None:
        self.tensors = tensors
        self.num_sites = len(tensors)
        self.num_bonds = self.num_sites - 1
        self.singular_values = singular_values
        self.tolerance = tolerance
        self.chi_max = chi_max
        self.dtype = tensors[0].dtype
        if len(tensors) != len(singular_values) - 1:
            raise ValueError(
                f"The number of singular values at the bonds ({len(singular_values)}) should be equal to "
                f"the number of sites plus one ({len(tensors) + 1})."
            )
        for i, tensor in enumerate(tensors):
            if tensor.ndim != 3:
                raise ValueError(f"A valid MPS tensor must have 3 legs while the one at site {i} has {tensor.ndim}.")
        for i, sv in enumerate(singular_values):
            norm = np.linalg.norm(sv)
            if abs(norm - 1) > tolerance:
                raise ValueError(f"The norm of each singular values list must be 1, instead the norm is {norm} at bond {i + 1}.")

    def __len__(self) -> int:
        return self.num_sites

    @property
    def bond_dimensions(self) -> List[int]:
        return [t.shape[-1] for t in self.tensors[:-1]]

    def copy(self) -> "ExplicitMPS":
        return ExplicitMPS(deepcopy(self.tensors), deepcopy(self.singular_values), self.tolerance, self.chi_max)

    def one_site_right_iso(self, site: int) -> np.ndarray:
        return self.tensors[site] * np.asarray(self.singular_values[site + 1])[None, None, :]

    def one_site_left_iso(self, site: int) -> np.ndarray:
        return self.tensors[site] * np.asarray(self.singular_values[site])[:, None, None]

    def right_canonical(self) -> CanonicalMPS:
        return CanonicalMPS([self.one_site_right_iso(i) for i in range(self.num_sites)], 0, self.tolerance,
                            self.chi_max)

    def left_canonical(self) -> CanonicalMPS:
        return CanonicalMPS([self.one_site_left_iso(i) for i in range(self.num_sites)], self.num_sites - 1,
                            self.tolerance, self.chi_max)

    def mixed_canonical(self, orth_centre: int) -> CanonicalMPS:
        if orth_centre not in range(self.num_sites):
            raise ValueError(
                "Orthogonality centre index must reside anywhere from site 0 to "
                f"{self.num_sites - 1}, the one given is at position {orth_centre}."
            )
        tensors = [self.one_site_left_iso(i) for i in range(orth_centre)]
        centre = self.one_site_left_iso(orth_centre) * np.asarray(self.singular_values[orth_centre + 1])[None, None, :]
        tensors.append(centre)
        tensors += [self.one_site_right_iso(i) for i in range(orth_centre + 1, self.num_sites)]
        return CanonicalMPS(tensors, orth_centre, self.tolerance, self.chi_max)
