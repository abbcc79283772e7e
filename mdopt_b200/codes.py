"""Duck-typed CSS code objects (the surface the decoder touches).

The reference takes ``qecstruct`` code objects (a Rust wheel that is not available here); the decoder only
calls ``len(code)``, ``num_{x,z}_logicals()``, ``num_{x,z}_stabs()`` and ``{x,z}_{stabs,logicals}_binary()``
whose results expose ``rows()/num_rows()/num_columns()`` (decoding.py:489-556, 624-634, 1244-1245).  Any
object with that surface works, including real ``qecstruct`` codes.
"""
from __future__ import annotations

from typing import List, Sequence


class BinaryMatrix:
    def __init__(self, num_columns: int, rows: Sequence[Sequence[int]]) -> None:
        self._n = int(num_columns)
        self._rows = [sorted(int(c) for c in r) for r in rows]

    def rows(self) -> List[List[int]]:
        return [list(r) for r in self._rows]

    def num_rows(self) -> int:
        return len(self._rows)

    def num_columns(self) -> int:
        return self._n


class CssCode:
    def __init__(self, n, x_stabs, z_stabs, x_logicals, z_logicals) -> None:
        self._n = int(n)
        self._xs, self._zs = BinaryMatrix(n, x_stabs), BinaryMatrix(n, z_stabs)
        self._xl, self._zl = BinaryMatrix(n, x_logicals), BinaryMatrix(n, z_logicals)

    def __len__(self) -> int:
        return self._n

    def num_x_logicals(self) -> int:
        return self._xl.num_rows()

    def num_z_logicals(self) -> int:
        return self._zl.num_rows()

    def num_x_stabs(self) -> int:
        return self._xs.num_rows()

    def num_z_stabs(self) -> int:
        return self._zs.num_rows()

    def x_stabs_binary(self) -> BinaryMatrix:
        return self._xs

    def z_stabs_binary(self) -> BinaryMatrix:
        return self._zs

    def x_logicals_binary(self) -> BinaryMatrix:
        return self._xl

    def z_logicals_binary(self) -> BinaryMatrix:
        return self._zl


def surface_code(lattice_size: int) -> CssCode:
    """``hypergraph_product(repetition_code(L), repetition_code(L))`` in the qubit/stabiliser order printed
    by the reference's 3x3 example (examples/decoding/quantum_surface.ipynb:79-94)."""
    L = int(lattice_size)
    h = [[i, i + 1] for i in range(L - 1)]
    n, m = L, L - 1
    xs = [sorted([i * n + j for j in h[c]] + [n * n + r * m + c for r in range(m) if i in h[r]])
          for i in range(n) for c in range(m)]
    zs = [sorted([i * n + j for i in h[r]] + [n * n + r * m + c for c in range(m) if j in h[c]])
          for r in range(m) for j in range(n)]
    return CssCode(n * n + m * m, xs, zs, [[n - 1 + i * n for i in range(n)]], [list(range(n))])
