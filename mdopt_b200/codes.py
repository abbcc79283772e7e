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


# --------------------------------------------------------------------------------------------------------------
# Classical codes and the hypergraph product: the constructors the reference's drivers take from ``qecstruct``
# (``repetition_code``, ``random_regular_code``, ``hypergraph_product``; quantum_surface.py:133-134,
# quantum_hypergraph_product.py:134-142).  ``qecstruct`` is a Rust wheel that is not available here: the qubit /
# stabiliser ORDER follows the convention its printed 3x3 surface code reveals; the logical
# representatives of a general product and the random stream of ``random_regular_code`` are this module's own
# (documented below) -- decoding is well defined for any valid choice, but class labels are not comparable with a
# run that used ``qecstruct`` objects.  Real ``qecstruct`` codes can still be passed to every decoder entry point.
# --------------------------------------------------------------------------------------------------------------


class LinearCode:
    """Classical binary linear code given by its parity-check rows (the slice of qecstruct.LinearCode the drivers use)."""

    def __init__(self, parity_rows: Sequence[Sequence[int]], num_bits: int) -> None:
        self._h = BinaryMatrix(num_bits, parity_rows)

    def par_mat(self) -> BinaryMatrix:
        return self._h

    def num_checks(self) -> int:
        return self._h.num_rows()

    def __len__(self) -> int:
        return self._h.num_columns()


def repetition_code(length: int) -> LinearCode:
    return LinearCode([[i, i + 1] for i in range(int(length) - 1)], int(length))


def random_regular_code(num_bits: int, num_checks: int, bit_degree: int, check_degree: int, rng=None) -> LinearCode:
    """A (bit_degree, check_degree)-regular LDPC code: every bit in ``bit_degree`` checks, every check on
    ``check_degree`` bits, drawn from the configuration model with rejection of repeated edges.  ``rng`` is a numpy
    Generator or a seed (the reference passes ``qecstruct.Rng(seed)``, whose stream cannot be reproduced)."""
    import numpy as np

    if num_bits * bit_degree != num_checks * check_degree:
        raise ValueError("The Tanner graph of the code must be bipartite.")   # quantum_hypergraph_product.py:136
    gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
    bit_stubs = np.repeat(np.arange(num_bits), bit_degree)
    for _ in range(10000):
        perm = gen.permutation(bit_stubs)
        rows = perm.reshape(num_checks, check_degree)
        if all(len(set(r)) == check_degree for r in rows):
            return LinearCode([sorted(int(b) for b in r) for r in rows], num_bits)
    raise RuntimeError("could not draw a simple regular Tanner graph")


def _gf2_rank_reduce(rows: List[int]) -> List[int]:
    """Row-reduced basis of the span of bit-mask rows (pivot = highest set bit)."""
    basis: List[int] = []
    for v in rows:
        for b in basis:
            v = min(v, v ^ b)
        if v:
            basis.append(v)
            basis.sort(reverse=True)
    return basis


def _gf2_kernel(rows: List[int], n: int) -> List[int]:
    """Basis of {x : row . x = 0 for every row} over GF(2), as bit masks over n columns."""
    pivots, red = {}, []
    for v in rows:                       # Gaussian elimination to reduced row echelon form
        for pc, r in pivots.items():
            if (v >> pc) & 1:
                v ^= red[r]
        if v:
            pc = v.bit_length() - 1
            for i in range(len(red)):
                if (red[i] >> pc) & 1:
                    red[i] ^= v
            pivots[pc] = len(red)
            red.append(v)
    free = [c for c in range(n) if c not in pivots]
    out = []
    for f in free:
        x = 1 << f
        for pc, r in pivots.items():
            if (red[r] >> f) & 1:
                x |= 1 << pc
        out.append(x)
    return out


def _logicals(commute_with: List[List[int]], modulo: List[List[int]], n: int) -> List[List[int]]:
    """Representatives of ker(commute_with) / rowspace(modulo): vectors of the kernel, lowest weight first, that are
    independent of the stabilisers of their own type and of each other."""
    mask = lambda sup: sum(1 << int(c) for c in sup)  # noqa: E731
    kernel = sorted(_gf2_kernel([mask(r) for r in commute_with], n), key=lambda v: (bin(v).count("1"), v))
    span = _gf2_rank_reduce([mask(r) for r in modulo])
    out = []
    for v in kernel:
        w = v
        for b in span:
            w = min(w, w ^ b)
        if w:
            span.append(w)
            span.sort(reverse=True)
            out.append([c for c in range(n) if (v >> c) & 1])
    return out


def hypergraph_product(code_a, code_b) -> CssCode:
    """Hypergraph product of two classical codes with parity-check matrices H_a (m_a x n_a) and H_b (m_b x n_b):
    qubits (i, j) -> i*n_b + j for i < n_a, j < n_b, then (r, c) -> n_a*n_b + r*m_b + c for r < m_a, c < m_b;
    X stabiliser (i, c): {(i, j): j in H_b[c]} + {(r, c): i in H_a[r]}; Z stabiliser (r, j): {(i, j): i in H_a[r]} +
    {(r, c): j in H_b[c]} -- the order of the reference's printed 3x3 surface code (quantum_surface.ipynb:79-94).
    Repetition x repetition keeps the printed logical representatives (last column / first row); other products
    get lowest-weight representatives of ker(H_Z) / rowspace(H_X) and ker(H_X) / rowspace(H_Z)."""
    ha, hb = code_a.par_mat(), code_b.par_mat()
    na, nb, ma, mb = ha.num_columns(), hb.num_columns(), ha.num_rows(), hb.num_rows()
    ra, rb = ha.rows(), hb.rows()
    n = na * nb + ma * mb
    xs = [sorted([i * nb + j for j in rb[c]] + [na * nb + r * mb + c for r in range(ma) if i in ra[r]])
          for i in range(na) for c in range(mb)]
    zs = [sorted([i * nb + j for i in ra[r]] + [na * nb + r * mb + c for c in range(mb) if j in rb[c]])
          for r in range(ma) for j in range(nb)]
    rep = lambda rows, nn: rows == [[i, i + 1] for i in range(nn - 1)]  # noqa: E731
    if na == nb and rep(ra, na) and rep(rb, nb):
        return CssCode(n, xs, zs, [[nb - 1 + i * nb for i in range(na)]], [list(range(nb))])
    return CssCode(n, xs, zs, _logicals(zs, xs, n), _logicals(xs, zs, n))
