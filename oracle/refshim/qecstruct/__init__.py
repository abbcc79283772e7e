"""Duck-typed stand-in for qecstruct 0.2.9 (Rust wheel, absent). TEST INFRASTRUCTURE ONLY.

Implements exactly the surface ``examples/decoding/decoding.py`` touches: ``CssCode``
with ``*_binary()`` matrices exposing ``rows()/num_rows()/num_columns()``, ``repetition_code`` and
``hypergraph_product`` following the convention reverse-engineered from the printed 3x3 code
(``examples/decoding/quantum_surface.ipynb:79-94``).
"""


class BinaryMatrix:
    def __init__(self, num_columns, rows):
        self._n = int(num_columns)
        self._rows = [sorted(int(c) for c in r) for r in rows]

    def rows(self):
        return [list(r) for r in self._rows]

    def num_rows(self):
        return len(self._rows)

    def num_columns(self):
        return self._n


class BinaryVector:  # import-time name only
    pass


class BinarySymmetricChannel:  # import-time name only
    pass


class Rng:  # import-time name only
    def __init__(self, *a, **k):
        pass


class LinearCode:
    def __init__(self, parity_rows, num_bits):
        self._h = BinaryMatrix(num_bits, parity_rows)

    def par_mat(self):
        return self._h

    def __len__(self):
        return self._h.num_columns()


class CssCode:
    def __init__(self, n, x_stabs, z_stabs, x_logicals, z_logicals):
        self._n = int(n)
        self._xs = BinaryMatrix(n, x_stabs)
        self._zs = BinaryMatrix(n, z_stabs)
        self._xl = BinaryMatrix(n, x_logicals)
        self._zl = BinaryMatrix(n, z_logicals)

    def __len__(self):
        return self._n

    def num_x_logicals(self):
        return self._xl.num_rows()

    def num_z_logicals(self):
        return self._zl.num_rows()

    def num_x_stabs(self):
        return self._xs.num_rows()

    def num_z_stabs(self):
        return self._zs.num_rows()

    def x_stabs_binary(self):
        return self._xs

    def z_stabs_binary(self):
        return self._zs

    def x_logicals_binary(self):
        return self._xl

    def z_logicals_binary(self):
        return self._zl


def repetition_code(length):
    return LinearCode([[i, i + 1] for i in range(length - 1)], length)


def hypergraph_product(code_a, code_b):
    """HGP of two classical codes; only the repetition x repetition case has pinned logicals."""
    ha, hb = code_a.par_mat(), code_b.par_mat()
    if ha.rows() != hb.rows() or ha.num_columns() != hb.num_columns():
        raise NotImplementedError("stand-in supports hypergraph_product(C, C) only")
    h = ha.rows()
    n, m = ha.num_columns(), ha.num_rows()
    nq = n * n + m * m
    x_stabs = []
    for i in range(n):
        for c in range(m):
            s = [i * n + j for j in h[c]] + [n * n + r * m + c for r in range(m) if i in h[r]]
            x_stabs.append(sorted(s))
    z_stabs = []
    for r in range(m):
        for j in range(n):
            s = [i * n + j for i in h[r]] + [n * n + r * m + c for c in range(m) if j in h[c]]
            z_stabs.append(sorted(s))
    x_logicals = [[n - 1 + i * n for i in range(n)]]
    z_logicals = [list(range(n))]
    return CssCode(nq, x_stabs, z_stabs, x_logicals, z_logicals)
