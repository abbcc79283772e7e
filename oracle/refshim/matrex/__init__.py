"""Stand-in for quicophy/matrex 0.0.1 (absent: no source, no wheel, no network). TEST INFRASTRUCTURE ONLY.

``msro(location_matrix) -> row order``. The real algorithm (matrix front minimisation) is not
available, so parity on the "Optimised" order is UNPINNED: the order is treated as an explicit input
shared by the oracle and the CUDA path.  ``ORDER_OVERRIDE`` lets a test inject an order; otherwise rows
are sorted by (first column, last column), the same rule ``mdopt_b200.ordering.msro`` ships.
"""
import numpy as np

ORDER_OVERRIDE = None


def msro(matrix):
    if ORDER_OVERRIDE is not None:
        return list(ORDER_OVERRIDE(matrix))
    m = np.asarray(matrix)
    keys = []
    for r in range(m.shape[0]):
        nz = np.nonzero(m[r])[0]
        keys.append((int(nz[0]) if nz.size else 0, int(nz[-1]) if nz.size else 0, r))
    return [k[2] for k in sorted(keys)]
