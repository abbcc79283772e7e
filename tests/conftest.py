import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def block_422_code(num_blocks):
    """``num_blocks`` disjoint [[4,2,2]] blocks: n = 4*num_blocks qubits and 4*num_blocks logical sites (2 X- and 2 Z-
    logicals per block) -- a small CSS code whose read-out has 2^(4*num_blocks) classes.  Returns the CssCode
    constructor arguments (n, x_stabs, z_stabs, x_logicals, z_logicals)."""
    xs, zs, xl, zl = [], [], [], []
    for b in range(num_blocks):
        o = 4 * b
        xs.append([o, o + 1, o + 2, o + 3])
        zs.append([o, o + 1, o + 2, o + 3])
        xl += [[o, o + 1], [o, o + 2]]
        zl += [[o, o + 2], [o, o + 1]]
    return 4 * num_blocks, xs, zs, xl, zl


# errors on the block codes that the oracle decodes without any truncation at the stated chi_max
BLOCK_CASES = [(2, 64, ["XXIIIIII", "IYIIIYZZ", "IIIIIIZI"]),
               (3, 256, ["XXIIIIIIIYII", "IYZZIIIIIIZI", "XIIIIXIIXIII"])]
