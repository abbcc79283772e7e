"""The C-ABI library loads on a CPU-only box and exports every symbol include/mdopt_b200.h declares."""
import ctypes
import os
import re

from mdopt_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdopt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdopt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    names = declared_symbols()
    assert set(names) == set(_abi.PROTOTYPES), (names, sorted(_abi.PROTOTYPES))
    lib = _lib.load()
    for name in names:
        assert getattr(lib, name) is not None
    assert lib.mdopt_b200_abi_version() == 1


def test_pure_host_queries():
    lib = _lib.load()
    assert [lib.mdopt_chi_capacity(c) for c in (1, 8, 9, 16, 33, 64, 65, 128, 200, 256, 257)] == \
        [8, 8, 16, 16, 64, 64, 128, 128, 256, 256, 0]
    assert [_abi.chi_capacity(c) for c in (1, 64, 65, 256, 257)] == [8, 64, 128, 256, 0]
    assert lib.mdopt_site_stride(64) == 64 * 2 * 64
    assert lib.mdopt_mps_program_workspace_bytes(64, 2) == 2 * 2 * 128 * 128 * 8
    # the large-chi tier adds its two work matrices (2chi x (2chi+1) and chi x (2chi+1)) to the workspace
    assert lib.mdopt_mps_program_workspace_bytes(128, 1) == (2 * 256 * 256 + 3 * 128 * 257) * 8


def test_desc_layout_matches_header():
    assert ctypes.sizeof(_abi.DecodeDesc) == 8 * 4 + 4 * 8 + 2 * 4 + 2 * 8
