"""Loads the host emulation of the CUDA engine (tests/emu/emu.cpp) for CPU-only logic tests.

TEST INFRASTRUCTURE ONLY.  The emulation compiles the same ``mdopt_b200/csrc/*.cuh`` sources with
``-DMDOPT_EMU`` so that pivoting, truncation rules, index maps and the program interpreter can be checked
against the oracle without a GPU.  The product package never loads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from mdopt_b200 import _abi, schedule

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "emu", "libmdopt_emu.so")
SRC = [os.path.join(HERE, "emu", "emu.cpp"), os.path.join(ROOT, "mdopt_b200", "csrc", "block_ops.cuh"),
       os.path.join(ROOT, "mdopt_b200", "csrc", "mps_core.cuh"), os.path.join(ROOT, "mdopt_b200", "csrc", "mono_core.cuh"),
       os.path.join(ROOT, "include", "mdopt_b200.h")]


def build(force=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in SRC):
        return SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMDOPT_EMU", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                           "-o", SO, SRC[0]])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def decode(program: schedule.Program, symbols, chi_max, cut, bias_p, renormalise=True, mode=schedule.MODE_FAST,
           trace_max=0, rank_tol=_abi.DEFAULT_RANK_TOL, kept_mask=0, readout=(1e-9, 1e-12), mono=False, cap=None):
    """``mono=True`` runs the monomial (warp-per-syndrome) engine of mono_core.cuh instead of the dense one."""
    if cap is None:
        cap = (_abi.mono_chi_capacity(max(chi_max, 1)) if mono else
               _abi.chi_capacity(max(chi_max, 1, _abi.readout_min_capacity(program.n_logical))))
    B = symbols.shape[0]
    ncls = program.n_logical if (mono and program.n_logical > 12) else 1 << program.n_logical
    stride = (4 + 2 * cap) if trace_max else 0
    desc = _abi.make_desc(program.n_sites, program.n_logical, chi_max, mode, renormalise, program.ops.size, cut,
                          bias_p, stride, trace_max, rank_tol=rank_tol, kept_mask=kept_mask,
                          readout_rel=readout[0], readout_abs=readout[1])
    dist = np.zeros((B, ncls))
    success = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    trace = np.zeros((B, max(trace_max, 1), max(stride, 1)))
    counters = np.zeros(8)
    syms = np.ascontiguousarray(symbols, dtype=np.uint8)
    if mono:
        rc = lib().emu_mono_decode_batch(C.byref(desc), cap, _ptr(program.ops), _ptr(program.wtab),
                                         _ptr(program.wdim), int(program.wtab.shape[0]), _ptr(syms), B, _ptr(dist),
                                         _ptr(success), _ptr(status), _ptr(trace), _ptr(counters))
    else:
        rc = lib().emu_decode_batch(C.byref(desc), cap, _ptr(program.ops), _ptr(program.wtab), _ptr(program.wdim),
                                    _ptr(syms), B, _ptr(dist), _ptr(success), _ptr(status), _ptr(trace),
                                    _ptr(counters))
    assert rc == 0
    return dist, success, status, trace, counters


def program(tensors, centre, ops, wtab, wdim, chi_max, cut, mode=schedule.MODE_SVD, cap=None, renormalise=False,
            trace_max=0):
    """Run an op stream on one MPS through the host build of the engine (the mps_program kernel's body)."""
    from mdopt_b200 import engine

    if cap is None:
        cap = _abi.chi_capacity(max(max(t.shape[0], t.shape[2]) for t in tensors))
    sites, bonds = engine.pack_mps(tensors, cap)
    ops = np.ascontiguousarray(ops, dtype=np.int32)
    stride = (4 + 2 * cap) if trace_max else 0
    desc = _abi.make_desc(len(tensors), 1, min(int(chi_max), 1 << 30), mode, renormalise, ops.size, cut, -1.0,
                          stride, trace_max)
    centre_io = np.array([centre], dtype=np.int32)
    status = np.zeros(1, dtype=np.int32)
    trace = np.zeros((max(trace_max, 1), max(stride, 1)))
    rc = lib().emu_mps_program(C.byref(desc), cap, _ptr(ops), _ptr(wtab), _ptr(wdim), _ptr(sites), _ptr(bonds),
                               _ptr(centre_io), None, None, _ptr(status), _ptr(trace) if trace_max else None)
    assert rc == 0
    if trace_max:
        return engine.unpack_mps(sites, bonds), int(centre_io[0]), int(status[0]), trace
    return engine.unpack_mps(sites, bonds), int(centre_io[0]), int(status[0])


def svd(a, chi_max, cut):
    a = np.ascontiguousarray(a, dtype=float)
    m, n = a.shape
    cap = _abi.chi_capacity((max(m, n) + 1) // 2)
    k = min(m, n)
    u, s, vh = np.zeros((m, k)), np.zeros(k), np.zeros((k, n))
    kept = C.c_int(0)
    bad = lib().emu_svd(cap, m, n, _ptr(a), int(chi_max), C.c_double(cut), _ptr(u), _ptr(s), _ptr(vh), C.byref(kept))
    return u, s, vh, kept.value, bad


def split(a):
    """The SVD-mode split of the portable tier on one matrix: (route, X J, order, sigma); routes: 0 orthogonal columns,
    1 one round of disjoint pair rotations, 2 greedy rounds, 3 cyclic sweeps."""
    a = np.ascontiguousarray(a, dtype=float)
    m, n = a.shape
    cap = _abi.chi_capacity((max(m, n) + 1) // 2)
    x = np.zeros((m, n))
    order = np.zeros(n, dtype=np.int32)
    sig = np.zeros(n)
    route = lib().emu_split(cap, m, n, _ptr(a), _ptr(x), _ptr(order), _ptr(sig))
    return route, x, order, sig


def qr(a, kmax, rel_tol=1e-14):
    a = np.ascontiguousarray(a, dtype=float)
    m, n = a.shape
    cap = _abi.chi_capacity((max(m, n) + 1) // 2)
    q, coeff = np.zeros((m, min(m, n))), np.zeros((min(m, n), n))
    more = C.c_int(0)
    K = lib().emu_qr(cap, m, n, _ptr(a), int(kmax), C.c_double(rel_tol), _ptr(q), _ptr(coeff), C.byref(more))
    return q.ravel()[: m * K].reshape(m, K), coeff[:K], K, more.value
