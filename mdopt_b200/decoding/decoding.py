"""MPS decoder for CSS codes on the GPU (mirror of ``examples/decoding/decoding.py`` for the hot path).

``decode_css`` keeps the reference's signature and return convention (decoding.py:1164-1404);
``decode_css_batch`` is the additive batched entry the GPU path is built around: all syndromes that share
the literal-character flags (has 'X', has 'Z') run the same error-independent program, one CTA each.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from mdopt_b200 import schedule

_PROGRAM_CACHE: Dict[Tuple, schedule.Program] = {}


def pauli_to_mps(pauli_string: str) -> str:
    """"I"->"00", "X"->"10", "Z"->"01", "Y"->"11", "E"->"++" (reference decoding.py:283-320)."""
    table = {"I": "00", "X": "10", "Z": "01", "Y": "11", "E": "++"}
    out = []
    for pauli in pauli_string:
        if pauli not in table:
            raise ValueError(f"Invalid Pauli encountered -- {pauli}.")
        out.append(table[pauli])
    return "".join(out)


def bitflip_bias(prob_bias: float = float(0.5)) -> np.ndarray:
    """One-site bias operator (reference decoding.py:49-89); applied inside the decode kernel."""
    if not 0 <= prob_bias <= 1:
        raise ValueError(f"The channel parameter `prob_bias` should be a probability, given {prob_bias}.")
    op = np.full((2, 2), np.sqrt(prob_bias))
    np.fill_diagonal(op, np.sqrt(1 - prob_bias))
    return op


_PAULI_BITS = {"I": 0, "X": 1, "Z": 2, "Y": 3}
_BITS_PAULI = "IXZY"


def multiply_pauli_strings(pauli1: str, pauli2: str) -> str:
    """Site-wise product of two Pauli strings, phases dropped (reference decoding.py:1042-1095)."""
    if len(pauli1) != len(pauli2):
        raise ValueError(f"The Pauli strings must have the same length, but got {len(pauli1)} and {len(pauli2)}.")
    return "".join(_BITS_PAULI[_PAULI_BITS[a] ^ _PAULI_BITS[b]] for a, b in zip(pauli1, pauli2))


def css_code_stabilisers(code) -> Tuple[List[str], List[str]]:
    """Stabilisers of a CSS code as Pauli strings (reference decoding.py:463-509).  As there, the rows of
    ``x_stabs_binary()`` (the checks that detect X errors) are written with the letter Z and the rows of
    ``z_stabs_binary()`` with X (:492-507)."""
    n = len(code)

    def strings(matrix, pauli):
        out = []
        for row in matrix.rows():
            support = {int(c) for c in row}
            out.append("".join(pauli if q in support else "I" for q in range(n)))
        return out

    return strings(code.x_stabs_binary(), "Z"), strings(code.z_stabs_binary(), "X")


def _state_errors(errors: Sequence[str], stabilisers: Sequence[str], rng) -> List[str]:
    """``multiply_by_stabiliser=True`` (decoding.py:1237-1240, 1480-1483): the product state is built from the
    error times one randomly drawn stabiliser; erasure errors are left alone.  The reference draws from numpy's
    global RNG; pass ``rng`` (a numpy Generator) for a reproducible draw."""
    out = []
    for err in errors:
        if "E" in err or err == "I" * len(err):
            out.append(err)
            continue
        pick = rng.integers(len(stabilisers)) if rng is not None else np.random.randint(len(stabilisers))
        out.append(multiply_pauli_strings(err, stabilisers[int(pick)]))
    return out


def css_code_constraint_sites(code):
    return schedule.code_site_lists(code)[0]


def css_code_logicals_sites(code):
    return schedule.code_site_lists(code)[1]


def _resolve_mode(mode: str, depolarising: bool, cut: float) -> str:
    """``"auto"``: the monomial warp engine for the bit-flip bias (product-state inputs and 0/1 constraint tensors keep
    every site tensor p-monomial, which the engine verifies; anything it hands back is re-run on the dense engine),
    the dense SVD mode for the depolarising bias (the two-site bias gates mix the basis).  An explicit ``"fast"`` with
    a cut that discards singular values is run as ``"svd"``: the QR steps do not apply ``cut``."""
    if mode == "auto":
        mode = "svd" if depolarising else "mono"
    if mode == "mono" and depolarising:
        mode = "svd"
    if mode == "fast" and cut > 1e-14:
        mode = "svd"
    return mode


def _effective_chi(chi_max, n_sites: int) -> int:
    """chi_max clipped to the largest Schmidt rank an n_sites-qubit chain can have (2^floor(N/2)): a larger
    chi_max never truncates, so short chains may keep the reference's default chi_max=1e4."""
    return int(min(int(chi_max), 1 << min(n_sites // 2, 30)))


def _code_key(code) -> Tuple:
    """Structural identity of a code object (object ids are recycled once a code is garbage collected)."""
    mats = (code.x_stabs_binary(), code.z_stabs_binary(), code.x_logicals_binary(), code.z_logicals_binary())
    return (len(code),) + tuple(tuple(tuple(int(c) for c in row) for row in m.rows()) for m in mats)


def _program_for(code, has_x, has_z, renormalise, strategy, orders, bias_type="Bitflip", bias_prob=0.1, num_runs=1):
    depol = bias_type != "Bitflip"
    key = (_code_key(code), has_x, has_z, bool(renormalise), strategy, int(num_runs),
           None if not orders else tuple(sorted((k, v if isinstance(v, str) else tuple(v)) for k, v in orders.items())),
           depol, float(bias_prob) if depol else None)
    if key not in _PROGRAM_CACHE:
        _PROGRAM_CACHE[key] = schedule.build_decode_program(code, has_x, has_z, renormalise, strategy, orders,
                                                            "Depolarising" if depol else "Bitflip", bias_prob,
                                                            num_runs)
    return _PROGRAM_CACHE[key]


def decode_css_batch(code, errors: Sequence[str], chi_max: int = int(1e4), cut: float = float(1e-17),
                     bias_type: str = "Depolarising", bias_prob: float = float(0.1), renormalise: bool = True,
                     silent: bool = True, contraction_strategy: str = "Naive", tolerance: float = float(1e-12),
                     orders: Optional[dict] = None, mode: str = "auto", return_status: bool = False,
                     multiply_by_stabiliser: bool = False, rng=None, num_runs: int = 1):
    """Decode many independent syndromes on the current GPU.

    Returns ``(dist float64[B, 2^k], success int32[B])`` (plus ``status int32[B]`` if asked); row b equals
    what ``decode_css(code, errors[b], ...)`` returns.  All-identity errors take the reference's early exit
    (``[1, 0, 0, 0], 1``, decoding.py:1225-1228).

    Codes with more than 12 logical sites take the reference's Dephasing-DMRG read-out (decoding.py:1382-1400,
    ``num_runs`` sweeps from |+...+>), run inside the monomial kernel: the first array is then ``bits float64[B, k]``
    -- the main component found, in the order of the reversed logical chain the reference's engine works on
    (``engine.mps``) -- and the second the overlap with |0...0> (1 or 0).  Bit-flip bias only.

    ``mode``: ``"mono"`` is the monomial warp engine (one warp per syndrome, site tensors as 2*chi (value, target)
    pairs; the structure is verified at every step and a syndrome that leaves it is re-run on the dense engine).
    ``"svd"`` runs the reference algorithm split by split on dense tensors (SVD with ``k = min(chi_max, #(s > cut))``,
    orthogonality checks in front of the Jacobi rotations); ``"svd-gram"`` is ``"svd"`` with the deterministic Gram
    check only; ``"fast"`` replaces non-truncating splits by rank-revealing QR steps (needs ``cut <= 1e-14``).
    ``"auto"`` (default) is ``"mono"`` for the bit-flip bias and ``"svd"`` for the depolarising bias.
    """
    from mdopt_b200 import engine

    depol = bias_type != "Bitflip"  # any other string takes the depolarising branch (decoding.py:1267-1283)
    if not 0 <= bias_prob <= 1:
        raise ValueError(f"The channel parameter should be a probability, given {bias_prob} at site 0.")
    k = code.num_x_logicals() + code.num_z_logicals()
    n = len(code)
    n_sites = 2 * n + k
    mode = _resolve_mode(mode, depol, cut)
    errors = list(errors)
    B = len(errors)
    if B == 0:
        empty = (np.zeros((0, k if k > 12 else 1 << k)), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
        return empty if return_status else empty[:2]
    n_out = k if k > 12 else (1 << k)
    dist = np.zeros((B, n_out))
    success = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    if k > 12 and depol:
        raise NotImplementedError("more than 12 logical sites: the Dephasing-DMRG read-out needs the bit-flip bias "
                                  "(the monomial engine); the depolarising bias is dense")
    # whole-batch character tests (decoding.py:1225-1231: trivial / "X" in error / "Z" in error), no per-string Python
    raw, flags = schedule.pack_errors_with_flags(errors, n)
    trivial = flags == 0
    has_x = (flags & schedule.FLAG_X) != 0
    has_z = (flags & schedule.FLAG_Z) != 0
    erased = (flags & schedule.FLAG_E) != 0
    if k > 12:
        success[trivial] = 1     # all bits 0, overlap 1 (the reference's early exit returns [1, 0, 0, 0], 1)
    else:
        dist[trivial, 0], success[trivial] = 1.0, 1
    default_mask = (1 << k) - 1
    groups: Dict[Tuple[bool, bool, int], np.ndarray] = {}
    plain = ~trivial & ~erased
    for hx in (False, True):
        for hz in (False, True):
            idx = np.nonzero(plain & (has_x == hx) & (has_z == hz))[0]
            if idx.size:
                groups[(hx, hz, default_mask)] = idx
    by_mask: Dict[Tuple[bool, bool, int], List[int]] = {}
    for b in np.nonzero(~trivial & erased)[0]:
        by_mask.setdefault((bool(has_x[b]), bool(has_z[b]), schedule.readout_kept_mask(errors[b], k)), []).append(int(b))
    for key, rows in by_mask.items():
        rows = np.asarray(rows)
        groups[key] = np.sort(np.concatenate([groups[key], rows])) if key in groups else rows
    if not groups:
        return (dist, success, status) if return_status else (dist, success)
    # the constraint groups stay gated by the characters of the ORIGINAL error (decoding.py:1230-1231 run before
    # the multiplication at :1237-1240); only the product state sees the multiplied string
    state_raw = raw
    if multiply_by_stabiliser:
        stabs_x, stabs_z = css_code_stabilisers(code)
        state_raw = schedule.pack_errors(_state_errors(errors, stabs_x + stabs_z, rng), n)
    eng = engine.get_engine(n_sites, k, _effective_chi(chi_max, n_sites))
    for (hx, hz, mask), idx in groups.items():
        prog = _program_for(code, hx, hz, renormalise, contraction_strategy, orders, bias_type, bias_prob, num_runs)
        syms = schedule.symbols_from_raw(state_raw[idx], k)
        d, s, st = eng.decode_host(prog, syms, cut=cut, bias_p=(-1.0 if depol else bias_prob),
                                   renormalise=renormalise, mode=mode,
                                   kept_mask=(0 if mask == default_mask else mask))
        dist[idx], success[idx], status[idx] = d, s, st
        if not silent:
            logging.info("decoded %d syndromes with flags X=%s Z=%s", len(idx), hx, hz)
    return (dist, success, status) if return_status else (dist, success)


def decode_css(code, error: str, num_runs: int = int(1), chi_max: int = int(1e4), cut: float = float(1e-17),
               bias_type: str = "Depolarising", bias_prob: float = float(0.1), renormalise: bool = True,
               multiply_by_stabiliser: bool = False, silent: bool = False, contraction_strategy: str = "Naive",
               optimiser: str = "Dephasing DMRG", tolerance: float = float(1e-12), orders: Optional[dict] = None,
               mode: str = "auto", rng=None):
    """Error-based decoding of a CSS code by MPS marginalisation (reference decoding.py:1164-1404).

    Returns ``(|marginal| over the 2^k logical classes, success)`` with index 0 = I, 1 = X, 2 = Z, 3 = Y for
    k = 2 logical sites, and ``([1.0, 0.0, 0.0, 0.0], 1)`` for an all-identity error.
    """
    if not silent:
        logging.info("Starting the decoding.")
    if error == "I" * len(error):
        if not silent:
            logging.info("No error detected.")
        return [1.0, 0.0, 0.0, 0.0], 1
    num_logicals = code.num_x_logicals() + code.num_z_logicals()
    num_sites = 2 * len(code) + num_logicals
    if 2 * len(error) != num_sites - num_logicals:
        raise ValueError(f"The error length is {2 * len(error)}, expected {num_sites - num_logicals}.")
    pauli_to_mps(error)  # validates the characters
    dist, success = decode_css_batch(code, [error], chi_max=chi_max, cut=cut, bias_type=bias_type,
                                     bias_prob=bias_prob, renormalise=renormalise, silent=True,
                                     contraction_strategy=contraction_strategy, tolerance=tolerance, orders=orders,
                                     mode=mode, multiply_by_stabiliser=multiply_by_stabiliser, rng=rng,
                                     num_runs=num_runs)
    if num_logicals > 12:
        # decoding.py:1382-1400 returns (engine, overlap); what callers read from the engine is engine.mps
        return DephasingDMRGResult([int(b) for b in dist[0]]), float(success[0])
    return dist[0], int(success[0])


class DephasingDMRGResult:
    """What ``decode_css`` hands back in place of the reference's ``DephasingDMRG`` engine for codes with more than 12
    logical sites: ``bits`` -- the main component the sweeps ended on, one bit per site of the reversed logical chain
    -- and ``mps``, the same as the bond-dimension-1 ``CanonicalMPS`` the reference's ``engine.mps`` is."""

    def __init__(self, bits: List[int]) -> None:
        self.bits = list(bits)

    @property
    def mps(self):
        from mdopt_b200.mps.canonical import CanonicalMPS

        tensors = [np.eye(2)[b].reshape(1, 2, 1) for b in self.bits]
        return CanonicalMPS(tensors, 0)


def decode_custom_batch(stabilizers: Sequence[str], x_logicals: Sequence[str], z_logicals: Sequence[str],
                        errors: Sequence[str], chi_max: int = int(1e4), cut: float = float(1e-17),
                        bias_type: str = "Depolarising", bias_prob: float = float(0.1), renormalise: bool = True,
                        silent: bool = True, contraction_strategy: str = "Naive", tolerance: float = float(1e-12),
                        orders: Optional[dict] = None, mode: str = "auto", return_status: bool = False,
                        multiply_by_stabiliser: bool = False, rng=None):
    """Batched ``decode_custom`` (reference decoding.py:1407-1628): a code given by stabiliser / logical Pauli
    strings; every constraint group is always applied, success uses the relative tolerance 1e-12 (:1600-1604)."""
    from mdopt_b200 import engine

    depol = bias_type != "Bitflip"
    if not 0 <= bias_prob <= 1:
        raise ValueError(f"The channel parameter should be a probability, given {bias_prob} at site 0.")
    k = len(x_logicals) + len(z_logicals)
    n = len(stabilizers[0])
    n_sites = 2 * n + k
    if not 1 <= k <= 12:
        raise NotImplementedError("the dense read-out supports 1..12 logical sites")
    mode = _resolve_mode(mode, depol, cut)
    errors = list(errors)
    B = len(errors)
    dist = np.zeros((B, 1 << k))
    success = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    default_mask = (1 << k) - 1
    groups: Dict[int, List[int]] = {}
    for b, err in enumerate(errors):
        if len(err) != n:
            raise ValueError(f"The error length is {2 * len(err)}, expected {2 * n}.")
        if err == "I" * n:
            dist[b, 0], success[b] = 1.0, 1   # early exit returns [1, 0, 0, 0] (decoding.py:1472-1475)
        else:
            groups.setdefault(schedule.readout_kept_mask(err, k) if "E" in err else default_mask, []).append(b)
    if groups:
        key = ("custom", tuple(stabilizers), tuple(x_logicals), tuple(z_logicals), bool(renormalise),
               contraction_strategy,
               None if not orders else tuple(sorted((kk, v if isinstance(v, str) else tuple(v)) for kk, v in orders.items())),
               depol, float(bias_prob) if depol else None)
        if key not in _PROGRAM_CACHE:
            _PROGRAM_CACHE[key] = schedule.build_custom_program(
                stabilizers, x_logicals, z_logicals, renormalise, contraction_strategy, orders,
                "Depolarising" if depol else "Bitflip", bias_prob)
        prog = _PROGRAM_CACHE[key]
        eng = engine.get_engine(n_sites, k, _effective_chi(chi_max, n_sites))
        state_errors = _state_errors(errors, list(stabilizers), rng) if multiply_by_stabiliser else errors
        for mask, idx in groups.items():
            syms = schedule.symbols_from_errors([state_errors[b] for b in idx], k)
            d, s, st = eng.decode_host(prog, syms, cut=cut, bias_p=(-1.0 if depol else bias_prob),
                                       renormalise=renormalise, mode=mode,
                                       kept_mask=(0 if mask == default_mask else mask), readout=(1e-12, 0.0))
            dist[idx], success[idx], status[idx] = d, s, st
    return (dist, success, status) if return_status else (dist, success)


def decode_custom(stabilizers: List[str], x_logicals: List[str], z_logicals: List[str], error: str,
                  num_runs: int = int(1), chi_max: int = int(1e4), cut: float = float(1e-17),
                  bias_type: str = "Depolarising", bias_prob: float = float(0.1), renormalise: bool = True,
                  multiply_by_stabiliser: bool = False, silent: bool = False, contraction_strategy: str = "Naive",
                  optimiser: str = "Dephasing DMRG", tolerance: float = float(1e-12), orders: Optional[dict] = None,
                  mode: str = "auto", rng=None):
    """Error-based decoding of a code given by Pauli strings (reference decoding.py:1407-1628)."""
    if not silent:
        logging.info("Starting the decoding.")
    if error == "I" * len(error):
        if not silent:
            logging.info("No error detected.")
        return [1.0, 0.0, 0.0, 0.0], 1
    num_logicals = len(x_logicals) + len(z_logicals)
    num_sites = len(stabilizers[0]) * 2 + num_logicals
    if 2 * len(error) != num_sites - num_logicals:
        raise ValueError(f"The error length is {2 * len(error)}, expected {num_sites - num_logicals}.")
    pauli_to_mps(error)
    dist, success = decode_custom_batch(stabilizers, x_logicals, z_logicals, [error], chi_max=chi_max, cut=cut,
                                        bias_type=bias_type, bias_prob=bias_prob, renormalise=renormalise,
                                        contraction_strategy=contraction_strategy, tolerance=tolerance,
                                        orders=orders, mode=mode, multiply_by_stabiliser=multiply_by_stabiliser,
                                        rng=rng)
    return dist[0], int(success[0])
