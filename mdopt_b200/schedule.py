"""Error-independent device programs for the decode path (host logic, no GPU needed).

A *program* is the int32 opcode stream ``libmdopt_b200`` interprets for every syndrome of a batch
(``include/mdopt_b200.h``): one ``OP_SWEEP`` per constraint string, in the order ``apply_constraints``
would apply them, then ``OP_MARGINAL``.  It depends only on the code and on which constraint groups are
active, so it is built once per (code, has-X, has-Z) and reused for the whole batch.  Reference call
sequence: ``examples/decoding/decoding.py:1285-1356``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from mdopt_b200.optimiser.utils import (COPY_LEFT, SWAP, XOR_BULK, XOR_LEFT, XOR_RIGHT, ConstraintString,
                                        front_minimising_order, msro)

OP_END, OP_SWEEP, OP_MOVE, OP_MARGINAL, OP_GATE2, OP_COMPRESS, OP_MARGINAL_DMRG = 0, 1, 2, 3, 4, 5, 6
MODE_FAST, MODE_SVD, MODE_SVD_GRAM, MODE_MONO = 0, 1, 2, 3
SYM_ZERO, SYM_ONE, SYM_PLUS = 0, 1, 2

# decoding.py:283-320; "E" (erasure) becomes "++"
_PAULI_BITS = {"I": (0, 0), "X": (1, 0), "Z": (0, 1), "Y": (1, 1), "E": (2, 2)}


class TensorTable:
    """De-duplicated table of MPO site tensors, padded to 16 doubles, plus (vL, vR, isometric) flags."""

    def __init__(self) -> None:
        self._ids: Dict[bytes, int] = {}
        self.data: List[np.ndarray] = []
        self.dims: List[Tuple[int, int, int, int]] = []

    @staticmethod
    def is_isometric(w: np.ndarray) -> bool:
        vl, vr = w.shape[0], w.shape[1]
        mat = np.transpose(w, (0, 2, 1, 3)).reshape(vl * 2, vr * 2)
        return bool(np.allclose(mat @ mat.T, np.eye(vl * 2), atol=1e-14))

    def add(self, w: np.ndarray) -> int:
        w = np.ascontiguousarray(w, dtype=np.float64)
        if w.ndim != 4:
            raise ValueError(f"A valid MPO tensor must have 4 legs while the one given has {w.ndim}.")
        if w.shape[2] != 2 or w.shape[3] != 2 or w.shape[0] > 2 or w.shape[1] > 2:
            raise ValueError(f"MPO tensor of shape {w.shape} exceeds the engine's (<=2, <=2, 2, 2) capacity")
        key = bytes(w.shape) + w.tobytes()
        if key not in self._ids:
            self._ids[key] = len(self.data)
            flat = np.zeros(16)
            flat[: w.size] = w.ravel()
            self.data.append(flat)
            self.dims.append((w.shape[0], w.shape[1], int(self.is_isometric(w)), 0))
        return self._ids[key]

    def add_gate(self, gate: np.ndarray) -> int:
        """Two-site gate with legs (pUL, pUR, pDL, pDR) = (j, l, n, o), stored as 16 doubles."""
        g = np.ascontiguousarray(gate, dtype=np.float64)
        if g.shape != (2, 2, 2, 2):
            raise ValueError("a two-site gate must have shape (2, 2, 2, 2)")
        key = b"gate" + g.tobytes()
        if key not in self._ids:
            self._ids[key] = len(self.data)
            self.data.append(g.ravel().copy())
            self.dims.append((2, 2, 0, 1))
        return self._ids[key]

    def arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        return (np.ascontiguousarray(np.stack(self.data), dtype=np.float64),
                np.ascontiguousarray(np.array(self.dims, dtype=np.int32)))


@dataclass
class Program:
    ops: np.ndarray            # int32 opcode stream
    wtab: np.ndarray           # float64 [n_tensors, 16]
    wdim: np.ndarray           # int32 [n_tensors, 4]
    n_sites: int
    n_logical: int
    n_strings: int = 0
    spans: List[Tuple[int, int]] = field(default_factory=list)


def code_site_lists(code):
    """Constraint-site groups of a CSS code object (duck type: SURVEY.md 8b).

    Returns ``((x_checks, z_checks), (x_logicals, z_logicals))``; every entry is
    ``[[first], [bulk xor sites], [swap sites], [last]]``.  X-type supports sit at ``2q + k``, Z-type at
    ``2q + k + 1`` with ``k`` logical sites in front (reference decoding.py:512-696).
    """
    k = code.num_x_logicals() + code.num_z_logicals()

    def supports(matrix, offset):
        return [[2 * int(q) + k + offset for q in sorted(row)] for row in matrix.rows()]

    def check_groups(sup):
        inner = sup[1:-1]
        return [[sup[0]], inner, [s for s in range(sup[0] + 1, sup[-1]) if s not in inner], [sup[-1]]]

    def logical_groups(index, sup):
        bulk = sup[:-1]
        return [[index], bulk, [s for s in range(index + 1, sup[-1]) if s not in bulk], [sup[-1]]]

    xs, zs = supports(code.x_stabs_binary(), 0), supports(code.z_stabs_binary(), 1)
    xl, zl = supports(code.x_logicals_binary(), 0), supports(code.z_logicals_binary(), 1)
    checks = ([check_groups(s) for s in xs], [check_groups(s) for s in zs])
    logicals = ([logical_groups(i, s) for i, s in enumerate(xl)],
                [logical_groups(len(xl) + i, s) for i, s in enumerate(zl)])
    return checks, logicals


def _pauli_support(pauli: str, offset: int) -> List[int]:
    """Sites (2q + offset for the X part, 2q + offset + 1 for the Z part) of a Pauli string (decoding.py:755-781)."""
    out = []
    for q, ch in enumerate(pauli):
        if ch not in _PAULI_BITS or ch == "E":
            raise ValueError(f"Invalid Pauli encountered -- {ch}.")
        x, z = _PAULI_BITS[ch]
        if x:
            out.append(2 * q + offset)
        if z:
            out.append(2 * q + offset + 1)
    return out


def custom_site_lists(stabilizers: Sequence[str], x_logicals: Sequence[str], z_logicals: Sequence[str]):
    """Constraint-site groups of a code given by Pauli strings (reference decoding.py:755-916): returns
    ``(checks, (x_logical_groups, z_logical_groups))`` in the same ``[[first], [bulk], [swap], [last]]`` form."""
    k = len(x_logicals) + len(z_logicals)

    def check_groups(sup):
        inner = sup[1:-1]
        return [[sup[0]], inner, [s for s in range(sup[0] + 1, sup[-1]) if s not in inner], [sup[-1]]]

    def logical_groups(index, sup):
        bulk = sup[:-1]
        return [[index], bulk, [s for s in range(index + 1, sup[-1]) if s not in bulk], [sup[-1]]]

    checks = [check_groups(_pauli_support(st, k)) for st in stabilizers]
    xl = [logical_groups(i, _pauli_support(lg, k)) for i, lg in enumerate(x_logicals)]
    zl = [logical_groups(len(x_logicals) + i, _pauli_support(lg, k)) for i, lg in enumerate(z_logicals)]
    return checks, (xl, zl)


def build_custom_program(stabilizers: Sequence[str], x_logicals: Sequence[str], z_logicals: Sequence[str],
                         renormalise: bool = True, strategy: str = "Naive", orders: Optional[dict] = None,
                         bias_type: str = "Bitflip", bias_prob: float = 0.1) -> Program:
    """Program for ``decode_custom`` (decoding.py:1407-1628): X logicals, Z logicals, then all checks in one
    group -- no gating by the error's characters."""
    k = len(x_logicals) + len(z_logicals)
    n_sites = 2 * len(stabilizers[0]) + k
    checks, (xl, zl) = custom_site_lists(stabilizers, x_logicals, z_logicals)
    orders = orders or {}
    table, ops, spans = TensorTable(), [], []
    if bias_type != "Bitflip":
        gid = table.add_gate(depolarising_bias(bias_prob))
        for site in range(k, n_sites - 1):
            ops.extend([OP_GATE2, site, gid, int(bool(renormalise))])
    logical_t = [COPY_LEFT, XOR_BULK, SWAP, XOR_RIGHT]
    check_t = [XOR_LEFT, XOR_BULK, SWAP, XOR_RIGHT]
    for strings, tensors, key in ((xl, logical_t, "xl"), (zl, logical_t, "zl"), (checks, check_t, "c")):
        ordered = order_strings(strings, n_sites, strategy, orders.get(key))
        append_strings(ops, table, ordered, tensors, n_sites, renormalise, spans)
    ops.extend([OP_MARGINAL, OP_END])
    wtab, wdim = table.arrays()
    return Program(np.array(ops, dtype=np.int32), wtab, wdim, n_sites, k, len(spans), spans)


def order_strings(strings, n_sites: int, strategy: str, order: Optional[Sequence[int]] = None):
    """"Optimised" reorders the strings (optimiser/utils.py:270-279); "Naive" keeps them."""
    if strategy != "Optimised":
        return list(strings)
    if order is None or isinstance(order, str):
        loc = np.zeros((len(strings), n_sites), dtype=int)
        for r, groups in enumerate(strings):
            for group in groups:
                for site in group:
                    loc[r, site] = 1
        if order is None:
            order = msro(loc)
        elif order == "front":
            order = front_minimising_order(loc)
        else:
            raise ValueError(f"unknown ordering {order!r} (None, 'front' or an explicit list of rows)")
    return [strings[i] for i in order]


def append_strings(ops: List[int], table: TensorTable, strings, tensors, n_sites: int, renormalise: bool,
                   spans: List[Tuple[int, int]], renorm_spectrum: bool = False) -> None:
    for groups in strings:
        cs = ConstraintString(tensors, groups)
        start, span = cs.start(), cs.span()
        if span < 2:
            raise ValueError("a constraint string must cover at least two sites")
        if start + span > n_sites:
            raise ValueError(
                "The length of the MPO should correspond to the number of sites where it is applied, "
                f"given MPO of length {span} and the starting site {start}, while the MPS has length {n_sites}."
            )
        ids = [table.add(cs.constraints[g]) for g in cs.tensor_slots()]
        ops.extend([OP_SWEEP, start, span, int(bool(renormalise)), int(bool(renorm_spectrum))] + ids)
        spans.append((start, span))


def depolarising_bias(prob_bias: float) -> np.ndarray:
    """Two-site bias operator (reference decoding.py:92-137; ``np.fill_diagonal`` on a 4-d array touches only
    the four [i,i,i,i]... entries with equal indices, i.e. [0,0,0,0] and [1,1,1,1])."""
    if not 0 <= prob_bias <= 1:
        raise ValueError(f"The channel parameter `prob_bias` should be a probability, given {prob_bias}.")
    op = np.full((2, 2, 2, 2), np.sqrt(prob_bias / 3))
    np.fill_diagonal(op, np.sqrt(1 - prob_bias))
    return op


def build_decode_program(code, has_x: bool, has_z: bool, renormalise: bool = True, strategy: str = "Naive",
                         orders: Optional[dict] = None, bias_type: str = "Bitflip",
                         bias_prob: float = 0.1, num_runs: int = 1) -> Program:
    """Program for ``decode_css`` on errors with the given literal-character flags (decoding.py:1230-1339).

    ``bias_type="Bitflip"`` is applied by the kernel while it builds the product state; any other value takes
    the reference's depolarising branch (decoding.py:1267-1283): centre to the first bias site, then one
    two-site gate + split on every adjacent pair of error sites (decoding.py:252-274)."""
    k = code.num_x_logicals() + code.num_z_logicals()
    n_sites = 2 * len(code) + k
    checks, logicals = code_site_lists(code)
    orders = orders or {}
    table, ops, spans = TensorTable(), [], []
    if bias_type != "Bitflip":
        gid = table.add_gate(depolarising_bias(bias_prob))
        for site in range(k, n_sites - 1):
            ops.extend([OP_GATE2, site, gid, int(bool(renormalise))])
    logical_t = [COPY_LEFT, XOR_BULK, SWAP, XOR_RIGHT]
    check_t = [XOR_LEFT, XOR_BULK, SWAP, XOR_RIGHT]
    plan = []
    if has_x:
        plan.append((logicals[0], logical_t, "xl"))
    if has_z:
        plan.append((logicals[1], logical_t, "zl"))
    if has_x:
        plan.append((checks[0], check_t, "xc"))
    if has_z:
        plan.append((checks[1], check_t, "zc"))
    for strings, tensors, key in plan:
        ordered = order_strings(strings, n_sites, strategy, orders.get(key))
        append_strings(ops, table, ordered, tensors, n_sites, renormalise, spans)
    if k > 12:   # decoding.py:1382-1400: the Dephasing-DMRG read-out, `num_runs` sweeps
        ops.extend([OP_MARGINAL_DMRG, max(int(num_runs), 1), OP_END])
    else:
        ops.extend([OP_MARGINAL, OP_END])
    if not table.data:
        table.add(SWAP)  # keep the table non-empty for programs without constraints
    wtab, wdim = table.arrays()
    return Program(np.array(ops, dtype=np.int32), wtab, wdim, n_sites, k, len(spans), spans)


def error_flags(error: str) -> Tuple[bool, bool, bool]:
    """(trivial, has_x, has_z) by literal characters, like the reference (decoding.py:1225-1231)."""
    return error == "I" * len(error), "X" in error, "Z" in error


def readout_kept_mask(error: str, n_logical: int) -> int:
    """Which leading sites survive the read-out of this error (bit j = site j is kept).

    Without erasures these are the ``n_logical`` logical sites.  The reference first marginalises the erased
    *qubit* indices as if they were *site* indices (decoding.py:1341-1347) and then everything from position
    ``n_logical`` of the shortened chain on (:1348-1356); together that keeps the first ``n_logical`` sites whose
    index is not the index of an erased qubit."""
    erased = {q for q, ch in enumerate(error) if ch == "E"}
    mask, kept, site = 0, 0, 0
    while kept < n_logical:
        if site not in erased:
            mask |= 1 << site
            kept += 1
        site += 1
        if site > 31:
            raise NotImplementedError("erasure pattern pushes the read-out beyond the first 32 sites")
    return mask


# Character tables for whole-batch packing (one table look-up per character, no per-string Python):
#   _FLAG[c]  : bit 0 = 'X', bit 1 = 'Z', bit 2 = 'Y', bit 3 = 'E', bit 7 = not a Pauli character ('I' -> 0)
#   _SYM16[c] : the two product-state symbols of the qubit as one little-endian uint16 (low byte = X site, high = Z site)
_FLAG = np.full(256, 128, dtype=np.uint8)
_SYM16 = np.zeros(256, dtype=np.uint16)
for _ch, _bits in _PAULI_BITS.items():
    _FLAG[ord(_ch)] = {"I": 0, "X": 1, "Z": 2, "Y": 4, "E": 8}[_ch]
    _SYM16[ord(_ch)] = _bits[0] | (_bits[1] << 8)
FLAG_X, FLAG_Z, FLAG_Y, FLAG_E = 1, 2, 4, 8


def pack_errors(errors: Sequence[str], n: int) -> np.ndarray:
    """uint8 [B, n] array of the Pauli characters of a batch of error strings, validated like the reference does
    (length: decoding.py:1249-1252, characters: :318).  One buffer for the whole batch: the per-string work of a
    100k-syndrome batch must stay out of Python loops."""
    return pack_errors_with_flags(errors, n)[0]


def pack_errors_with_flags(errors: Sequence[str], n: int):
    """(raw uint8 [B, n], flags uint8 [B]): the characters and, per string, the OR of their _FLAG bits -- which is all
    the reference's literal-character tests need (all-'I': flags == 0; "X" in error: flags & FLAG_X; ...)."""
    B = len(errors)
    lengths = set(map(len, errors))
    if lengths - {n}:
        bad = next(e for e in errors if len(e) != n)
        raise ValueError(f"The error length is {2 * len(bad)}, expected {2 * n}.")
    try:
        raw = np.frombuffer("".join(errors).encode("ascii"), dtype=np.uint8).reshape(B, n)
    except UnicodeEncodeError:
        bad = next(ch for e in errors for ch in e if ord(ch) > 127)
        raise ValueError(f"Invalid Pauli encountered -- {bad}.") from None
    flags = np.bitwise_or.reduce(_FLAG[raw], axis=1) if B else np.zeros(0, dtype=np.uint8)
    if (flags & 128).any():
        b = int(np.argmax(flags & 128))
        q = int(np.argmax(_FLAG[raw[b]] & 128))
        raise ValueError(f"Invalid Pauli encountered -- {errors[b][q]}.")
    return raw, flags


def symbols_from_raw(raw: np.ndarray, n_logical: int) -> np.ndarray:
    """uint8 [B, k + 2n] product-state symbols from packed Pauli characters: '+' on the logical sites, then the bits."""
    B, n = raw.shape
    out = np.empty((B, n_logical + 2 * n), dtype=np.uint8)
    out[:, :n_logical] = SYM_PLUS
    out[:, n_logical:] = _SYM16[raw].view(np.uint8).reshape(B, 2 * n)   # little-endian: X-site symbol, Z-site symbol
    return out


def symbols_from_errors(errors: Sequence[str], n_logical: int) -> np.ndarray:
    """uint8 [B, k + 2n] product-state symbols: '+' on the logical sites, then the Pauli bits."""
    return symbols_from_raw(pack_errors(errors, len(errors[0])), n_logical)
