"""GPU runtime of the decode path: device buffers, program upload, kernel launches.

Everything numerical happens inside ``libmdopt_b200.so``; this module only moves bytes and launches.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from mdopt_b200 import _abi, _lib, schedule

_MODES = {"fast": schedule.MODE_FAST, "svd": schedule.MODE_SVD, "svd-gram": schedule.MODE_SVD_GRAM,
          "mono": schedule.MODE_MONO}
MONO_FALLBACK_MODE = "svd-gram"   # dense mode that re-runs syndromes the monomial engine hands back


def mode_id(mode) -> int:
    if isinstance(mode, str):
        if mode not in _MODES:
            raise ValueError(f"mode must be 'mono', 'fast', 'svd' or 'svd-gram', given {mode!r}")
        return _MODES[mode]
    return int(mode)


def program_is_monomial_candidate(program: schedule.Program) -> bool:
    """The monomial engine interprets sweeps, moves and the read-out; two-site gates and compression are dense."""
    ops, pc = program.ops, 0
    while pc < len(ops):
        op = int(ops[pc])
        if op == schedule.OP_END:
            return True
        if op == schedule.OP_SWEEP:
            pc += 5 + int(ops[pc + 2])
        elif op == schedule.OP_MOVE:
            pc += 2
        elif op == schedule.OP_MARGINAL:
            pc += 1
        elif op == schedule.OP_MARGINAL_DMRG:
            pc += 2
        else:
            return False
    return True


def _vp(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(torch) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class DeviceProgram:
    program: schedule.Program
    ops: "object"
    wtab: "object"
    wdim: "object"


class DecodeEngine:
    """Owns the per-device workspace and staging buffers; decodes batches that share a program.

    One engine serves every ``chi_max`` up to its compiled capacity (``chi_max`` travels in the descriptor of each
    call).  The workspaces of the two kernels -- the dense CTA-per-syndrome engine and the monomial warp-per-syndrome
    engine -- are allocated on first use and freed by ``release()``.  An engine owns ONE workspace, work counter and
    counter block per kernel: launches on it must be serialised on one stream (do not share an engine between
    concurrently running streams)."""

    def __init__(self, n_sites: int, n_logical: int, chi_max: int, device=None, max_batch: int = 1 << 16) -> None:
        self.lib, self.torch = _lib.require_gpu()
        torch = self.torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_sites, self.n_logical, self.chi_max = int(n_sites), int(n_logical), int(chi_max)
        self.cap = 0 if n_logical > 12 else _abi.chi_capacity(max(self.chi_max, _abi.readout_min_capacity(n_logical)))
        self.mono_cap = _abi.mono_chi_capacity(self.chi_max)
        if self.cap == 0 and self.mono_cap == 0:
            raise _lib.MdoptCudaError(
                f"chi_max={chi_max} exceeds the largest compiled capacity ({_abi.MONO_CAPACITIES[-1]})"
            )
        if not 1 <= self.n_logical <= 64:
            raise NotImplementedError("the read-outs support 1..64 logical sites")
        # more than 12 logical sites: Dephasing-DMRG read-out of the monomial engine (k doubles per syndrome)
        self.n_out = self.n_logical if self.n_logical > 12 else (1 << self.n_logical)
        if self.n_logical > 12:
            self.cap = 0   # the dense engine's read-out stops at 12 logical sites
        self.counters = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.workspace = None
        self.mono_workspace = None
        self.n_ctas = self.n_warps = 0
        self.max_batch = 0
        self._alloc(max_batch)
        self._programs: Dict[int, DeviceProgram] = {}

    def _ensure_dense(self) -> None:
        if self.workspace is not None:
            return
        if self.cap == 0:
            raise _lib.MdoptCudaError(
                f"chi_max={self.chi_max} exceeds the largest capacity of the dense engine ({_abi.MAX_CAPACITY})")
        torch = self.torch
        with torch.cuda.device(self.device):
            self.n_ctas = int(self.lib.mdopt_decode_num_ctas(self.cap))
            if self.n_ctas <= 0:
                raise _lib.MdoptCudaError("decode kernel cannot be made resident on this device")
            # every resident CTA owns one MPS slice; at large chi the slices are sized against free HBM
            free_bytes, _total = torch.cuda.mem_get_info()
            per_cta = int(self.lib.mdopt_decode_workspace_bytes(self.cap, self.n_sites, 1))
            self.n_ctas = max(1, min(self.n_ctas, int(0.8 * free_bytes) // max(per_cta, 1)))
            self.ws_bytes = int(self.lib.mdopt_decode_workspace_bytes(self.cap, self.n_sites, self.n_ctas))
            self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def _ensure_mono(self) -> None:
        if self.mono_workspace is not None:
            return
        if self.mono_cap == 0:
            raise _lib.MdoptCudaError(f"chi_max={self.chi_max} exceeds the monomial engine's largest capacity")
        torch = self.torch
        with torch.cuda.device(self.device):
            self.n_warps = int(self.lib.mdopt_mono_num_warps(self.mono_cap))
            if self.n_warps <= 0:
                raise _lib.MdoptCudaError("monomial decode kernel cannot be made resident on this device")
            free_bytes, _total = torch.cuda.mem_get_info()
            per_warp = int(self.lib.mdopt_mono_workspace_bytes(self.mono_cap, self.n_sites, 1))
            self.n_warps = max(1, min(self.n_warps, int(0.8 * free_bytes) // max(per_warp, 1)))
            self.mono_ws_bytes = int(self.lib.mdopt_mono_workspace_bytes(self.mono_cap, self.n_sites, self.n_warps))
            self.mono_workspace = torch.empty(self.mono_ws_bytes, dtype=torch.uint8, device=self.device)

    def release(self) -> None:
        """Free the device workspaces (they come back on the next launch)."""
        self.workspace = None
        self.mono_workspace = None

    def _alloc(self, batch: int) -> None:
        torch = self.torch
        ncls = self.n_out
        self.max_batch = batch
        self.d_sym = torch.empty((batch, self.n_sites), dtype=torch.uint8, device=self.device)
        self.d_dist = torch.empty((batch, ncls), dtype=torch.float64, device=self.device)
        self.d_success = torch.empty(batch, dtype=torch.int32, device=self.device)
        self.d_status = torch.empty(batch, dtype=torch.int32, device=self.device)
        self.h_sym = torch.empty((batch, self.n_sites), dtype=torch.uint8, pin_memory=True)
        self.h_dist = torch.empty((batch, ncls), dtype=torch.float64, pin_memory=True)
        self.h_success = torch.empty(batch, dtype=torch.int32, pin_memory=True)
        self.h_status = torch.empty(batch, dtype=torch.int32, pin_memory=True)

    def upload(self, program: schedule.Program) -> DeviceProgram:
        key = id(program)
        if key not in self._programs:
            torch = self.torch
            self._programs[key] = DeviceProgram(
                program,
                torch.from_numpy(program.ops).to(self.device),
                torch.from_numpy(program.wtab).to(self.device),
                torch.from_numpy(program.wdim).to(self.device),
            )
        return self._programs[key]

    def _desc(self, program: schedule.Program, cut, bias_p, renormalise, mode, trace_stride=0, trace_max=0,
              rank_tol=_abi.DEFAULT_RANK_TOL, kept_mask=0, readout=(1e-9, 1e-12)):
        return _abi.make_desc(self.n_sites, self.n_logical, self.chi_max, mode_id(mode), renormalise,
                              program.ops.size, cut, bias_p, trace_stride, trace_max, rank_tol=rank_tol,
                              kept_mask=kept_mask, readout_rel=readout[0], readout_abs=readout[1])

    # ---- inputs already resident in HBM -----------------------------------------------------------
    def decode_device(self, program, d_symbols, cut=1e-17, bias_p=0.1, renormalise=True, mode="fast",
                      d_dist=None, d_success=None, d_status=None, trace_max=0, rank_tol=_abi.DEFAULT_RANK_TOL):
        """Launch the fused decode kernel on device-resident symbols; returns device tensors (async)."""
        torch = self.torch
        dp = self.upload(program)
        batch = int(d_symbols.shape[0])
        ncls = self.n_out
        if d_dist is None:
            d_dist = torch.empty((batch, ncls), dtype=torch.float64, device=self.device)
        if d_success is None:
            d_success = torch.empty(batch, dtype=torch.int32, device=self.device)
        if d_status is None:
            d_status = torch.empty(batch, dtype=torch.int32, device=self.device)
        mono = mode_id(mode) == schedule.MODE_MONO
        cap = self.mono_cap if mono else self.cap
        stride = 4 + 2 * cap if trace_max else 0
        d_trace = torch.zeros((batch, trace_max, stride), dtype=torch.float64, device=self.device) if trace_max else None
        desc = self._desc(program, cut, bias_p, renormalise, mode, stride, trace_max, rank_tol)
        with torch.cuda.device(self.device):
            if mono:
                # no automatic dense re-run here: the caller sees MDOPT_ST_NOT_MONOMIAL in d_status (decode_host
                # and decode_css_batch do the re-run)
                self._ensure_mono()
                rc = self.lib.mdopt_mono_decode_batch_device(
                    C.byref(desc), self.mono_cap, self.n_warps, _vp(dp.ops), _vp(dp.wtab), _vp(dp.wdim),
                    dp.wtab.shape[0], _vp(d_symbols), batch, _vp(d_dist), _vp(d_success), _vp(d_status), _vp(d_trace),
                    _vp(self.counters), _vp(self.mono_workspace), self.mono_ws_bytes, _stream(torch))
            else:
                self._ensure_dense()
                rc = self.lib.mdopt_decode_batch_device(
                    C.byref(desc), self.cap, self.n_ctas, _vp(dp.ops), _vp(dp.wtab), _vp(dp.wdim), dp.wtab.shape[0],
                    _vp(d_symbols), batch, _vp(d_dist), _vp(d_success), _vp(d_status), _vp(d_trace),
                    _vp(self.counters), _vp(self.workspace), self.ws_bytes, _stream(torch))
        _lib.check(rc, "mdopt_mono_decode_batch_device" if mono else "mdopt_decode_batch_device")
        return d_dist, d_success, d_status, d_trace

    # ---- host buffers in, host buffers out (the drop-in call) --------------------------------------
    def decode_host(self, program, symbols: np.ndarray, cut=1e-17, bias_p=0.1, renormalise=True, mode="fast",
                    rank_tol=_abi.DEFAULT_RANK_TOL, kept_mask=0, readout=(1e-9, 1e-12)):
        """H2D copy of the symbols, fused decode kernel, D2H of (dist, success, status); synchronous.

        ``mode="mono"`` runs the monomial warp engine and re-runs, on the dense engine, every syndrome it hands back
        with MDOPT_ST_NOT_MONOMIAL (none for a bit-flip-bias decode); both stages are GPU kernels."""
        torch = self.torch
        dp = self.upload(program)
        batch = int(symbols.shape[0])
        if batch > self.max_batch:
            self._alloc(batch)
        self.h_sym[:batch].numpy()[...] = symbols
        mono = mode_id(mode) == schedule.MODE_MONO
        if self.n_logical > 12 and not (mono and program_is_monomial_candidate(program) and not kept_mask):
            raise NotImplementedError(
                "more than 12 logical sites: the Dephasing-DMRG read-out runs on the monomial engine only "
                "(bit-flip bias, no erased logical-index qubits)")
        if mono and (kept_mask or not program_is_monomial_candidate(program)):
            mono, mode = False, MONO_FALLBACK_MODE   # erased logical-index qubits / two-site gates: dense engine
        desc = self._desc(program, cut, bias_p, renormalise, mode, rank_tol=rank_tol, kept_mask=kept_mask,
                          readout=readout)
        with torch.cuda.device(self.device):
            if mono:
                self._ensure_mono()
                rc = self.lib.mdopt_mono_decode_batch_host(
                    C.byref(desc), self.mono_cap, self.n_warps, _vp(dp.ops), _vp(dp.wtab), _vp(dp.wdim),
                    dp.wtab.shape[0], _vp(self.h_sym), batch, _vp(self.h_dist), _vp(self.h_success),
                    _vp(self.h_status), _vp(self.d_sym), _vp(self.d_dist), _vp(self.d_success), _vp(self.d_status),
                    _vp(self.counters), _vp(self.mono_workspace), self.mono_ws_bytes, _stream(torch))
            else:
                self._ensure_dense()
                rc = self.lib.mdopt_decode_batch_host(
                    C.byref(desc), self.cap, self.n_ctas, _vp(dp.ops), _vp(dp.wtab), _vp(dp.wdim), dp.wtab.shape[0],
                    _vp(self.h_sym), batch, _vp(self.h_dist), _vp(self.h_success), _vp(self.h_status),
                    _vp(self.d_sym), _vp(self.d_dist), _vp(self.d_success), _vp(self.d_status), _vp(self.counters),
                    _vp(self.workspace), self.ws_bytes, _stream(torch))
        _lib.check(rc, "mdopt_mono_decode_batch_host" if mono else "mdopt_decode_batch_host")
        dist = self.h_dist[:batch].numpy().copy()
        success = self.h_success[:batch].numpy().copy()
        status = self.h_status[:batch].numpy().copy()
        if mono:
            redo = np.nonzero(status == _abi.ST_NOT_MONOMIAL)[0]
            self.last_mono_fallbacks = int(redo.size)
            if redo.size and self.n_logical > 12:
                raise NotImplementedError("a syndrome left the monomial structure and has more than 12 logical sites")
            if redo.size:
                d2, s2, st2 = self.decode_host(program, np.ascontiguousarray(symbols[redo]), cut=cut, bias_p=bias_p,
                                               renormalise=renormalise, mode=MONO_FALLBACK_MODE, rank_tol=rank_tol,
                                               kept_mask=kept_mask, readout=readout)
                dist[redo], success[redo], status[redo] = d2, s2, st2
        return dist, success, status

    def read_counters(self) -> Dict[str, float]:
        c = self.counters.cpu().numpy()
        keys = ["flops_qr", "flops_jacobi", "flops_apply", "bytes", "steps", "truncating_steps", "jacobi_sweeps",
                "qr_reflectors", "cyc_qr", "cyc_formq", "cyc_jacobi", "cyc_apply", "cyc_io", "cyc_coef", "cyc_total", "cyc_pad"]
        return {k: float(v) for k, v in zip(keys, c)}


_ENGINES: Dict[Tuple, DecodeEngine] = {}


def get_engine(n_sites: int, n_logical: int, chi_max: int) -> DecodeEngine:
    """Cached engine, keyed on the COMPILED capacities (a sweep over chi_max = 20, 30, 40, 64 shares one engine and one
    workspace; ``chi_max`` itself travels in the descriptor of every call)."""
    lib, torch = _lib.require_gpu()
    cap = 0 if n_logical > 12 else _abi.chi_capacity(max(int(chi_max), _abi.readout_min_capacity(n_logical)))
    key = (torch.cuda.current_device(), n_sites, n_logical, cap, _abi.mono_chi_capacity(int(chi_max)))
    if key not in _ENGINES:
        _ENGINES[key] = DecodeEngine(n_sites, n_logical, int(chi_max), max_batch=4096)
    eng = _ENGINES[key]
    eng.chi_max = int(chi_max)
    return eng


def clear_engines() -> None:
    """Drop every cached engine and its device workspaces."""
    for eng in _ENGINES.values():
        eng.release()
    _ENGINES.clear()


# --------------------------------------------------------------------------------------------------
# Single-state programs on caller-owned MPS tensors (mps_mpo_contract, apply_constraints, moves, marginal)
# --------------------------------------------------------------------------------------------------


def pack_mps(tensors: Sequence[np.ndarray], cap: int):
    n = len(tensors)
    stride = cap * 2 * cap
    sites = np.zeros((n, stride))
    bonds = np.ones(n + 1, dtype=np.int32)
    for i, t in enumerate(tensors):
        t = np.asarray(t)
        if t.ndim != 3:
            raise ValueError(f"A valid MPS tensor must have 3 legs while the one at site {i} has {t.ndim}.")
        if np.iscomplexobj(t):
            raise NotImplementedError("the GPU engine is real FP64; complex MPS are not supported")
        if t.shape[1] != 2:
            raise NotImplementedError("the GPU engine supports physical dimension 2 only")
        if t.shape[0] > cap or t.shape[2] > cap:
            raise _lib.MdoptCudaError(f"bond dimension {max(t.shape[0], t.shape[2])} exceeds capacity {cap}")
        if i and t.shape[0] != bonds[i]:
            raise ValueError(f"bond mismatch between sites {i-1} and {i}")
        bonds[i], bonds[i + 1] = t.shape[0], t.shape[2]
        sites[i, : t.size] = np.ascontiguousarray(t, dtype=np.float64).ravel()
    return sites, bonds


def unpack_mps(sites: np.ndarray, bonds: np.ndarray) -> List[np.ndarray]:
    return [sites[i, : bonds[i] * 2 * bonds[i + 1]].reshape(bonds[i], 2, bonds[i + 1]).copy()
            for i in range(sites.shape[0])]


def run_program_on_mps(tensors, centre: int, ops: np.ndarray, wtab: np.ndarray, wdim: np.ndarray, chi_max: int,
                       cut: float, renormalise: bool, mode="svd", n_logical: int = 1, want_dist: bool = False,
                       trace_max: int = 0, cap: Optional[int] = None):
    """Upload an MPS, run a program on it in place with one CTA, download the result.

    Without an explicit ``cap`` the compiled capacity is chosen from the input bonds and ``chi_max`` and raised
    (64 -> 128 -> 256) when an intermediate bond outgrows it, so an untruncated contraction behaves like the
    reference's up to bond dimension 256."""
    if mode == "fast" and cut > 1e-14:
        mode = "svd"   # the QR steps of "fast" never apply the user's cut (rule k = min(chi_max, #(s > cut)))
    max_bond = max(max(t.shape[0], t.shape[2]) for t in tensors)
    if cap is not None:
        caps = [int(cap)]
    else:
        need = _abi.readout_min_capacity(n_logical) if want_dist else 1
        first = _abi.chi_capacity(max(min(int(chi_max), 64), max_bond, need))
        if first == 0:
            raise _lib.MdoptCudaError(f"MPS exceeds the largest compiled capacity ({_abi.MAX_CAPACITY})")
        limit = _abi.chi_capacity(min(int(chi_max), _abi.MAX_CAPACITY)) or _abi.MAX_CAPACITY
        caps = [c for c in _abi.CAPACITIES if first <= c <= max(first, limit)]
    for i, c in enumerate(caps):
        res = _run_program_at_capacity(tensors, centre, ops, wtab, wdim, chi_max, cut, renormalise, mode, n_logical,
                                       want_dist, trace_max, c)
        if res["status"] != _abi.ST_CAPACITY:
            return res
    raise _lib.MdoptCudaError(
        f"a tensor exceeded the engine capacity (bond dimension > {caps[-1]} or MPO bond > 2)")


def _run_program_at_capacity(tensors, centre, ops, wtab, wdim, chi_max, cut, renormalise, mode, n_logical, want_dist,
                             trace_max, cap):
    lib, torch = _lib.require_gpu()
    sites, bonds = pack_mps(tensors, cap)
    n = len(tensors)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_sites = torch.from_numpy(sites).to(dev)
    d_bonds = torch.from_numpy(bonds).to(dev)
    d_centre = torch.tensor([int(centre)], dtype=torch.int32, device=dev)
    d_status = torch.zeros(1, dtype=torch.int32, device=dev)
    d_ops = torch.from_numpy(np.ascontiguousarray(ops, dtype=np.int32)).to(dev)
    d_wtab = torch.from_numpy(wtab).to(dev)
    d_wdim = torch.from_numpy(wdim).to(dev)
    ncls = 1 << n_logical
    d_dist = torch.zeros(ncls, dtype=torch.float64, device=dev) if want_dist else None
    d_succ = torch.zeros(1, dtype=torch.int32, device=dev) if want_dist else None
    stride = 4 + 2 * cap if trace_max else 0
    d_trace = torch.zeros((max(trace_max, 1), max(stride, 1)), dtype=torch.float64, device=dev) if trace_max else None
    ws_bytes = int(lib.mdopt_mps_program_workspace_bytes(cap, 1))
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    desc = _abi.make_desc(n, n_logical, min(int(chi_max), 1 << 30), mode_id(mode), renormalise, len(ops), cut, -1.0,
                          stride, trace_max)
    rc = lib.mdopt_mps_program_device(C.byref(desc), cap, _vp(d_ops), _vp(d_wtab), _vp(d_wdim), wtab.shape[0], 1,
                                      _vp(d_sites), _vp(d_bonds), _vp(d_centre), _vp(d_dist), _vp(d_succ),
                                      _vp(d_status), _vp(d_trace), _vp(d_ws), ws_bytes, _stream(torch))
    _lib.check(rc, "mdopt_mps_program_device")
    torch.cuda.synchronize()
    status = int(d_status.item())
    if status == _abi.ST_CAPACITY:
        return {"status": status}
    out = unpack_mps(d_sites.cpu().numpy(), d_bonds.cpu().numpy())
    res = {"tensors": out, "centre": int(d_centre.item()), "status": status}
    if want_dist:
        res["dist"] = d_dist.cpu().numpy()
        res["success"] = int(d_succ.item())
    if trace_max:
        res["trace"] = d_trace.cpu().numpy()
    return res


def run_constraints(mps, strings, logical_tensors, chi_max, cut, renormalise, strategy, order=None, mode="svd"):
    from mdopt_b200.mps.canonical import CanonicalMPS

    table, ops, spans = schedule.TensorTable(), [], []
    ordered = schedule.order_strings(strings, mps.num_sites, strategy, order)
    schedule.append_strings(ops, table, ordered, logical_tensors, mps.num_sites, renormalise, spans)
    ops.append(schedule.OP_END)
    wtab, wdim = table.arrays()
    centre = mps.orth_centre
    if centre is None:
        centre = mps.resolve_orth_centre()
    res = run_program_on_mps(mps.tensors, centre, np.array(ops, dtype=np.int32), wtab, wdim, chi_max, cut,
                             renormalise, mode=mode)
    return CanonicalMPS(res["tensors"], res["centre"], mps.tolerance, mps.chi_max)
