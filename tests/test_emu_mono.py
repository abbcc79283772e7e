"""CPU-tier checks of the MONOMIAL warp engine (mdopt_b200/csrc/mono_core.cuh compiled for the host, lane loops run
sequentially) and of the canonical tie-break that makes truncating decodes well-posed.

Fixtures (tests/golden, written by oracle/make_golden.py from the UNMODIFIED reference):
  d5_chi64_bitflip_1024            the first 1024 errors of the bench batch, stock LAPACK tie-breaks
  d5_chi64_bitflip_1024_canonical  the same with oracle/structured_svd.py installed at the reference's own SVD hook
                                   (xp.linalg.svd, mdopt/utils/utils.py:80): the rule the engine implements
"""
import numpy as np
import pytest

import emu_harness as H
import mps_oracle as O
import structured_svd as S
from conftest import BLOCK_CASES, block_422_code, load_golden
from mdopt_b200 import _abi, schedule
from mdopt_b200.codes import surface_code

MONO = schedule.MODE_MONO


def mono_decode(code, errors, chi, cut=1e-17, bias=0.1, strategy="Optimised", renormalise=True, trace_max=0, cap=None):
    """Group by literal-character flags like decode_css_batch, run the host build of the warp engine."""
    k = code.num_x_logicals() + code.num_z_logicals()
    dist = np.zeros((len(errors), 1 << k))
    succ = np.zeros(len(errors), dtype=int)
    stat = np.zeros(len(errors), dtype=int)
    traces = [None] * len(errors)
    groups = {}
    for b, e in enumerate(errors):
        triv, hx, hz = schedule.error_flags(e)
        if triv:
            dist[b, 0], succ[b] = 1.0, 1
        else:
            groups.setdefault((hx, hz), []).append(b)
    for (hx, hz), idx in groups.items():
        prog = schedule.build_decode_program(code, hx, hz, renormalise, strategy, None, "Bitflip", bias)
        sym = schedule.symbols_from_errors([errors[b] for b in idx], prog.n_logical)
        d, s, st, tr, _ = H.decode(prog, sym, chi, cut, bias, renormalise, MONO, trace_max=trace_max, mono=True, cap=cap)
        dist[idx], succ[idx], stat[idx] = d, s, st
        for j, b in enumerate(idx):
            traces[b] = tr[j]
    return dist, succ, stat, traces


@pytest.mark.parametrize("name", ["readme_3x3", "d3_chi64_bitflip", "d3_chi64_bitflip_cut0", "gating_3x3"])
def test_non_truncating_configs_match_reference(name):
    g = load_golden(name)
    kw = g["kwargs"]
    rows = [r for r in g["results"] if not ("X" in r["error"] and "Z" in r["error"] and name == "gating_3x3")]
    dist, succ, stat, _ = mono_decode(surface_code(g["lattice_size"]), [r["error"] for r in rows], kw["chi_max"],
                                      kw.get("cut", 1e-17), kw["bias_prob"])
    assert (stat == 0).all()
    for b, r in enumerate(rows):
        assert succ[b] == r["success"], r["error"]
        np.testing.assert_allclose(dist[b], r["dist"], rtol=1e-10, atol=1e-14)  # north-star tolerance


def test_headline_config_equals_reference_with_canonical_tie_break():
    """d=5, chi_max=64, 1024 errors of the bench batch (907 non-trivial, ~43 truncating splits each): the engine equals
    the unmodified reference run with the canonical tie-break at its SVD hook to 1e-10 -- flags identical."""
    g = load_golden("d5_chi64_bitflip_1024_canonical")
    assert g["tie_break"] == "canonical" and g["lapack_calls_under_hook"] == 0
    kw = g["kwargs"]
    dist, succ, stat, _ = mono_decode(surface_code(5), g["errors"], kw["chi_max"], kw["cut"], kw["bias_prob"])
    ref = np.array(g["dist"])
    assert (stat == 0).all()
    assert (succ == np.array(g["success"])).all()
    rel = np.abs(dist - ref).max(axis=1) / ref.max(axis=1)
    assert rel.max() < 1e-10, rel.max()


def test_headline_config_flags_against_stock_reference():
    """Against the reference with LAPACK's own tie-breaks the marginals move inside the degenerate-truncation spread
    (the reference's gesdd and gesvd differ by 8.5e-2); the success flags must agree wherever the
    margin |v0 - max v| exceeds that spread."""
    g = load_golden("d5_chi64_bitflip_1024")
    assert g["tie_break"] == "lapack"
    kw = g["kwargs"]
    dist, succ, stat, _ = mono_decode(surface_code(5), g["errors"], kw["chi_max"], kw["cut"], kw["bias_prob"])
    ref, ok = np.array(g["dist"]), np.array(g["success"])
    spread = np.abs(dist - ref).max(axis=1)
    mism = np.nonzero(succ != ok)[0]
    report = [(g["errors"][b], float(abs(ref[b][0] - ref[b].max())), float(spread[b])) for b in mism]
    # measured here: 3 of 1024 flags differ (syndromes 219, 268, 950), each a near-tie whose margin |v0 - max v| in the
    # reference (0, 4.0e-2, 1.5e-2) is inside the distance between the two results for that very syndrome.  A mismatch
    # with a margin outside its spread would be a decoding difference, not a tie-break difference: that fails.
    bad = [r for r in report if r[1] > r[2]]
    assert not bad, bad
    assert len(mism) <= 5, report
    assert spread.max() < 0.15


def test_canonical_hook_on_the_oracle_port():
    """The numpy port with the hook reproduces the hooked reference (so the port + hook can stand in for it on the
    GPU box, where /root/reference does not exist)."""
    g = load_golden("d5_chi64_bitflip_1024_canonical")
    kw = dict(g["kwargs"])
    picks = [b for b, e in enumerate(g["errors"]) if e.count("X") >= 2][:2]
    code = O.surface_code(5)
    for b in picks:
        with S.installed() as stats:
            stats.update(structured=0, lapack=0)
            d, ok = O.decode_css(code, g["errors"][b], contraction_strategy="Optimised", **kw)
        assert stats["lapack"] == 0
        np.testing.assert_allclose(d, g["dist"][b], rtol=1e-10, atol=1e-14)
        assert ok == g["success"][b]


def test_truncating_d3_chi8_equals_hooked_oracle():
    g = load_golden("d3_chi8_bitflip")
    errors = [r["error"] for r in g["results"]]
    dist, succ, stat, _ = mono_decode(surface_code(3), errors, 8)
    code = O.surface_code(3)
    for b, e in enumerate(errors):
        if e == "I" * 13:
            continue
        with S.installed():
            d, ok = O.decode_css(code, e, chi_max=8, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                 contraction_strategy="Optimised")
        assert stat[b] == 0 and succ[b] == ok
        np.testing.assert_allclose(dist[b], d, rtol=1e-10, atol=1e-14)


def test_sweep_spectra_match_the_hooked_oracle_split_by_split():
    err = "IXIIIIIXIIIII"
    dist, succ, stat, traces = mono_decode(surface_code(3), [err], 8, trace_max=400)
    rec = []
    with S.installed():
        O.decode_css(O.surface_code(3), err, chi_max=8, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                     contraction_strategy="Optimised", record=rec)
    sweeps = [row for row in traces[0] if row[0] > 0 and row[3] == 0]
    assert len(sweeps) == len(rec)
    for row, ref in zip(sweeps, rec):
        kept = int(row[2])
        assert kept == len(ref)
        np.testing.assert_allclose(np.sort(row[4:4 + kept])[::-1], np.sort(ref)[::-1], rtol=1e-10, atol=1e-300)


def test_mono_equals_dense_engine_without_truncation():
    code = surface_code(3)
    errors = O.seeded_errors(13, 0.12, 40, 5, "Bitflip")
    dist, succ, stat, _ = mono_decode(code, errors, 64)
    for b, e in enumerate(errors):
        triv, hx, hz = schedule.error_flags(e)
        if triv:
            continue
        prog = schedule.build_decode_program(code, hx, hz, True, "Optimised", None, "Bitflip", 0.1)
        sym = schedule.symbols_from_errors([e], prog.n_logical)
        d, s, st, _, _ = H.decode(prog, sym, 64, 1e-17, 0.1, True, schedule.MODE_SVD_GRAM)
        np.testing.assert_allclose(dist[b], d[0], rtol=1e-12, atol=1e-15)
        assert succ[b] == s[0] and stat[b] == st[0] == 0


@pytest.mark.parametrize("cap", [16, 128, 512])
def test_capacity_independent(cap):
    """The same chi_max on a larger compiled capacity (other sort width / record size) gives the same numbers."""
    code = surface_code(3)
    errors = ["IXIIIIIXIIIII", "XIIIIIIIIIIIX", "IIXIIZIIIIIII"]
    base, s0, st0, _ = mono_decode(code, errors, 8, cap=8)
    big, s1, st1, _ = mono_decode(code, errors, 8, cap=cap)
    assert (st0 == 0).all() and (st1 == 0).all() and (s0 == s1).all()
    np.testing.assert_allclose(big, base, rtol=1e-13, atol=1e-16)


def test_naive_strategy_no_renormalise_and_xz_errors():
    code, ocode = surface_code(3), O.surface_code(3)
    for e, strat, ren in [("IIXIIZIIIIIII", "Naive", True), ("ZIXIIIIIIIIIZ", "Optimised", False),
                          ("YIXIIIIIIIIII", "Naive", False)]:
        dist, succ, stat, _ = mono_decode(code, [e], 1 << 14 if False else 64, strategy=strat, renormalise=ren)
        with S.installed():
            d, ok = O.decode_css(ocode, e, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                 renormalise=ren, contraction_strategy=strat)
        assert stat[0] == 0 and succ[0] == ok
        np.testing.assert_allclose(dist[0], d, rtol=1e-10, atol=1e-14)


def test_erasure_errors_with_the_default_read_out():
    g = load_golden("erasure_3x3")
    code, ocode = surface_code(3), O.surface_code(3)
    for r in g["results"]:
        e = r["error"]
        if schedule.readout_kept_mask(e, 2) != 3:
            continue   # erasures of qubits 0 / 1 change the read-out sites: dense engine (checked in test_emu_engine)
        dist, succ, stat, _ = mono_decode(code, [e], 64)
        assert stat[0] == 0 and succ[0] == r["success"]
        # X+Z errors truncate inside degenerate multiplets at chi_max = 64: exact against the reference only with
        # the same tie-break
        with S.installed():
            d, ok = O.decode_css(ocode, e, contraction_strategy="Optimised", **g["kwargs"])
        np.testing.assert_allclose(dist[0], d, rtol=1e-10, atol=1e-13)
        if not O.decode_truncates(ocode, e, contraction_strategy="Optimised", **g["kwargs"]):
            np.testing.assert_allclose(dist[0], r["dist"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("name", ["five_qubit_bitflip"])
def test_decode_custom_program(name):
    g = load_golden(name)
    c, kw = g["code"], g["kwargs"]
    prog = schedule.build_custom_program(c["stabilizers"], c["x_logicals"], c["z_logicals"], True,
                                         kw["contraction_strategy"], None, "Bitflip", kw["bias_prob"])
    for r in g["results"]:
        e = r["error"]
        if e == "I" * 5 or "E" in e:
            continue
        sym = schedule.symbols_from_errors([e], prog.n_logical)
        d, s, st, _, _ = H.decode(prog, sym, 64, kw["cut"], kw["bias_prob"], True, MONO, mono=True,
                                  readout=(1e-12, 0.0))
        assert st[0] == 0 and s[0] == r["success"], e
        np.testing.assert_allclose(d[0], r["dist"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("num_blocks, chi, errors", BLOCK_CASES)
def test_many_logical_sites_readout(num_blocks, chi, errors):
    """8 and 12 logical sites (256 / 4096 classes): the path enumeration of the monomial read-out."""
    n, xs, zs, xl, zl = block_422_code(num_blocks)
    from mdopt_b200.codes import CssCode
    code = CssCode(n, xs, zs, xl, zl)
    ocode = O.CssCode(n, xs, zs, xl, zl)
    for e in errors:
        dist, succ, stat, _ = mono_decode(code, [e], chi, strategy="Naive")
        d, ok = O.decode_css(ocode, e, chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                             contraction_strategy="Naive")
        assert stat[0] == 0 and succ[0] == ok
        np.testing.assert_allclose(dist[0], d, rtol=1e-10, atol=1e-14)


def test_structure_violations_are_reported_not_decoded():
    """Programs / read-outs outside the monomial structure end with ST_NOT_MONOMIAL and NaN output."""
    code = surface_code(3)
    e = "IIXIIIIIIIIII"
    prog = schedule.build_decode_program(code, True, False, True, "Naive", None, "Depolarising", 0.1)  # two-site gates
    sym = schedule.symbols_from_errors([e], prog.n_logical)
    d, s, st, _, _ = H.decode(prog, sym, 64, 1e-17, -1.0, True, MONO, mono=True)
    assert st[0] == _abi.ST_NOT_MONOMIAL and np.isnan(d[0]).all() and s[0] == 0
    # a Hadamard-like MPO tensor needs two targets per row
    had = np.zeros((1, 1, 2, 2)); had[0, 0] = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    table = schedule.TensorTable()
    from mdopt_b200.optimiser.utils import XOR_LEFT, XOR_RIGHT
    xor = [table.add(XOR_LEFT), table.add(XOR_RIGHT)]   # entangle sites 2, 3 first: on a product state every row has
    ids = [table.add(had), table.add(had)]              # a single target whatever the tensor
    ops = np.array([schedule.OP_SWEEP, 2, 2, 1, 0] + xor + [schedule.OP_SWEEP, 2, 2, 1, 0] + ids +
                   [schedule.OP_MARGINAL, schedule.OP_END], dtype=np.int32)
    wtab, wdim = table.arrays()
    prog2 = schedule.Program(ops, wtab, wdim, prog.n_sites, prog.n_logical)
    d, s, st, _, _ = H.decode(prog2, sym, 64, 1e-17, 0.1, True, MONO, mono=True)
    assert st[0] == _abi.ST_NOT_MONOMIAL


def test_zero_norm_state_gives_nan_like_the_reference():
    """bias 0 and an error that violates a check: the state loses all weight; NaN + ST_ZERO_NORM, as the reference's
    unguarded renormalisation (canonical.py:987-989)."""
    code = surface_code(3)
    prog = schedule.build_decode_program(code, True, False, True, "Optimised", None, "Bitflip", 0.0)
    sym = schedule.symbols_from_errors(["IIXIIIIIIIIII"], prog.n_logical)
    d, s, st, _, _ = H.decode(prog, sym, 64, 1e-17, 0.0, True, MONO, mono=True)
    assert st[0] == _abi.ST_ZERO_NORM and np.isnan(d[0]).all() and s[0] == 0


def test_structured_svd_hook_unit():
    """The hook is an SVD (reconstruction, orthonormal kept vectors) with the documented order; LAPACK otherwise."""
    rng = np.random.default_rng(3)
    t = np.linalg.qr(rng.standard_normal((12, 6)))[0].T           # 6 orthonormal rows of length 12
    q = np.zeros((8, 6))
    for r, (c, v) in enumerate([(0, 1.0), (1, 2.0), (0, 1.0), (2, 2.0), (3, 0.5), (1, 0.0), (4, 2.0), (5, 1e-30)]):
        q[r, c] = v
    theta = q @ t
    u, s, vh = S.structured_svd(theta)
    np.testing.assert_allclose((u * s) @ vh, theta, atol=1e-14)
    k = int(np.count_nonzero(s))
    np.testing.assert_allclose(u[:, :k].T @ u[:, :k], np.eye(k), atol=1e-14)
    # sigma: columns 1, 2, 4 tie at 2.0 -> ordered by first supporting row (1, 3, 6); then sqrt(2), 0.5; 1e-30 dropped
    np.testing.assert_allclose(s[:5], [2.0, 2.0, 2.0, np.sqrt(2.0), 0.5], rtol=1e-14)
    assert [int(np.nonzero(u[:, j])[0][0]) for j in range(5)] == [1, 3, 6, 0, 4]
    assert k == 5
    dense = rng.standard_normal((6, 9))
    u2, s2, v2 = S.structured_svd(dense)
    np.testing.assert_allclose(s2, np.linalg.svd(dense, compute_uv=False), rtol=1e-13)


# ---- more than 12 logical sites: the Dephasing-DMRG read-out (decoding.py:1382-1400) -------------------------------


def test_dephasing_dmrg_restatement_equals_the_reference():
    """The numpy restatement (oracle) against the unmodified reference's DephasingDMRG on seeded positive targets with
    a unique main component (the reference's own test problem, tests/optimiser/test_dephasing_dmrg.py:199-300)."""
    g = load_golden("dephasing_dmrg")
    for c in g["cases"]:
        t = [np.array(x) for x in c["tensors"]]
        assert O.dephasing_dmrg(t, 1) == c["bits_1"]
        assert O.dephasing_dmrg(t, 2) == c["bits_2"]


def _steane7():
    from mdopt_b200.codes import CssCode

    s7 = load_golden("dephasing_dmrg")["steane7"]
    args = (49, s7["x_stabs"], s7["z_stabs"], s7["x_logicals"], s7["z_logicals"])
    return s7, CssCode(*args), O.CssCode(*args)


def test_decode_with_14_logical_sites_equals_the_reference():
    """decode_css on 7 Steane blocks (14 logical sites) by the unmodified reference (canonical SVD tie-break installed):
    the in-kernel read-out ends on the same bitstring, overlap included; so does the oracle port."""
    s7, code, ocode = _steane7()
    kw = s7["kwargs"]
    for row in s7["decodes"]:
        e = row["error"]
        prog = schedule.build_decode_program(code, "X" in e, "Z" in e, True, "Naive", None, "Bitflip", kw["bias_prob"], 1)
        assert prog.ops[-3] == schedule.OP_MARGINAL_DMRG
        sym = schedule.symbols_from_errors([e], prog.n_logical)
        d, s, st, _, _ = H.decode(prog, sym, kw["chi_max"], kw["cut"], kw["bias_prob"], True, MONO, mono=True)
        assert st[0] == 0 and d.shape == (1, 14)
        assert d[0].astype(int).tolist() == row["bits"] and float(s[0]) == row["overlap"], e
    e = s7["decodes"][2]["error"]
    with S.installed():
        bits, ov = O.decode_css(ocode, e, contraction_strategy="Naive", **kw)
    assert bits == s7["decodes"][2]["bits"] and ov == s7["decodes"][2]["overlap"]


def test_dmrg_sweeps_parameter_and_x_only_errors():
    """num_runs sweeps; an X-only error leaves the Z-logical sites unconstrained: their weights tie exactly and the
    first maximal configuration (bit 0) is kept -- the rule of the snap's argmax (dephasing_dmrg.py:245)."""
    s7, code, ocode = _steane7()
    e = "X" + "I" * 48
    for runs in (1, 2):
        prog = schedule.build_decode_program(code, True, False, True, "Naive", None, "Bitflip", 0.1, runs)
        sym = schedule.symbols_from_errors([e], prog.n_logical)
        d, s, st, _, _ = H.decode(prog, sym, 64, 1e-17, 0.1, True, MONO, mono=True)
        bits, ov = O.decode_css(ocode, e, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                contraction_strategy="Naive", num_runs=runs)
        assert st[0] == 0 and d[0].astype(int).tolist() == bits and float(s[0]) == ov == 1.0
