"""CPU-tier checks of the engine LOGIC: the CUDA headers compiled for the host (tests/emu, -DMDOPT_EMU)
against the oracle / reference fixtures.  This validates pivoting, truncation rules, index maps, frames and
the program interpreter; the GPU tier (test_gpu_*.py) validates the kernels themselves."""
import numpy as np
import pytest

import emu_harness as H
import mps_oracle as O
from conftest import BLOCK_CASES, block_422_code, load_golden
from mdopt_b200 import schedule
from mdopt_b200.codes import surface_code


def emu_decode(code, errors, chi, cut, bias, mode, strategy="Optimised", renormalise=True, trace_max=0,
               bias_type="Bitflip"):
    out = []
    for e in errors:
        triv, hx, hz = schedule.error_flags(e)
        if triv:
            out.append((np.array([1.0, 0, 0, 0]), 1, 0, None, None))
            continue
        prog = schedule.build_decode_program(code, hx, hz, renormalise, strategy, None, bias_type, bias)
        sym = schedule.symbols_from_errors([e], prog.n_logical)
        mask = schedule.readout_kept_mask(e, prog.n_logical)
        dist, succ, st, tr, cnt = H.decode(prog, sym, chi, cut, bias if bias_type == "Bitflip" else -1.0,
                                           renormalise, mode, trace_max=trace_max,
                                           kept_mask=0 if mask == (1 << prog.n_logical) - 1 else mask)
        out.append((dist[0], int(succ[0]), int(st[0]), tr[0], cnt))
    return out


@pytest.mark.parametrize("mode", [schedule.MODE_FAST, schedule.MODE_SVD])
@pytest.mark.parametrize("name", ["readme_3x3", "d3_chi64_bitflip", "d3_chi64_bitflip_cut0"])
def test_non_truncating_configs_match_reference(name, mode):
    g = load_golden(name)
    kw = g["kwargs"]
    rows = g["results"][:40]
    res = emu_decode(surface_code(g["lattice_size"]), [r["error"] for r in rows], kw["chi_max"],
                     kw.get("cut", 1e-17), kw["bias_prob"], mode)
    for r, (dist, ok, st, _, _) in zip(rows, res):
        assert st == 0
        assert ok == r["success"], r["error"]
        np.testing.assert_allclose(dist, r["dist"], rtol=1e-10, atol=1e-14)  # north-star tolerance


def test_gating_cases():
    g = load_golden("gating_3x3")
    res = emu_decode(surface_code(3), [r["error"] for r in g["results"]], 64, 1e-17, 0.1, schedule.MODE_FAST)
    for r, (dist, ok, st, _, _) in zip(g["results"], res):
        assert ok == r["success"]
        both = "X" in r["error"] and "Z" in r["error"]  # X+Z at chi=64 truncates inside degenerate multiplets
        np.testing.assert_allclose(dist, r["dist"], rtol=5e-2 if both else 1e-10, atol=1e-13)


def test_truncating_config_success_flags_and_first_truncated_spectrum():
    g = load_golden("d3_chi8_bitflip")
    rows = g["results"]
    res = emu_decode(surface_code(3), [r["error"] for r in rows], 8, 1e-17, 0.1, schedule.MODE_SVD, trace_max=400)
    agree = sum(int(ok == r["success"]) for r, (_, ok, _, _, _) in zip(rows, res))
    assert agree == len(rows)
    # singular values of every sweep split up to and including the first truncating one agree with the oracle
    ocode = O.surface_code(3)
    checked = 0
    for r, (_, _, _, tr, _) in zip(rows, res):
        if r["error"] == "I" * 13:
            continue
        rec = []
        O.decode_css(ocode, r["error"], chi_max=8, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                     contraction_strategy="Optimised", record=rec)
        sweeps = [row for row in tr if row[0] > 0 and row[3] == 0]
        assert len(sweeps) == len(rec)
        for row, ref in zip(sweeps, rec):
            kept = int(row[2])
            top = ref[0]
            big = ref > 1e-12 * top
            assert kept >= int(big.sum()) or kept == 8
            np.testing.assert_allclose(row[4:4 + int(big.sum())][: kept], ref[big][: kept], rtol=1e-10, atol=1e-13 * top)
            checked += 1
            if len(ref) == 8 and row[1] > 8 and row[4 + 8] > 1e-10 * top:
                break  # past the first genuinely truncating split the states may differ (degenerate cuts)
    assert checked > 50


@pytest.mark.parametrize("shape", [(8, 8), (16, 12), (12, 16), (64, 64), (31, 17)])
def test_jacobi_svd_matches_lapack(shape):
    rng = np.random.default_rng(sum(shape))
    a = rng.standard_normal(shape)
    if shape[1] > shape[0]:
        a = a  # wide matrices are legal (extra columns converge to zero)
    u, s, vh, k, bad = H.svd(a, 10 ** 4, 1e-12)
    ref = np.linalg.svd(a, compute_uv=False)
    assert bad == 0 and k == min(shape)
    np.testing.assert_allclose(s[:k], ref[:k], rtol=1e-10)
    np.testing.assert_allclose((u * s) @ vh, a, atol=1e-12)
    np.testing.assert_allclose(u.T @ u, np.eye(k), atol=1e-12)


def test_svd_truncation_rule_matches_reference_fixture():
    doc = load_golden("svd_cases")
    rng = np.random.default_rng(doc["rng_seed"])
    for case in doc["cases"]:
        a = rng.standard_normal((case["m"], case["n"]))
        if max(a.shape) > 64:
            continue  # emulated 128-column Jacobi is slow on the CPU; covered on the GPU tier
        u, s, vh, k, bad = H.svd(a if a.shape[0] >= a.shape[1] else a.T.copy(), case["chi_max"], case["cut"])
        assert k == len(case["s"])
        np.testing.assert_allclose(s[:k], case["s"], rtol=1e-10)


@pytest.mark.parametrize("m,n,r", [(16, 16, 5), (32, 20, 7), (64, 64, 30), (128, 128, 40), (100, 128, 64)])
def test_rank_revealing_qr(m, n, r):
    rng = np.random.default_rng(m * n + r)
    a = rng.standard_normal((m, r)) @ rng.standard_normal((r, n))
    q, coeff, K, more = H.qr(a, min(64, (max(m, n) + 1) // 2))
    assert more == 0 and K == r
    np.testing.assert_allclose(q @ coeff, a, atol=1e-12 * np.abs(a).max() * max(m, n))
    np.testing.assert_allclose(q.T @ q, np.eye(K), atol=1e-13)


def test_qr_flags_genuine_truncation():
    rng = np.random.default_rng(5)
    a = rng.standard_normal((32, 32))
    _, _, K, more = H.qr(a, 16)
    assert K == 16 and more == 1


def test_fast_and_svd_modes_agree_without_truncation():
    code = surface_code(3)
    errs = O.seeded_errors(13, 0.15, 12, 99, "Bitflip")
    fast = emu_decode(code, errs, 64, 1e-17, 0.1, schedule.MODE_FAST)
    full = emu_decode(code, errs, 64, 1e-17, 0.1, schedule.MODE_SVD)
    for (d0, s0, *_), (d1, s1, *_) in zip(fast, full):
        assert s0 == s1
        np.testing.assert_allclose(d0, d1, rtol=1e-11, atol=1e-14)


def test_naive_strategy_and_no_renormalise():
    code, ocode = surface_code(3), O.surface_code(3)
    for e in ["IIXIIIIIIIIII", "IXIIIIIXIIIII"]:
        for strategy in ("Naive", "Optimised"):
            for ren in (True, False):
                (dist, ok, st, _, _), = emu_decode(code, [e], 64, 1e-17, 0.1, schedule.MODE_FAST, strategy, ren)
                ref, rok = O.decode_css(ocode, e, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                        renormalise=ren, contraction_strategy=strategy)
                assert ok == rok
                np.testing.assert_allclose(dist, ref, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("mode", [schedule.MODE_FAST, schedule.MODE_SVD])
def test_depolarising_bias_matches_reference(mode):
    """decoding.py:193-276 (two-site bias + split on every adjacent pair) through the OP_GATE2 path."""
    g = load_golden("d3_chi64_depol")
    kw = g["kwargs"]
    rows = [r for r in g["results"] if not ("X" in r["error"] and "Z" in r["error"])][:20]  # X+Z truncates at chi=64
    res = emu_decode(surface_code(3), [r["error"] for r in rows], kw["chi_max"], kw["cut"], kw["bias_prob"], mode,
                     bias_type="Depolarising")
    for r, (dist, ok, st, _, _) in zip(rows, res):
        assert st == 0 and ok == r["success"], r["error"]
        np.testing.assert_allclose(dist, r["dist"], rtol=1e-10, atol=1e-14)


def test_pair_rotations_agree_with_cyclic_sweeps():
    """X+Z depolarising syndromes are the rotating splits: after a failed orthogonality probe the flagged columns pair
    up and one round of disjoint rotations settles the split (block_ops.cuh pair_fix, layout 4 of Engine::step).  The
    same decodes with that shortcut switched off run the cyclic Jacobi sweeps; both follow the canonical order, so
    the marginals must agree closely, and the shortcut must actually be the one taken."""
    g = load_golden("d3_chi64_depol")
    kw = g["kwargs"]
    rows = [r for r in g["results"] if "X" in r["error"] and "Z" in r["error"]][:6]
    assert rows
    errors = [r["error"] for r in rows]
    args = (surface_code(3), errors, kw["chi_max"], kw["cut"], kw["bias_prob"], schedule.MODE_SVD)
    H.lib().emu_set_pair_fix_off(1)
    try:
        swept = emu_decode(*args, bias_type="Depolarising")
    finally:
        H.lib().emu_set_pair_fix_off(0)
    H.lib().emu_set_pair_fix_off(2)   # no single-round shortcut: greedy rounds, each followed by the probe
    try:
        greedy = emu_decode(*args, bias_type="Depolarising")
    finally:
        H.lib().emu_set_pair_fix_off(0)
    paired = emu_decode(*args, bias_type="Depolarising")
    for (d0, ok0, st0, _, _), (d2, ok2, st2, _, _) in zip(swept, greedy):
        assert st2 == 0 and ok0 == ok2
        np.testing.assert_allclose(d2, d0, rtol=1e-8, atol=1e-13)
    sweeps_off = sum(r[4][6] for r in swept)     # counters[6] = Jacobi sweeps + pair rounds
    sweeps_on = sum(r[4][6] for r in paired)
    print("sweeps with / without the pair rotations:", sweeps_on, sweeps_off)
    assert sweeps_on < 0.8 * sweeps_off          # a pair round counts 1, the sweeps it replaces at least 2
    for (d0, ok0, st0, _, _), (d1, ok1, st1, _, _) in zip(swept, paired):
        assert st0 == 0 and st1 == 0 and ok0 == ok1
        np.testing.assert_allclose(d1, d0, rtol=1e-8, atol=1e-13)


def test_erasure_errors_match_reference():
    """'E' -> '++' and the reference's erased-qubit marginal (decoding.py:1341-1347), incl. erasures of qubits 0/1."""
    g = load_golden("erasure_3x3")
    kw = g["kwargs"]
    res = emu_decode(surface_code(3), [r["error"] for r in g["results"]], kw["chi_max"], kw["cut"], kw["bias_prob"],
                     schedule.MODE_FAST)
    for r, (dist, ok, st, _, _) in zip(g["results"], res):
        assert st == 0 and ok == r["success"], r["error"]
        both = "X" in r["error"] and "Z" in r["error"]
        np.testing.assert_allclose(dist, r["dist"], rtol=5e-2 if both else 1e-10, atol=1e-13)


@pytest.mark.parametrize("name", ["five_qubit_bitflip", "five_qubit_depol"])
def test_decode_custom_program(name):
    """decode_custom's program (X logicals, Z logicals, all checks, relative success tolerance 1e-12)."""
    g = load_golden(name)
    c, kw = g["code"], g["kwargs"]
    depol = kw["bias_type"] != "Bitflip"
    prog = schedule.build_custom_program(c["stabilizers"], c["x_logicals"], c["z_logicals"], True,
                                         kw["contraction_strategy"], None, kw["bias_type"], kw["bias_prob"])
    assert prog.n_sites == 12 and prog.n_logical == 2 and prog.n_strings == 6
    for r in g["results"]:
        e = r["error"]
        if e == "I" * 5:
            continue
        sym = schedule.symbols_from_errors([e], 2)
        mask = schedule.readout_kept_mask(e, 2)
        dist, succ, st, _, _ = H.decode(prog, sym, 64, kw["cut"], -1.0 if depol else kw["bias_prob"], True,
                                        schedule.MODE_FAST, kept_mask=0 if mask == 3 else mask, readout=(1e-12, 0.0))
        assert st[0] == 0 and int(succ[0]) == r["success"], e
        np.testing.assert_allclose(dist[0], r["dist"], rtol=1e-10, atol=1e-14)


def _random_canonical_mps(rng, n, chi):
    """Right-canonical random MPS with bonds min(chi, 2^i, 2^(n-i)) and centre 0 (oracle object)."""
    dims = [min(chi, 2 ** i, 2 ** (n - i)) for i in range(n + 1)]
    tensors = []
    for i in range(n):
        a = rng.standard_normal((dims[i], 2 * dims[i + 1]))
        q, _ = np.linalg.qr(a.T)
        tensors.append(q.T.reshape(dims[i], 2, dims[i + 1]))
    tensors[0] = tensors[0] / np.linalg.norm(tensors[0])
    return O.MPS(tensors, 0)


@pytest.mark.parametrize("mode", [schedule.MODE_FAST, schedule.MODE_SVD])
def test_large_chi_tier_contract_matches_oracle(mode):
    """Capacity 128 (work matrices outside shared memory): an untruncated XOR string on a bond-80 MPS grows
    bonds past 64; the dense state must equal the oracle's mps_mpo_contract (contractor.py:177-378)."""
    rng = np.random.default_rng(5)
    n = 14
    omps = _random_canonical_mps(rng, n, 80)
    mpo = [O.XOR_LEFT] + [O.XOR_BULK, O.SWAP] * 3 + [O.XOR_BULK] * 4 + [O.XOR_RIGHT]
    start = 1
    ref = O.mps_mpo_contract(omps, mpo, start, chi_max=10 ** 4, cut=1e-14)
    table = schedule.TensorTable()
    ids = [table.add(np.asarray(t)) for t in mpo]
    ops = [schedule.OP_SWEEP, start, len(mpo), 0, 0] + ids + [schedule.OP_END]
    wtab, wdim = table.arrays()
    tensors, centre, status = H.program([t.copy() for t in omps.tensors], 0, ops, wtab, wdim, 10 ** 4, 1e-14,
                                        mode=mode, cap=128)
    assert status == 0 and centre == start + len(mpo) - 1
    assert max(t.shape[2] for t in tensors) > 64
    ours = O.MPS(tensors, centre)
    np.testing.assert_allclose(np.abs(ours.dense()), np.abs(ref.dense()), atol=1e-12)
    # the capacity-64 build reports the overflow instead of truncating silently
    small = _random_canonical_mps(rng, n, 64)
    _, _, status64 = H.program([t.copy() for t in small.tensors], 0, ops, wtab, wdim, 10 ** 4, 1e-14, mode=mode,
                               cap=64)
    assert status64 == 2  # MDOPT_ST_CAPACITY


def test_compress_program_matches_reference():
    """OP_COMPRESS through the host build of the engine against the reference's compress() outputs: bond
    dimensions, truncation error per bond (from the traced spectra) and the dense state."""
    g = load_golden("compress_case")
    tensors = [np.array(t) for t in g["tensors"]]
    n = len(tensors)
    table = schedule.TensorTable()
    table.add(np.asarray(O.IDENTITY))
    wtab, wdim = table.arrays()
    for c in g["cases"]:
        ops = []
        for bond in range(n - 1):
            ops += [schedule.OP_COMPRESS, bond, int(c["renormalise"])]
        ops.append(schedule.OP_END)
        out, centre, status, trace = H.program([t.copy() for t in tensors], 0, ops, wtab, wdim, c["chi_max"], c["cut"],
                                               mode=schedule.MODE_SVD, cap=16, trace_max=4 * n)
        assert status == 0 and centre == n - 2
        assert [t.shape[2] for t in out[:-1]] == c["bond_dimensions"]
        errs = []
        for row in trace:
            r, cc, kept, is_move = (int(v) for v in row[:4])
            if r == 0 and cc == 0:
                break
            if not is_move:
                sv = row[4:4 + min(r, cc)]
                errs.append(float(np.sum(sv[kept:] ** 2)))
        np.testing.assert_allclose(errs, c["errors"], rtol=1e-9, atol=1e-15)
        dense = O.MPS(out, centre).dense(flatten=True)
        sign = np.sign(np.dot(dense, c["dense"]))
        np.testing.assert_allclose(sign * dense, c["dense"], atol=1e-11)


@pytest.mark.parametrize("num_blocks,chi,errors", BLOCK_CASES)
def test_many_logical_sites_readout(num_blocks, chi, errors):
    """8 and 12 logical sites (256 / 4096 classes): the dense read-out meets in the middle; every class
    probability must equal the oracle's (decoding.py:1351-1379) on errors that never truncate."""
    from mdopt_b200.codes import CssCode

    args = block_422_code(num_blocks)
    code, ocode = CssCode(*args), O.CssCode(*args)
    kw = dict(chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised",
              tolerance=1e-12)
    for (dist, ok, st, _, _), e in zip(emu_decode(code, errors, chi, 1e-17, 0.1, schedule.MODE_FAST), errors):
        ref, ref_ok = O.decode_css(ocode, e, **kw)
        assert st == 0 and len(dist) == 1 << (4 * num_blocks)
        assert ok == ref_ok
        np.testing.assert_allclose(dist, np.asarray(ref, float), rtol=1e-10, atol=1e-14)


def test_cluster_default_cut_is_svd_mode_parity():
    """cut = 1e-8 at d=3 through the host build in MODE_SVD (what decode_css_batch selects for cut > 1e-14)."""
    code, ocode = surface_code(3), O.surface_code(3)
    errors = [e for e in O.seeded_errors(13, 0.08, 60, 77, "Bitflip") if "X" in e][:6]
    kw = dict(chi_max=64, cut=1e-8, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised",
              tolerance=1e-12)
    for (dist, ok, st, _, _), e in zip(emu_decode(code, errors, 64, 1e-8, 0.1, schedule.MODE_SVD), errors):
        ref, ref_ok = O.decode_css(ocode, e, **kw)
        assert st == 0 and ok == ref_ok
        np.testing.assert_allclose(dist, np.asarray(ref, float), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("mode", [schedule.MODE_SVD, schedule.MODE_SVD_GRAM])
def test_dense_engine_uses_the_canonical_tie_break(mode):
    """The dense engine orders singular values like the monomial engine and oracle/structured_svd.py (sigma descending,
    ties by first non-zero row): on truncating d=5 / chi_max=64 syndromes it equals the unmodified reference run with
    the canonical tie-break at its SVD hook to 1e-10."""
    g = load_golden("d5_chi64_bitflip_1024_canonical")
    idx = [i for i, e in enumerate(g["errors"]) if e != "I" * 41][:40]
    prog = schedule.build_decode_program(surface_code(5), True, False, True, "Optimised", None, "Bitflip", 0.1)
    sym = schedule.symbols_from_errors([g["errors"][i] for i in idx], 2)
    d, s, st, _, _ = H.decode(prog, sym, 64, 1e-17, 0.1, True, mode)
    ref, ok = np.array(g["dist"])[idx], np.array(g["success"])[idx]
    assert (st == 0).all() and (s == ok).all()
    assert (np.abs(d - ref).max(axis=1) / ref.max(axis=1)).max() < 1e-10


def _orthogonal_columns(rng, m, n, scales):
    q, _ = np.linalg.qr(rng.standard_normal((m, n)))
    return q * np.asarray(scales)[None, :]


def _check_split(a, x, order, sig):
    """X = A J with J orthogonal and mutually orthogonal columns: same Gram operator from the left, orthogonal columns,
    column norms = the singular values of A in the reported order."""
    np.testing.assert_allclose(x @ x.T, a @ a.T, rtol=0, atol=1e-12 * np.linalg.norm(a) ** 2)
    norms = np.linalg.norm(x, axis=0)
    gram = x.T @ x
    big = norms > 1e-12 * norms.max()
    cos = gram[np.ix_(big, big)] / np.outer(norms[big], norms[big])
    assert np.max(np.abs(cos - np.eye(big.sum()))) < 1e-11
    np.testing.assert_allclose(sig[order], norms[order], rtol=1e-12, atol=1e-300)
    sv = np.linalg.svd(a, compute_uv=False)
    got = sig[order][:len(sv)]
    keep = sv > 1e-9 * sv[0]
    np.testing.assert_allclose(got[keep], sv[keep], rtol=1e-10)
    assert np.all(np.diff(got) <= 1e-10 * got[0])


def test_split_routes_of_the_portable_tier():
    """The split of the SVD modes, route by route (block_ops.cuh: ortho_probe, pair_fix, jacobi_columns):
    orthogonal columns pass the probe; disjoint non-orthogonal pairs -- also very unequal ones, whose larger column
    the probe does not flag -- take one round of pair rotations; a cluster of three takes the greedy rounds or the
    sweeps; a random dense matrix takes the cyclic sweeps.  Every route must deliver the singular values."""
    rng = np.random.default_rng(2024)
    m = n = 32
    scales = np.exp(rng.uniform(-6, 0, n))
    base = _orthogonal_columns(rng, m, n, scales)

    route, x, order, sig = H.split(base)
    assert route == 0
    _check_split(base, x, order, sig)

    pairs = base.copy()
    for a, b, w in ((1, 7, 0.4), (3, 20, -1.3), (10, 11, 0.05), (30, 2, 2.0)):
        pairs[:, b] += w * pairs[:, a] * (scales[b] / scales[a])
    route, x, order, sig = H.split(pairs)
    assert route == 1, route
    _check_split(pairs, x, order, sig)

    unequal = _orthogonal_columns(rng, m, n, np.where(np.arange(n) == 5, 1e-7, 1.0))
    unequal[:, 5] += 3e-8 * unequal[:, 9]        # the small column leans on a large one: only the small one is flagged
    route, x, order, sig = H.split(unequal)
    assert route == 1, route
    _check_split(unequal, x, order, sig)

    cluster = base.copy()
    cluster[:, 4] += 0.5 * cluster[:, 8] * (scales[4] / scales[8])
    cluster[:, 12] += 0.7 * cluster[:, 8] * (scales[12] / scales[8])
    route, x, order, sig = H.split(cluster)
    assert route in (2, 3), route
    _check_split(cluster, x, order, sig)

    dense = rng.standard_normal((m, n))
    route, x, order, sig = H.split(dense)
    assert route in (2, 3), route
    _check_split(dense, x, order, sig)


def test_split_pairs_every_column():
    """More than half of the columns flagged -- every column has a partner, as behind a two-site bias gate: still one
    round of pair rotations (the candidate tables live in the flat scratch matrix, so their size is not a limit)."""
    rng = np.random.default_rng(77)
    m = n = 64
    scales = np.exp(rng.uniform(-3, 0, n))
    a = _orthogonal_columns(rng, m, n, scales)
    perm = rng.permutation(n)
    for k in range(0, n, 2):
        p, q = perm[k], perm[k + 1]
        a[:, q] += rng.uniform(0.2, 1.5) * a[:, p] * (scales[q] / scales[p])
    route, x, order, sig = H.split(a)
    assert route == 1, route
    _check_split(a, x, order, sig)
