"""GPU parity tests of the monomial warp engine, through the C ABI (mdopt_mono_decode_batch_host / _device).

The headline config is pinned on 1024 syndromes decoded by the UNMODIFIED reference (tests/golden, written by
oracle/make_golden.py big): with the canonical tie-break installed at the reference's own SVD hook the GPU must
match to 1e-10 on every syndrome, truncating or not; against the stock LAPACK tie-breaks the success flags may
differ only on near-ties whose margin lies inside the distance between the two results.
"""
import numpy as np
import pytest

import mps_oracle as O
import structured_svd as S
from conftest import BLOCK_CASES, block_422_code, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mdopt_b200 import _lib, engine, schedule

    _lib.require_gpu()
    from mdopt_b200.codes import CssCode, surface_code
    from mdopt_b200.decoding import decode_css, decode_css_batch, decode_custom_batch

    return dict(surface_code=surface_code, decode_css=decode_css, decode_css_batch=decode_css_batch,
                decode_custom_batch=decode_custom_batch, CssCode=CssCode, engine=engine, schedule=schedule,
                torch=torch)


def test_headline_config_equals_reference_with_canonical_tie_break(api):
    g = load_golden("d5_chi64_bitflip_1024_canonical")
    assert g["tie_break"] == "canonical" and g["lapack_calls_under_hook"] == 0
    dist, succ, status = api["decode_css_batch"](api["surface_code"](5), g["errors"], contraction_strategy="Optimised",
                                                 mode="mono", return_status=True, **g["kwargs"])
    ref = np.array(g["dist"])
    assert (status == 0).all()
    assert api["engine"].get_engine(84, 2, 64).last_mono_fallbacks == 0
    assert (succ == np.array(g["success"])).all()
    rel = np.abs(dist - ref).max(axis=1) / ref.max(axis=1)
    print("d5/chi64 x1024 vs reference+canonical tie-break: max rel", rel.max())
    assert rel.max() < 1e-10   # north-star tolerance, on truncating syndromes too


def test_headline_config_flags_against_stock_reference(api):
    g = load_golden("d5_chi64_bitflip_1024")
    assert g["tie_break"] == "lapack"
    dist, succ = api["decode_css_batch"](api["surface_code"](5), g["errors"], contraction_strategy="Optimised",
                                         **g["kwargs"])   # mode="auto" -> mono
    ref, ok = np.array(g["dist"]), np.array(g["success"])
    spread = np.abs(dist - ref).max(axis=1)
    mism = np.nonzero(succ != ok)[0]
    report = [(g["errors"][b], float(abs(ref[b][0] - ref[b].max())), float(spread[b])) for b in mism]
    print(f"success-flag agreement with the stock reference: {len(ok) - len(mism)}/{len(ok)}; mismatches "
          f"(error, reference margin, distance): {report}; max |delta| {spread.max():.3e}")
    assert not [r for r in report if r[1] > r[2]], report
    assert len(mism) <= 5 and spread.max() < 0.15


@pytest.mark.parametrize("name", ["readme_3x3", "d3_chi64_bitflip", "d3_chi64_bitflip_cut0"])
def test_non_truncating_configs_match_reference(api, name):
    g = load_golden(name)
    errors = [r["error"] for r in g["results"]]
    dist, success, status = api["decode_css_batch"](api["surface_code"](g["lattice_size"]), errors,
                                                    contraction_strategy="Optimised", mode="mono",
                                                    return_status=True, **dict(g["kwargs"]))
    assert (status == 0).all()
    for b, r in enumerate(g["results"]):
        assert int(success[b]) == r["success"], r["error"]
        np.testing.assert_allclose(dist[b], r["dist"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("chi", [8, 16, 20, 128, 256, 512])
def test_truncating_and_large_capacities(api, chi):
    """Every compiled capacity of the warp engine (sort width, record size).  Up to chi_max = 128 against the numpy
    port with the canonical hook; above that the port needs minutes per decode, so chi_max = 256 / 512 (column sorts of
    512 / 1024 keys, 16 / 32 per lane) are checked against the host build of the same engine (std::sort), which the
    CPU tier pins against the reference."""
    import emu_harness as H

    L = 3 if chi <= 20 else 5
    n = {3: 13, 5: 41}[L]
    errors = [e for e in O.seeded_errors(n, 0.08, 60, 17, "Bitflip") if e != "I" * n][: (10 if L == 3 else 3)]
    kw = dict(chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    code = api["surface_code"](L)
    dist, succ, status = api["decode_css_batch"](code, errors, contraction_strategy="Optimised", mode="mono",
                                                 return_status=True, **kw)
    assert (status == 0).all()
    ocode = O.surface_code(L)
    prog = api["schedule"].build_decode_program(code, True, False, True, "Optimised", None, "Bitflip", 0.1)
    for b, e in enumerate(errors):
        if chi <= 128 and b < 2:
            with S.installed():
                d, ok = O.decode_css(ocode, e, contraction_strategy="Optimised", **kw)
        else:
            sym = api["schedule"].symbols_from_errors([e], prog.n_logical)
            dd, ss, st, _, _ = H.decode(prog, sym, chi, 1e-17, 0.1, True, api["schedule"].MODE_MONO, mono=True)
            d, ok = dd[0], int(ss[0])
        assert succ[b] == ok
        np.testing.assert_allclose(dist[b], d, rtol=1e-10, atol=1e-14)


def test_mono_and_dense_engine_agree_without_truncation(api):
    errors = O.seeded_errors(13, 0.12, 200, 5, "Bitflip")
    code = api["surface_code"](3)
    kw = dict(chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised")
    d0, s0 = api["decode_css_batch"](code, errors, mode="mono", **kw)
    d1, s1 = api["decode_css_batch"](code, errors, mode="svd-gram", **kw)
    assert (s0 == s1).all()
    np.testing.assert_allclose(d0, d1, rtol=1e-12, atol=1e-15)


def test_xz_errors_erasures_and_custom_codes(api):
    g = load_golden("erasure_3x3")
    code, ocode = api["surface_code"](3), O.surface_code(3)
    errors = [r["error"] for r in g["results"]] + ["IIXIIZIIIIIII", "YIXIIIIIIIIII", "ZIIIIIIIIIIIX"]
    dist, succ, status = api["decode_css_batch"](code, errors, contraction_strategy="Optimised", return_status=True,
                                                 **g["kwargs"])
    assert (status == 0).all()
    for b, e in enumerate(errors):
        with S.installed():
            d, ok = O.decode_css(ocode, e, contraction_strategy="Optimised", **g["kwargs"])
        if api["schedule"].readout_kept_mask(e, 2) == 3:   # others take the dense engine (different tie-break)
            np.testing.assert_allclose(dist[b], d, rtol=1e-10, atol=1e-13)
        assert succ[b] == ok
    g5 = load_golden("five_qubit_bitflip")
    c, kw = g5["code"], dict(g5["kwargs"])
    errs = [r["error"] for r in g5["results"]]
    d, s = api["decode_custom_batch"](c["stabilizers"], c["x_logicals"], c["z_logicals"], errs, mode="mono", **kw)
    for b, r in enumerate(g5["results"]):
        assert int(s[b]) == r["success"]
        np.testing.assert_allclose(d[b], r["dist"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("num_blocks, chi, errors", BLOCK_CASES)
def test_many_logical_sites_readout(api, num_blocks, chi, errors):
    n, xs, zs, xl, zl = block_422_code(num_blocks)
    code, ocode = api["CssCode"](n, xs, zs, xl, zl), O.CssCode(n, xs, zs, xl, zl)
    kw = dict(chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Naive")
    dist, succ, status = api["decode_css_batch"](code, errors, mode="mono", return_status=True, **kw)
    assert (status == 0).all()
    for b, e in enumerate(errors):
        d, ok = O.decode_css(ocode, e, **kw)
        assert succ[b] == ok
        np.testing.assert_allclose(dist[b], d, rtol=1e-10, atol=1e-14)


def test_structure_violation_falls_back_to_the_dense_engine(api):
    """A program with two-site gates is not monomial: mode='mono' on the engine API re-runs it densely, the raw
    C-ABI call reports MDOPT_ST_NOT_MONOMIAL."""
    from mdopt_b200 import _abi

    sch, eng_mod, torch = api["schedule"], api["engine"], api["torch"]
    code = api["surface_code"](3)
    e = "IIXIIIIIIIIII"
    prog = sch.build_decode_program(code, True, False, True, "Naive", None, "Depolarising", 0.1)
    sym = sch.symbols_from_errors([e, e], prog.n_logical)
    eng = eng_mod.get_engine(prog.n_sites, prog.n_logical, 64)
    d_dist, d_succ, d_stat, _ = eng.decode_device(prog, torch.from_numpy(sym).cuda(), cut=1e-17, bias_p=-1.0,
                                                  mode="mono")
    torch.cuda.synchronize()
    assert (d_stat.cpu().numpy() == _abi.ST_NOT_MONOMIAL).all() and torch.isnan(d_dist).all()
    dist, succ, status = eng.decode_host(prog, sym, cut=1e-17, bias_p=-1.0, mode="mono")
    ref, ok = O.decode_css(O.surface_code(3), e, chi_max=64, cut=1e-17, bias_type="Depolarising", bias_prob=0.1)
    assert (status == 0).all() and succ[0] == ok
    np.testing.assert_allclose(dist[0], ref, rtol=1e-9, atol=1e-12)


def test_full_bench_batch_properties(api):
    """Size-independent properties at the bench size: unit norm, X/Z symmetry of bit-flip decodes, duplicates decode
    bit-identically whatever warp takes them, all status codes OK, zero dense fallbacks."""
    errors = O.seeded_errors(41, 0.05, 8192, 123, "Bitflip")
    errors = errors + errors[:512]
    code = api["surface_code"](5)
    dist, succ, status = api["decode_css_batch"](code, errors, chi_max=64, cut=1e-17, bias_type="Bitflip",
                                                 bias_prob=0.1, contraction_strategy="Optimised", return_status=True)
    assert (status == 0).all()
    nt = np.array([e != "I" * 41 for e in errors])
    np.testing.assert_allclose(np.linalg.norm(dist[nt], axis=1), 1.0, rtol=1e-12)
    np.testing.assert_array_equal(dist[nt][:, 0], dist[nt][:, 2])
    np.testing.assert_array_equal(dist[nt][:, 1], dist[nt][:, 3])
    np.testing.assert_array_equal(dist[:512], dist[8192:])
    np.testing.assert_array_equal(succ[:512], succ[8192:])
    assert api["engine"].get_engine(84, 2, 64).last_mono_fallbacks == 0


def test_dephasing_dmrg_readout_for_14_logical_sites(api):
    """decode_css with more than 12 logical sites (decoding.py:1382-1400) against the unmodified reference's
    (bits of engine.mps, overlap) on 7 Steane blocks; the return convention mirrors (engine, overlap)."""
    s7 = load_golden("dephasing_dmrg")["steane7"]
    code = api["CssCode"](49, s7["x_stabs"], s7["z_stabs"], s7["x_logicals"], s7["z_logicals"])
    kw = dict(s7["kwargs"])
    errors = [r["error"] for r in s7["decodes"]] + ["I" * 49]
    bits, overlap, status = api["decode_css_batch"](code, errors, contraction_strategy="Naive", return_status=True, **kw)
    assert bits.shape == (len(errors), 14) and (status == 0).all()
    for b, r in enumerate(s7["decodes"]):
        assert bits[b].astype(int).tolist() == r["bits"] and float(overlap[b]) == r["overlap"], r["error"]
    assert not bits[-1].any() and overlap[-1] == 1
    eng, ov = api["decode_css"](code, errors[1], contraction_strategy="Naive", silent=True, **kw)
    assert eng.bits == s7["decodes"][1]["bits"] and ov == s7["decodes"][1]["overlap"]
    assert eng.mps.bond_dimensions == [1] * 13
    with pytest.raises(NotImplementedError):
        api["decode_css_batch"](code, errors[:1], chi_max=64, bias_type="Depolarising")


def test_hypergraph_product_ldpc_code_chi512(api):
    """BASELINE.json configs[3]-shaped work: HGP of a (3,4)-regular LDPC code (n = 400 qubits, 32 logical sites, 832 MPS
    sites), bit-flip errors, chi_max = 512, Dephasing-DMRG read-out -- against the host build of the same engine."""
    import emu_harness as H
    from mdopt_b200.codes import hypergraph_product, random_regular_code

    classical = random_regular_code(16, 12, 3, 4, 7)
    code = hypergraph_product(classical, classical)
    k = code.num_x_logicals() + code.num_z_logicals()
    assert len(code) == 400 and k == 32
    rng = np.random.default_rng(5)
    errors = ["".join("X" if rng.random() < 0.01 else "I" for _ in range(400)) for _ in range(6)]
    errors = [e for e in errors if "X" in e][:3]
    kw = dict(chi_max=512, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised")
    bits, overlap, status = api["decode_css_batch"](code, errors, return_status=True, **kw)
    assert (status == 0).all() and bits.shape == (len(errors), 32)
    prog = api["schedule"].build_decode_program(code, True, False, True, "Optimised", None, "Bitflip", 0.1, 1)
    for b, e in enumerate(errors):
        sym = api["schedule"].symbols_from_errors([e], k)
        d, s, st, _, _ = H.decode(prog, sym, 512, 1e-17, 0.1, True, api["schedule"].MODE_MONO, mono=True)
        assert st[0] == 0 and d[0].astype(int).tolist() == bits[b].astype(int).tolist() and int(s[0]) == int(overlap[b])


@pytest.mark.parametrize("mode", ["svd-gram", "svd"])
def test_dense_engine_equals_reference_with_canonical_tie_break(api, mode):
    """The dense CTA engine (the monomial engine's fallback) applies the same canonical order: 1 024 syndromes of the
    headline config against the unmodified reference + hook, and against the monomial engine, to 1e-10."""
    g = load_golden("d5_chi64_bitflip_1024_canonical")
    code = api["surface_code"](5)
    dist, succ, status = api["decode_css_batch"](code, g["errors"], contraction_strategy="Optimised", mode=mode,
                                                 return_status=True, **g["kwargs"])
    ref = np.array(g["dist"])
    assert (status == 0).all() and (succ == np.array(g["success"])).all()
    rel = np.abs(dist - ref).max(axis=1) / ref.max(axis=1)
    print(f"dense {mode}: max rel vs reference+canonical {rel.max():.2e}")
    assert rel.max() < 1e-10
    mono, msucc = api["decode_css_batch"](code, g["errors"], contraction_strategy="Optimised", mode="mono", **g["kwargs"])
    assert (msucc == succ).all()
    np.testing.assert_allclose(mono, dist, rtol=1e-10, atol=1e-14)
