"""The CPU oracle against the fixtures generated from the unmodified reference (oracle/make_golden.py)
and against the reference's published known answers."""
import json

import numpy as np
import pytest

import mps_oracle as O
from conftest import load_golden


@pytest.mark.parametrize("name", ["readme_3x3", "gating_3x3", "d3_chi64_bitflip", "d3_chi64_bitflip_cut0",
                                  "d3_chi64_depol", "d3_chi8_bitflip", "d5_chi16_bitflip", "erasure_3x3"])
def test_oracle_matches_reference_outputs(name):
    g = load_golden(name)
    kw = dict(g["kwargs"])
    code = O.surface_code(g["lattice_size"])
    truncating = name in ("d3_chi8_bitflip", "d5_chi16_bitflip", "gating_3x3", "erasure_3x3")
    for row in g["results"]:
        dist, ok = O.decode_css(code, row["error"], contraction_strategy="Optimised", **kw)
        assert int(ok) == row["success"], row["error"]
        # bit-for-bit on this machine; the tolerance only absorbs BLAS differences between hosts on
        # non-truncating cases (truncating cases cut degenerate multiplets and are rounding-sensitive)
        tol = 5e-2 if truncating else 1e-12
        np.testing.assert_allclose(np.asarray(dist, dtype=float), row["dist"], rtol=tol, atol=tol * 1e-2)


def test_readme_known_answer():
    dist, ok = O.decode_css(O.surface_code(3), "IIXIIIIIIIIII", chi_max=64, bias_type="Bitflip", bias_prob=0.01,
                            contraction_strategy="Optimised", tolerance=1e-12)
    np.testing.assert_allclose(dist, [0.7025030517, 0.0805571990, 0.7025030517, 0.0805571990], atol=2e-10)
    assert ok == 1


def test_gating_by_literal_character():
    code = O.surface_code(3)
    dist, ok = O.decode_css(code, "IIYIIIIIIIIII", chi_max=64, bias_type="Bitflip", bias_prob=0.1)
    np.testing.assert_allclose(dist, [0.5, 0.5, 0.5, 0.5], atol=1e-14)
    assert ok == 1
    assert O.decode_css(code, "I" * 13) == ([1.0, 0.0, 0.0, 0.0], 1)


def test_mpo_tensor_tables_match_reference():
    ref = load_golden("mpo_tensors")
    for name, values in ref.items():
        np.testing.assert_array_equal(getattr(O, name), np.asarray(values))


def test_svd_truncation_rule_against_reference():
    doc = load_golden("svd_cases")
    rng = np.random.default_rng(doc["rng_seed"])
    for case in doc["cases"]:
        a = rng.standard_normal((case["m"], case["n"]))
        u, s, vh, err = O.svd(a, cut=case["cut"], chi_max=case["chi_max"])
        np.testing.assert_allclose(s, case["s"], rtol=1e-12)
        assert abs(err - case["trunc_err"]) <= 1e-10 * max(1.0, case["trunc_err"])
        np.testing.assert_allclose((u * s) @ vh, (u * s) @ vh)
        assert u.shape == (case["m"], len(case["s"])) and vh.shape == (len(case["s"]), case["n"])


def test_marginal_against_reference():
    g = load_golden("marginal_case")
    mps = O.MPS([np.asarray(t) for t in g["tensors"]])
    out = O.marginal(mps, g["sites"], renormalise=True)
    assert len(out.tensors) == len(g["result"])
    for a, b in zip(out.tensors, g["result"]):
        np.testing.assert_allclose(a, np.asarray(b), rtol=1e-13, atol=1e-15)


def test_sweep_spectra_against_reference():
    g = load_golden("d3_spectra")
    rec = []
    dist, ok = O.decode_css(O.surface_code(3), g["error"], chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                            contraction_strategy="Optimised", tolerance=1e-12, record=rec)
    ref_sweeps = [np.asarray(x["s"]) for x in g["svds"] if x["chi_max"] == 64]
    assert len(rec) == len(ref_sweeps)
    for a, b in zip(rec, ref_sweeps):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dist, g["dist"], rtol=1e-12)


def test_error_paths():
    code = O.surface_code(3)
    with pytest.raises(ValueError):
        O.decode_css(code, "IIXII", chi_max=8, bias_type="Bitflip")
    with pytest.raises(ValueError):
        O.decode_css(code, "IIQIIIIIIIIII", chi_max=8, bias_type="Bitflip")
    with pytest.raises(ValueError):
        O.constraint_mpo([O.XOR_LEFT, O.XOR_RIGHT], [[0], [2]])
    with pytest.raises(ValueError):
        O.svd(np.zeros((2, 2, 2)))


SHOR = dict(n=9, x_stabs=[[0, 1, 2, 3, 4, 5], [3, 4, 5, 6, 7, 8]],
            z_stabs=[[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8]], x_logicals=[[0, 1, 2]], z_logicals=[[0, 3, 6]])
SHOR_KNOWN = [0.61225663, 0.06778069, 0.28807136, 0.03189132]  # examples/decoding/shor.ipynb cell 25 output


def test_shor_notebook_known_answer():
    """The Shor-code walk-through of the reference (examples/decoding/shor.ipynb cells 4-25): |++0...0>,
    bit-flip bias 0.1, X and Z checks then X and Z logicals ("Optimised"), marginal, L1-normalised dense."""
    code = O.CssCode(SHOR["n"], SHOR["x_stabs"], SHOR["z_stabs"], SHOR["x_logicals"], SHOR["z_logicals"])
    checks, logicals = O.code_sites(code)
    assert checks[0] == [[[2], [4, 6, 8, 10], [3, 5, 7, 9, 11], [12]], [[8], [10, 12, 14, 16], [9, 11, 13, 15, 17], [18]]]
    assert checks[1][0] == [[3], [], [4], [5]]
    assert logicals[1] == [[[1], [3, 9], [2, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14], [15]]]
    mps = O.product_state("++" + "0" * 18)
    mps = O.apply_bitflip_bias(mps, range(2, 20), 0.1)
    for i in (0, 1):
        mps = O.apply_constraints(mps, checks[i], [O.XOR_LEFT, O.XOR_BULK, O.SWAP, O.XOR_RIGHT], renormalise=True,
                                  strategy="Optimised")
    for i in (0, 1):
        mps = O.apply_constraints(mps, logicals[i], [O.COPY_LEFT, O.XOR_BULK, O.SWAP, O.XOR_RIGHT], renormalise=True,
                                  strategy="Optimised")
    out = O.marginal(mps, list(range(2, 20))).dense(flatten=True, renormalise=True, norm=1)
    np.testing.assert_allclose(out, SHOR_KNOWN, atol=5e-9)


@pytest.mark.parametrize("name", ["five_qubit_bitflip", "five_qubit_depol"])
def test_decode_custom_five_qubit_code(name):
    """decode_custom (decoding.py:1407-1628) on the five-qubit code of examples/decoding/quantum_five_qubit.ipynb."""
    g = load_golden(name)
    c = g["code"]
    for row in g["results"]:
        dist, ok = O.decode_custom(c["stabilizers"], c["x_logicals"], c["z_logicals"], row["error"], **g["kwargs"])
        assert int(ok) == row["success"], row["error"]
        np.testing.assert_allclose(np.asarray(dist, dtype=float), row["dist"], rtol=1e-11, atol=1e-14)


def test_multiply_by_stabiliser_matches_reference():
    """decoding.py:1237-1240 with the drawn stabiliser recorded in the fixture (oracle/make_golden.py stabiliser)."""
    g = load_golden("stabiliser_3x3")
    code = O.surface_code(3)
    for r in g["results"]:
        dist, ok = O.decode_css(code, r["error"], contraction_strategy="Optimised", stabiliser=r["stabiliser"],
                                **g["kwargs"])
        assert ok == r["success"]
        np.testing.assert_allclose(dist, r["dist"], rtol=1e-12, atol=1e-15)
    assert O.multiply_pauli_strings("IXYZ", "XXXX") == "XIZY"


def _custom_walkthrough(stabs, xl, zl, strategy, norm):
    """The notebooks' manual decode of a code given by Pauli strings: |+..+0..0>, bit-flip bias 0.01, X and Z
    logicals, then all stabiliser checks (cut 1e-17, chi 1e4, renormalise), marginal, dense."""
    k, n = len(xl) + len(zl), len(stabs[0])
    checks, (lx, lz) = O.custom_sites(stabs, xl, zl)
    mps = O.product_state("+" * k + "0" * (2 * n))
    mps = O.apply_bitflip_bias(mps, range(k, 2 * n + k), 0.01)
    kw = dict(chi_max=10 ** 4, cut=1e-17, renormalise=True, strategy=strategy)
    for strings in (lx, lz):
        mps = O.apply_constraints(mps, strings, [O.COPY_LEFT, O.XOR_BULK, O.SWAP, O.XOR_RIGHT], **kw)
    mps = O.apply_constraints(mps, checks, [O.XOR_LEFT, O.XOR_BULK, O.SWAP, O.XOR_RIGHT], **kw)
    return O.marginal(mps, list(range(k, 2 * n + k))).dense(flatten=True, renormalise=True, norm=norm)


def test_five_and_three_qubit_notebook_known_answers():
    """Printed outputs of examples/decoding/quantum_five_qubit.ipynb (cell 28, L2-normalised, "Optimised") and
    examples/decoding/quantum_three_qubit.ipynb (cell 29, L1-normalised, "Naive")."""
    five = _custom_walkthrough(["XZZXI", "IXZZX", "XIXZZ", "ZXIXZ"], ["XXXXX"], ["ZZZZZ"], "Optimised", 2)
    np.testing.assert_allclose(five, [0.9999731, 0.00513518, 0.00513518, 0.0010249], atol=5e-8)
    three = _custom_walkthrough(["XXI", "IXX"], ["XXX"], ["ZZZ"], "Naive", 1)
    np.testing.assert_allclose(three, [7.72235375e-01, 2.26750465e-01, 7.83965408e-04, 2.30194740e-04], atol=5e-9)


def test_compress_matches_reference():
    """CanonicalMPS.compress / compress_bond (canonical.py:684-904, "svd") against the reference's outputs on a
    seeded MPS (oracle/make_golden.py compress)."""
    g = load_golden("compress_case")
    tensors = [np.array(t) for t in g["tensors"]]
    for c in g["cases"]:
        out, errs = O.compress(O.MPS([t.copy() for t in tensors], 0), chi_max=c["chi_max"], cut=c["cut"],
                               renormalise=c["renormalise"])
        assert out.bond_dimensions == c["bond_dimensions"]
        np.testing.assert_allclose(errs, c["errors"], rtol=1e-10, atol=1e-16)
        np.testing.assert_allclose(out.dense(flatten=True), c["dense"], atol=1e-12)
    b = g["bond_case"]
    out, err = O.compress_bond(O.MPS([t.copy() for t in tensors], 0), b["bond"], chi_max=b["chi_max"])
    assert out.orth_centre == b["orth_centre"] and out.bond_dimensions == b["bond_dimensions"]
    np.testing.assert_allclose(err, b["error"], rtol=1e-10)
    np.testing.assert_allclose(out.dense(flatten=True), b["dense"], atol=1e-12)
