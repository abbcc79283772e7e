"""GPU parity tests of the standalone batched kernels (K2 svd, K1 site-apply, K3 read-out) and the
single-state API (mps_mpo_contract, apply_constraints, move_orth_centre) against the oracle."""
import numpy as np
import pytest

import mps_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    assert torch.cuda.is_available()
    from mdopt_b200 import _lib

    _lib.require_gpu()
    return torch


def test_svd_against_reference_fixture(torch_mod):
    from mdopt_b200.utils import svd

    doc = load_golden("svd_cases")
    rng = np.random.default_rng(doc["rng_seed"])
    for case in doc["cases"]:
        a = rng.standard_normal((case["m"], case["n"]))
        u, s, vh, err = svd(a, cut=case["cut"], chi_max=case["chi_max"], return_truncation_error=True)
        assert len(s) == len(case["s"])
        np.testing.assert_allclose(s, case["s"], rtol=1e-10)  # north-star tolerance on singular values
        assert abs(err - case["trunc_err"]) <= 1e-9 * max(1.0, case["trunc_err"])
        k = len(s)
        np.testing.assert_allclose(u.T @ u, np.eye(k), atol=1e-12)
        np.testing.assert_allclose(vh @ vh.T, np.eye(k), atol=1e-9)
        full_u, full_s, full_vh = np.linalg.svd(a, full_matrices=False)
        np.testing.assert_allclose((u * s) @ vh, (full_u[:, :k] * full_s[:k]) @ full_vh[:k], atol=1e-11)


def test_svd_batch_and_edge_cases(torch_mod):
    from mdopt_b200.utils import split_two_site_tensor, svd, svd_batch

    rng = np.random.default_rng(1)
    mats = rng.standard_normal((300, 32, 24))
    u, s, vh, kept, trunc, status = svd_batch(mats, cut=1e-12, chi_max=10)
    assert (status == 0).all() and (kept == 10).all()
    for b in (0, 17, 299):
        ref = np.linalg.svd(mats[b], compute_uv=False)
        np.testing.assert_allclose(s[b][:10], ref[:10], rtol=1e-10)
        np.testing.assert_allclose(trunc[b], np.sum(ref[10:] ** 2), rtol=1e-9)
    # renormalise, rank deficiency, cut counting (tests/utils/test_utils.py:19-81 of the reference)
    low = rng.standard_normal((20, 3)) @ rng.standard_normal((3, 16))
    u1, s1, v1, _ = svd(low, cut=1e-10)
    assert len(s1) == 3
    u2, s2, v2, _ = svd(low, cut=1e-10, renormalise=True)
    np.testing.assert_allclose(np.linalg.norm(s2), 1.0, atol=1e-14)
    with pytest.raises(ValueError):
        svd(np.zeros((2, 2, 2)))
    with pytest.raises(RuntimeError):
        svd(np.full((3, 3), np.nan))
    theta = rng.standard_normal((4, 2, 2, 6))
    a, sv, b = split_two_site_tensor(theta, chi_max=5)
    assert a.shape == (4, 2, 5) and b.shape == (5, 2, 6)
    ref_a, ref_s, ref_b, _ = O.split_two_site_tensor(theta, chi_max=5)
    np.testing.assert_allclose(sv, ref_s, rtol=1e-10)
    np.testing.assert_allclose(np.einsum("ijm,m,mkl->ijkl", a, sv, b), np.einsum("ijm,m,mkl->ijkl", ref_a, ref_s, ref_b),
                               atol=1e-11)
    with pytest.raises(ValueError):
        split_two_site_tensor(np.zeros((2, 2, 2)))


def test_site_apply_kernel_matches_einsum(torch_mod):
    import ctypes as C

    from mdopt_b200 import _lib

    torch = torch_mod
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for w, name in [(O.XOR_BULK, "xor"), (O.SWAP, "swap"), (O.XOR_RIGHT, "xor_right"), (O.IDENTITY, "id")]:
        vl, vr = w.shape[0], w.shape[1]
        batch, rows, mr, bn = 37, 24, 12, 9
        m_in = rng.standard_normal((batch, rows, vl, mr))
        a = rng.standard_normal((batch, mr, 2, bn))
        ref = np.einsum("brqm,mpn,qwpd->brdwn", m_in, a[0] * 0 + 1, w) if False else \
            np.stack([np.einsum("rqm,mpn,qwpd->rdwn", m_in[b], a[b], w) for b in range(batch)])
        d_m = torch.from_numpy(m_in).cuda()
        d_a = torch.from_numpy(a).cuda()
        d_w = torch.from_numpy(np.ascontiguousarray(w)).cuda()
        d_out = torch.zeros((batch, rows, 2, vr, bn), dtype=torch.float64, device="cuda")
        rc = lib.mdopt_site_apply_batch_device(16, batch, rows, vl, vr, mr, bn, C.c_void_p(d_m.data_ptr()),
                                               C.c_void_p(d_a.data_ptr()), C.c_void_p(d_w.data_ptr()),
                                               C.c_void_p(d_out.data_ptr()), C.c_void_p(0))
        assert rc == 0, name
        torch.cuda.synchronize()
        np.testing.assert_allclose(d_out.cpu().numpy(), ref, rtol=1e-13, atol=1e-13)


def test_site_apply_tma_dmma_kernel_matches_einsum(torch_mod):
    """Tile-aligned shapes take the TMA-staged FP64-MMA kernel (site_apply_mma.cuh); same contraction, same answer."""
    import ctypes as C

    from mdopt_b200 import _lib

    torch = torch_mod
    lib = _lib.load()
    rng = np.random.default_rng(8)
    for cap, rows, mr, bn in [(64, 128, 64, 64), (64, 64, 32, 32), (32, 32, 16, 32), (64, 16, 4, 64)]:
        for w in (O.XOR_BULK, O.SWAP, O.XOR_RIGHT, O.IDENTITY, O.XOR_LEFT):
            vl, vr = w.shape[0], w.shape[1]
            if rows > 2 * cap or vl * mr > 2 * cap:
                continue
            batch = 300 if rows == 128 else 41
            m_in = rng.standard_normal((batch, rows, vl, mr))
            a = rng.standard_normal((batch, mr, 2, bn))
            ref = np.einsum("brqm,bmpn,qwpd->brdwn", m_in, a, w)
            d_m, d_a = torch.from_numpy(m_in).cuda(), torch.from_numpy(a).cuda()
            d_w = torch.from_numpy(np.ascontiguousarray(w)).cuda()
            d_out = torch.full((batch, rows, 2, vr, bn), float("nan"), dtype=torch.float64, device="cuda")
            rc = lib.mdopt_site_apply_batch_device(cap, batch, rows, vl, vr, mr, bn, C.c_void_p(d_m.data_ptr()),
                                                   C.c_void_p(d_a.data_ptr()), C.c_void_p(d_w.data_ptr()),
                                                   C.c_void_p(d_out.data_ptr()), C.c_void_p(0))
            assert rc == 0
            torch.cuda.synchronize()
            np.testing.assert_allclose(d_out.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)


def test_readout_kernel(torch_mod):
    import ctypes as C

    from mdopt_b200 import _lib

    torch = torch_mod
    lib = _lib.load()
    rng = np.random.default_rng(4)
    v = rng.standard_normal((1000, 4))
    v[0] = [0.5, -0.5, 0.5, 0.5]          # exact tie: identity is in the MAP set
    v[1] = [0.5, 0.5 + 1e-11, 0.1, 0.1]   # inside the tolerance
    v[2] = [0.5, 0.5 + 1e-7, 0.1, 0.1]    # outside
    v[3] = [np.nan, 1, 1, 1]
    d_v = torch.from_numpy(v).cuda()
    d_dist = torch.zeros((1000, 4), dtype=torch.float64, device="cuda")
    d_s = torch.zeros(1000, dtype=torch.int32, device="cuda")
    assert lib.mdopt_readout_batch_device(1000, 4, C.c_void_p(d_v.data_ptr()), 1, C.c_void_p(d_dist.data_ptr()),
                                          C.c_void_p(d_s.data_ptr()), C.c_void_p(0)) == 0
    torch.cuda.synchronize()
    dist, succ = d_dist.cpu().numpy(), d_s.cpu().numpy()
    for b in range(4, 1000):
        ref, ok = O.readout(v[b] / np.linalg.norm(v[b]))
        np.testing.assert_allclose(dist[b], ref, rtol=1e-14)
        assert succ[b] == ok
    assert list(succ[:4]) == [1, 1, 0, 0]


def _random_canonical_mps(rng, n, chi):
    tensors = []
    left = 1
    for i in range(n):
        right = 1 if i == n - 1 else int(min(chi, 2 ** min(i + 1, n - 1 - i)))
        tensors.append(rng.standard_normal((left, 2, right)))
        left = right
    mps = O.MPS(tensors, n - 1)
    # bring into canonical form with the oracle (centre 0), then normalise
    mps.orth_centre = n - 1
    for i in range(n - 1, 0, -1):
        t = mps.tensors[i]
        q, r = np.linalg.qr(t.reshape(t.shape[0], -1).T)
        mps.tensors[i] = q.T.reshape(-1, 2, t.shape[2])
        mps.tensors[i - 1] = np.tensordot(mps.tensors[i - 1], r.T, (2, 0))
    mps.tensors[0] = mps.tensors[0] / np.linalg.norm(mps.tensors[0])
    mps.orth_centre = 0
    return mps


@pytest.mark.parametrize("mode", ["svd", "fast"])
def test_mps_mpo_contract_against_oracle(torch_mod, mode):
    """tests/contractor/test_contractor.py:152-208 of the reference, with the logical-constraint MPOs."""
    from mdopt_b200.contractor import mps_mpo_contract
    from mdopt_b200.mps import CanonicalMPS

    rng = np.random.default_rng(11)
    for trial in range(4):
        n = 8
        omps = _random_canonical_mps(rng, n, 8)
        start = int(rng.integers(0, 3))
        length = int(rng.integers(3, n - start + 1))
        mid = [O.XOR_BULK if rng.random() < 0.5 else O.SWAP for _ in range(length - 2)]
        mpo = [O.XOR_LEFT] + mid + [O.XOR_RIGHT]
        ref = O.mps_mpo_contract(omps, mpo, start, chi_max=10 ** 4, cut=1e-14)
        ours = mps_mpo_contract(CanonicalMPS([t.copy() for t in omps.tensors], 0), mpo, start, chi_max=10 ** 4,
                                cut=1e-14, mode=mode)
        assert ours.orth_centre == start + length - 1
        np.testing.assert_allclose(np.abs(ours.dense()), np.abs(ref.dense()), atol=1e-12)
        # canonical form: isometries on both sides of the centre
        for i, t in enumerate(ours.tensors):
            if i < ours.orth_centre:
                g = np.einsum("ijk,ijl->kl", t, t)
                np.testing.assert_allclose(g, np.eye(g.shape[0]), atol=1e-12)
            elif i > ours.orth_centre:
                g = np.einsum("ijk,ljk->il", t, t)
                np.testing.assert_allclose(g, np.eye(g.shape[0]), atol=1e-12)
    with pytest.raises(ValueError):
        mps_mpo_contract(CanonicalMPS(omps.tensors, 0), [np.zeros((2, 2, 2))] * 2, 0)
    with pytest.raises(ValueError):
        mps_mpo_contract(CanonicalMPS(omps.tensors, 0), [O.XOR_LEFT] + [O.XOR_BULK] * 8 + [O.XOR_RIGHT], 0)


@pytest.mark.parametrize("bond,cap_expected", [(80, 256), (40, 128)])
def test_large_chi_tier_contract(torch_mod, bond, cap_expected):
    """Untruncated XOR string on a 16-site MPS: bonds double past 64, the engine escalates to the 128 / 256
    capacities and the dense state equals the oracle's (contractor.py:177-378) to 1e-12."""
    from mdopt_b200.contractor import mps_mpo_contract
    from mdopt_b200.mps import CanonicalMPS

    rng = np.random.default_rng(21 + bond)
    n = 16
    omps = _random_canonical_mps(rng, n, bond)
    mpo = [O.XOR_LEFT] + [O.XOR_BULK, O.SWAP] * 4 + [O.XOR_BULK] * 4 + [O.XOR_RIGHT]
    for mode in ("fast", "svd"):
        ref = O.mps_mpo_contract(omps, mpo, 1, chi_max=10 ** 4, cut=1e-14)
        ours = mps_mpo_contract(CanonicalMPS([t.copy() for t in omps.tensors], 0), mpo, 1, chi_max=10 ** 4,
                                cut=1e-14, mode=mode)
        top = max(ours.bond_dimensions)
        assert cap_expected // 2 < top <= cap_expected, top
        np.testing.assert_allclose(np.abs(ours.dense()), np.abs(ref.dense()), atol=1e-12)


def test_move_orth_centre_and_truncated_contract(torch_mod):
    from mdopt_b200.contractor import mps_mpo_contract
    from mdopt_b200.mps import CanonicalMPS

    rng = np.random.default_rng(12)
    omps = _random_canonical_mps(rng, 10, 16)
    ours = CanonicalMPS([t.copy() for t in omps.tensors], 0)
    moved = ours.move_orth_centre(6, renormalise=False)
    assert moved.orth_centre == 6
    np.testing.assert_allclose(np.abs(moved.dense()), np.abs(omps.dense()), atol=1e-12)
    back = moved.move_orth_centre(2, renormalise=False)
    assert back.orth_centre == 2
    np.testing.assert_allclose(np.abs(back.dense()), np.abs(omps.dense()), atol=1e-12)
    # truncating sweep: singular values of every split agree with the oracle (generic spectrum, no ties)
    mpo = [O.XOR_LEFT] + [O.XOR_BULK, O.SWAP] * 3 + [O.XOR_RIGHT]
    rec = []
    ref = O.mps_mpo_contract(omps, mpo, 1, chi_max=6, cut=1e-17, record=rec)
    out = mps_mpo_contract(ours, mpo, 1, chi_max=6, cut=1e-17, mode="svd")
    np.testing.assert_allclose(np.abs(out.dense()), np.abs(ref.dense()), atol=1e-10)
    assert out.bond_dimensions == ref.bond_dimensions


def test_apply_constraints_against_dense_projector(torch_mod):
    """tests/optimiser/test_utils.py:397-462 of the reference: |+...+> with one XOR constraint keeps exactly
    the basis states of even parity on the constrained bits."""
    from mdopt_b200.mps.utils import create_simple_product_state
    from mdopt_b200.optimiser.utils import SWAP, XOR_BULK, XOR_LEFT, XOR_RIGHT, apply_constraints

    n = 8
    bits = [1, 4, 6]
    string = [[bits[0]], [bits[1]], [s for s in range(bits[0] + 1, bits[-1]) if s not in bits], [bits[-1]]]
    mps = create_simple_product_state(n, which="+")
    out = apply_constraints(mps, [string], [XOR_LEFT, XOR_BULK, SWAP, XOR_RIGHT], renormalise=True, silent=True)
    dense = out.dense()
    for idx in range(2 ** n):
        word = format(idx, f"0{n}b")
        parity = sum(int(word[b]) for b in bits) % 2
        if parity:
            assert abs(dense[idx]) < 1e-13
        else:
            assert abs(abs(dense[idx]) - 1 / np.sqrt(2 ** (n - 1))) < 1e-12


def test_marginal_tail_on_gpu(torch_mod):
    """CanonicalMPS.marginal for the decoder's case (contiguous tail) against the oracle
    (reference canonical.py:906-991; hand case tests/mps/test_canonical.py:851-865 style)."""
    from mdopt_b200.mps import CanonicalMPS

    rng = np.random.default_rng(21)
    for ren in (True, False):
        omps = _random_canonical_mps(rng, 7, 8)
        ours = CanonicalMPS([t.copy() for t in omps.tensors], 0)
        ref = O.marginal(omps.copy(), list(range(2, 7)), renormalise=ren)
        got = ours.marginal(list(range(2, 7)), renormalise=ren)
        assert got is ours and len(got.tensors) == 2 and got.orth_centre == 1
        np.testing.assert_allclose(got.dense(), ref.dense(), rtol=1e-12, atol=1e-14)
    # any other site set runs the same chain through the library tier (mdopt_b200/generic.py)
    omps = _random_canonical_mps(rng, 6, 8)
    ours = CanonicalMPS([t.copy() for t in omps.tensors], 0)
    ref = O.marginal(omps.copy(), [0, 3], renormalise=False)
    got = ours.marginal([0, 3])
    assert len(got.tensors) == len(ref.tensors) == 4
    np.testing.assert_allclose(got.dense(), ref.dense(), rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        ours.marginal([99])


def test_shor_notebook_known_answer_on_gpu(torch_mod):
    """examples/decoding/shor.ipynb cells 4-25 through the drop-in API: product state, bias, apply_constraints
    (GPU) for checks and logicals, marginal (GPU), L1-normalised dense = the notebook's printed vector."""
    from mdopt_b200.codes import CssCode
    from mdopt_b200.decoding.decoding import bitflip_bias, css_code_constraint_sites, css_code_logicals_sites
    from mdopt_b200.contractor import apply_one_site_operator
    from mdopt_b200.mps.utils import create_custom_product_state
    from mdopt_b200.optimiser.utils import COPY_LEFT, SWAP, XOR_BULK, XOR_LEFT, XOR_RIGHT, apply_constraints

    code = CssCode(9, [[0, 1, 2, 3, 4, 5], [3, 4, 5, 6, 7, 8]], [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8]],
                   [[0, 1, 2]], [[0, 3, 6]])
    mps = create_custom_product_state("++" + "0" * 18)
    for site in range(2, 20):
        mps.tensors[site] = apply_one_site_operator(mps.tensors[site], bitflip_bias(0.1))
    checks, logicals = css_code_constraint_sites(code), css_code_logicals_sites(code)
    for i in (0, 1):
        mps = apply_constraints(mps, checks[i], [XOR_LEFT, XOR_BULK, SWAP, XOR_RIGHT], renormalise=True,
                                strategy="Optimised", silent=True)
    for i in (0, 1):
        mps = apply_constraints(mps, logicals[i], [COPY_LEFT, XOR_BULK, SWAP, XOR_RIGHT], renormalise=True,
                                strategy="Optimised", silent=True)
    out = mps.marginal(list(range(2, 20))).dense(flatten=True, renormalise=True, norm=1)
    np.testing.assert_allclose(np.abs(out), [0.61225663, 0.06778069, 0.28807136, 0.03189132], atol=5e-9)


@pytest.mark.parametrize("stabs,xl,zl,strategy,norm,known", [
    (["XZZXI", "IXZZX", "XIXZZ", "ZXIXZ"], ["XXXXX"], ["ZZZZZ"], "Optimised", 2,
     [0.9999731, 0.00513518, 0.00513518, 0.0010249]),            # quantum_five_qubit.ipynb cell 28
    (["XXI", "IXX"], ["XXX"], ["ZZZ"], "Naive", 1,
     [7.72235375e-01, 2.26750465e-01, 7.83965408e-04, 2.30194740e-04]),  # quantum_three_qubit.ipynb cell 29
])
def test_custom_code_notebook_known_answers_on_gpu(torch_mod, stabs, xl, zl, strategy, norm, known):
    """The notebooks' manual walk-through of a code given by Pauli strings, through the drop-in API on the GPU."""
    from mdopt_b200 import schedule
    from mdopt_b200.contractor import apply_one_site_operator
    from mdopt_b200.decoding.decoding import bitflip_bias
    from mdopt_b200.mps.utils import create_custom_product_state
    from mdopt_b200.optimiser.utils import COPY_LEFT, SWAP, XOR_BULK, XOR_LEFT, XOR_RIGHT, apply_constraints

    k, n = len(xl) + len(zl), len(stabs[0])
    checks, (lx, lz) = schedule.custom_site_lists(stabs, xl, zl)
    mps = create_custom_product_state("+" * k + "0" * (2 * n))
    for site in range(k, 2 * n + k):
        mps.tensors[site] = apply_one_site_operator(mps.tensors[site], bitflip_bias(0.01))
    kw = dict(chi_max=10 ** 4, cut=1e-17, renormalise=True, strategy=strategy, silent=True)
    for strings in (lx, lz):
        mps = apply_constraints(mps, strings, [COPY_LEFT, XOR_BULK, SWAP, XOR_RIGHT], **kw)
    mps = apply_constraints(mps, checks, [XOR_LEFT, XOR_BULK, SWAP, XOR_RIGHT], **kw)
    out = mps.marginal(list(range(k, 2 * n + k))).dense(flatten=True, renormalise=True, norm=norm)
    np.testing.assert_allclose(np.abs(out), known, atol=5e-8)


def test_compress_on_gpu(torch_mod):
    """CanonicalMPS.compress / compress_bond (canonical.py:684-904, "svd") on the GPU against the reference's
    outputs for a seeded MPS: bond dimensions, truncation error per bond, dense state."""
    from mdopt_b200.mps import CanonicalMPS

    g = load_golden("compress_case")
    tensors = [np.array(t) for t in g["tensors"]]
    for c in g["cases"]:
        mps = CanonicalMPS([t.copy() for t in tensors], 0)
        out, errs = mps.compress(chi_max=c["chi_max"], cut=c["cut"], renormalise=c["renormalise"],
                                 return_truncation_errors=True)
        assert out.bond_dimensions == c["bond_dimensions"]
        np.testing.assert_allclose(errs, c["errors"], rtol=1e-9, atol=1e-15)
        dense = out.dense(flatten=True)
        np.testing.assert_allclose(np.sign(np.dot(dense, c["dense"])) * dense, c["dense"], atol=1e-11)
        assert mps.compress(chi_max=c["chi_max"])[1] == [None]
    b = g["bond_case"]
    out, err = CanonicalMPS([t.copy() for t in tensors], 0).compress_bond(b["bond"], chi_max=b["chi_max"],
                                                                          return_truncation_error=True)
    assert out.orth_centre == b["orth_centre"] and out.bond_dimensions == b["bond_dimensions"]
    np.testing.assert_allclose(err, b["error"], rtol=1e-9)
    dense = out.dense(flatten=True)
    np.testing.assert_allclose(np.sign(np.dot(dense, b["dense"])) * dense, b["dense"], atol=1e-11)
    # centre on the last site: the chain is processed mirrored, errors come back in bond order
    c = g["cases"][0]
    n = len(tensors)
    right = CanonicalMPS([t.copy() for t in tensors], 0).move_orth_centre(n - 1, renormalise=False)
    out2, errs2 = right.compress(chi_max=c["chi_max"], cut=c["cut"], return_truncation_errors=True)
    assert len(errs2) == n - 1 and max(out2.bond_dimensions) <= c["chi_max"]
    mirrored = CanonicalMPS([np.ascontiguousarray(np.transpose(t)) for t in reversed(right.tensors)], 0)
    out3, errs3 = mirrored.compress(chi_max=c["chi_max"], cut=c["cut"], return_truncation_errors=True)
    np.testing.assert_allclose(errs2, errs3[::-1], rtol=1e-9, atol=1e-15)
    with pytest.raises(ValueError):
        right.compress_bond(n - 1)
    with pytest.raises(ValueError):
        right.compress(strategy="lu")


@pytest.mark.gpu
def test_svd_kernel_split_routes():
    """`svd` / `split_two_site_tensor` (svd_kernel, cyclic one-sided Jacobi) on the matrix shapes the decode produces and
    on a dense one: orthogonal columns spanning six decades, disjoint non-orthogonal pairs -- equal, very unequal,
    every column --, a cluster of three, random dense.  Singular values against LAPACK, reconstruction, orthonormal
    factors.  (The probe / pair-rotation routes of the decode's own splits are covered by the d=5 depolarising tests
    in test_gpu_decode.py and, route by route, by the host build in test_emu_engine.py.)"""
    from mdopt_b200.utils import svd_batch

    rng = np.random.default_rng(2024)
    m = n = 64

    def ortho(scales):
        q, _ = np.linalg.qr(rng.standard_normal((m, n)))
        return q * np.asarray(scales)[None, :]

    scales = np.exp(rng.uniform(-6, 0, n))
    base = ortho(scales)
    pairs = base.copy()
    for a, b, w in ((1, 7, 0.4), (3, 20, -1.3), (10, 11, 0.05), (30, 2, 2.0), (40, 63, 0.8)):
        pairs[:, b] += w * pairs[:, a] * (scales[b] / scales[a])
    unequal = ortho(np.where(np.arange(n) == 5, 1e-7, 1.0))
    unequal[:, 5] += 3e-8 * unequal[:, 9]
    every = ortho(np.exp(rng.uniform(-3, 0, n)))
    perm = rng.permutation(n)
    for k in range(0, n, 2):
        every[:, perm[k + 1]] += rng.uniform(0.2, 1.5) * every[:, perm[k]]
    cluster = base.copy()
    cluster[:, 4] += 0.5 * cluster[:, 8] * (scales[4] / scales[8])
    cluster[:, 12] += 0.7 * cluster[:, 8] * (scales[12] / scales[8])
    dense = rng.standard_normal((m, n))
    mats = np.stack([base, pairs, unequal, every, cluster, dense])
    u, s, vh, kept, trunc, status = svd_batch(mats, cut=0.0, chi_max=n)
    assert (status == 0).all()
    for i, a in enumerate(mats):
        sv = np.linalg.svd(a, compute_uv=False)
        k = int(kept[i])
        big = sv > 1e-9 * sv[0]
        np.testing.assert_allclose(s[i][:len(sv)][big], sv[big], rtol=1e-10, err_msg=f"matrix {i}")
        np.testing.assert_allclose((u[i][:, :k] * s[i][:k]) @ vh[i][:k], a, atol=1e-12 * sv[0], err_msg=f"matrix {i}")
        np.testing.assert_allclose(u[i][:, :k].T @ u[i][:, :k], np.eye(k), atol=1e-10, err_msg=f"matrix {i}")
