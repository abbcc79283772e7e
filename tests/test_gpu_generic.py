"""The library tier (mdopt_b200/generic.py): inputs outside the specialised engines' domain -- complex dtypes, MPO
bond > 2, arbitrary-site marginals -- follow the reference's own tests (tests/contractor/test_contractor.py:152-208,
tests/mps/test_canonical.py:748-865) on the GPU."""
import numpy as np
import pytest

import mps_oracle as O

pytestmark = pytest.mark.gpu


def random_mps(rng, n, chi, complex_=True, phys=2):
    dims = [min(chi, phys ** i, phys ** (n - i)) for i in range(n + 1)]
    ts = []
    for i in range(n):
        shape = (dims[i], phys, dims[i + 1])
        ts.append(rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if complex_ else 0))
    return ts


def dense(ts):
    out = ts[0]
    for t in ts[1:]:
        out = np.tensordot(out, t, (-1, 0))
    return out.reshape(-1)


def apply_mpo_dense(psi, mpo, start, n):
    """(vL, vR, pU, pD) tensors applied to a dense n-qubit state on sites start.. (pU contracts with the state,
    contractor.py:215-217)."""
    state = psi.reshape([2] * n)
    op = mpo[0][0]                                   # (vR, pU, pD)
    for w in mpo[1:]:
        op = np.tensordot(op, w, ([0], [0]))         # (..., vR', pU, pD) with the new bond last-3
        op = np.moveaxis(op, -3, 0)
    op = op[0]                                       # legs (pU0, pD0, pU1, pD1, ...)
    k = len(mpo)
    ups, downs = list(range(0, 2 * k, 2)), list(range(1, 2 * k, 2))
    out = np.tensordot(op, state, (ups, list(range(start, start + k))))   # (pD0..pDk-1, rest)
    out = np.moveaxis(out, list(range(k)), list(range(start, start + k)))
    return out.reshape(-1)


def test_mps_mpo_contract_complex_random_mpo_bond_3_to_9():
    """tests/contractor/test_contractor.py:152-208: random 4..8-site state, random complex MPO with bonds 3..9:
    dense(result) == (MPO as a matrix) @ psi, and the sites left of the new centre are left isometries."""
    from mdopt_b200.contractor.contractor import mps_mpo_contract
    from mdopt_b200.mps.canonical import CanonicalMPS

    rng = np.random.default_rng(11)
    for _ in range(6):
        n = int(rng.integers(4, 9))
        raw = random_mps(rng, n, 1 << 10)
        mps = CanonicalMPS(raw, n - 1).move_orth_centre(0)     # library tier: complex input; canonicalises the chain
        psi = dense(mps.tensors)
        length = int(rng.integers(2, n + 1))
        start = int(rng.integers(0, n - length + 1))
        bonds = [1] + [int(rng.integers(3, 10)) for _ in range(length - 1)] + [1]
        mpo = [rng.standard_normal((bonds[i], bonds[i + 1], 2, 2)) + 1j * rng.standard_normal((bonds[i], bonds[i + 1], 2, 2))
               for i in range(length)]
        out = mps_mpo_contract(mps, mpo, start_site=start, chi_max=int(1e4), cut=1e-17)
        assert out.orth_centre == start + length - 1
        expect = apply_mpo_dense(psi, mpo, start, n)
        np.testing.assert_allclose(dense(out.tensors), expect, atol=1e-7 * max(1.0, np.linalg.norm(expect)))
        for i in range(out.orth_centre):
            t = out.tensors[i]
            g = np.einsum("ijk,ijl->kl", t.conj(), t)
            np.testing.assert_allclose(g, np.eye(g.shape[0]), atol=1e-9)
        with pytest.raises(ValueError):
            mps_mpo_contract(mps, mpo + [mpo[-1]] * n, start_site=start)
        with pytest.raises(ValueError):
            mps_mpo_contract(mps, [mpo[0][0]] + mpo[1:], start_site=start)


def test_marginal_of_arbitrary_sites_matches_reference_fixture(golden):
    from mdopt_b200.mps.canonical import CanonicalMPS

    g = golden("marginal_case")   # reference CanonicalMPS.marginal([1, 3], renormalise=True) on a seeded MPS
    mps = CanonicalMPS([np.array(t) for t in g["tensors"]])
    out = mps.marginal(g["sites"], renormalise=True)
    assert out is mps and out.num_sites == len(g["result"]) and out.orth_centre == out.num_sites - 1
    for a, b in zip(out.tensors, g["result"]):
        np.testing.assert_allclose(a, np.array(b), rtol=1e-12, atol=1e-14)
    rng = np.random.default_rng(2)   # first site, complex state
    ts = random_mps(rng, 5, 8)
    ref = O.marginal(O.MPS([t.copy() for t in ts], None, int(1e4)), [0, 2], renormalise=False)
    got = CanonicalMPS([t.copy() for t in ts]).marginal([0, 2])
    assert len(got.tensors) == len(ref.tensors)
    for a, b in zip(got.tensors, ref.tensors):
        np.testing.assert_allclose(a, b, atol=1e-12)


def test_svd_of_complex_and_large_matrices():
    from mdopt_b200.utils.utils import split_two_site_tensor, svd

    rng = np.random.default_rng(5)
    a = rng.standard_normal((40, 24)) + 1j * rng.standard_normal((40, 24))
    u, s, vh, err = svd(a, cut=1e-12, chi_max=10, return_truncation_error=True)
    ref = np.linalg.svd(a, compute_uv=False)
    np.testing.assert_allclose(s, ref[:10], rtol=1e-12)
    assert err == pytest.approx(float(np.sum(ref[10:] ** 2)), rel=1e-10)
    big = rng.standard_normal((300, 200))
    u, s, vh, _ = svd(big, cut=1e-12, chi_max=int(1e4))
    np.testing.assert_allclose((u * s) @ vh, big, atol=1e-10)
    t = rng.standard_normal((6, 2, 2, 6)) + 1j * rng.standard_normal((6, 2, 2, 6))
    ul, sv, vr = split_two_site_tensor(t, chi_max=12, cut=1e-12)
    np.testing.assert_allclose(np.einsum("ijm,m,mkl->ijkl", ul, sv, vr), t, atol=1e-12)


def test_random_circuit_layer_and_compress_complex():
    """examples/random_circuit/mps-rand-circ.py:57-79: a brickwork layer as an MPO (bond 4) applied with
    mps_mpo_contract(renormalise=True), then compress(chi_max) -- complex128 end to end, against the numpy port."""
    from mdopt_b200.contractor.contractor import mps_mpo_contract
    from mdopt_b200.mps.canonical import CanonicalMPS

    rng = np.random.default_rng(7)
    n = 8
    ts = [np.array([1.0, 0.0], dtype=complex).reshape(1, 2, 1) for _ in range(n)]
    mpo = []
    for _ in range(0, n, 2):   # a two-site unitary split by QR into (1, 4, 2, 2) and (4, 1, 2, 2)
        gate = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0].reshape(2, 2, 2, 2)
        mat = gate.transpose(0, 2, 1, 3).reshape(4, 4)          # rows (pU0, pD0), columns (pU1, pD1)
        ql, rr = np.linalg.qr(mat)
        mpo.append(ql.reshape(2, 2, 4).transpose(2, 0, 1).reshape(1, 4, 2, 2))
        mpo.append(rr.reshape(4, 1, 2, 2))
    ref = O.mps_mpo_contract(O.MPS([t.copy() for t in ts], 0, int(1e4)), mpo, start_site=0, renormalise=True,
                             chi_max=int(1e4), cut=1e-12)
    out = mps_mpo_contract(CanonicalMPS([t.copy() for t in ts], 0), mpo, start_site=0, renormalise=True,
                           chi_max=int(1e4), cut=1e-12)
    assert out.bond_dimensions == [t.shape[2] for t in ref.tensors[:-1]]
    # compare states up to the global gauge the two SVD libraries are free to choose (a phase per bond cancels)
    a, b = dense(out.tensors), dense(ref.tensors)
    np.testing.assert_allclose(abs(np.vdot(a, b)), np.linalg.norm(a) * np.linalg.norm(b), rtol=1e-12)
    np.testing.assert_allclose(np.abs(a), np.abs(b), atol=1e-12)
    comp, errs = out.compress(chi_max=2, cut=1e-12, renormalise=True, return_truncation_errors=True)
    assert max(comp.bond_dimensions) <= 2 and len(errs) == n - 1


def test_random_circuit_driver_matches_the_reference_fixture(golden):
    """examples/random_circuit/mps-rand-circ.py on a seeded 8-qubit circuit: the fidelity tail per layer (contract with
    renormalise=True, then compress from the nearer border) equals the unmodified reference's."""
    from mdopt_b200.drivers import random_circuit as rc

    g = golden("random_circuit")
    state, tails, _ = rc.run(g["num_qubits"], g["bond_dim"], g["num_layers"], g["seed"])
    np.testing.assert_allclose(tails, g["fidelity_tail"], rtol=1e-10)
    assert [int(b) for b in state.bond_dimensions] == g["bond_dimensions"]
    assert state.tensors[0].dtype == np.complex128


def test_qr_strategy_and_svd_advanced(golden):
    """`qr` (pivoted QR, utils.py:141-212, :370-383) and the `qr` / `svd_advanced` compression strategies
    (canonical.py:762-833), following the reference's tests (tests/utils/test_utils.py: reconstruction, truncation
    count by |diag R| > cut and chi_max, renormalisation)."""
    from mdopt_b200.mps.canonical import CanonicalMPS
    from mdopt_b200.utils.utils import qr, split_two_site_tensor

    rng = np.random.default_rng(21)
    a = rng.standard_normal((20, 12))
    q, r, err = qr(a, cut=1e-12, chi_max=int(1e4), return_truncation_error=True)
    np.testing.assert_allclose(q @ r, a, atol=1e-12)
    np.testing.assert_allclose(q.T @ q, np.eye(12), atol=1e-12)
    assert err < 1e-12
    low = rng.standard_normal((20, 3)) @ rng.standard_normal((3, 12))        # rank 3
    q, r, _ = qr(low, cut=1e-9)
    assert q.shape == (20, 3) and r.shape == (3, 12)
    np.testing.assert_allclose(q @ r, low, atol=1e-10)
    q, r, err = qr(a, chi_max=5, return_truncation_error=True)
    assert q.shape == (20, 5) and r.shape == (5, 12) and err == pytest.approx(np.linalg.norm(a - q @ r))
    import scipy.linalg
    _, r_ref, _ = scipy.linalg.qr(a, pivoting=True, mode="economic")
    q_full, r_full, _ = qr(a, cut=0.0)
    np.testing.assert_allclose(np.sort(np.abs(np.diag(r_ref)))[::-1], np.abs(np.diag(r_ref)), rtol=1e-12)
    # same pivot order as LAPACK's geqp3 on a generic matrix: |R| agrees entry by entry up to the row signs
    piv = np.argsort(-np.linalg.norm(a, axis=0))[:1]
    assert np.argmax(np.abs(r_full[0])) == piv[0]
    t = rng.standard_normal((5, 2, 2, 5)) + 1j * rng.standard_normal((5, 2, 2, 5))
    ql, rr, e1, e2 = split_two_site_tensor(t, chi_max=int(1e4), cut=1e-12, strategy="qr", return_truncation_error=True)
    assert e2 is None
    np.testing.assert_allclose(np.einsum("ijm,mkl->ijkl", ql, rr), t, atol=1e-12)
    # compression strategies against the numpy port's SVD compression: same truncated state up to the gauge
    g = golden("compress_case")
    tensors = [np.array(x) for x in g["tensors"]]
    for strategy in ("svd_advanced", "qr"):
        mps = CanonicalMPS([x.copy() for x in tensors], 0, 1e-12, int(1e4))
        out, errs = mps.compress(chi_max=100, cut=1e-14, strategy=strategy, return_truncation_errors=True)
        full = CanonicalMPS([x.copy() for x in tensors], 0).dense()
        np.testing.assert_allclose(np.abs(out.dense()), np.abs(full), atol=1e-10)
        assert len(errs) == len(tensors) - 1
    mps = CanonicalMPS([x.copy() for x in tensors], 0, 1e-12, int(1e4))
    out, err = mps.compress_bond(4, chi_max=5, cut=1e-17, strategy="svd_advanced", return_truncation_error=True)
    assert out.tensors[4].shape[2] <= 5 and err >= 0.0
    with pytest.raises(ValueError):
        mps.compress_bond(4, strategy="nope")
