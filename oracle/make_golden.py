"""Generate tests/golden/*.json by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=oracle/refshim:/root/reference python oracle/make_golden.py

Third-party modules the reference imports but this image lacks are replaced by the stand-ins in
oracle/refshim (opt_einsum -> numpy.einsum with the reference's explicit path; matrex.msro -> the
(first site, last site) row sort; qecstruct -> duck-typed CSS code objects).  Every fixture stores the
inputs (error strings, kwargs), the reference outputs, and for one case the singular values of every
sweep split, so tests can pin oracle/mps_oracle.py against the reference without re-importing it.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import qecstruct as qec  # noqa: E402  (refshim)
import mdopt.utils.utils as ref_utils  # noqa: E402
from examples.decoding.decoding import decode_css  # noqa: E402
from mdopt.optimiser import utils as ref_opt  # noqa: E402
from mdopt.mps.canonical import CanonicalMPS  # noqa: E402

import mps_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def surface(L):
    return qec.hypergraph_product(qec.repetition_code(L), qec.repetition_code(L))


def run_batch(name, L, errors, **kw):
    code = surface(L)
    t0 = time.time()
    rows = []
    for e in errors:
        dist, ok = decode_css(code, e, silent=True, contraction_strategy="Optimised", **kw)
        rows.append({"error": e, "dist": [float(x) for x in dist], "success": int(ok)})
    doc = {"name": name, "lattice_size": L, "kwargs": kw, "results": rows,
           "reference_seconds": round(time.time() - t0, 2)}
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(doc, fh, indent=0)
    print(name, len(rows), "decodes", doc["reference_seconds"], "s")


# erasure errors ('E' -> '++', decoding.py:283-320, 1341-1347), incl. erasures of qubits 0 / 1 whose indices
# the reference hands to marginal() as site indices
ERASURE_CASES = ["IIEIIIIIIIIII", "IIXIIEIIIIIII", "EIXIIIIIIIIII", "IEIIZIIIIIIII", "EEIIIIIIIIIII", "IIEIIIIIIIIIE",
                 "XIIEIIIYIIIEI", "IIIIIIXEIIIII", "ZEIIIIIIIIIIE", "IIXIIEIIEIIIZ"]


FIVE_QUBIT = dict(stabilizers=["XZZXI", "IXZZX", "XIXZZ", "ZXIXZ"], x_logicals=["XXXXX"], z_logicals=["ZZZZZ"])


def run_custom(name, errors, **kw):
    """decode_custom on the five-qubit code (examples/decoding/quantum_five_qubit.ipynb)."""
    from examples.decoding.decoding import decode_custom

    rows = []
    for e in errors:
        dist, ok = decode_custom(FIVE_QUBIT["stabilizers"], FIVE_QUBIT["x_logicals"], FIVE_QUBIT["z_logicals"], e,
                                 silent=True, **kw)
        rows.append({"error": e, "dist": [float(x) for x in dist], "success": int(ok)})
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump({"name": name, "code": FIVE_QUBIT, "kwargs": kw, "results": rows}, fh, indent=0)
    print(name, len(rows), "decodes")


def five_qubit_errors():
    singles = ["I" * q + p + "I" * (4 - q) for q in range(5) for p in "XYZ"]
    return singles + ["XXIII", "IZIZI", "YIIIX", "IIXZI", "ZZZII", "XIYIZ", "IIIII", "EIIIX"]


def _pool_decode(args):
    """Worker of the ``big`` mode: one decode by the unmodified reference, optionally with the canonical
    tie-break installed at the reference's own SVD hook (``xp.linalg.svd``, utils.py:80)."""
    L, e, kw, hooked = args
    import structured_svd as S
    code = surface(L)
    if hooked:
        S.STATS.update(structured=0, lapack=0)
        with S.installed():
            dist, ok = decode_css(code, e, silent=True, contraction_strategy="Optimised", **kw)
        extra = S.STATS["lapack"]
    else:
        dist, ok = decode_css(code, e, silent=True, contraction_strategy="Optimised", **kw)
        extra = 0
    return [float(x) for x in dist], int(ok), int(extra)


def run_big(name, L, errors, nproc, hooked, **kw):
    """Large fixtures, decoded under multiprocessing.Pool like the reference's own driver
    (quantum_surface.py:242-246).  Compact layout: parallel lists instead of one dict per syndrome."""
    import multiprocessing as mp

    t0 = time.time()
    with mp.Pool(nproc) as pool:
        rows = pool.map(_pool_decode, [(L, e, kw, hooked) for e in errors], chunksize=4)
    doc = {"name": name, "lattice_size": L, "kwargs": kw, "tie_break": "canonical" if hooked else "lapack",
           "errors": list(errors), "dist": [r[0] for r in rows], "success": [r[1] for r in rows],
           "lapack_calls_under_hook": int(sum(r[2] for r in rows)),
           "reference_seconds": round(time.time() - t0, 1), "processes": nproc}
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(doc, fh)
    print(name, len(rows), "decodes", doc["reference_seconds"], "s", "lapack under hook:",
          doc["lapack_calls_under_hook"], flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        # >= 1000 syndromes of the headline config (the first 1024 errors of the bench batch)
        # from the unmodified reference, the same with the canonical tie-break injected at the reference's SVD hook,
        # and >= 64 depolarising syndromes
        nproc = int(sys.argv[2]) if len(sys.argv) > 2 else 6
        kw = dict(chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
        errs = O.seeded_errors(41, 0.05, 1024, 123, "Bitflip")
        run_big("d5_chi64_bitflip_1024_canonical", 5, errs, nproc, True, **kw)
        run_big("d5_chi64_bitflip_1024", 5, errs, nproc, False, **kw)
        kwd = dict(chi_max=64, cut=1e-17, bias_type="Depolarising", bias_prob=0.1, tolerance=1e-12)
        run_big("d5_chi64_depol_64", 5, O.seeded_errors(41, 0.05, 64, 321, "Depolarising"), nproc, False, **kwd)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "custom":
        run_custom("five_qubit_bitflip", five_qubit_errors(), chi_max=int(1e4), cut=1e-17, bias_type="Bitflip",
                   bias_prob=0.01, contraction_strategy="Optimised")
        run_custom("five_qubit_depol", five_qubit_errors(), chi_max=int(1e4), cut=1e-17, bias_type="Depolarising",
                   bias_prob=0.1, contraction_strategy="Naive")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dmrg":
        # Dephasing DMRG from |+...+> on seeded positive targets with a unique main component
        # (the reference's own test problem, tests/optimiser/test_dephasing_dmrg.py:199-300): target tensors in, the
        # bitstring engine.mps ends on out, after one and after two sweeps
        from mdopt.mps.utils import create_simple_product_state
        from mdopt.optimiser.dephasing_dmrg import DephasingDMRG

        rng = np.random.default_rng(31)
        cases = []
        for n, chi in [(8, 4), (13, 6), (16, 8), (14, 3), (20, 5)]:
            dims = [min(chi, 2 ** i, 2 ** (n - i)) for i in range(n + 1)]
            tensors = [rng.uniform(0.05, 1.0, size=(dims[i], 2, dims[i + 1])) for i in range(n)]
            target = CanonicalMPS([t.copy() for t in tensors], n - 1, 1e-12, int(1e4))
            out = {}
            for runs in (1, 2):
                eng = DephasingDMRG(mps=create_simple_product_state(n, which="+"), mps_target=target.copy(),
                                    chi_max=int(1e4), cut=1e-17, mode="LA", silent=True)
                eng.run(num_iter=runs)
                assert eng.mps.bond_dimensions == [1] * (n - 1)
                out[runs] = [int(np.argmax(np.abs(t.reshape(-1)))) for t in eng.mps.tensors]
            cases.append({"tensors": [t.tolist() for t in tensors], "bits_1": out[1], "bits_2": out[2]})
            print("dmrg", n, chi, out[1], out[2])
        # decode_css with more than 12 logical sites (decoding.py:1382-1400): 7 Steane blocks = 14 logical sites,
        # errors with X and Z letters so that every logical site is constrained (unique main component); the SVDs
        # run with the canonical tie-break installed (chi_max = 64 truncates)
        import structured_svd as S

        rows = [[3, 4, 5, 6], [1, 2, 5, 6], [0, 2, 4, 6]]
        xs, zs, xl, zl = [], [], [], []
        for b in range(7):
            o = 7 * b
            for r in rows:
                xs.append([o + c for c in r])
                zs.append([o + c for c in r])
            xl.append([o, o + 1, o + 2])
            zl.append([o, o + 1, o + 2])
        code = qec.CssCode(49, xs, zs, xl, zl)

        def err(**kw):
            e = ["I"] * 49
            for q, p in kw.items():
                e[int(q[1:])] = p
            return "".join(e)

        errors = [err(q0="X", q8="Z"), err(q3="X", q4="X", q20="Z"), err(q10="Y", q30="X", q44="Z"),
                  err(q0="X", q1="X", q2="X", q47="Z"), err(q5="Z", q6="Z", q12="X", q13="X", q40="X")]
        kw = dict(chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
        decodes = []
        for e in errors:
            with S.installed():
                eng, ov = decode_css(code, e, silent=True, contraction_strategy="Naive", **kw)
            bits = [int(np.argmax(np.abs(t.reshape(-1)))) for t in eng.mps.tensors]
            decodes.append({"error": e, "bits": bits, "overlap": float(ov)})
            print("decode", e[:20], bits, ov)
        with open(os.path.join(OUT, "dephasing_dmrg.json"), "w") as fh:
            json.dump({"cases": cases, "steane7": {"x_stabs": xs, "z_stabs": zs, "x_logicals": xl, "z_logicals": zl,
                                                   "kwargs": kw, "decodes": decodes}}, fh)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "randcirc":
        # examples/random_circuit/mps-rand-circ.py:48-83 on a seeded 8-qubit circuit: fidelity tail per layer
        from scipy.stats import unitary_group

        from mdopt.contractor.contractor import mps_mpo_contract
        from mdopt.mps.utils import create_simple_product_state

        nq, chi, nl, seed = 8, 6, 6, 11
        rng = np.random.default_rng(seed)
        st = create_simple_product_state(num_sites=nq, phys_dim=2, which="0", form="Right-canonical", tolerance=np.inf)
        tails = []
        for k in range(nl):
            layer = unitary_group.rvs(4, size=nq // 2 - k % 2, random_state=rng)
            mpo = []
            for gate in layer:   # random_circuit.py:9-71
                q, r = np.linalg.qr(gate)
                mpo += [q.reshape((1, q.shape[1], 2, 2)), r.reshape((r.shape[0], 1, 2, 2))]
            st = mps_mpo_contract(mps=st, mpo=mpo, start_site=k % 2, renormalise=True, chi_max=1e4, cut=1e-12,
                                  inplace=False)
            st, errs = st.compress(chi_max=chi, cut=1e-12, renormalise=True, strategy="svd",
                                   return_truncation_errors=True)
            tails.append(float(np.prod([1 - e for e in errs])))
        with open(os.path.join(OUT, "random_circuit.json"), "w") as fh:
            json.dump({"num_qubits": nq, "bond_dim": chi, "num_layers": nl, "seed": seed, "fidelity_tail": tails,
                       "bond_dimensions": [int(b) for b in st.bond_dimensions]}, fh)
        print("random_circuit", tails)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "stabiliser":
        # multiply_by_stabiliser=True (decoding.py:1237-1240): the draw comes from numpy's global RNG, so seed it,
        # record which stabiliser np.random.choice returns and replay the same seed for the decode
        from examples.decoding.decoding import css_code_stabilisers
        code = surface(3)
        sx, sz = css_code_stabilisers(code)
        kw = dict(chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
        rows = []
        for seed, e in enumerate(["IIXIIIIIIIIII", "IIXIIZIIIIIII", "ZIIIIIIIIIIIZ", "IYIIIIXIIIIII"]):
            np.random.seed(seed)
            pick = str(np.random.choice(sx + sz))
            np.random.seed(seed)
            dist, ok = decode_css(code, e, multiply_by_stabiliser=True, silent=True,
                                  contraction_strategy="Optimised", **kw)
            rows.append({"error": e, "stabiliser": pick, "dist": [float(x) for x in dist], "success": int(ok)})
        with open(os.path.join(OUT, "stabiliser_3x3.json"), "w") as fh:
            json.dump({"name": "stabiliser_3x3", "lattice_size": 3, "kwargs": kw, "results": rows}, fh, indent=0)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "compress":
        # CanonicalMPS.compress / compress_bond (canonical.py:684-904, "svd") on a seeded real MPS with a generic
        # spectrum: inputs, truncation errors, bond dimensions and the dense result
        rng = np.random.default_rng(77)
        n, chi = 9, 16
        dims = [min(chi, 2 ** i, 2 ** (n - i)) for i in range(n + 1)]
        tensors = []
        for i in range(n):
            q, _ = np.linalg.qr(rng.standard_normal((dims[i], 2 * dims[i + 1])).T)
            tensors.append(q.T.reshape(dims[i], 2, dims[i + 1]))
        tensors[0] = tensors[0] / np.linalg.norm(tensors[0])
        cases = []
        for (chi_max, cut, ren) in [(6, 1e-17, False), (6, 1e-17, True), (100, 0.05, False), (3, 1e-17, True)]:
            mps = CanonicalMPS([t.copy() for t in tensors], 0, 1e-12, int(1e4))
            out, errs = mps.compress(chi_max=chi_max, cut=cut, renormalise=ren, return_truncation_errors=True)
            cases.append({"chi_max": chi_max, "cut": cut, "renormalise": ren, "errors": [float(e) for e in errs],
                          "bond_dimensions": [int(b) for b in out.bond_dimensions],
                          "dense": [float(x) for x in out.dense(flatten=True)]})
        one = CanonicalMPS([t.copy() for t in tensors], 0, 1e-12, int(1e4))
        cb, err = one.compress_bond(bond=4, chi_max=5, cut=1e-17, renormalise=False, return_truncation_error=True)
        doc = {"tensors": [t.tolist() for t in tensors], "cases": cases,
               "bond_case": {"bond": 4, "chi_max": 5, "error": float(err), "orth_centre": int(cb.orth_centre),
                             "bond_dimensions": [int(b) for b in cb.bond_dimensions],
                             "dense": [float(x) for x in cb.dense(flatten=True)]}}
        with open(os.path.join(OUT, "compress_case.json"), "w") as fh:
            json.dump(doc, fh)
        print("compress_case", [c["bond_dimensions"] for c in cases])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "bigchi":
        # large-chi tier (capacity 128): d=5 still truncates at chi_max=128, so this pins success flags and
        # the marginals to the degenerate-truncation tolerance, like d5_chi64_bitflip
        errs = [e for e in O.seeded_errors(41, 0.05, 40, 321, "Bitflip") if "X" in e][:5]
        run_batch("d5_chi128_bitflip", 5, errs, chi_max=128, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                  tolerance=1e-12)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "erasure":
        run_batch("erasure_3x3", 3, ERASURE_CASES, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                  tolerance=1e-12)
        return
    # config 1: README minimal example (README.md:38-61)
    run_batch("readme_3x3", 3, ["IIXIIIIIIIIII"], chi_max=64, bias_type="Bitflip", bias_prob=0.01,
              tolerance=1e-12)
    # character-gating quirk cases
    run_batch("gating_3x3", 3, ["IIYIIIIIIIIII", "IIXIIIIIIIIII", "IIXIIZIIIIIII", "YIXIIIIIIIIII",
                                "IIIIZIIIIIIII", "XXXXXXXXXXXXX", "ZIIIIIIIIIIIX"],
              chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    n3, n5 = 13, 41
    run_batch("d3_chi64_bitflip", 3, O.seeded_errors(n3, 0.05, 96, 123, "Bitflip"),
              chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    run_batch("d3_chi64_bitflip_cut0", 3, O.seeded_errors(n3, 0.10, 48, 100, "Bitflip"),
              chi_max=64, cut=0.0, bias_type="Bitflip", bias_prob=0.1, tolerance=0.0)
    run_batch("d3_chi64_depol", 3, O.seeded_errors(n3, 0.08, 48, 11, "Depolarising"),
              chi_max=64, cut=1e-17, bias_type="Depolarising", bias_prob=0.1, tolerance=1e-12)
    run_batch("d3_chi8_bitflip", 3, O.seeded_errors(n3, 0.08, 48, 7, "Bitflip"),
              chi_max=8, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    run_batch("d5_chi64_bitflip", 5, O.seeded_errors(n5, 0.05, 24, 123, "Bitflip"),
              chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    run_batch("d5_chi16_bitflip", 5, O.seeded_errors(n5, 0.05, 12, 5, "Bitflip"),
              chi_max=16, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    run_batch("erasure_3x3", 3, ERASURE_CASES, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
              tolerance=1e-12)

    run_custom("five_qubit_bitflip", five_qubit_errors(), chi_max=int(1e4), cut=1e-17, bias_type="Bitflip",
               bias_prob=0.01, contraction_strategy="Optimised")
    run_custom("five_qubit_depol", five_qubit_errors(), chi_max=int(1e4), cut=1e-17, bias_type="Depolarising",
               bias_prob=0.1, contraction_strategy="Naive")

    # exact MPO tensor tables (reference values: mdopt/optimiser/utils.py:23-43)
    tables = {k: getattr(ref_opt, k).tolist() for k in
              ["IDENTITY", "COPY_LEFT", "XOR_BULK", "XOR_LEFT", "XOR_RIGHT", "SWAP"]}
    with open(os.path.join(OUT, "mpo_tensors.json"), "w") as fh:
        json.dump(tables, fh)

    # reference svd / split on seeded matrices (singular values + truncation counts)
    rng = np.random.default_rng(2024)
    cases = []
    for (m, n, chi, cut) in [(8, 8, 4, 1e-12), (16, 32, 8, 1e-3), (32, 16, 100, 0.5), (64, 64, 64, 1e-17),
                             (128, 128, 64, 1e-17), (20, 128, 16, 1e-12)]:
        a = rng.standard_normal((m, n))
        u, s, vh, err = ref_utils.svd(a, cut=cut, chi_max=chi, renormalise=False, return_truncation_error=True)
        cases.append({"m": m, "n": n, "chi_max": chi, "cut": cut,
                      "s": [float(x) for x in s], "trunc_err": float(err)})
    with open(os.path.join(OUT, "svd_cases.json"), "w") as fh:
        json.dump({"rng_seed": 2024, "note": "matrices are rng.standard_normal((m,n)) drawn in this order from default_rng(2024)", "cases": cases}, fh)

    # marginal hand case (tests/mps/test_canonical.py:851-865 style): small random MPS, reference marginal
    tens = [rng.standard_normal((1, 2, 3)), rng.standard_normal((3, 2, 4)), rng.standard_normal((4, 2, 2)),
            rng.standard_normal((2, 2, 1))]
    mps = CanonicalMPS([t.copy() for t in tens])
    out = mps.marginal([1, 3], renormalise=True)
    with open(os.path.join(OUT, "marginal_case.json"), "w") as fh:
        json.dump({"tensors": [t.tolist() for t in tens], "sites": [1, 3],
                   "result": [t.tolist() for t in out.tensors]}, fh)

    # sweep spectra of one d=3 decode from the reference itself (hook on split_two_site_tensor's svd)
    spectra = []
    orig = ref_utils.svd

    def spy(mat, cut=1e-12, chi_max=int(1e4), renormalise=False, return_truncation_error=False):
        res = orig(mat, cut=cut, chi_max=chi_max, renormalise=renormalise,
                   return_truncation_error=return_truncation_error)
        spectra.append({"shape": list(mat.shape), "chi_max": int(chi_max), "s": [float(x) for x in res[1]]})
        return res

    ref_utils.svd = spy
    try:
        err = "IXIIIIIXIIIII"
        dist, ok = decode_css(surface(3), err, silent=True, contraction_strategy="Optimised", chi_max=64,
                              cut=1e-17, bias_type="Bitflip", bias_prob=0.1, tolerance=1e-12)
    finally:
        ref_utils.svd = orig
    with open(os.path.join(OUT, "d3_spectra.json"), "w") as fh:
        json.dump({"error": err, "dist": [float(x) for x in dist], "success": int(ok), "svds": spectra}, fh)
    print("done")


if __name__ == "__main__":
    main()
