"""CPU tier: host-side logic (schedule builder, constraint strings, sharding + the gloo reduction)."""
import os
import socket

import numpy as np
import pytest

import mps_oracle as O
from mdopt_b200 import distributed as D
from mdopt_b200 import schedule
from conftest import load_golden
from mdopt_b200.codes import surface_code
from mdopt_b200.drivers import quantum_surface as drv
from mdopt_b200.optimiser import utils as U


def test_mpo_tensors_equal_reference_tables():
    from conftest import load_golden

    ref = load_golden("mpo_tensors")
    for name, values in ref.items():
        np.testing.assert_array_equal(getattr(U, name), np.asarray(values))


def test_isometry_flags():
    flags = {n: schedule.TensorTable.is_isometric(getattr(U, n))
             for n in ["IDENTITY", "XOR_BULK", "SWAP", "XOR_LEFT", "XOR_RIGHT", "COPY_LEFT"]}
    assert flags == {"IDENTITY": True, "XOR_BULK": True, "SWAP": True, "XOR_LEFT": True, "XOR_RIGHT": False,
                     "COPY_LEFT": False}


def test_site_lists_match_oracle_and_notebook():
    code = surface_code(3)
    checks, logicals = schedule.code_site_lists(code)
    ochecks, ologicals = O.code_sites(O.surface_code(3))
    assert checks == ochecks and logicals == ologicals
    # printed in the reference's notebook (quantum_surface.ipynb cells 10/13/19)
    assert checks[0][0] == [[2], [4], [3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19], [20]]
    assert logicals[0][0][0] == [0] and logicals[0][0][1] == [6, 12] and logicals[0][0][3] == [18]
    assert logicals[1][0] == [[1], [3, 5], [2, 4, 6], [7]]


def test_program_layout():
    code = surface_code(3)
    prog = schedule.build_decode_program(code, True, False, True, "Optimised")
    assert prog.n_sites == 28 and prog.n_logical == 2 and prog.n_strings == 7
    ops = prog.ops.tolist()
    pc, spans = 0, []
    while ops[pc] == schedule.OP_SWEEP:
        start, span = ops[pc + 1], ops[pc + 2]
        assert ops[pc + 3] == 1 and ops[pc + 4] == 0
        ids = ops[pc + 5: pc + 5 + span]
        assert all(0 <= i < prog.wtab.shape[0] for i in ids)
        assert prog.wdim[ids[-1]][1] == 1 and prog.wdim[ids[0]][0] == 1  # MPO bond closes at both ends
        spans.append((start, span))
        pc += 5 + span
    assert ops[pc:pc + 2] == [schedule.OP_MARGINAL, schedule.OP_END]
    assert spans == prog.spans
    assert spans[0] == (0, 19)  # the X logical first, then the checks ordered by (first, last) site
    assert [s[0] for s in spans[1:]] == sorted(s[0] for s in spans[1:])
    none = schedule.build_decode_program(code, False, False)
    assert none.ops.tolist() == [schedule.OP_MARGINAL, schedule.OP_END]
    depol = schedule.build_decode_program(code, False, False, True, "Naive", None, "Depolarising", 0.1)
    ops = depol.ops.tolist()
    assert ops[:4] == [schedule.OP_GATE2, 2, 0, 1] and ops[-6:-2] == [schedule.OP_GATE2, 26, 0, 1]
    assert len(ops) == 4 * 25 + 2
    np.testing.assert_allclose(depol.wtab[0].reshape(2, 2, 2, 2), O.depolarising_bias(0.1))


def test_symbols_and_flags():
    syms = schedule.symbols_from_errors(["IXZY"], 2)
    assert syms.tolist() == [[2, 2, 0, 0, 1, 0, 0, 1, 1, 1]]
    assert schedule.symbols_from_errors(["EX"], 1).tolist() == [[2, 2, 2, 1, 0]]
    assert schedule.readout_kept_mask("IIXE", 2) == 0b11
    assert schedule.readout_kept_mask("EIXI", 2) == 0b110
    assert schedule.readout_kept_mask("EEXI", 2) == 0b1100
    assert schedule.readout_kept_mask("IEIE", 2) == 0b101
    assert schedule.error_flags("IIII") == (True, False, False)
    assert schedule.error_flags("IYII") == (False, False, False)
    assert schedule.error_flags("XYZI") == (False, True, True)
    with pytest.raises(ValueError):
        schedule.symbols_from_errors(["IQ"], 1)
    with pytest.raises(ValueError):
        schedule.symbols_from_errors(["II", "I"], 1)


def test_constraint_string_errors():
    t = [U.XOR_LEFT, U.XOR_BULK, U.SWAP, U.XOR_RIGHT]
    with pytest.raises(ValueError):
        U.ConstraintString([], [[0]])
    with pytest.raises(ValueError):
        U.ConstraintString(t, [])
    with pytest.raises(ValueError):
        U.ConstraintString(t, [[0], [1]])
    with pytest.raises(ValueError):
        U.ConstraintString(t, [[0], [1], [1], [2]])
    with pytest.raises(ValueError):
        U.ConstraintString(t, [[0], [1], [], [3]])
    cs = U.ConstraintString(t, [[3], [5], [4, 6], [7]])
    assert cs.span() == 5 and cs.start() == 3 and cs.tensor_slots() == [0, 2, 1, 2, 3]
    assert [m.shape for m in cs.mpo()] == [(1, 2, 2, 2), (2, 2, 2, 2), (2, 2, 2, 2), (2, 2, 2, 2), (2, 1, 2, 2)]


def test_seeded_errors_match_reference_generator():
    ours = drv.generate_errors(5, 0.05, 20, "Bitflip", 123)
    assert ours == O.seeded_errors(41, 0.05, 20, 123, "Bitflip")
    from conftest import load_golden

    assert ours == [r["error"] for r in load_golden("d5_chi64_bitflip")["results"]][:20]


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            parts = [D.shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_bounds(n_items, rank, world)
    rng = np.random.default_rng(5)
    success = (rng.random(n_items) < 0.8).astype(np.int32)
    rows = rng.random((n_items, 4))
    rows[3] = np.nan
    total = D.reduce_counts(D.local_counts(rows[lo:hi], success[lo:hi]))
    gathered = D.gather_rows(rows[lo:hi])
    ok = (total.tolist() == [n_items, int(success.sum()), 1]) and np.array_equal(
        np.nan_to_num(gathered, nan=-1.0), np.nan_to_num(rows, nan=-1.0))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_count_reduction():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, 37, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_stabiliser_helpers_follow_reference_conventions():
    """css_code_stabilisers writes x_stabs_binary rows with Z and z_stabs_binary rows with X (decoding.py:492-507);
    every stabiliser the reference drew for the fixture is in that list."""
    from mdopt_b200.codes import surface_code
    from mdopt_b200.decoding import css_code_stabilisers, multiply_pauli_strings

    code = surface_code(3)
    first, second = css_code_stabilisers(code)
    assert all(set(s) <= {"I", "Z"} for s in first) and all(set(s) <= {"I", "X"} for s in second)
    assert len(first) == code.num_x_stabs() and len(second) == code.num_z_stabs()
    g = load_golden("stabiliser_3x3")
    for r in g["results"]:
        assert r["stabiliser"] in first + second
    assert multiply_pauli_strings("IXYZ", "XXXX") == "XIZY"
    assert multiply_pauli_strings("XZ", "ZX") == "YY"
    with pytest.raises(ValueError):
        multiply_pauli_strings("XX", "X")


def test_code_constructors_give_valid_css_codes():
    """hypergraph_product / repetition_code / random_regular_code (the qecstruct constructors the drivers use)."""
    from mdopt_b200.codes import hypergraph_product, random_regular_code, repetition_code, surface_code

    s3, ref = hypergraph_product(repetition_code(3), repetition_code(3)), surface_code(3)
    assert s3.x_stabs_binary().rows() == ref.x_stabs_binary().rows()
    assert s3.z_stabs_binary().rows() == ref.z_stabs_binary().rows()
    assert s3.x_logicals_binary().rows() == [[2, 5, 8]] and s3.z_logicals_binary().rows() == [[0, 1, 2]]

    def mat(rows, n):
        m = np.zeros((len(rows), n), dtype=int)
        for i, r in enumerate(rows):
            m[i, r] = 1
        return m

    def rank(m):
        m, r = m.copy() % 2, 0
        for c in range(m.shape[1]):
            piv = [i for i in range(r, m.shape[0]) if m[i, c]]
            if not piv:
                continue
            m[[r, piv[0]]] = m[[piv[0], r]]
            for i in range(m.shape[0]):
                if i != r and m[i, c]:
                    m[i] ^= m[r]
            r += 1
        return r

    ldpc = random_regular_code(16, 12, 3, 4, 7)   # the classical code of quantum_hypergraph_product.py:134-142
    h = mat(ldpc.par_mat().rows(), 16)
    assert (h.sum(axis=0) == 3).all() and (h.sum(axis=1) == 4).all()
    assert random_regular_code(16, 12, 3, 4, 7).par_mat().rows() == ldpc.par_mat().rows()   # seeded
    with pytest.raises(ValueError):
        random_regular_code(16, 11, 3, 4, 7)
    for a, b in [(repetition_code(3), repetition_code(4)), (ldpc, ldpc)]:
        c = hypergraph_product(a, b)
        n = len(c)
        hx, hz = mat(c.x_stabs_binary().rows(), n), mat(c.z_stabs_binary().rows(), n)
        lx, lz = mat(c.x_logicals_binary().rows(), n), mat(c.z_logicals_binary().rows(), n)
        k = n - rank(hx) - rank(hz)
        assert not (hx @ hz.T % 2).any() and not (lx @ hz.T % 2).any() and not (lz @ hx.T % 2).any()
        assert len(lx) == len(lz) == k == c.num_x_logicals() == c.num_z_logicals()
        assert rank(np.vstack([hx, lx])) == rank(hx) + k and rank(lx @ lz.T % 2) == k
    assert len(hypergraph_product(ldpc, ldpc)) == 400


def test_front_minimising_order():
    from mdopt_b200.optimiser.utils import front_minimising_order, msro

    rng = np.random.default_rng(4)
    loc = np.zeros((12, 40), dtype=int)
    for r in range(12):
        lo = rng.integers(0, 30)
        loc[r, lo:lo + rng.integers(2, 10)] = 1
    order = front_minimising_order(loc)
    assert sorted(order) == list(range(12))

    def max_front(o):
        spans = [(np.nonzero(loc[r])[0][0], np.nonzero(loc[r])[0][-1]) for r in range(12)]
        cover, ahead, worst = np.zeros(40, int), np.zeros(40, int), 0
        for lo, hi in spans:
            ahead[lo:hi + 1] += 1
        for r in o:
            lo, hi = spans[r]
            cover[lo:hi + 1] += 1
            ahead[lo:hi + 1] -= 1
            worst = max(worst, int(np.count_nonzero((cover > 0) & (ahead > 0))))
        return worst

    assert max_front(order) <= max_front(msro(loc))
    code = __import__("mdopt_b200.codes", fromlist=["surface_code"]).surface_code(3)
    prog_a = schedule.build_decode_program(code, True, False, True, "Optimised", {"xc": "front"})
    prog_b = schedule.build_decode_program(code, True, False, True, "Optimised")
    assert prog_a.n_strings == prog_b.n_strings and sorted(prog_a.spans) == sorted(prog_b.spans)
