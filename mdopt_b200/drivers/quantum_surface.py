"""Batch driver for surface-code decoding on B200s (drop-in for examples/decoding/quantum_surface.py).

Same command line (reference :60-125), same seeded error generation (:128-146), same result dictionary and
pickle file name (:257-291).  The multiprocessing.Pool fan-out (:242-246) is replaced by one batched GPU
decode per rank; under torchrun the error list is sharded contiguously over the ranks and only the failure
counts cross NVLink.

    python -m mdopt_b200.drivers.quantum_surface --lattice_size 5 --bond_dim 64 --error_rate 0.05 \
        --bias_prob 0.1 --num_experiments 1000 --error_model Bitflip --seed 123 --num_processes 1 \
        --silent True --tolerance 1e-8 --cut 1e-17
"""
from __future__ import annotations

import argparse
import logging
import os
import pickle

import numpy as np


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(description="Launch decoding of quantum surface code.")
    parser.add_argument("--lattice_size", type=int, required=True, help="Lattice size for the surface code.")
    parser.add_argument("--bond_dim", type=int, required=True, help="Maximum bond dimension to keep during contraction.")
    parser.add_argument("--error_rate", type=float, required=True, help="The error rate.")
    parser.add_argument("--bias_prob", type=float, required=True, help="The decoder bias probability.")
    parser.add_argument("--num_experiments", type=int, required=True, help="The number of experiments to run.")
    parser.add_argument("--error_model", type=str, required=True, help="The error model to use.")
    parser.add_argument("--seed", type=int, required=True, help="The seed for the random number generator.")
    parser.add_argument("--num_processes", type=int, required=True,
                        help="Kept for CLI compatibility; GPUs are selected by the launcher (torchrun).")
    parser.add_argument("--silent", type=bool, required=True, help="Whether to silence the output.")
    parser.add_argument("--tolerance", type=float, required=True, help="The numerical tolerance for the MPS within the decoder.")
    parser.add_argument("--cut", type=float, required=True, help="Singular values smaller than that will be discarded in the SVD.")
    parser.add_argument("--output_dir", type=str, default=".", help="Where to write the pickle (extension).")
    return parser.parse_args(argv)


def generate_pauli_error_string(num_qubits, error_rate, error_model="Depolarising", rng=None):
    """Bitflip / Phaseflip / Depolarising error strings.  Bit-for-bit the reference's generator for the
    seeded models (decoding.py:977-1040); Depolarising draws the Pauli from `rng` too (the reference uses the
    unseeded global numpy RNG there, decoding.py:1018, which cannot be reproduced)."""
    rng = np.random.default_rng() if rng is None else rng
    chars = []
    for _ in range(num_qubits):
        if error_model == "Depolarising":
            chars.append("XYZ"[int(rng.integers(3))] if rng.random() < error_rate else "I")
        elif error_model == "Bitflip":
            chars.append("X" if rng.random() < error_rate else "I")
        elif error_model == "Phaseflip":
            chars.append("Z" if rng.random() < error_rate else "I")
        else:
            raise ValueError(f"Unsupported error model {error_model}.")
    return "".join(chars)


def generate_errors(lattice_size, error_rate, num_experiments, error_model, seed):
    """One spawned generator per experiment (reference quantum_surface.py:128-146)."""
    seed_seq = np.random.SeedSequence(seed)
    n = lattice_size ** 2 + (lattice_size - 1) ** 2
    return [generate_pauli_error_string(n, error_rate, error_model, np.random.default_rng(seed_seq.spawn(1)[0]))
            for _ in range(num_experiments)]


def run_experiment(lattice_size, chi_max, error_rate, bias_prob, num_experiments, error_model, seed, errors, silent,
                   num_processes=1, tolerance=1e-8, cut=1e-8):
    """Decode this rank's shard and assemble the reference's result dictionary (quantum_surface.py:207-268)."""
    from mdopt_b200 import _abi
    from mdopt_b200 import distributed as D
    from mdopt_b200.codes import surface_code
    from mdopt_b200.decoding import decode_css_batch

    import torch
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = D.shard_bounds(num_experiments, rank, world)
    code = surface_code(lattice_size)
    shard = errors[lo:hi]
    # bias_type = error_model, unchanged, like the reference (quantum_surface.py:157-169): in decode_css any string but
    # "Bitflip" takes the depolarising-bias branch (decoding.py:1267-1283), "Phaseflip" included
    kw = dict(chi_max=chi_max, cut=cut, bias_type=error_model, bias_prob=bias_prob, renormalise=True, silent=silent,
              contraction_strategy="Optimised", tolerance=tolerance)
    dist_rows, success, status = decode_css_batch(code, shard, return_status=True, **kw)
    dist_rows = np.array(dist_rows, dtype=np.float64)
    failure = (1 - success).astype(np.float64)
    # retry path (quantum_surface.py:170-192): the reference re-decodes a syndrome whose decode raised with
    # multiply_by_stabiliser=True and records NaN if that fails too.  Here "raised" is a per-syndrome status that
    # invalidates the result (capacity); a zero-norm state is not an exception in the reference either (NaN vector,
    # success 0, counted as a failure).
    redo = np.nonzero(status == _abi.ST_CAPACITY)[0]
    if redo.size:
        logging.info("Trying to decode %d syndromes with multiply_by_stabiliser=True.", redo.size)
        d2, s2, st2 = decode_css_batch(code, [shard[b] for b in redo], return_status=True,
                                       multiply_by_stabiliser=True, rng=np.random.default_rng([seed, rank]), **kw)
        dist_rows[redo], failure[redo] = d2, (1 - s2)
        lost = redo[st2 == _abi.ST_CAPACITY]
        dist_rows[lost], failure[lost] = np.nan, np.nan
        if lost.size:
            logging.error("Decoding has not been completed for %d syndromes.", lost.size)
    device = torch.device("cuda", torch.cuda.current_device())
    counts = D.reduce_counts(D.local_counts(dist_rows, success), device)
    all_rows = D.gather_rows(dist_rows, device)
    all_fail = D.gather_rows(failure.reshape(-1, 1), device).reshape(-1)
    if not silent and rank == 0:
        logging.info("decoded %d syndromes, %d successes, %d NaN", *[int(c) for c in counts])
    return {
        "logicals_distributions": [row for row in all_rows],
        "failures": [int(f) if f == f else float("nan") for f in all_fail],
        "lattice_size": lattice_size, "chi_max": chi_max, "error_rate": error_rate, "bias_prob": bias_prob,
        "error_model": error_model, "seed": seed, "tolerance": tolerance, "cut": cut,
    }


def save_experiment_data(data, lattice_size, chi_max, error_rate, error_model, bias_prob, num_experiments, seed,
                         tolerance, cut, output_dir="."):
    error_model = error_model.replace(" ", "")
    file_key = (f"latticesize{lattice_size}_bonddim{chi_max}_errorrate{error_rate}_errormodel{error_model}"
                f"_bias_prob{bias_prob}_numexperiments{num_experiments}_tolerance{tolerance}_cut{cut}_seed{seed}.pkl")
    path = os.path.join(output_dir, file_key)
    with open(path, "wb") as fh:
        pickle.dump(data, fh)
    fails = np.asarray(data["failures"], dtype=float)
    sem = float(np.nanstd(fails, ddof=1) / np.sqrt(np.sum(~np.isnan(fails)))) if fails.size > 1 else 0.0
    logging.info(f"Saved data for {file_key} with {np.nanmean(fails) * 100:.2f}±{sem * 100:.2f}% failure rate.")
    return path


def main(argv=None):
    import torch
    import torch.distributed as dist

    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    args = parse_arguments(argv)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    errors = generate_errors(args.lattice_size, args.error_rate, args.num_experiments, args.error_model, args.seed)
    data = run_experiment(args.lattice_size, args.bond_dim, args.error_rate, args.bias_prob, args.num_experiments,
                          args.error_model, args.seed, errors, args.silent, args.num_processes, args.tolerance,
                          args.cut)
    if not dist.is_initialized() or dist.get_rank() == 0:
        save_experiment_data(data, args.lattice_size, args.bond_dim, args.error_rate, args.error_model,
                             args.bias_prob, args.num_experiments, args.seed, args.tolerance, args.cut,
                             args.output_dir)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
