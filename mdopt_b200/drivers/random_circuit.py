"""Random brickwork-circuit MPS simulation on the GPU (mirror of examples/random_circuit/mps-rand-circ.py:40-90 with
examples/random_circuit/random_circuit.py:9-71): BASELINE.json configs[4].

Every layer of Haar-random two-qubit gates becomes an MPO (each gate split by QR into (1, 4, 2, 2) and (4, 1, 2, 2)),
is applied with ``mps_mpo_contract(renormalise=True, chi_max=1e4, cut=1e-12)`` and the state is compressed back with
``compress(chi_max=BOND_DIM, cut=1e-12, renormalise=True, strategy="svd")``; the product of the per-bond fidelities
``1 - truncation error`` is recorded per layer, as in the reference.  The state is complex128 and the MPO bond is 4,
so the contraction runs through the library tier of this package (``mdopt_b200/generic.py``: cuBLAS / cuSOLVER on the
device), not through the decode kernels.  The state is a single MPS: with N GPUs the driver runs N independent
circuit instances (replicas), there is nothing to shard.

    python -m mdopt_b200.drivers.random_circuit --num_qubits 40 --bond_dim 1024 --num_layers 20 --seed 1
"""
from __future__ import annotations

import argparse
import json
import time
from typing import List

import numpy as np


def create_mpo(gate: np.ndarray, phys_dim: int = 2) -> List[np.ndarray]:
    """Two-site MPO from a two-qubit gate by QR (random_circuit.py:9-47)."""
    if len(gate.shape) != 2:
        raise ValueError(
            f"The number of dimensions of the gate is {len(gate.shape)},"
            "but the number of dimensions expected is 2."
        )
    q, r = np.linalg.qr(gate)
    return [q.reshape((1, q.shape[1], phys_dim, phys_dim)), r.reshape((r.shape[0], 1, phys_dim, phys_dim))]


def create_mpo_from_layer(layer, phys_dim: int = 2) -> List[np.ndarray]:
    """MPO of a layer of two-site gates (random_circuit.py:50-71)."""
    mpo: List[np.ndarray] = []
    for gate in layer:
        mpo += create_mpo(gate=gate, phys_dim=phys_dim)
    return mpo


def run(num_qubits: int, bond_dim: int, num_layers: int, seed: int = 0, max_seconds: float = 1e9):
    """The reference's loop (mps-rand-circ.py:48-83).  Returns (state, fidelity tail per layer, seconds per layer)."""
    from scipy.stats import unitary_group

    from mdopt_b200.contractor.contractor import mps_mpo_contract
    from mdopt_b200.mps.canonical import CanonicalMPS

    rng = np.random.default_rng(seed)
    state = CanonicalMPS([np.array([1.0, 0.0], dtype=complex).reshape(1, 2, 1) for _ in range(num_qubits)], 0,
                         tolerance=np.inf)
    tails, seconds = [], []
    t_start = time.time()
    for k in range(num_layers):
        t0 = time.time()
        layer = unitary_group.rvs(4, size=num_qubits // 2 - k % 2, random_state=rng)
        if layer.ndim == 2:
            layer = layer[None]
        mpo = create_mpo_from_layer(layer)
        state = mps_mpo_contract(mps=state, mpo=mpo, start_site=k % 2, renormalise=True, chi_max=int(1e4),
                                 cut=1e-12, inplace=False)
        state, errors = state.compress(chi_max=bond_dim, cut=1e-12, renormalise=True, strategy="svd",
                                       return_truncation_errors=True)
        tails.append(float(np.prod([1 - e for e in errors])))
        seconds.append(time.time() - t0)
        if time.time() - t_start > max_seconds:
            break
    return state, tails, seconds


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_qubits", type=int, default=40)
    ap.add_argument("--bond_dim", type=int, default=1024)
    ap.add_argument("--num_layers", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max_seconds", type=float, default=1e9)
    args = ap.parse_args(argv)
    state, tails, seconds = run(args.num_qubits, args.bond_dim, args.num_layers, args.seed, args.max_seconds)
    print(json.dumps({"config": f"random_circuit_q{args.num_qubits}_chi{args.bond_dim}_layers{args.num_layers}",
                      "layers_done": len(tails), "max_bond": int(max(state.bond_dimensions)),
                      "seconds_per_layer": [round(s, 3) for s in seconds], "fidelity_tail": tails,
                      "dtype": str(state.tensors[0].dtype)}))


if __name__ == "__main__":
    main()
