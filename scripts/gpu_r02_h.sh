#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -q > gpurun_out/r02h_gputests.log 2>&1; tail -3 gpurun_out/r02h_gputests.log
timeout 400 python -m mdopt_b200.drivers.random_circuit --num_qubits 40 --bond_dim 256 --num_layers 16 --seed 1 --max_seconds 240 > gpurun_out/r02h_config5_chi256.json 2> gpurun_out/r02h_c5a.err; echo "c5a exit $?"; cat gpurun_out/r02h_config5_chi256.json; tail -2 gpurun_out/r02h_c5a.err
timeout 500 python -m mdopt_b200.drivers.random_circuit --num_qubits 40 --bond_dim 1024 --num_layers 20 --seed 1 --max_seconds 300 > gpurun_out/r02h_config5_chi1024.json 2> gpurun_out/r02h_c5b.err; echo "c5b exit $?"; cat gpurun_out/r02h_config5_chi1024.json; tail -2 gpurun_out/r02h_c5b.err
timeout 300 python scripts/mono_profile.py --batch 16384 --launches 1 > gpurun_out/r02h_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02h_mono \
    python scripts/mono_profile.py --batch 16384 --launches 1 > gpurun_out/r02h_ncu_full.log 2>&1
cat gpurun_out/r02h_profile_plain.log
