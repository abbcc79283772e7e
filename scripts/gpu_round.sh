#!/bin/bash
# The GPU pass of a round, as run through `gpurun --timeout 3300 -- 'bash scripts/gpu_round.sh'` (one B200):
# the GPU test tier, the default bench line, the launch list of the same command, one full ncu capture of each
# decode kernel (each only after the same command exited 0 without ncu).  Outputs land in gpurun_out/; what is judged
# is copied into profiles/ by hand.  Set SKIP_NCU=1 for tests + bench only.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/round_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/round_gputests.log
tail -5 gpurun_out/round_gputests.log
# the measured agreement figures the parity section of DESIGN.md quotes
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_mono.py -m gpu -q -s -k "depolarising_d5 or truncating_configs or headline_config or stock" 2>&1 | grep -i "deviation\|agreement\|mismatch\|max rel\|flags" | cut -c1-300 > gpurun_out/round_parity_figures.log; cat gpurun_out/round_parity_figures.log
timeout 900 python bench.py > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err; echo "bench exit $?"
tail -c 2500 gpurun_out/round_bench.json; tail -3 gpurun_out/round_bench.err
timeout 600 python bench.py --workload config3 --lattice 5 --chi 64 --batch 1184 --steps 2 --warmup 1 > gpurun_out/round_config3_d5_chi64.json 2>> gpurun_out/round_bench.err
cut -c1-160 gpurun_out/round_config3_d5_chi64.json
python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/round_depol_cycle_split.json 2>> gpurun_out/round_bench.err; cut -c1-400 gpurun_out/round_depol_cycle_split.json
[ -n "$SKIP_NCU" ] && exit 0
# launch list of the bench command (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/round_bench_short.json 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/round_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/round_ncu_launches.log 2>&1
# headline kernel (monomial warp engine)
timeout 300 python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/round_mono_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/round_mono \
    python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/round_ncu_mono.log 2>&1
# dense engine on the depolarising workload (the plain run is the counters line above)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/round_dense \
    python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/round_ncu_dense.log 2>&1
ls -la gpurun_out | tail -12
