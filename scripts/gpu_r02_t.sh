#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/r02t_depol_counters.json 2>gpurun_out/r02t.err || exit 1
cat gpurun_out/r02t_depol_counters.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/r02t_dense \
    python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/r02t_ncu.log 2>&1
tail -3 gpurun_out/r02t_ncu.log
ls -la gpurun_out/
