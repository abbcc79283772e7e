#!/bin/bash
# 8-GPU scaling points (weak: 100k per GPU; strong: one 100k batch sharded), torchrun like the driver
set -x
mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02f_bench_n$n.json 2> gpurun_out/r02f_bench_n$n.err; echo "n$n exit $?"
head -c 400 gpurun_out/r02f_bench_n$n.json; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong > gpurun_out/r02f_bench_n8_strong.json 2> gpurun_out/r02f_bench_n8_strong.err; echo "strong exit $?"
head -c 400 gpurun_out/r02f_bench_n8_strong.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 -m pytest tests/test_gpu_decode.py -q -k "distributed or multi" > gpurun_out/r02f_dist_tests.log 2>&1; tail -3 gpurun_out/r02f_dist_tests.log
