set -x
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 2500 gpurun_out/bench_default.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json | cut -c1-600
