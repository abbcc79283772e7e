set -x
nproc
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
python bench.py --steps 2 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --batch 512 --no-cpu-baseline > gpurun_out/launch_ncu.log 2>&1
wc -l gpurun_out/launches.csv
