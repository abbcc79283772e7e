set -x
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 2 --warmup 1 --batch 1024 --no-cpu-baseline | cut -c1-400
timeout 200 python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/r01_decode_kernel python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_ncu.log 2>&1
timeout 200 python bench.py --steps 2 --warmup 1 --batch 296 --no-cpu-baseline > gpurun_out/launch_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --batch 296 --no-cpu-baseline > gpurun_out/launch_ncu.log 2>&1
wc -l gpurun_out/r01_launches.csv
