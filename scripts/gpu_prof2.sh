set -x
python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/prof_decode2 python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log | cut -c1-300
