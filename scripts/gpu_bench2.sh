timeout 300 python bench.py --steps 1 --warmup 1 --batch 340 --no-cpu-baseline
