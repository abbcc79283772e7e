timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --steps 1 --warmup 1 --batch 340 --no-cpu-baseline
