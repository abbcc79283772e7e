set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 2 --warmup 1 --batch 1024 --no-cpu-baseline
