set -x
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -12
bash scripts/gpu_variants.sh main redux 2>&1 | grep -v "^+"
