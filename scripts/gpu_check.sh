set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu 2>&1 | tail -30
