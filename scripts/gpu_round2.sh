set -x
timeout 900 python -m pytest tests -x -q -m gpu --durations=6 2>&1 | tail -12
timeout 300 python bench.py --steps 3 --warmup 3 --batch 2072 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/iter_bench.json").read().strip().splitlines()[-1])
print("VALUE", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
PY
timeout 400 python scripts/gpu_config3_demo.py 7 128 148 > gpurun_out/demo_d7_chi128.json 2> gpurun_out/demo_d7.err; cat gpurun_out/demo_d7_chi128.json; tail -2 gpurun_out/demo_d7.err
