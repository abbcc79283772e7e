export MDOPT_B200_LIB=$PWD/build/variants/lib$1.so
MDOPT_TEST_MODE=svd timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for mode in svd; do
timeout 400 python bench.py --mode $mode --steps 2 --warmup 2 --batch 2072 --no-cpu-baseline > gpurun_out/mode_$mode.json 2> gpurun_out/mode_$mode.err
python - $mode <<'PY'
import json, sys
m = sys.argv[1]
d = json.loads(open(f"gpurun_out/mode_{m}.json").read().strip().splitlines()[-1]); r = d["roofline"]; cs = r["cycle_split"]; t = cs["cyc_total"]
print("MODE", m, "value %.1f" % d["value"], {k[4:]: round(v / t, 3) for k, v in cs.items() if k != "cyc_total"}, "sweeps", r["jacobi_sweeps"], "steps", r["steps_per_launch"], d["counts"])
PY
done
