timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py --steps 2 --warmup 2 --batch 2072 --no-cpu-baseline > gpurun_out/mode_svd.json 2> gpurun_out/mode_svd.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/mode_svd.json").read().strip().splitlines()[-1]); r = d["roofline"]; cs = r["cycle_split"]; t = cs["cyc_total"]
print("MAIN value %.1f" % d["value"], {k[4:]: round(v / t, 3) for k, v in cs.items() if k != "cyc_total"}, d["counts"])
PY
