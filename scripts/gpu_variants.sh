# bench several builds of the chi=64 kernels back to back: scripts/gpu_variants.sh NAME...
for v in "$@"; do
  export MDOPT_B200_LIB=$PWD/build/variants/lib$v.so
  timeout 200 python -m pytest tests/test_gpu_decode.py -x -q -m gpu -k "non_truncating or fresh_seeds or truncating_configs" 2>&1 | tail -1
  timeout 300 python bench.py --steps 3 --warmup 3 --batch 2072 --no-cpu-baseline > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open(f"gpurun_out/var_{v}.json").read().strip().splitlines()[-1])
r = d["roofline"]; cs = r["cycle_split"]; tot = cs["cyc_total"]
print("VARIANT", v, "value %.1f" % d["value"], "cyc/reflector %.0f" % (cs["cyc_qr"] / r["qr_reflectors"]),
      {k[4:]: round(x / tot, 3) for k, x in cs.items() if k not in ("cyc_total",)}, flush=True)
PY
done
