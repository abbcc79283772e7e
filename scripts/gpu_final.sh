set -x
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
tail -c 600 gpurun_out/r01_bench.json
timeout 300 python bench.py --mode fast --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_fast.json 2> gpurun_out/r01_bench_fast.err
timeout 300 python bench.py --mode svd-gram --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_gram.json 2> gpurun_out/r01_bench_gram.err
timeout 200 python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/r01_decode_kernel_final python bench.py --steps 1 --warmup 1 --batch 168 --no-cpu-baseline > gpurun_out/prof_ncu.log 2>&1
timeout 200 python bench.py --steps 2 --warmup 1 --batch 296 --no-cpu-baseline > gpurun_out/launch_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --batch 296 --no-cpu-baseline > gpurun_out/launch_ncu.log 2>&1
timeout 300 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import mps_oracle as O
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch
errs = O.seeded_errors(41, 0.05, 20000, 777, "Bitflip")
t = time.time()
d, s, st = decode_css_batch(surface_code(5), errs, chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised", return_status=True)
dt = time.time() - t
print("BIG", len(errs), "syndromes", f"{dt:.1f}s", f"{len(errs)/dt:.0f}/s", "status!=0:", int((st != 0).sum()), "nan rows:", int(np.isnan(d).any(axis=1).sum()), "success rate:", float(s.mean()), "norm err:", float(np.abs(np.linalg.norm(d, axis=1) - 1).max()))
PY
