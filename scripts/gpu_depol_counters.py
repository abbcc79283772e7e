"""Cycle split of the dense decode kernel on the depolarising workload (d=5, chi_max=64): the leader-thread clock
counters of decode_kernel<64> per phase, per two-site step."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import bench
from mdopt_b200 import engine
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
code = surface_code(L)
errors = bench.seeded_depolarising_errors(len(code), 0.05, batch, 123)
both = [e for e in errors if "X" in e and "Z" in e]       # one flag group: a single launch to read counters from
kw = dict(chi_max=chi, cut=1e-17, bias_type="Depolarising", bias_prob=0.1, contraction_strategy="Optimised", mode="svd",
          return_status=True)
decode_css_batch(code, both[:16], **kw)
torch.cuda.synchronize()
t0 = time.time()
d, s, st = decode_css_batch(code, both, **kw)
torch.cuda.synchronize()
dt = time.time() - t0
eng = next(iter(engine._ENGINES.values()))
c = eng.read_counters()
steps = c["steps"]
out = {"config": f"d{L}_depol_chi{chi}", "syndromes": len(both), "seconds": round(dt, 3), "rate": round(len(both) / dt, 1),
       "steps_per_syndrome": steps / len(both), "jacobi_sweeps_per_step": c["jacobi_sweeps"] / steps,
       "cycles_per_step": {k: round(v / steps) for k, v in c.items() if k.startswith("cyc_")},
       "flops_per_step": {k: round(c[k] / steps) for k in ("flops_qr", "flops_jacobi", "flops_apply")},
       "status_not_ok": int((st != 0).sum())}
if "--raw" in sys.argv:
    out["raw"] = {k: float(v) for k, v in c.items()}
print(json.dumps(out))
