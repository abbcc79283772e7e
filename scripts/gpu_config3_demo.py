"""One batch of BASELINE.json configs[2]-shaped work: d=9 surface code, depolarising noise, chi_max=256, on one
B200.  Sanity only (the oracle needs hours per syndrome at this size): unit norms, status codes, throughput."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import torch  # noqa: E402

import mps_oracle as O  # noqa: E402  (seeded error generator only)
from mdopt_b200.codes import surface_code  # noqa: E402
from mdopt_b200.decoding import decode_css_batch  # noqa: E402

d, chi, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model = sys.argv[4] if len(sys.argv) > 4 else "Depolarising"
mode = sys.argv[5] if len(sys.argv) > 5 else "auto"
code = surface_code(d)
errs = [e for e in O.seeded_errors(len(code), 0.05, 4 * batch, 2024, model) if set(e) != {"I"}][:batch]
kw = dict(chi_max=chi, cut=1e-17, bias_type=model, bias_prob=0.1, contraction_strategy="Optimised",
          return_status=True, mode=mode)
torch.cuda.synchronize()
t0 = time.time()
dist, success, status = decode_css_batch(code, errs, **kw)
torch.cuda.synchronize()
dt = time.time() - t0
out = {"config": f"surface_d{d}_{model.lower()}_p0.05_bias0.1_chi{chi}", "n_qubits": len(code), "n_sites": 2 * len(code) + 2,
       "mode": mode, "batch": len(errs), "seconds": round(dt, 2), "syndromes_per_s": round(len(errs) / dt, 4),
       "status_not_ok": int((status != 0).sum()), "nan_rows": int(np.isnan(dist).any(axis=1).sum()),
       "success_rate": float(success.mean()), "max_norm_error": float(np.abs(np.linalg.norm(dist, axis=1) - 1).max())}
print(json.dumps(out))
