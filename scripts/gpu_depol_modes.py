"""Depolarising noise at d=5, chi_max=64: throughput of the three modes (one B200)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, torch
import mps_oracle as O
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch
code = surface_code(5)
errs = [e for e in O.seeded_errors(41, 0.05, 2000, 99, "Depolarising") if set(e) != {"I"}][:1184]
res = {}
for mode in ("fast", "svd", "svd-gram"):
    kw = dict(chi_max=64, cut=1e-17, bias_type="Depolarising", bias_prob=0.1, contraction_strategy="Optimised", return_status=True, mode=mode)
    decode_css_batch(code, errs[:8], **kw)
    torch.cuda.synchronize(); t = time.time()
    d, s, st = decode_css_batch(code, errs, **kw)
    torch.cuda.synchronize(); dt = time.time() - t
    res[mode] = d
    print("DEPOL", mode, len(errs), "syndromes", f"{dt:.2f}s", f"{len(errs)/dt:.1f}/s", "status!=0:", int((st != 0).sum()), "success", float(s.mean()), flush=True)
print("max |svd - fast|", float(np.max(np.abs(res["svd"] - res["fast"]))), "agree flags", "n/a")
