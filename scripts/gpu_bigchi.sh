set -x
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16
timeout 600 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import mps_oracle as O
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch
errs = [e for e in O.seeded_errors(41, 0.05, 400, 777, "Bitflip") if "X" in e]
for chi, nb in ((64, 148), (128, 148), (256, 148)):
    kw = dict(chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised", return_status=True)
    decode_css_batch(surface_code(5), errs[:2], **kw)
    torch.cuda.synchronize(); t = time.time()
    d, s, st = decode_css_batch(surface_code(5), errs[:nb], **kw)
    torch.cuda.synchronize(); dt = time.time() - t
    print("TIER chi", chi, nb, "syndromes", f"{dt:.2f}s", f"{nb/dt:.1f}/s", "status!=0:", int((st != 0).sum()), "success", float(s.mean()), flush=True)
PY
