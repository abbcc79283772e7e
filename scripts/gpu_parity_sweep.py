"""Wide parity sweep on the GPU box: d=3, many seeded errors of every kind, default mode, against the oracle."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import mps_oracle as O
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch

code, ocode = surface_code(3), O.surface_code(3)
rng = np.random.default_rng(4242)
worst = {}
for bias_type, model, n in (("Bitflip", "Bitflip", 400), ("Bitflip", "Depolarising", 300), ("Depolarising", "Depolarising", 200)):
    errs = O.seeded_errors(13, 0.12, n, int(rng.integers(1 << 30)), model)
    # sprinkle erasures into a few
    errs = [("E" + e[1:]) if (i % 17 == 0 and set(e) != {"I"}) else e for i, e in enumerate(errs)]
    kw = dict(chi_max=64, cut=1e-17, bias_type=bias_type, bias_prob=0.1, contraction_strategy="Optimised", tolerance=1e-12)
    t = time.time()
    dist, succ, st = decode_css_batch(code, errs, return_status=True, **kw)
    dt = time.time() - t
    bad_flag = 0; mx = 0.0; nexact = 0; ntr = 0
    for b, e in enumerate(errs):
        ref, ok = O.decode_css(ocode, e, **kw)
        if set(e) == {"I"}:
            continue
        if O.decode_truncates(ocode, e, **kw):
            ntr += 1
            bad_flag += int(succ[b]) != ok
            continue
        nexact += 1
        bad_flag += int(succ[b]) != ok
        mx = max(mx, float(np.max(np.abs(dist[b] - np.asarray(ref, float)) / np.max(np.asarray(ref, float)))))
    print(f"SWEEP bias={bias_type} errors={model}: {len(errs)} syndromes, gpu {dt:.2f}s, status!=0 {int((st != 0).sum())}, "
          f"non-truncating {nexact}: max rel err {mx:.2e}; truncating {ntr}; flag mismatches {bad_flag}", flush=True)
