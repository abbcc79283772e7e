import subprocess, sys, os
sys.path.insert(0, "oracle")
import mps_oracle as O
errors = O.seeded_errors(13, 0.15, 64, 99, "Bitflip")
code = r'''
import sys
sys.path.insert(0, ".")
from mdopt_b200.codes import surface_code
from mdopt_b200.decoding import decode_css_batch
e = sys.argv[1]; mode = sys.argv[2]
d, s = decode_css_batch(surface_code(3), [e], chi_max=64, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised", mode=mode)
print("OK", e, mode, d[0], s[0])
'''
open("/tmp/one.py", "w").write(code)
procs = []
for mode in ("fast", "svd"):
    for e in sorted(set(errors)):
        procs.append((e, mode, subprocess.Popen([sys.executable, "/tmp/one.py", e, mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        if len(procs) >= 8:
            for (e2, m2, p) in procs:
                out, err = p.communicate()
                print(out.strip() if p.returncode == 0 else f"FAIL {e2} {m2} :: {err.strip().splitlines()[-1]}", flush=True)
            procs = []
for (e2, m2, p) in procs:
    out, err = p.communicate()
    print(out.strip() if p.returncode == 0 else f"FAIL {e2} {m2} :: {err.strip().splitlines()[-1]}", flush=True)
