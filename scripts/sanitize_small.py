"""Small decodes through every new kernel path (mono engine at three capacities, DMRG read-out, dense fallback) for
compute-sanitizer memcheck."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
sys.path.insert(0, "tests")
import numpy as np

import mps_oracle as O
from conftest import load_golden
from mdopt_b200.codes import CssCode, surface_code
from mdopt_b200.decoding import decode_css_batch

errs = O.seeded_errors(13, 0.1, 48, 3, "Bitflip") + ["IIXIIZIIIIIII", "EIXIIIIIIIIII"]
for chi in (8, 64):
    d, s, st = decode_css_batch(surface_code(3), errs, chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                contraction_strategy="Optimised", return_status=True)
    print("d3 chi", chi, "status", np.bincount(st), "success", int(s.sum()))
e5 = [e for e in O.seeded_errors(41, 0.05, 40, 123, "Bitflip") if "X" in e][:24]
for chi in (64, 128):
    d, s, st = decode_css_batch(surface_code(5), e5, chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1,
                                contraction_strategy="Optimised", return_status=True)
    print("d5 chi", chi, "status", np.bincount(st), "success", int(s.sum()))
s7 = load_golden("dephasing_dmrg")["steane7"]
code = CssCode(49, s7["x_stabs"], s7["z_stabs"], s7["x_logicals"], s7["z_logicals"])
b, ov, st = decode_css_batch(code, [r["error"] for r in s7["decodes"]], contraction_strategy="Naive", return_status=True,
                             **s7["kwargs"])
print("steane7 overlaps", ov.tolist(), "status", st.tolist())
