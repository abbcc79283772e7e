"""One warm-up and N timed launches of the monomial decode kernel on a seeded d=5 / chi_max=64 batch; the command
the ncu captures under profiles/ are taken with (plain run first, then the same command under ncu)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mdopt_b200 import engine, schedule  # noqa: E402
from mdopt_b200.codes import surface_code  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--mode", default="mono")
ap.add_argument("--lattice", type=int, default=5)
ap.add_argument("--chi", type=int, default=64)
args = ap.parse_args()

code = surface_code(args.lattice)
n = len(code)
errors = [e for e in bench.seeded_errors(n, 0.05, args.batch, 123) if e != "I" * n]
prog = schedule.build_decode_program(code, True, False, True, "Optimised")
eng = engine.get_engine(prog.n_sites, prog.n_logical, args.chi)
d_sym = torch.from_numpy(schedule.symbols_from_errors(errors, prog.n_logical)).cuda()
times = []
for it in range(args.launches + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.decode_device(prog, d_sym, cut=1e-17, bias_p=0.1, renormalise=True, mode=args.mode)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
c = eng.read_counters()
st = out[2].cpu().numpy()
print(json.dumps({"nontrivial": len(errors), "ms": times, "rate_nontrivial_per_s": len(errors) / (min(times[1:]) * 1e-3),
                  "bytes_per_syndrome": c["bytes"] / len(errors), "steps_per_syndrome": c["steps"] / len(errors),
                  "status_not_ok": int((st != 0).sum()), "warps": getattr(eng, "n_warps", 0),
                  "success_rate": float(out[1].float().mean().item())}))
