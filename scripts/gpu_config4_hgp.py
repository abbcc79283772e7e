"""BASELINE.json configs[3]: hypergraph-product code of a random (3,4)-regular LDPC code (n = 16, m = 12 as in
examples/decoding/quantum_hypergraph_product.py:134-142 -> 400 qubits, 32 logical sites, 832 MPS sites), seeded
bit-flip errors, chi_max = 512, Dephasing-DMRG read-out, on one B200 through decode_css_batch."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from mdopt_b200 import engine  # noqa: E402
from mdopt_b200.codes import hypergraph_product, random_regular_code  # noqa: E402
from mdopt_b200.decoding import decode_css_batch  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 888
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rate = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
classical = random_regular_code(16, 12, 3, 4, 123)
code = hypergraph_product(classical, classical)
ss = np.random.SeedSequence(123)
errors = []
for child in ss.spawn(batch):
    rng = np.random.default_rng(child)
    errors.append("".join("X" if rng.random() < rate else "I" for _ in range(len(code))))
kw = dict(chi_max=chi, cut=1e-17, bias_type="Bitflip", bias_prob=0.1, contraction_strategy="Optimised", return_status=True)
decode_css_batch(code, errors[:8], **kw)
torch.cuda.synchronize()
t0 = time.time()
bits, overlap, status = decode_css_batch(code, errors, **kw)
torch.cuda.synchronize()
dt = time.time() - t0
eng = next(iter(engine._ENGINES.values()))
c = eng.read_counters()
print(json.dumps({"config": f"hgp_ldpc16x12_bitflip_p{rate}_bias0.1_chi{chi}", "n_qubits": len(code),
                  "logical_sites": int(bits.shape[1]), "mps_sites": 2 * len(code) + int(bits.shape[1]), "batch": batch,
                  "nontrivial": int(sum(e != "I" * len(code) for e in errors)), "seconds": round(dt, 3),
                  "syndromes_per_s": round(batch / dt, 2), "status_not_ok": int((status != 0).sum()),
                  "overlap_mean": float(np.mean(overlap)), "resident_warps": eng.n_warps,
                  "two_site_steps": c["steps"], "truncating_steps": c["truncating_steps"],
                  "site_record_bytes": c["bytes"]}))
