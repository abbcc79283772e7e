#!/usr/bin/env python
"""Stand-alone micro-benchmarks of the batched kernels behind the C ABI (K1 theta-build, K2 truncated SVD,
K3 read-out), each against the roofline that bounds it.  Prints one JSON object per kernel.

    python scripts/bench_kernels.py            # on a B200
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mdopt_b200 import _lib  # noqa: E402
from mdopt_b200.optimiser.utils import XOR_BULK  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def main():
    lib, _ = _lib.require_gpu()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    t = timed(lambda: torch.matmul(a, b), iters=3, warm=1)
    fp64_peak = 2.0 * n ** 3 / t / 1e12
    del a, b
    vp = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    out = []

    # K1: theta-build at chi=64, XOR_BULK (contractor.py:328-342); working set > L2
    chi, batch = 64, 6000
    rows, vl, vr, mr, bn = chi, 2, 2, chi, chi
    m_in = torch.randn(batch, rows, vl * mr, dtype=torch.float64, device="cuda")
    a_site = torch.randn(batch, mr, 2, bn, dtype=torch.float64, device="cuda")
    w = torch.from_numpy(np.ascontiguousarray(XOR_BULK)).cuda()
    theta = torch.empty(batch, rows, 2, vr, bn, dtype=torch.float64, device="cuda")
    t = timed(lambda: _lib.check(lib.mdopt_site_apply_batch_device(64, batch, rows, vl, vr, mr, bn, vp(m_in), vp(a_site),
                                                                   vp(w), vp(theta), C.c_void_p(0)), "site_apply"))
    nbytes = 8.0 * batch * (rows * vl * mr + mr * 2 * bn + rows * 2 * vr * bn)
    flops = 2.0 * batch * rows * bn * mr * 4  # XOR_BULK has 4 non-zeros
    # 16 FLOP/B is above the FP64 machine balance (35.5 TFLOP/s / 6.55 TB/s = 5.4 FLOP/B): FP64-bound
    out.append({"kernel": "K1 site_apply_mma_kernel<64> (theta-build, TMA-staged operands, FP64 mma tiles)",
                "batch": batch, "ms": t * 1e3,
                "roofline": {"bound": "tensor", "achieved": flops / t / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": flops / t / 1e12 / fp64_peak},
                "hbm": {"achieved_gbs": nbytes / t / 1e9, "peak_gbs": hbm_peak, "frac": nbytes / t / 1e9 / hbm_peak},
                "bytes_per_item": nbytes / batch, "flops_per_item": flops / batch})
    del m_in, a_site, theta

    # K2: truncated SVD, the reference's move shape 128x128 and chi_max=64 (46.1 MFLOP in LAPACK's count)
    for (m, nn_, lapack_flops) in [(128, 128, 46.1e6), (64, 64, 6 * 64 * 64 * 64 + 20 * 64 ** 3)]:
        batch = 1184
        mats = torch.randn(batch, m, nn_, dtype=torch.float64, device="cuda")
        k = min(m, nn_)
        u = torch.empty(batch, m, k, dtype=torch.float64, device="cuda")
        s = torch.empty(batch, k, dtype=torch.float64, device="cuda")
        vh = torch.empty(batch, k, nn_, dtype=torch.float64, device="cuda")
        kept = torch.empty(batch, dtype=torch.int32, device="cuda")
        trunc = torch.empty(batch, dtype=torch.float64, device="cuda")
        status = torch.empty(batch, dtype=torch.int32, device="cuda")
        cap = 64 if m > 64 else 32
        t = timed(lambda: _lib.check(lib.mdopt_svd_batch_device(cap, batch, m, nn_, vp(mats), 64, 1e-17, 0, vp(u), vp(s),
                                                                vp(vh), vp(kept), vp(trunc), vp(status),
                                                                C.c_void_p(0)), "svd"), iters=3, warm=1)
        out.append({"kernel": f"K2 svd_kernel ({m}x{nn_}, random dense, chi_max=64)", "batch": batch, "ms": t * 1e3,
                    "matrices_per_s": batch / t,
                    "roofline": {"bound": "tensor", "achieved": lapack_flops * batch / t / 1e12, "peak": fp64_peak,
                                 "unit": "TFLOP/s", "frac": lapack_flops * batch / t / 1e12 / fp64_peak,
                                 "note": "FLOPs in LAPACK's gesdd convention (SURVEY 8d); the Jacobi kernel executes more"},
                    "not_converged": int((status != 0).sum().item())})

    # K3: read-out
    batch = 1 << 24
    v = torch.randn(batch, 4, dtype=torch.float64, device="cuda")
    dist = torch.empty_like(v)
    succ = torch.empty(batch, dtype=torch.int32, device="cuda")
    t = timed(lambda: _lib.check(lib.mdopt_readout_batch_device(batch, 4, vp(v), 1, vp(dist), vp(succ), C.c_void_p(0)),
                                 "readout"))
    nbytes = batch * (32 + 32 + 4.0)
    out.append({"kernel": "K3 readout_kernel", "batch": batch, "ms": t * 1e3,
                "roofline": {"bound": "hbm", "achieved": nbytes / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": nbytes / t / 1e9 / hbm_peak}})
    for o in out:
        o["fp64_peak_tflops_measured"] = fp64_peak
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
