#!/usr/bin/env python
"""Headline benchmark: decoded syndromes/s for the d=5 surface code at chi_max=64 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

One *step* decodes one batch of seeded bit-flip error strings (p = 0.05, decoder bias 0.1, cut 1e-17,
"Optimised" order) per GPU.  `value` is whole-job syndromes/s with the product-state symbols already
resident in HBM; `e2e` is the same batch through the public `decode_css_batch` call with host strings in
and host arrays out (symbol packing, H2D, kernel, D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LATTICE, CHI_MAX, ERROR_RATE, BIAS, CUT, SEED = 5, 64, 0.05, 0.1, 1e-17, 123
WORKLOAD = f"surface_d{LATTICE}_bitflip_p{ERROR_RATE}_bias{BIAS}_chi{CHI_MAX}_cut{CUT}_optimised"
REF_FLOP_PER_NONTRIVIAL = 47.98e9 + 1.70e9  # LAPACK-count of the reference's call sequence
# dram__bytes_read.sum + dram__bytes_write.sum of decode_kernel<64> from the `ncu --set full` capture summarised
# in profiles/r01_decode_kernel_ncu_summary.md: 12.23 GB for a launch of 143 non-trivial syndromes (default mode)
DRAM_BYTES_PER_NONTRIVIAL = 12.23e9 / 143


def seeded_errors(n_qubits, rate, count, seed, offset=0):
    """One child generator per syndrome (examples/decoding/quantum_surface.py:130-137), bit-flip model."""
    ss = np.random.SeedSequence(seed)
    out = []
    for child in ss.spawn(offset + count)[offset:]:
        rng = np.random.default_rng(child)
        out.append("".join("X" if rng.random() < rate else "I" for _ in range(n_qubits)))
    return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores (the reference itself is Python and
# does not travel to the GPU box; oracle/mps_oracle.py is pinned bit-for-bit against it).
# ------------------------------------------------------------------------------------------------------
def _cpu_init():
    """One BLAS thread per worker (the reference's cluster convention, quantum_surface_cc.sh:58).  The env
    var alone is not enough: the parent has usually initialised its BLAS pool before the fork."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits

        _cpu_init.limit = threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        pass


def _cpu_worker(err):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mps_oracle as O

    code = _cpu_worker.code if hasattr(_cpu_worker, "code") else O.surface_code(LATTICE)
    _cpu_worker.code = code
    dist, ok = O.decode_css(code, err, chi_max=CHI_MAX, cut=CUT, bias_type="Bitflip", bias_prob=BIAS,
                            contraction_strategy="Optimised", tolerance=1e-12)
    return int(ok)


def cpu_rate(errors, cores, budget_s=25.0):
    """Decode `errors` on `cores` worker processes for at most ~budget_s seconds of measured time; returns
    (syndromes/s, seconds, n_done, n_success).  Bounded so the default bench always ends within minutes."""
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"  # the reference's own cluster convention (quantum_surface_cc.sh:58)
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_init)
    try:
        warm = pool.map_async(_cpu_worker, errors[:cores])  # imports, BLAS init, first decode
        warm.get(timeout=180)
        t0 = time.perf_counter()
        done, oks = 0, 0
        it = pool.imap_unordered(_cpu_worker, errors, chunksize=1)
        while done < len(errors):
            remaining = budget_s - (time.perf_counter() - t0)
            try:
                oks += it.next(timeout=max(remaining, 0.01))
                done += 1
            except mp.TimeoutError:
                break
        dt = time.perf_counter() - t0
    finally:
        pool.terminate()
        pool.join()
    return done / dt, dt, done, oks


def cpu_sample_main():
    """`bench.py --cpu-sample`: the bounded CPU sample, in its own interpreter (no torch, no CUDA context, so
    forking the worker pool is safe).  Prints one JSON object."""
    cores = min(len(os.sched_getaffinity(0)), 64)
    n = LATTICE * LATTICE + (LATTICE - 1) ** 2
    sample = seeded_errors(n, ERROR_RATE, 2 * cores, SEED)
    r, dt, done, _ = cpu_rate(sample, cores)
    print(json.dumps({"value": r, "unit": "syndromes/s", "cores": cores, "kind": "port",
                      "sample": f"{done} of the first {len(sample)} seeded syndromes of the workload in {dt:.1f}s, "
                                f"oracle port of the reference algorithm, multiprocessing.Pool({cores}), "
                                "one BLAS thread per worker"}), flush=True)


def cpu_baseline_subprocess():
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample"], capture_output=True,
                             text=True, timeout=300)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": "syndromes/s", "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(len(os.sched_getaffinity(0)), 64)
    n = LATTICE * LATTICE + (LATTICE - 1) ** 2
    per_step = 2 * cores
    errors = seeded_errors(n, ERROR_RATE, per_step * (args.steps + args.warmup), SEED)
    rates, times = [], []
    for step in range(args.warmup + args.steps):
        chunk = errors[step * per_step:(step + 1) * per_step]
        r, dt, done, _ = cpu_rate(chunk, cores)
        if step >= args.warmup:
            rates.append(done)
            times.append(dt)
    total = sum(rates) / sum(times)
    line = {
        "impl": "reference", "metric": "decoded syndromes/sec", "value": total, "unit": "syndromes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": per_step, "seed": SEED, "done_per_step": rates},
        "cpu_baseline": {"value": total, "unit": "syndromes/s", "cores": cores, "kind": "port",
                         "sample": f"up to {per_step} seeded syndromes per step x {args.steps} steps (25 s cap "
                                   f"per step), multiprocessing.Pool({cores}), one BLAS thread per worker"},
        "e2e": {"value": total, "unit": "syndromes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag = gpu_index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for nm, val in zip(names, parts[2:6]):
                    if val.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (MEASURED_PEAKS.json, driver-written), else the profiling
    guide's fallback."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as fh:
            doc = json.load(fh)
        for key in ("hbm_gbs_burst", "hbm_gbs", "hbm_copy_gbs", "hbm_gb_s"):
            if key in doc:
                return float(doc[key]), f"MEASURED_PEAKS.json:{key}"
        for key, val in doc.items():
            if "hbm" in key.lower() and isinstance(val, (int, float)):
                return float(val), f"MEASURED_PEAKS.json:{key}"
    except (OSError, ValueError):
        pass
    return 6553.0, "fallback (B200_PROFILING.md / BASELINE.md: 6553 GB/s measured copy bandwidth)"


def roofline_head(achieved_tflops, fp64_peak, achieved_gbs):
    """The kernel is reported against whichever of its two ceilings it is closer to: the FP64 pipe (QR mode:
    Householder / Jacobi arithmetic) or HBM (SVD mode with the orthogonality fast path: every step reads one site
    tensor and writes one, with ~10 flops per matrix entry in between)."""
    hbm_peak, src = hbm_peak_gbs()
    f_flop = achieved_tflops / fp64_peak if fp64_peak else 0.0
    f_hbm = achieved_gbs / hbm_peak
    if f_hbm >= f_flop:
        return {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": f_hbm,
                "peak_source": src, "fp64": {"achieved_tflops": achieved_tflops, "peak_tflops": fp64_peak,
                                             "frac": f_flop}}
    return {"bound": "tensor", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": f_flop,
            "hbm": {"achieved_gbs": achieved_gbs, "peak_gbs": hbm_peak, "frac": f_hbm, "peak_source": src}}


def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 8192^3 burst (FP64 peak is not in MEASURED_PEAKS.json)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mdopt_b200 import _lib, engine, schedule
    from mdopt_b200.codes import surface_code
    from mdopt_b200.decoding import decode_css_batch

    _lib.require_gpu()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    code = surface_code(LATTICE)
    n = len(code)
    k = code.num_x_logicals() + code.num_z_logicals()
    n_sites = 2 * n + k
    B = args.batch
    # contiguous shard of the seeded error list per rank (quantum_surface.py's generator, rank-offset)
    errors = seeded_errors(n, ERROR_RATE, B, SEED, offset=rank * B)
    nontriv = [e for e in errors if e != "I" * n]
    assert all("Z" not in e and "Y" not in e for e in nontriv)
    prog = schedule.build_decode_program(code, True, False, True, "Optimised")
    eng = engine.get_engine(n_sites, k, CHI_MAX)
    syms = schedule.symbols_from_errors(nontriv, k)
    d_syms = torch.from_numpy(syms).to(device)
    d_dist = torch.empty((len(nontriv), 1 << k), dtype=torch.float64, device=device)
    d_succ = torch.empty(len(nontriv), dtype=torch.int32, device=device)
    d_stat = torch.empty(len(nontriv), dtype=torch.int32, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def step_device():
        eng.decode_device(prog, d_syms, cut=CUT, bias_p=BIAS, renormalise=True, mode=args.mode, d_dist=d_dist,
                          d_success=d_succ, d_status=d_stat)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_fp64_peak(torch, device) if rank == 0 else 0.0

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms, counters = [], None
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        counters = eng.read_counters()
    barrier()
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())

    # end-to-end through the public API: host strings in, host arrays out (pack + H2D + kernel + D2H)
    for _ in range(max(1, min(args.warmup, 2))):
        decode_css_batch(code, errors, chi_max=CHI_MAX, cut=CUT, bias_type="Bitflip", bias_prob=BIAS,
                         contraction_strategy="Optimised", mode=args.mode)
    barrier()
    e2e_s = []
    succ_host = None
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist_host, succ_host = decode_css_batch(code, errors, chi_max=CHI_MAX, cut=CUT, bias_type="Bitflip",
                                                bias_prob=BIAS, contraction_strategy="Optimised", mode=args.mode)
        torch.cuda.synchronize()
        e2e_s.append(time.perf_counter() - t0)
    barrier()
    t_e2e = torch.tensor([sum(e2e_s)], dtype=torch.float64, device=device)
    # the one collective of the path: per-shard counts [decoded, successes, NaN rows]
    counts = torch.tensor([len(errors), int(succ_host.sum()), int(np.isnan(dist_host).any(axis=1).sum())],
                          dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    status_bad = int((d_stat != 0).sum().item())

    if rank == 0:
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e_value = world * B * args.steps / float(t_e2e.item())
        flops = counters["flops_qr"] + counters["flops_jacobi"] + counters["flops_apply"]
        kernel_s = (sum(step_ms) / args.steps) * 1e-3
        achieved = flops / kernel_s / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess()
        line = {
            "metric": "decoded syndromes/sec", "value": value, "unit": "syndromes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "nontrivial_per_gpu": len(nontriv), "seed": SEED,
                       "mode": args.mode, "l2": "256 MB flush write between timed steps; per-step working set "
                       f"{eng.ws_bytes / 1e6:.0f} MB > 126 MB L2", "n_ctas": eng.n_ctas},
            "e2e": {"value": e2e_value, "unit": "syndromes/s", "h2d_bytes_per_step": int(syms.nbytes),
                    "d2h_bytes_per_step": int(len(nontriv) * ((1 << k) * 8 + 8))},
            "gpu_launches": args.steps,
            "roofline": roofline_head(achieved, fp64_peak, counters["bytes"] / kernel_s / 1e9) | {
                         "traffic": DRAM_BYTES_PER_NONTRIVIAL * len(nontriv),
                         "traffic_note": "bytes per launch, scaled from the ncu capture in profiles/ (85 MB per "
                                         "non-trivial syndrome; the algorithmic site-tensor bytes are a little higher, "
                                         "L2 absorbs the rest)",
                         "kernel": "decode_kernel<64>", "note": "achieved = the kernel's own counters per launch "
                         "(site-tensor bytes read + written, or FLOPs) / CUDA-event duration; FP64 peak = cuBLAS "
                         "DGEMM 8192^3 measured in this run", "flops_per_launch": flops,
                         "flops_split": {k2: counters[k2] for k2 in ("flops_qr", "flops_jacobi", "flops_apply")},
                         "algorithmic_bytes_per_launch": counters["bytes"],
                         "achieved_gbs": counters["bytes"] / kernel_s / 1e9,
                         "steps_per_launch": counters["steps"], "truncating_steps": counters["truncating_steps"],
                         "cycle_split": {k2: counters[k2] for k2 in counters if k2.startswith("cyc_")},
                         "qr_reflectors": counters["qr_reflectors"], "jacobi_sweeps": counters["jacobi_sweeps"],
                         "reference_algorithm_equiv_tflops": REF_FLOP_PER_NONTRIVIAL * len(nontriv) / kernel_s / 1e12},
            "clocks": sampler.summary(),
            "counts": {"decoded": int(counts[0]), "success": int(counts[1]), "nan": int(counts[2]),
                       "status_not_ok_rank0": status_bad},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="syndromes per GPU per step")
    ap.add_argument("--mode", default="svd", choices=["fast", "svd", "svd-gram"],
                    help="svd (default): the reference algorithm with the orthogonality fast path; fast: QR steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_sample:
        cpu_sample_main()
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
