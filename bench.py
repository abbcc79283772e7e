#!/usr/bin/env python
"""Headline benchmark: decoded syndromes/s for the d=5 surface code at chi_max=64 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

One *step* decodes one batch of seeded bit-flip error strings (p = 0.05, decoder bias 0.1, cut 1e-17,
"Optimised" order) per GPU: by default the 100 000-syndrome batch BASELINE.json configs[1] names, one such shard per
GPU (weak scaling); `--scaling strong` shards ONE 100 000-syndrome batch over the GPUs instead.  `value` is whole-job
syndromes/s with the product-state symbols already resident in HBM; `e2e` is the same batch through the public
`decode_css_batch` call with host strings in and host arrays out (string packing, H2D, kernel, D2H inside the timed
region).  The CPU arm times the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) on the host cores;
the same run compares the GPU results of the CPU-sample syndromes with it (`parity`).
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
CONFIG2_BATCH = 100_000


def seeded_errors(n_qubits, rate, count, seed, offset=0):
    """One child generator per syndrome (examples/decoding/quantum_surface.py:130-137), bit-flip model."""
    ss = np.random.SeedSequence(seed, n_children_spawned=offset)
    table = np.array(["I", "X"])
    out = []
    for child in ss.spawn(count):
        # rng.random(n) draws the same doubles as n successive rng.random() calls (decoding.py:1018-1022)
        out.append("".join(table[(np.random.default_rng(child).random(n_qubits) < rate).astype(np.intp)]))
    return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) on the host cores, one process per core with
# one BLAS thread each (the reference's own cluster convention, quantum_surface_cc.sh:58).  If oracle/_ref did not
# travel, the numpy port of the same algorithm (oracle/mps_oracle.py, pinned against the reference) is timed instead
# and the line says kind = "port".
# ------------------------------------------------------------------------------------------------------
def _cpu_init():
    """One BLAS thread per worker.  The env var alone is not enough: the parent has usually initialised its BLAS
    pool before the fork."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits

        _cpu_init.limit = threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        pass


def _cpu_decoder():
    """(kind, decode(err, canonical) -> (dist, ok)) bound to the reference if present, else to the port."""
    if hasattr(_cpu_decoder, "cached"):
        return _cpu_decoder.cached
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref
    import structured_svd as S

    # the config-3 sample (`--workload config3 --lattice L --chi X`, L <= 5) travels to the sub-interpreter in the
    # environment: depolarising bias, its lattice and chi_max
    depol = os.environ.get("MDOPT_BENCH_CPU_BIAS", "Bitflip") == "Depolarising"
    lattice = int(os.environ.get("MDOPT_BENCH_CPU_LATTICE", LATTICE))
    chi = int(os.environ.get("MDOPT_BENCH_CPU_CHI", CHI_MAX))
    kw = dict(chi_max=chi, cut=CUT, bias_type="Depolarising" if depol else "Bitflip", bias_prob=BIAS,
              contraction_strategy="Optimised", tolerance=1e-12)
    ref = make_ref.import_reference()
    if ref is not None:
        decode_css, qec = ref
        code = qec.hypergraph_product(qec.repetition_code(lattice), qec.repetition_code(lattice))
        kind, fn = "reference", (lambda e: decode_css(code, e, silent=True, **kw))
    else:
        import mps_oracle as O

        code = O.surface_code(lattice)
        kind, fn = "port", (lambda e: O.decode_css(code, e, **kw))

    def decode(err, canonical=False):
        if canonical:   # the same decode with the canonical tie-break at the SVD hook (parity leg, never timed)
            with S.installed():
                d, ok = fn(err)
        else:
            d, ok = fn(err)
        return [float(x) for x in d], int(ok)

    _cpu_decoder.cached = (kind, decode)
    return _cpu_decoder.cached


def _cpu_worker(err):
    return (err,) + _cpu_decoder()[1](err)


def _cpu_worker_canonical(err):
    return (err,) + _cpu_decoder()[1](err, canonical=True)


def cpu_rate(errors, cores, budget_s=25.0, parity=False):
    """Decode `errors` on `cores` worker processes for at most ~budget_s seconds of measured time; returns
    (syndromes/s, seconds, n_done, rows) with rows = [(error, dist, ok)].  Bounded so the default bench always ends
    within minutes.  With `parity`, the finished syndromes are decoded once more (untimed) with the canonical
    tie-break installed, for the in-run parity check."""
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_init)
    rows, canon = [], []
    try:
        warm = pool.map_async(_cpu_worker, errors[:cores])  # imports, BLAS init, first decode
        warm.get(timeout=300)
        t0 = time.perf_counter()
        it = pool.imap_unordered(_cpu_worker, errors, chunksize=1)
        while len(rows) < len(errors):
            remaining = budget_s - (time.perf_counter() - t0)
            try:
                rows.append(it.next(timeout=max(remaining, 0.01)))
            except mp.TimeoutError:
                break
        dt = time.perf_counter() - t0
        if parity and rows:
            try:
                canon = pool.map_async(_cpu_worker_canonical, [r[0] for r in rows]).get(timeout=120)
            except Exception:  # noqa: BLE001
                canon = []
    finally:
        pool.terminate()
        pool.join()
    return len(rows) / dt, dt, len(rows), rows, canon


def cpu_sample_main():
    """`bench.py --cpu-sample`: the bounded CPU sample, in its own interpreter (no torch, no CUDA context, so
    forking the worker pool is safe).  Prints one JSON object."""
    cores = min(len(os.sched_getaffinity(0)), 64)
    depol = os.environ.get("MDOPT_BENCH_CPU_BIAS", "Bitflip") == "Depolarising"
    lattice = int(os.environ.get("MDOPT_BENCH_CPU_LATTICE", LATTICE))
    n = lattice * lattice + (lattice - 1) ** 2
    sample = (seeded_depolarising_errors if depol else seeded_errors)(n, ERROR_RATE, 2 * cores, SEED)
    kind = _cpu_decoder()[0]
    # (the canonical tie-break hook states the monomial structure of the bit-flip workload; not asked for config 3)
    r, dt, done, rows, canon = cpu_rate(sample, cores, parity=not depol)
    print(json.dumps({"value": r, "unit": "syndromes/s", "cores": cores, "kind": kind,
                      "sample": f"{done} of the first {len(sample)} seeded syndromes of the workload in {dt:.1f}s, "
                                + ("the unmodified reference (oracle/_ref)" if kind == "reference" else
                                   "numpy port of the reference algorithm")
                                + f", multiprocessing.Pool({cores}), one BLAS thread per worker",
                      "rows": rows, "rows_canonical": canon}), flush=True)


def cpu_baseline_subprocess(env=None):
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample"], capture_output=True,
                             text=True, timeout=600, env=dict(os.environ, **(env or {})))
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": "syndromes/s", "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}


def parity_report(rows, gpu_lookup):
    """Compare CPU rows [(error, dist, ok)] with the GPU results of the same error strings."""
    if not rows:
        return None
    mism, worst = 0, 0.0
    for err, dist, ok in rows:
        g_dist, g_ok = gpu_lookup(err)
        mism += int(int(ok) != int(g_ok))
        d = np.asarray(dist, dtype=float)
        worst = max(worst, float(np.max(np.abs(d - g_dist)) / max(np.max(d), 1e-300)))
    return {"n": len(rows), "flag_mismatch": mism, "max_rel": worst}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(len(os.sched_getaffinity(0)), 64)
    n = LATTICE * LATTICE + (LATTICE - 1) ** 2
    per_step = 2 * cores
    errors = seeded_errors(n, ERROR_RATE, per_step * (args.steps + args.warmup), SEED)
    kind = _cpu_decoder()[0]
    rates, times = [], []
    for step in range(args.warmup + args.steps):
        chunk = errors[step * per_step:(step + 1) * per_step]
        r, dt, done, _, _ = cpu_rate(chunk, cores)
        if step >= args.warmup:
            rates.append(done)
            times.append(dt)
    total = sum(rates) / sum(times)
    line = {
        "impl": "reference", "metric": "decoded syndromes/sec", "value": total, "unit": "syndromes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": per_step, "seed": SEED, "done_per_step": rates},
        "cpu_baseline": {"value": total, "unit": "syndromes/s", "cores": cores, "kind": kind,
                         "sample": f"up to {per_step} seeded syndromes per step x {args.steps} steps (25 s cap "
                                   f"per step), " + ("the unmodified reference from oracle/_ref" if kind == "reference"
                                                     else "numpy port of the reference algorithm")
                                   + f", multiprocessing.Pool({cores}), one BLAS thread per worker"},
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


def traffic_per_launch(kernel, n_nontrivial):
    """DRAM bytes per launch of the dominant kernel from the `ncu --set full` capture of this round, if one has been
    summarised under profiles/ (bytes per non-trivial syndrome x the syndromes of a launch); None otherwise -- the
    bench never invents the number."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            doc = json.load(fh)
        ent = doc.get(kernel)
        if ent:
            return ent["dram_bytes_per_nontrivial_syndrome"] * n_nontrivial, ent.get("source")
    except (OSError, ValueError, KeyError):
        pass
    return None, None


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
    # contiguous shard of the seeded error list per rank (quantum_surface.py's generator, rank-offset)
    if args.scaling == "strong":
        from mdopt_b200.distributed import shard_bounds
        lo, hi = shard_bounds(args.batch, rank, world)
        B, total = hi - lo, args.batch
        errors = seeded_errors(n, ERROR_RATE, B, SEED, offset=lo)
    else:
        B, total = args.batch, args.batch * world
        errors = seeded_errors(n, ERROR_RATE, B, SEED, offset=rank * B)
    mode = args.mode
    mono = mode == "mono"
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
        eng.decode_device(prog, d_syms, cut=CUT, bias_p=BIAS, renormalise=True, mode=mode, d_dist=d_dist,
                          d_success=d_succ, d_status=d_stat)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_fp64_peak(torch, device) if (rank == 0 and not mono) else 0.0

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
    status_dev = d_stat.cpu().numpy()
    gpu_dist_dev = d_dist.cpu().numpy()
    gpu_succ_dev = d_succ.cpu().numpy()

    # end-to-end through the public API: host strings in, host arrays out (pack + H2D + kernel + D2H)
    kw = dict(chi_max=CHI_MAX, cut=CUT, bias_type="Bitflip", bias_prob=BIAS, contraction_strategy="Optimised",
              mode=mode)
    for _ in range(max(1, min(args.warmup, 2))):
        decode_css_batch(code, errors, **kw)
    barrier()
    e2e_s = []
    succ_host = dist_host = None
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist_host, succ_host, stat_host = decode_css_batch(code, errors, return_status=True, **kw)
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
    status_bad = int((status_dev != 0).sum()) + int((stat_host != 0).sum())
    # the device-resident run and the host-API run decode the same strings: they must agree bit for bit
    nt_idx = np.array([i for i, e in enumerate(errors) if e != "I" * n])
    same_paths = bool(np.array_equal(gpu_dist_dev, dist_host[nt_idx]) and np.array_equal(gpu_succ_dev, succ_host[nt_idx]))

    if rank == 0:
        value = total * args.steps / (total_ms * 1e-3)
        e2e_value = total * args.steps / float(t_e2e.item())
        kernel_s = (sum(step_ms) / args.steps) * 1e-3
        # chi = 64 launches the lock-step variant (28 warps per CTA) unless MDOPT_MONO_LOCKSTEP=0 selects the
        # free-running one-warp CTAs (mono_kernels.cuh launch_mono_decode)
        lockstep = os.environ.get("MDOPT_MONO_LOCKSTEP", "28")
        mono_kernel = "mono_decode_kernel<64>" if lockstep == "0" else f"mono_decode_lockstep_kernel<64, {lockstep}>"
        kernel = mono_kernel if mono else "decode_kernel<64>"
        hbm_peak, hbm_src = hbm_peak_gbs()
        if mono:
            achieved_gbs = counters["bytes"] / kernel_s / 1e9
            roof = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                    "note": "achieved = site records read + written per launch (10 bytes per (value, target) pair, "
                            "counted by the kernel itself) / CUDA-event duration of the launch.  The kernel is bound "
                            "by instruction issue, not by HBM (profiles/r02_mono_decode_kernel_ncu_summary.md: ~1 850 "
                            "warp instructions and ~2.7 KB of records per two-site step), so this fraction is small "
                            "by construction"}
        else:
            flops = counters["flops_qr"] + counters["flops_jacobi"] + counters["flops_apply"]
            roof = roofline_head(flops / kernel_s / 1e12, fp64_peak, counters["bytes"] / kernel_s / 1e9)
            roof["flops_per_launch"] = flops
        traffic, traffic_src = traffic_per_launch(kernel, len(nontriv))
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": kernel,
                     "algorithmic_bytes_per_launch": counters["bytes"],
                     "algorithmic_bytes_per_nontrivial_syndrome": counters["bytes"] / max(len(nontriv), 1),
                     "steps_per_launch": counters["steps"], "truncating_steps_per_launch": counters["truncating_steps"],
                     "two_site_steps_per_s": counters["steps"] / kernel_s,
                     "reference_algorithm_equiv_tflops": REF_FLOP_PER_NONTRIVIAL * len(nontriv) / kernel_s / 1e12})
        cpu = parity = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess()
            lookup_idx = {e: i for i, e in reversed(list(enumerate(errors)))}

            def gpu_lookup(err):
                i = lookup_idx[err]
                return dist_host[i], succ_host[i]

            rows, canon = cpu.pop("rows", []), cpu.pop("rows_canonical", [])
            parity = {"vs_cpu_arm": parity_report(rows, gpu_lookup),
                      "vs_cpu_arm_with_canonical_tie_break": parity_report(canon, gpu_lookup),
                      "device_resident_equals_host_api": same_paths,
                      "note": "vs_cpu_arm: the timed CPU decodes (LAPACK tie-breaks; marginals differ inside the "
                              "degenerate-truncation spread, DESIGN.md section 6); with the canonical tie-break at the "
                              "SVD hook (oracle/structured_svd.py) the comparison is well-posed: 1e-10 is the bar"}
        n_warps = getattr(eng, "n_warps", 0)
        line = {
            "metric": "decoded syndromes/sec", "value": value, "unit": "syndromes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "batch_total": total,
                       "nontrivial_per_gpu": len(nontriv), "seed": SEED, "mode": mode,
                       "l2": "256 MB flush write between timed steps; the resident MPS records ("
                             + (f"{eng.mono_ws_bytes / 1e6:.0f}" if mono else f"{eng.ws_bytes / 1e6:.0f}")
                             + " MB) exceed the 126 MB L2",
                       "resident": {"warps": n_warps} if mono else {"ctas": eng.n_ctas}},
            "nontrivial_rate": value * len(nontriv) / max(B, 1),
            "e2e": {"value": e2e_value, "unit": "syndromes/s", "h2d_bytes_per_step": int(syms.nbytes),
                    "d2h_bytes_per_step": int(len(nontriv) * ((1 << k) * 8 + 8))},
            # kernels of ours inside the timed region: per step the table set-up kernel + the decode kernel (mono),
            # or the decode kernel alone (dense modes)
            "gpu_launches": args.steps * (2 if mono else 1),
            "roofline": roof,
            "clocks": sampler.summary(),
            "counts": {"decoded": int(counts[0]), "success": int(counts[1]), "nan": int(counts[2]),
                       "status_not_ok_rank0": status_bad,
                       "dense_fallbacks_rank0": int(getattr(eng, "last_mono_fallbacks", 0))},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if parity:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def seeded_depolarising_errors(n_qubits, rate, count, seed, offset=0):
    """Depolarising error strings from seeded per-syndrome generators (the reference draws the Pauli
    from the unseeded global RNG, decoding.py:1018, so the letter comes from the syndrome's own generator here)."""
    ss = np.random.SeedSequence(seed, n_children_spawned=offset)
    out = []
    for child in ss.spawn(count):
        rng = np.random.default_rng(child)
        out.append("".join("XYZ"[int(rng.integers(3))] if rng.random() < rate else "I" for _ in range(n_qubits)))
    return out


def run_config3(args):
    """`--workload config3` (BASELINE.json configs[2]): d = 9 surface code, depolarising noise AND depolarising bias,
    chi_max = 256 -- the rotation-bound regime, on the dense CTA engine (the two-site bias gates leave the monomial
    structure).  One step decodes `--batch` syndromes (default: one per SM); the CPU reference needs hours per
    syndrome at this size, so no CPU sample is taken."""
    import torch
    import torch.distributed as dist

    from mdopt_b200 import _lib, engine
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
    L, chi = args.lattice, args.chi
    code = surface_code(L)
    n = len(code)
    B = args.batch
    errors = seeded_depolarising_errors(n, ERROR_RATE, B, SEED, offset=rank * B)
    kw = dict(chi_max=chi, cut=CUT, bias_type="Depolarising", bias_prob=BIAS, contraction_strategy="Optimised",
              mode="svd", return_status=True)
    fp64_peak = measure_fp64_peak(torch, device) if rank == 0 else 0.0
    for _ in range(args.warmup):
        decode_css_batch(code, errors, **kw)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times, flops, nbytes, steps = [], 0.0, 0.0, 0.0
    for _ in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d, s, st = decode_css_batch(code, errors, **kw)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        for eng in engine._ENGINES.values():   # counters of the last launch (one flag group per launch)
            c = eng.read_counters()
            flops += c["flops_qr"] + c["flops_jacobi"] + c["flops_apply"]
            nbytes += c["bytes"]
            steps += c["steps"]
    t = torch.tensor([sum(times)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        total = B * world * args.steps
        value = total / float(t.item())
        hbm_peak, src = hbm_peak_gbs()
        cpu, parity = None, None
        if L <= 5 and not args.no_cpu_baseline:
            # at d <= 5 the unmodified reference decodes a depolarising syndrome in seconds: bounded sample of the same
            # seeded workload on the host cores (own interpreter), and the GPU results of exactly those strings
            cpu = cpu_baseline_subprocess({"MDOPT_BENCH_CPU_BIAS": "Depolarising", "MDOPT_BENCH_CPU_LATTICE": str(L),
                                           "MDOPT_BENCH_CPU_CHI": str(chi)})
            rows = cpu.pop("rows", []) or []
            cpu.pop("rows_canonical", None)
            if rows:
                gd, gs, _ = decode_css_batch(code, [r[0] for r in rows], **kw)
                look = {r[0]: (gd[i], int(gs[i])) for i, r in enumerate(rows)}
                parity = {"vs_cpu_arm": parity_report(rows, lambda e: look[e]),
                          "note": "the timed CPU decodes of the unmodified reference (stock LAPACK tie-breaks) against "
                                  "the GPU results of the same strings; max_rel is relative to the largest marginal"}
        line = {
            "metric": "decoded syndromes/sec", "value": value, "unit": "syndromes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(t.item()) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"surface_d{L}_depolarising_p{ERROR_RATE}_bias{BIAS}_chi{chi}_cut{CUT}_optimised",
                       "batch_per_gpu": B, "seed": SEED, "mode": "svd (dense CTA engine)",
                       "note": "timed through decode_css_batch (host strings in, host arrays out); the flag groups "
                               "(X only / Z only / X and Z / none) are separate launches; last-launch counters"},
            "e2e": {"value": value, "unit": "syndromes/s", "h2d_bytes_per_step": int(B * (2 * n + 2)),
                    "d2h_bytes_per_step": int(B * 40)},
            "gpu_launches": args.steps * 4,
            "roofline": {"bound": "tensor", "achieved": flops / sum(times) / 1e12 / max(world, 1), "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": (flops / sum(times) / 1e12) / fp64_peak if fp64_peak else None,
                         "traffic": None, "kernel": f"decode_kernel<{256 if chi > 128 else (128 if chi > 64 else 64)}>",
                         "note": "FLOPs counted by the kernel (Jacobi rotations + dot products, contractions) of the "
                                 "last launch of each step / wall time of the step; peak = cuBLAS DGEMM 8192^3 measured "
                                 "in this run", "hbm": {"algorithmic_bytes": nbytes, "peak_gbs": hbm_peak, "source": src},
                         "two_site_steps": steps},
            "clocks": sampler.summary(),
            "counts": {"decoded": int(len(errors)), "success": int(s.sum()), "nan": int(np.isnan(d).any(axis=1).sum()),
                       "status_not_ok_rank0": int((st != 0).sum())},
            "cpu_baseline": cpu if cpu is not None else {
                "value": None, "unit": "syndromes/s", "cores": 0, "kind": "reference",
                "sample": "not taken: one decode of this configuration by the reference exceeds the 30 s bound by "
                          "orders of magnitude (SURVEY.md 8(d): O(10^2) TFLOP per decode); `--lattice 5` takes one"},
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None,
                    help="syndromes per step: per GPU (weak scaling) or in total (--scaling strong); default 100000 "
                         "(config2) / 148 (config3)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the headline): d=5 bit-flip chi_max=64; config3: d=9 depolarising chi_max=256")
    ap.add_argument("--lattice", type=int, default=9, help="config3 only: lattice size")
    ap.add_argument("--chi", type=int, default=256, help="config3 only: chi_max")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU decodes its own --batch shard; strong: ONE --batch sharded over the GPUs")
    ap.add_argument("--mode", default="mono", choices=["mono", "fast", "svd", "svd-gram"],
                    help="mono (default): the monomial warp engine; svd / svd-gram / fast: the dense CTA engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.batch is None:
        args.batch = CONFIG2_BATCH if args.workload == "config2" else 148
    if args.cpu_sample:
        cpu_sample_main()
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "config3":
        run_config3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
