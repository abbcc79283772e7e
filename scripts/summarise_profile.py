#!/usr/bin/env python
"""Summarise an ncu report of the decode kernel into markdown (run here, no GPU needed).

    python scripts/summarise_profile.py gpurun_out/prof.ncu-rep build/inst_chi64.o decode_kernelILi64 > profiles/xx.md
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
metrics = dict(zip(hdr, vals))
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print(f"# ncu summary: `{os.path.basename(rep)}` — kernel `{metrics.get('Kernel Name', kern)[:80]}`\n")
print("| metric | value |\n|---|---|")
for w in want:
    for h in hdr:
        if h == w:
            print(f"| `{h}` | {metrics[h]} |")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        if float(metrics[h] or 0) >= 0.2:
            print(f"| `{h}` | {float(metrics[h]):.2f} |")

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = dis.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
addr2loc, cur, ingroup = {}, None, False
ours = ("block_ops.cuh", "mps_core.cuh", "kernels.cuh")
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = m.group(1).split("/")[-1]
        if not ingroup:
            cur, ingroup = (f, int(m.group(2))), True
        elif cur[0] not in ours and f in ours:
            cur = (f, int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        ingroup = False
        addr2loc[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
agg, tot, base = collections.defaultdict(collections.Counter), 0, None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    base = a if base is None else base
    loc = addr2loc.get(a - base)
    n = int(r[isamp] or 0)
    tot += n
    agg[loc]["samples"] += n
    agg[loc]["inst"] += int(r[iinst] or 0)
    for nm, i in idx.items():
        agg[loc][nm] += int(r[i] or 0)
print(f"\n## Hottest source lines ({tot} warp samples)\n")
print("| share | warp instr (M) | file:line | top stalls | source |\n|---|---|---|---|---|")
cache = {}
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top = sorted(((c[n], n.replace("stall_", "")) for n in names), reverse=True)[:3]
    text = ""
    if key and key[0] in ours:
        if key[0] not in cache:
            cache[key[0]] = open(os.path.join(ROOT, "mdopt_b200", "csrc", key[0])).read().split("\n")
        text = cache[key[0]][key[1] - 1].strip()[:80].replace("|", "\\|")
    stalls = " ".join(f"{n}:{100 * v / max(c['samples'], 1):.0f}%" for v, n in top)
    print(f"| {100 * c['samples'] / tot:.1f}% | {c['inst'] / 1e6:.0f} | {key[0] if key else '?'}:{key[1] if key else ''} | {stalls} | `{text}` |")
