#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title" [algorithmic_bytes]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
alg_bytes = float(sys.argv[4]) if len(sys.argv) > 4 else None


def page(p):
    txt = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_lsu.sum",
]
lines = [f"# {title}", "", f"Source: `{rep}` (ncu --set full --clock-control none --import-source on; one launch, cold cache, serialised).", ""]
kern = [v for h, v in zip(hdr, vals) if h == "Kernel Name"]
if kern:
    lines += [f"Kernel: `{kern[0]}`", ""]
lines += ["| metric | value | unit |", "|---|---:|---|"]
for w in want:
    if w in m:
        lines.append(f"| {w} | {m[w][1]} | {m[w][0]} |")
try:
    rd = float(m["dram__bytes_read.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][0]]
    wr = float(m["dram__bytes_write.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][0]]
    lines += ["", f"DRAM traffic per launch = {rd + wr:.0f} B (read {rd:.0f} + write {wr:.0f})."]
    if alg_bytes:
        lines.append(f"Algorithmic bytes per launch = {alg_bytes:.0f} B -> traffic / algorithmic = {(rd + wr) / alg_bytes:.3f}.")
    wf = float(m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][1])
    bc = float(m["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][1])
    lines.append(f"Shared-memory bank conflicts: {bc:.0f} of {wf:.0f} wavefronts = {100 * bc / wf:.2f} %.")
except Exception as e:  # pragma: no cover
    lines.append(f"(derived figures unavailable: {e})")
lines += ["", "## Warp stall reasons (per issued instruction)", "", "| reason | ratio |", "|---|---:|"]
for h, u, v in zip(hdr, units, vals):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        try:
            if float(v) > 0.02:
                lines.append(f"| {h[34:].replace('_per_issue_active.ratio', '')} | {float(v):.3f} |")
        except ValueError:
            pass
src = page("source")
hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
sh, data = src[hi], src[hi + 1:]
ix = {h: i for i, h in enumerate(sh)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


agg, cnt = defaultdict(float), defaultdict(float)
for r in data:
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    agg[name] += f(r, "# Samples")
    cnt[name] += f(r, "Instructions Executed")
tot = sum(agg.values()) or 1
lines += ["", "## Sampled stalls and executed warp-instructions by SASS opcode", "", "| opcode | samples | share | warp-instructions executed |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
    lines.append(f"| {k} | {v:.0f} | {100 * v / tot:.1f}% | {cnt[k]:.0f} |")
lines += ["", "## Hottest instructions", "", "| SASS | samples | top stall reasons |", "|---|---:|---|"]
keys = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:12]:
    st = sorted(((k[6:], f(r, k)) for k in keys if f(r, k) > 0), key=lambda kv: -kv[1])[:3]
    lines.append(f"| `{r[ix['Source']][:80]}` | {f(r, '# Samples'):.0f} | {', '.join(f'{a} {b:.0f}' for a, b in st)} |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
