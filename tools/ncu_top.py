"""Summarise an .ncu-rep: headline metrics per captured launch and the top stall sites (source page).

    python tools/ncu_top.py gpurun_out/prof.ncu-rep [n_top]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.per_cycle_active"]
if "Kernel Name" in hdr:
    i = hdr.index("Kernel Name")
    for k, r in enumerate(rows[2:]):
        print(f"launch {k}: {r[i][:110]}")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {units[i]:10s} {[r[i] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for k in range(len(secs) - 1):
    h = rows[secs[k] + 1]
    data = rows[secs[k] + 2:secs[k + 1]]
    ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
    print(f"\n== launch {k}: {len(data)} SASS instructions, {tot} samples")
    agg = {}
    for r in data:
        for i, c in stall:
            if r[i].isdigit():
                agg[c] = agg.get(c, 0) + int(r[i])
    print("   stall totals:", sorted(((v, c[6:]) for c, v in agg.items() if v), reverse=True)[:8])
    byop = {}
    for r in data:
        if not r[isamp].isdigit():
            continue
        toks = [t for t in r[ia].split() if not t.startswith("@")]
        op = toks[0] if toks else "?"
        d = byop.setdefault(op, {"n": 0, "samples": 0, "ex": 0})
        d["n"] += 1
        d["samples"] += int(r[isamp])
        d["ex"] += int(r[iex]) if r[iex].isdigit() else 0
        for i, c in stall:
            if r[i].isdigit() and int(r[i]):
                d[c[6:]] = d.get(c[6:], 0) + int(r[i])
    print("   by opcode (static count, executed, samples, top stalls):")
    for op, d in sorted(byop.items(), key=lambda kv: -kv[1]["samples"])[:14]:
        st = sorted(((v, k) for k, v in d.items() if k not in ("n", "samples", "ex")), reverse=True)[:4]
        print(f"     {op:22s} n={d['n']:5d} ex={d['ex']:>11d} samples={d['samples']:6d} ({100 * d['samples'] / max(tot, 1):4.1f}%) {st}")
    for r in sorted(data, key=lambda r: -int(r[isamp]) if r[isamp].isdigit() else 0)[:ntop]:
        st = sorted(((int(r[i]), c[6:]) for i, c in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / max(tot, 1):5.1f}% ex={r[iex]:>9s} {r[ia].strip()[:64]:64s} {st}")
