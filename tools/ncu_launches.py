"""Summarise an ncu per-launch CSV (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv
--log-file X.csv ...`) as a markdown table: launches, total time and share per kernel, DRAM bytes per launch when captured.

    python tools/ncu_launches.py gpurun_out/launches.csv "title line" > profiles/rNN_launches.md
"""
import csv
import re
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("srb::", "").strip()
    a = agg.setdefault(name, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["ids"].add(r["ID"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        a["ns"] += v * scale
    elif m == "dram__bytes_read.sum":
        a["rd"] += v * scale
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * scale
total = sum(a["ns"] for a in agg.values()) or 1.0
has_dram = any(a["rd"] or a["wr"] for a in agg.values())
print(f"# {title}\n# shares, not absolutes (ncu times are cold-cache and serialised)\n")
print("| kernel | launches | total ms | share |" + (" DRAM read MB / launch | DRAM write MB / launch |" if has_dram else ""))
print("|---|---|---|---|" + ("---|---|" if has_dram else ""))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    n = len(a["ids"])
    extra = f" {a['rd'] / n / 1e6:.1f} | {a['wr'] / n / 1e6:.1f} |" if has_dram else ""
    print(f"| `{name}` | {n} | {a['ns'] / 1e6:.3f} | {100 * a['ns'] / total:.1f} % |{extra}")
if has_dram:
    conv = {k: a for k, a in agg.items() if "conv3x3" in k}
    n = sum(len(a["ids"]) for a in conv.values())
    if n:
        print(f"\ntcgen05 conv launches: {n}, mean DRAM traffic {sum(a['rd'] + a['wr'] for a in conv.values()) / n / 1e6:.1f} MB per launch")
