"""Attribute the executed-instruction counts and stall samples of one captured launch to CUDA source lines.

ncu's CSV export has the per-SASS-instruction counters but not the source correlation, so the line table comes from
`nvdisasm -g` on the cubin inside the built library (same build as the capture).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <launch index> [n_top] [--src csrc/conv_tc.cu]
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launch = sys.argv[1], int(sys.argv[2])
ntop = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 40
lib = glob.glob(os.path.join(ROOT, "super-resolution*", "srb200", "libsrb200.so"))[0]

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(secs) >= 2 and all(rows[secs[i]][1] == rows[secs[i + 1]][1] for i in range(0, len(secs) - 1, 2)) and len(secs) % 2 == 0:
    launch *= 2                      # this ncu version prints every launch twice
secs.append(len(rows))
kname = rows[secs[launch]][1]
h = rows[secs[launch] + 1]
data = rows[secs[launch] + 2:secs[launch + 1]]
ia, iex, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = int(data[0][ia], 16)
print("kernel:", kname[:110])

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
# demangled template arguments -> mangled-name fragment, e.g. conv3x3_tc_kernel<6, false> -> conv3x3_tc_kernelILi6ELb0E
m = re.match(r"(?:void )?(?:\w+::)*(\w+)(?:<(.*?)>)?\(", kname)
frag = m.group(1)
if m.group(2):
    def mangle(arg):
        arg = arg.strip()
        mm = re.match(r"\((int|bool)\)(-?\d+)", arg)
        if mm:
            return f"L{'i' if mm.group(1) == 'int' else 'b'}{mm.group(2)}E"
        return {"float": "f", "unsigned char": "h"}.get(arg, "")
    frag += "I" + "".join(mangle(a) for a in m.group(2).split(","))
line_of = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    if frag not in dis:
        continue
    cur, active, fname = None, False, None
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            active = frag in ln
            continue
        if not active:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if mm:
            line_of[int(mm.group(1), 16)] = cur
agg = {}
tot_ex = tot_s = 0.0
for r in data:
    off = int(r[ia], 16) - base
    ex, sm = float(r[iex] or 0), float(r[isamp] or 0)
    key = line_of.get(off, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0, 0])
    a[0] += ex; a[1] += sm; a[2] += 1
    tot_ex += ex; tot_s += sm
srcs = {}
def text(key):
    f, l = key
    if f not in srcs:
        cands = glob.glob(os.path.join(ROOT, "super-resolution*", "csrc", f))
        srcs[f] = open(cands[0]).read().splitlines() if cands else []
    return srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
print(f"total executed warp-instructions {tot_ex:.0f}, samples {tot_s:.0f}, sass {len(data)}")
print(f"{'exec%':>6} {'smp%':>6} {'sass':>5}  line")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print(f"{a[0] / tot_ex * 100:6.2f} {a[1] / max(tot_s, 1) * 100:6.2f} {a[2]:5d}  {key[0]}:{key[1]}  {text(key)}")
