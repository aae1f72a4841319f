"""Per-kernel SASS evidence of the built library: which kernels carry tcgen05 MMAs (UTCHMMA / UTC*MMA), TMEM loads (LDTM),
TMA loads / stores (UTMALDG / UTMASTG / UBLKCP), cp.async (LDGSTS), packed fp32x2 math (FFMA2 / FMUL2 / FADD2) and - as a
negative check - legacy tensor-core instructions (HMMA).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt      (no GPU needed: cuobjdump reads the .so)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200", "srb200", "libsrb200.so")
PATTERNS = [("UTC*MMA", r"\bUTC\w*MMA"), ("  .2CTA", r"\bUTC\w*MMA\S*\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
            ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("FFMA2", r"\bFFMA2"), ("FMUL2/FADD2", r"\bF(MUL|ADD)2"), ("HMMA", r"\bHMMA"),
            ("instr", r"^\s+/\*[0-9a-f]{4,}\*/")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k, _ in PATTERNS}
            continue
        if cur is None:
            continue
        for k, pat in PATTERNS:
            if re.search(pat, line):
                kernels[cur][k] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} ({', '.join(arch)}; cuobjdump -sass, instruction counts per kernel)")
    print("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async -> LDGSTS; HMMA = warp-level mma.sync:\n"
          "# none in the convolution / attention kernels; psnr_ssim_mma_kernel uses it on purpose (M = 16 banded-Toeplitz filters, DESIGN 4.5)")
    hdr = f"{'kernel':92s}" + "".join(f"{k:>12s}" for k, _ in PATTERNS)
    print(hdr)
    tot = {k: 0 for k, _ in PATTERNS}
    for name in sorted(kernels, key=demangle):
        short = re.sub(r"\((int|bool)\)", "", demangle(name))
        short = re.sub(r"\(.*", "", short).replace("void srb::", "").replace("srb::", "")
        c = kernels[name]
        print(f"{short[:92]:92s}" + "".join(f"{c[k]:12d}" for k, _ in PATTERNS))
        for k in tot:
            tot[k] += c[k]
    print(f"{'TOTAL (' + str(len(kernels)) + ' kernels)':92s}" + "".join(f"{tot[k]:12d}" for k, _ in PATTERNS))


if __name__ == "__main__":
    main()
