"""Per-source-line view of an ncu report (needs -lineinfo and --import-source on): executed warp-instructions and stall
samples of every source line, largest first.
usage: python tools/ncu_by_line.py report.ncu-rep [per-unit divisor] [top-n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, items = "?", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {}
            for i, n in enumerate(r):
                hdr.setdefault(n, i)
            hnames = r
            continue
        if hdr is None or len(r) < len(hnames) or not r[0]:  # rows without a line number are the SASS instructions of the line above
            continue
        try:
            n = int(r[hdr["Instructions Executed"]] or 0)
            s = int(r[hdr["# Samples"]] or 0)
        except ValueError:
            continue
        if n == 0 and s == 0:
            continue
        stalls = []
        for k, i in hdr.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                v = int(r[i] or 0)
                if v:
                    stalls.append((v, k.replace("stall_", "")))
        stalls.sort(reverse=True)
        wf = r[hdr["L1 Wavefronts Shared"]] if "L1 Wavefronts Shared" in hdr else "0"
        items.append((n, s, cur, r[0], r[1].strip()[:70], stalls[:3], int(wf or 0)))
    tot = sum(i[0] for i in items)
    tots = sum(i[1] for i in items)
    print(f"total {tot / div:.1f} instr per unit, {tots} samples")
    for n, s, f, ln, src, st, wf in sorted(items, key=lambda i: -(i[0] / max(tot, 1) + i[1] / max(tots, 1)))[:topn]:
        print(f"{f}:{ln:>4} instr {n / div:7.1f} ({100 * n / tot:4.1f}%) samp {100 * s / max(tots, 1):4.1f}% wf {wf / div:6.1f} | {' '.join(f'{k}:{v}' for v, k in st)} | {src}")


if __name__ == "__main__":
    main()
