"""Summarise an ncu report's SASS page: executed warp-instructions by opcode, stall samples by reason, and the
hottest instructions.  usage: python tools/ncu_sass_mix.py report.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # the report holds one table per profiled launch; take the first
    blocks = out.split('"Kernel Name"')
    blk = blocks[1 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
    lines = blk.split("\n")
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    h = rows[0]
    ci = {n: i for i, n in enumerate(h)}
    ops = collections.Counter()
    stalls = collections.Counter()
    samples_by_op = collections.Counter()
    total = 0
    smem_wf = collections.Counter()
    for r in rows[1:]:
        if len(r) < len(h):
            continue
        sass = r[ci["Source"]].strip()
        toks = sass.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.rstrip(";")
        n = int(r[ci["Instructions Executed"]] or 0)
        ops[op.split(".")[0]] += n
        total += n
        s = int(r[ci["# Samples"]] or 0)
        samples_by_op[op.split(".")[0]] += s
        for k in h:
            if k.startswith("stall_") and "Not Issued" not in k:
                stalls[k] += int(r[ci[k]] or 0)
        smem_wf[op.split(".")[0]] += int(r[ci["L1 Wavefronts Shared"]] or 0)
    print(f"total warp-instructions {total}")
    for op, n in ops.most_common(40):
        print(f"  {op:12s} {n:12d} {100.0 * n / total:5.1f}%   samples {samples_by_op[op]:7d}  smem-wavefronts {smem_wf[op]}")
    ts = sum(stalls.values())
    print("stall samples:")
    for k, n in stalls.most_common(12):
        print(f"  {k:28s} {n:8d} {100.0 * n / max(ts, 1):5.1f}%")


if __name__ == "__main__":
    main()
