"""Summarise an ncu report per CUDA source line: share of executed instructions and of stall samples.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n] [kernel-regex] [sort: inst|stall]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3 and sys.argv[3]:
    cmd += ["-k", "regex:" + sys.argv[3]]
by_stall = len(sys.argv) > 4 and sys.argv[4] == "stall"
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hdr]
ii, sa = h.index("Instructions Executed"), h.index("# Samples")
lines = []
for r in rows[hdr + 1:]:
    if len(r) <= ii or r[2] != "-":
        continue   # keep the per-CUDA-line aggregate rows (Address == '-')
    try:
        lines.append((float(r[ii]), float(r[sa] or 0), r[0], r[1].strip()[:100]))
    except ValueError:
        pass
ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"total warp-instructions {ti:.3e}, samples {ts:.0f}")
key = (lambda l: l[1]) if by_stall else (lambda l: l[0])
for n, s, ln, src in sorted(lines, key=key, reverse=True)[:top]:
    print(f"inst {n / ti * 100:5.1f}%  stall {s / max(ts, 1) * 100:5.1f}%  L{ln:>4}  {src}")
