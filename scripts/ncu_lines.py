"""Aggregate the warp-stall samples of an ncu capture per CUDA source line:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv ; python scripts/ncu_lines.py X.csv [top]
Prints the lines with the most samples (all / not-issued) and the SASS instructions that carry them."""
import csv, sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
per_line = defaultdict(lambda: [0, 0, ""])
per_sass = []
cur_file, cur_line, cur_src, hdr = "", None, "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1]
        continue
    try:
        a, n = int(r[4]), int(r[5])
    except (ValueError, IndexError):
        continue
    key = (cur_file, cur_line)
    per_line[key][0] += a; per_line[key][1] += n; per_line[key][2] = cur_src
    per_sass.append((a, n, cur_file, cur_line, r[3].strip()))
total = sum(v[0] for v in per_line.values()) or 1
print(f"total samples {total}")
for (f, l), (a, n, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a:7d} {100 * a / total:5.1f}%  not-issued {n:7d}  {f}:{l}  {src.strip()[:110]}")
print("---- top SASS")
for a, n, f, l, s in sorted(per_sass, key=lambda t: -t[0])[:top]:
    print(f"{a:7d} {100 * a / total:5.1f}%  {f}:{l}  {s[:100]}")
