"""Build the per-GEMM-class table from one ncu pass over scripts/gemm_classes.py:
    python scripts/ncu_gemm_table.py <ncu --csv log> <stdout of gemm_classes.py>
Launches are matched to classes by position (REP launches per class); the last launch of each class is reported."""
import csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
recs = {}
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr) or "gemm" not in r[col["Kernel Name"]]:
        continue
    recs.setdefault(int(r[col["ID"]]), {})[r[col["Metric Name"]]] = (float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]])
ids = sorted(recs)
order = [l.split() for l in open(sys.argv[2]) if l.strip() and l.split()[0].count(".") == 1 and len(l.split()) == 5]
print("# per-call-site ncu metrics of the tcgen05 GEMM kernels (one ds2 transformer block, batch = M / 135), last of 4 launches")
print(f"# {'class':14s} {'M':>6s} {'N':>5s} {'K':>5s} {'time us':>8s} {'TFLOP/s':>8s} {'tensor pipe active %':>21s}")
pos = 0
for name, m, n, k, rep in order:
    m, n, k, rep = int(m), int(n), int(k), int(rep)
    chunk = ids[pos:pos + rep]
    pos += rep
    if len(chunk) < rep:
        break
    rec = recs[chunk[-1]]
    t, unit = rec["gpu__time_duration.sum"]
    us = t / 1000.0 if unit.startswith("ns") else (t * 1000.0 if unit.startswith("ms") else t)
    pipe = rec["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]
    print(f"  {name:14s} {m:6d} {n:5d} {k:5d} {us:8.1f} {2.0 * m * n * k / us / 1e6:8.0f} {pipe:21.1f}")
