"""Aggregate the warp-stall samples of `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA line.
usage: ncu_lines.py file.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Line No")
iS = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
cur, per, src, tot = None, collections.OrderedDict(), {}, 0
for r in rows:
    if len(r) < iS + 1 or r[0] == "Line No":
        continue
    if r[0].isdigit():
        cur = int(r[0]); src[cur] = r[1]
        continue
    try:
        n = int(r[iS])
    except ValueError:
        continue
    d = per.setdefault(cur, [0, collections.Counter()])
    d[0] += n; tot += n
    for i, h in stall_cols:
        try:
            d[1][h] += int(r[i])
        except ValueError:
            pass
print("total samples", tot)
for ln, (n, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_n]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in c.most_common(3))
    print(f"{ln:4d} {100 * n / tot:5.1f}%  {src.get(ln, '').strip()[:88]:88s} | {top}")
