"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name).split("::")[-1]
    rows.append((int(r["ID"]), name, us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[skip:skip + take]
agg = collections.OrderedDict()
for _, n, us in rows:
    c, t = agg.get(n, (0, 0.0)); agg[n] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
print(f"launches {len(rows)}  total {tot/1e3:.3f} ms (ids {rows[0][0]}..{rows[-1][0]})")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:48s} {c:5d} {t:10.1f} us  {100*t/tot:5.1f}%  avg {t/c:7.2f} us")
