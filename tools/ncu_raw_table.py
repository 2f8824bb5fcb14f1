"""Compact table from `ncu --page raw --csv`: one row per launch with the metrics that matter for a roofline reading."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("lts__t_sector_hit_rate.pct", "l2hit"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]
print("kernel".ljust(34), " ".join(c[1].rjust(8) for c in cols))
for r in data:
    name = r[idx["Kernel Name"]][:60]
    vals = []
    for m, _ in cols:
        v = r[idx[m]] if m in idx else ""
        try:
            f = float(v.replace(",", ""))
            u = units[idx[m]]
            if u == "byte": f /= 1e6
            if u == "Kbyte": f /= 1e3
            if u == "Gbyte": f *= 1e3
            if u in ("ns", "nsecond"): f /= 1e3
            if u in ("ms", "msecond"): f *= 1e3
            vals.append(f"{f:8.1f}")
        except Exception:
            vals.append(v[:8].rjust(8))
    print(name[5:39].ljust(34), " ".join(vals))
