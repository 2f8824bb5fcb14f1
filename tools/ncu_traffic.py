"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,... --csv` pass over
tools/run_pipeline_once.py: one table row per launch (summary on stdout) and, with a second argument, the per-kernel DRAM
bytes per launch as JSON — the file bench.py reads for `roofline.traffic` (profiles/r02_traffic.json).
usage: ncu_traffic.py <metrics.csv> [out.json] [skip_launches]"""
import collections, csv, json, re, sys

M = {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "ns",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram", "smsp__inst_executed.sum": "inst"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "ns": 1.0, "usecond": 1e3, "us": 1e3,
        "msecond": 1e6, "ms": 1e6}
# kernel -> the label of the library's own timers (what bench.py's `kernels` / `roofline` use)
LABEL = [("kp_weighted_c1", "kp_weighted_c1_kernel"), ("kp_weighted", "kp_weighted_kernel"), ("gemm_tf32", "gemm_tf32_kernel"), ("nb_query", "nb_query_kernel"),
         ("max_pool", "max_pool_kernel"), ("norm_seg_apply", "norm_seg_apply_kernel"),
         ("norm_seg_groups", "norm_seg_groups_kernel")]


def label_of(name):
    if name.startswith("gemm_nrm_f16_kernel"):
        m = re.search(r"<\(?(?:int\))?(\d), \(?(?:int\))?(\d)>", name)
        return "gemm_nrm_stats_kernel" if (m and m.group(2) == "1") else "gemm_nrm_apply_kernel"
    for key, lab in LABEL:
        if name.startswith(key):
            return lab
    return None


launches = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = M.get(r["Metric Name"])
    if k is None:
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"]).split("::")[-1]
    d = launches.setdefault(int(r["ID"]), {"name": name})
    d[k] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(launches.values())[skip:]
print(f"{'kernel':44s} {'us':>8s} {'rdMB':>8s} {'wrMB':>8s} {'GB/s':>7s} {'tensor%':>8s} {'issue%':>7s} {'dram%':>6s} {'Minstr':>8s}")
agg = collections.OrderedDict()
for d in rows:
    us = d.get("ns", 0.0) / 1e3
    tot = d.get("rd", 0.0) + d.get("wr", 0.0)
    print(f"{re.sub(r'[(].*', '', d['name'])[:44]:44s} {us:8.1f} {d.get('rd', 0) / 1e6:8.1f} {d.get('wr', 0) / 1e6:8.1f} "
          f"{tot / max(us, 1e-9) / 1e3:7.0f} {d.get('tensor', 0):8.1f} {d.get('issue', 0):7.1f} {d.get('dram', 0):6.1f} {d.get('inst', 0) / 1e6:8.1f}")
    lab = label_of(d["name"])
    if lab:
        a = agg.setdefault(lab, {"bytes": 0.0, "launches": 0, "us": 0.0, "inst": 0.0})
        a["bytes"] += tot; a["launches"] += 1; a["us"] += us; a["inst"] += d.get("inst", 0.0)
if len(sys.argv) > 2 and sys.argv[2] != "-":
    out = {"workload": "kitti_pair_kfe_encoder", "pairs_per_call": 8,
           "source": "ncu metrics pass of tools/run_pipeline_once.py 8 1, summarised by tools/ncu_traffic.py "
                     "(profiles/r02_ncu_traffic_summary.txt, r02_ncu_traffic.csv), round 2 final tree",
           "kernels": {lab: {"dram_bytes_per_launch": a["bytes"] / a["launches"], "launches": a["launches"],
                             "sum_us_under_ncu": a["us"], "warp_instructions": a["inst"]} for lab, a in agg.items()}}
    with open(sys.argv[2], "w") as f:
        json.dump(out, f, indent=1)
