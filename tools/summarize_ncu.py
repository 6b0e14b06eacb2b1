"""Condense `ncu --page raw --csv` exports into the small JSON summaries committed under profiles/.
    python tools/summarize_ncu.py <tag>      (reads gpurun_out/<tag>_*.csv, writes profiles/<tag>_ncu_*.json)"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = {k: r[hdr.index(k)] for k in KEYS if k in hdr}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in hdr:
                d[k + "_bytes"] = to_bytes(r[hdr.index(k)], units[hdr.index(k)])
        out.append(d)
    return out


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg, total = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("pn2::<unnamed>::", "")
        t = float(r[mv].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    return {"total_us": total, "launches": sum(a[0] for a in agg.values()),
            "kernels": [{"kernel": n, "launches": c, "us": round(t, 1), "share": round(t / total, 4)}
                        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}


def main():
    tag = sys.argv[1]
    g = os.path.join(ROOT, "gpurun_out")
    p = os.path.join(ROOT, "profiles")
    traffic = {}
    for part in ("gemm", "rest"):
        path = os.path.join(g, f"{tag}_{part}_raw.csv")
        if not os.path.exists(path):
            continue
        rows = rows_of(path)
        json.dump(rows, open(os.path.join(p, f"{tag}_ncu_{part}.json"), "w"), indent=1)
        per = collections.defaultdict(list)
        for r in rows:
            name = r["Kernel Name"]
            key = ("gemm_tc_kernel" if ("gemm_tc" in name or "wgrad_tc" in name) else "gemm_kernel" if "gemm_kernel" in name else
                   "fps_kernel" if "fps_" in name else name.split("(")[0].split("::")[-1])
            per[key].append(r.get("dram__bytes_read.sum_bytes", 0) + r.get("dram__bytes_write.sum_bytes", 0))
        for k, v in per.items():
            traffic[k] = sum(v) / len(v)
        if part == "gemm":  # the capture holds exactly one step's GEMM launches (tools/profile_step.sh)
            traffic["gemm_launches_captured"] = len(rows)
            traffic["gemm_dram_bytes_per_step"] = sum(r.get("dram__bytes_read.sum_bytes", 0) + r.get("dram__bytes_write.sum_bytes", 0)
                                                      for r in rows)
            traffic["gemm_us_per_step_under_ncu"] = sum(float(r["gpu__time_duration.sum"].replace(",", "")) for r in rows)
            tp = [float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"].replace(",", "")) for r in rows]
            tt = [float(r["gpu__time_duration.sum"].replace(",", "")) for r in rows]
            traffic["gemm_tensor_pipe_active_pct_time_weighted"] = sum(a * b for a, b in zip(tp, tt)) / max(sum(tt), 1e-9)
    if traffic:
        json.dump(traffic, open(os.path.join(p, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    lpath = os.path.join(g, f"{tag}_launches.csv")
    if os.path.exists(lpath):
        json.dump(launches(lpath), open(os.path.join(p, f"{tag}_ncu_launch_summary.json"), "w"), indent=1)
    print("traffic", traffic)


if __name__ == "__main__":
    main()
