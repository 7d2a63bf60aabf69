#!/usr/bin/env python
"""tools/ncu_summary.py <launches.csv> <bench.json> [out.json]: per-kernel totals of an `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list taken over ONE bench step of the exact pipeline (the last batch in the
list), with the DP fill kernel's DRAM traffic per DP cell (cells from the bench line of the same run).  Writes a JSON summary for
profiles/ (bench.py's roofline.traffic reads it) and prints a table."""
import collections
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        m = re.search(r"(k_\w+)", r[ki])
        name = m.group(1) if m else r[ki][:40]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        launches.setdefault(int(r[ii]), {"name": name})[r[mi]] = v
    seq = [launches[k] for k in sorted(launches)]
    starts = [i for i, l in enumerate(seq) if l["name"] == "k_xe_reset"]
    seq = seq[starts[-1]:] if starts else seq
    agg = collections.OrderedDict()
    for l in seq:
        a = agg.setdefault(l["name"], {"launches": 0, "ms": 0.0, "dram_read_GB": 0.0, "dram_write_GB": 0.0})
        a["launches"] += 1
        a["ms"] += l.get("gpu__time_duration.sum", 0.0) / 1e6
        a["dram_read_GB"] += l.get("dram__bytes_read.sum", 0.0) / 1e9
        a["dram_write_GB"] += l.get("dram__bytes_write.sum", 0.0) / 1e9
    tot = sum(a["ms"] for a in agg.values())
    bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    w = bench["work_per_step"]
    cells = w["seed_dp_cells"] + w["mate_dp_cells"]
    print(f"{'kernel':22s} {'n':>6s} {'ms':>9s} {'share':>6s} {'rd GB':>8s} {'wr GB':>8s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["ms"]):
        a["share"] = a["ms"] / tot
        print(f"{k:22s} {a['launches']:6d} {a['ms']:9.2f} {100 * a['share']:5.1f}% {a['dram_read_GB']:8.2f} {a['dram_write_GB']:8.2f}")
    fill = agg.get("k_dp_fill_h", {})
    out = {"source": "profiles/" + sys.argv[1].split("/")[-1] + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one step; "
                     "per-launch times are cold-cache and serialised: shares, not absolutes)",
           "units_in_step": bench["config"]["batch"], "dp_cells_in_step": cells, "kernels": agg, "total_ms": tot,
           "dram_bytes_per_cell": (fill.get("dram_read_GB", 0) + fill.get("dram_write_GB", 0)) * 1e9 / cells if cells else None,
           "dram_write_bytes_per_cell": fill.get("dram_write_GB", 0) * 1e9 / cells if cells else None}
    print("DP fill: DRAM bytes per cell", out["dram_bytes_per_cell"], "(written:", out["dram_write_bytes_per_cell"], ")")
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
