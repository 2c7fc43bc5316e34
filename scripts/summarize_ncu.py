#!/usr/bin/env python
"""Turn the CSV pages exported from an `ncu --set full` capture (`ncu -i x.ncu-rep --page raw --csv`) into the
summaries committed under profiles/: one row per kernel launch with the counters the roofline discussion uses, and
`r2_ncu_summary.json` (DRAM bytes per launch per kernel) that bench.py reads for `roofline.traffic`.

    python scripts/summarize_ncu.py gpurun_out/r2_ncu_fit_raw.csv gpurun_out/r2_ncu_aux_raw.csv
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def main(paths):
    rows_out, traffic = [], {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("bark::", "").replace("void ", "").strip()
            name = re.sub(r"<.*", "", name)
            out = {"kernel": name}
            for k in KEEP:
                if k in col:
                    out[k] = r[col[k]]
                    out[k + " [unit]"] = units[col[k]]
            rows_out.append(out)
            if "dram__bytes_read.sum" in col:
                b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                    to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                ms = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
                tu = units[col["gpu__time_duration.sum"]]
                ms *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(tu, 1)
                t = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches": 0, "ms_under_ncu": 0.0})
                t["dram_bytes_per_launch"] += b; t["launches"] += 1; t["ms_under_ncu"] += ms
    for t in traffic.values():
        t["dram_bytes_per_launch"] /= t["launches"]; t["ms_under_ncu"] /= t["launches"]
    keys = ["kernel"] + [k for k in KEEP] 
    with open(os.path.join(ROOT, "profiles", "r2_ncu_full_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys + ["units: " + "; ".join(f"{k}={rows_out[0].get(k + ' [unit]', '')}" for k in KEEP)])
        for o in rows_out:
            w.writerow([o.get(k, "") for k in keys])
    traffic["_source"] = "ncu --set full --clock-control none, one launch per kernel (scripts/summarize_ncu.py); per-launch DRAM read + write"
    with open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    for k, t in traffic.items():
        if k != "_source":
            print(f"{k:28s} dram {t['dram_bytes_per_launch'] / 1e6:10.1f} MB/launch  {t['ms_under_ncu']:8.3f} ms (under ncu)  x{t['launches']}")


if __name__ == "__main__":
    main(sys.argv[1:])
