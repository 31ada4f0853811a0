#!/usr/bin/env python3
"""profiles/rN_traffic.json from an ncu summary CSV (tools/ncu_summary.py): DRAM bytes and warp instructions per
launch of every kernel, keyed on a hash of the kernel sources so that bench.py never reuses numbers captured for
other code (roofline.traffic / roofline.int_issue).
Usage: python tools/make_traffic.py profiles/r2_v1_full.csv profiles/r2_traffic.json --frames N --batch B --templates T
"""
import argparse
import csv
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("out")
    ap.add_argument("--frames", type=float, required=True, help="MFCC frames per launch of the profiled configuration")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--templates", type=int, default=20)
    ap.add_argument("--source", default=None)
    a = ap.parse_args()
    import bench
    rows = list(csv.reader(open(a.csv)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    kern = {}
    for r in rows[2:]:
        name = r[0]
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        ms = float(r[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[col["gpu__time_duration.sum"]], 1.0)
        kern.setdefault(name, {"dram_bytes_per_launch": rd + wr, "ms": ms, "warp_inst_per_launch": float(r[col["smsp__inst_executed.sum"]]),
                               "active_threads_per_inst": float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]])})
    mf = [v for k, v in kern.items() if k.startswith("mfcc_kernel")][0]
    out = {"source": a.source or ("ncu --set full, %s" % a.csv), "kernel_source_sha": bench.kernel_source_sha(),
           "batch": a.batch, "templates": a.templates, "frames_per_launch": a.frames,
           "mfcc_warp_inst_per_frame": mf["warp_inst_per_launch"] / a.frames,
           "summary": os.path.splitext(a.csv)[0] + ".md", "kernels": kern}
    json.dump(out, open(a.out, "w"), indent=1)
    print("wrote", a.out, "sha", out["kernel_source_sha"], "inst/frame %.1f" % out["mfcc_warp_inst_per_frame"])


if __name__ == "__main__":
    main()
