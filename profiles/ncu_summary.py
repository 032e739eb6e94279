"""Condense an `ncu --page raw --csv` dump into one line per captured launch (the metrics the roofline needs).
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_summary.py raw.csv"""
import csv
import sys

WANT = [("us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tc_pipe_%", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed"),  # sees tcgen05.mma (UTCHMMA)
        ("utchmma_Gop", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum"),
        ("L2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1_%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("%-46s %-10s " % ("kernel", "grid") + " ".join("%10s" % n for n, _ in WANT))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("isc::", "").replace("void ", "").split("(")[0][:46]
    vals = []
    for n, key in WANT:
        if key not in idx:
            vals.append("-")
            continue
        v, u = r[idx[key]].replace(",", ""), units[idx[key]]
        try:
            f = float(v)
            if u in ("ns", "nsecond"):
                f /= 1e3
            if u == "byte":
                f /= 1e6
            if u == "Kbyte":
                f /= 1e3
            if u == "Gbyte":
                f *= 1e3
            if n == "utchmma_Gop":
                f /= 1e9
            vals.append("%.1f" % f)
        except ValueError:
            vals.append(v[:10])
    print("%-46s %-10s " % (name, r[idx["Grid Size"]].replace(" ", "")[:10]) + " ".join("%10s" % v for v in vals))
