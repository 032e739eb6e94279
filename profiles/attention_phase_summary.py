"""Phase table of the attention kernel from an `ncu --set full --import-source on` capture.

    ncu -i step.ncu-rep --page source --csv --kernel-name regex:attention > src.csv
    ncu -i step.ncu-rep --page raw --csv > raw.csv
    python profiles/attention_phase_summary.py src.csv raw.csv

Phases are the SASS ranges between the kernel's CTA-wide barriers (BAR.SYNC); per phase: warp-state samples (time share)
and warp instructions executed; plus the instruction mix and the stall-reason totals of the whole kernel."""
import collections
import csv
import sys

PHASES = [  # the kernel's barrier-delimited phases in program order (attention_kernel<PREC, 3>, csrc/kernels_step.cu)
    "query set-up (all global loads requested up front; rings already streaming)",
    "scores: content rows (register queries, treduce8) + feature-ring prime + sentiment rows",
    "softmax (content rows on warps 0-2, sentiment rows on warps 7-5)",
    "weighted sum through the per-warp rings",
    "cross-warp sum, sentiment context, stores, exit",
]
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active"]


def main(src_csv, raw_csv=None):
    if raw_csv:
        rows = list(csv.reader(open(raw_csv)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        r = next(r for r in rows[2:] if "attention_kernel" in r[ix["Kernel Name"]])
        for k in RAW:
            if k in ix:
                print("%-78s %s %s" % (k, r[ix[k]], units[ix[k]]))
        print()
    rows = list(csv.reader(open(src_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[h + 1:]:
        if not r or r[0] == "Kernel Name":  # a second capture of the same kernel follows: the first one is summarised
            break
        if len(r) >= len(hdr) - 2:
            body.append(r)

    def num(r, name):
        try:
            return float(r[ix[name]].replace(",", ""))
        except (ValueError, IndexError, KeyError):
            return 0.0

    tot_s = sum(num(r, "# Samples") for r in body)
    tot_i = sum(num(r, "Instructions Executed") for r in body)
    print("warp instructions executed: %d; warp-state samples: %d" % (tot_i, tot_s))
    mix = collections.Counter()
    for r in body:
        op = r[ix["Source"]].strip().split()
        op = [o for o in op if not o.startswith("@")]
        if op:
            mix[op[0].split(".")[0]] += num(r, "Instructions Executed")
    print("instruction mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in mix.most_common(12)))
    stalls = collections.Counter()
    for n in hdr:
        if n.startswith("stall_") and "Not Issued" not in n:
            stalls[n[6:]] = sum(num(r, n) for r in body)
    print("stall reasons (samples): " + ", ".join("%s %d" % (k, v) for k, v in stalls.most_common(8)))
    print()
    cuts = [i for i, r in enumerate(body) if "BAR.SYNC" in r[ix["Source"]]]
    bounds = [0] + [c + 1 for c in cuts] + [len(body)]
    names = list(PHASES)
    if len(bounds) > 2 and bounds[1] < 50:  # a barrier right after the entry sequence
        names.insert(0, "kernel entry")
    print("%-88s %-12s %8s %6s %12s" % ("phase", "SASS lines", "samples", "share", "warp instr"))
    for p in range(len(bounds) - 1):
        a, b = bounds[p], bounds[p + 1]
        if a >= b:
            continue
        s = sum(num(r, "# Samples") for r in body[a:b])
        n = sum(num(r, "Instructions Executed") for r in body[a:b])
        name = names[p] if p < len(names) else "tail"
        print("%-88s %-12s %8d %5.1f%% %12d" % (name, "%d-%d" % (a, b - 1), s, 100 * s / tot_s, n))


if __name__ == "__main__":
    main(*sys.argv[1:3])
