"""Summarise an `ncu --metrics <tcgen05 counters> --csv` capture of the gemm_tc launches of one beam-3 call (profiles/calls/*.sh):
per launch the kernel time, the UTCHMMA math ops (sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum = 2*M*N*K*passes incl.
tile padding), flops / time, and the tensor-pipe activity (sm__pipe_tc_cycles_active, % of elapsed).
Usage: python profiles/tc_pipe_summary.py gpurun_out/r02/tc_metrics.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = next(r for r in rows if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
recs = collections.OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr):
        continue
    key = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void isc::tc::", ""), r[ix["Grid Size"]].replace(" ", ""))
    recs.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
print("%-4s %-44s %-12s %8s %12s %9s %8s %8s" % ("id", "kernel <PASSES,BN,ACT,EPI,CG,AF,H16,SK>", "grid", "us", "utchmma ops", "TFLOP/s", "tc_pipe%", "inst_tc"))
for (i, k, g), m in recs.items():
    us = m.get("gpu__time_duration.sum", 0) / 1e3
    ops = m.get("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", 0)
    print("%-4d %-44s %-12s %8.1f %12.3e %9.0f %8.1f %8d" % (
        i, k[:44], g, us, ops, ops / us / 1e6 if us else 0,
        m.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", -1), m.get("sm__inst_executed_pipe_tc.sum", 0)))
