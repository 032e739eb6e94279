"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals/shares and the
kernel sequence of decode step t=1. Usage: python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("isc::", "").replace("<unnamed>::", "").replace("void ", "")
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
        seq.append((name, val, row.get("Grid Size")))
    return seq


def main(path):
    seq = load(path)
    agg = collections.OrderedDict()
    for name, val, _ in seq:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print("total %.1f us over %d launches" % (tot, len(seq)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s n=%4d total %10.1f us  avg %8.1f  share %.3f" % (k[:58], v[0], v[1], v[1] / v[0], v[1] / tot))
    # a decode step ends with its merge kernel (beam search) — the one kernel every step has exactly once
    ends = [i for i, s in enumerate(seq) if "beam_merge" in s[0] or "beam_select" in s[0] or "greedy_merge" in s[0]]
    if len(ends) > 2:
        print("decode step t=1:")
        for s in seq[ends[0] + 1:ends[1] + 1]:
            print("   %-50s %9.1f us grid %s" % (s[0][:50], s[1], s[2]))
        print("   step total %.1f us" % sum(s[1] for s in seq[ends[0] + 1:ends[1] + 1]))
        first_gemm = next(i for i, s in enumerate(seq) if "gemm_tc" in s[0] or "gemm_simt" in s[0])
        first_step = next(i for i, s in enumerate(seq) if "embed_pack" in s[0] or "beam_init" in s[0])
        print("prologue (first GEMM .. beam init): %.1f us" % sum(s[1] for s in seq[first_gemm:first_step]))
        print("decode step t=0 (one row per image): %.1f us" % sum(s[1] for s in seq[first_step + 1:ends[0] + 1]))
        print("prologue head:")
        for s in seq[first_gemm:first_gemm + 6]:
            print("   %-50s %9.1f us grid %s" % (s[0][:50], s[1], s[2]))

if __name__ == "__main__":
    main(sys.argv[1])
