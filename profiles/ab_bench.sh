#!/bin/bash
# A/B builds of libisc_b200.so on the same box: bash profiles/ab_bench.sh [rounds] lib1.so lib2.so ...
# (each line: library, captions/s, ms per call, per-class kernel ms of the eager pass)
N=$1; shift
for i in $(seq 1 $N); do
  for L in "$@"; do
    ISC_B200_LIB=$PWD/$L timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$L', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
  done
done
