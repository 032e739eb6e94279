#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_q.log
bash profiles/ab_bench.sh 2 ab/lib_prev.so ab/lib_f2.so > $O/ab_q.log 2>&1; cat $O/ab_q.log
for L in ab/lib_prev.so ab/lib_f2.so; do ISC_B200_LIB=$PWD/$L timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('bf16 $L', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; done
