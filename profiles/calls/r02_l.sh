#!/bin/bash
# round-2 GPU call L: tensor-core weighted sum in the attention kernel: tests, A/B, ncu
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_l.log
bash profiles/ab_bench.sh 2 ab/lib_h.so ab/lib_j.so > $O/ab_m.log 2>&1; cat $O/ab_l.log
for L in ab/lib_h.so ab/lib_j.so; do ISC_B200_LIB=$PWD/$L timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('bf16 $L', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; done
timeout 300 python profiles/prof_step.py bf16x3 1024 1 > $O/plain_m.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o $O/attn_v8 python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_m.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_m.log
