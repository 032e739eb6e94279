#!/bin/bash
# round-2 GPU call D: register-query attention + fused fp16 feature output + MUFU LSTM cell: tests, A/B against lib_a / r01
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_d.log
bash profiles/ab_bench.sh 2 ab/lib_r01.so ab/lib_a.so ab/lib_b.so > $O/ab_d.log 2>&1; cat $O/ab_d.log
for L in ab/lib_a.so ab/lib_b.so; do ISC_B200_LIB=$PWD/$L timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('bf16 $L', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; done > $O/ab_d_bf16.log 2>&1; cat $O/ab_d_bf16.log
python profiles/prof_step.py bf16x3 1024 1 > $O/plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o $O/attn16_v2 python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_d.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_d.log
