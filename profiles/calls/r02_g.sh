#!/bin/bash
# round-2 GPU call G: attention v5 (unrolled set-up, fast softmax) + fp16 shard path: tests, A/B, launch list, ncu
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_g.log
bash profiles/ab_bench.sh 2 ab/lib_d.so ab/lib_e.so > $O/ab_g.log 2>&1; cat $O/ab_g.log
timeout 300 python profiles/prof_step.py bf16x3 1024 1 > $O/plain_g.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_g.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_g1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o $O/attn_v5 python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_g.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_g.log
