#!/bin/bash
# round-2 GPU call H: coalesced LSTM epilogue, compile-time fp16 epilogue switch, attention v6: tests, A/B, launch list, cider
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_h.log
bash profiles/ab_bench.sh 2 ab/lib_e.so ab/lib_f.so > $O/ab_h.log 2>&1; cat $O/ab_h.log
timeout 300 python profiles/prof_step.py bf16x3 1024 1 > $O/plain_h.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_h.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_h1.log 2>&1
python profiles/summarize_launches.py $O/launches_h.csv | tail -14
timeout 300 python profiles/cider_bench.py 512 50 > $O/cider_bench.json 2> $O/cider_bench.err; cat $O/cider_bench.json
