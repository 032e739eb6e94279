#!/bin/bash
O=gpurun_out/r02/ad; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_rl.csv python profiles/train_bench.py rl 512 1 5 > $O/ncu.log 2>&1
echo "ncu rc=$?"; wc -l $O/launches_train_rl.csv
