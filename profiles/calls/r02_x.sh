#!/bin/bash
O=gpurun_out/r02/x; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_detector.py -m gpu -q -x > $O/pytest_x.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_x.log
for i in 1 2 3; do timeout 300 python profiles/train_bench.py xe 256 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140; done
for i in 1 2; do timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_xe.csv python profiles/train_bench.py xe 256 1 > $O/ncu_x.log 2>&1
echo "ncu rc=$?"
