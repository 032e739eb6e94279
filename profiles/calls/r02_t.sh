#!/bin/bash
# training launch list (XE iteration, 256 rows) + the switched-variant parity test
O=gpurun_out/r02/t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k switched > $O/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_t.log
timeout 300 python profiles/train_bench.py xe 256 3 > $O/train_plain.log 2>&1 && tail -1 $O/train_plain.log | cut -c1-300 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_xe.csv python profiles/train_bench.py xe 256 1 > $O/ncu_t.log 2>&1
echo "ncu rc=$?"; wc -l $O/launches_train_xe.csv
