#!/bin/bash
# split-K GEMM: parity + training timing on/off
O=gpurun_out/r02/u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q -x -k "test_gemm or train" > $O/pytest_u.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_u.log
for sk in 1 0 1 0; do ISC_SPLITK=$sk timeout 300 python profiles/train_bench.py xe 256 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140; done
ISC_SPLITK=1 timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140
ISC_SPLITK=0 timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_xe.csv python profiles/train_bench.py xe 256 1 > $O/ncu_u.log 2>&1
echo "ncu rc=$?"
