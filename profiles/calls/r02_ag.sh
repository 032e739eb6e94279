#!/bin/bash
O=gpurun_out/r02/ag; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_detector.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log | cut -c1-300
for i in 1 2; do timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140; done
