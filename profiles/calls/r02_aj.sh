#!/bin/bash
O=gpurun_out/r02/aj; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "tiled or one_pass" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 200 python profiles/loader_pipeline.py 8192 512 3 pin 2>&1 | grep -v Warning | grep "epoch [0-9]"
