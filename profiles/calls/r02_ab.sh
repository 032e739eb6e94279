#!/bin/bash
O=gpurun_out/r02/ab; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_loader.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python profiles/loader_pipeline.py 4096 512 3 pin 2>&1 | grep -v Warning | grep "epoch [0-9]"
timeout 600 python profiles/loader_pipeline.py 8192 512 3 pin 2>&1 | grep -v Warning | grep "epoch [0-9]"
timeout 600 python profiles/loader_pipeline.py 4096 512 3 2>&1 | grep -v Warning | grep "epoch [0-9]"
