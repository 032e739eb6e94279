#!/bin/bash
O=gpurun_out/r02/y; mkdir -p $O
ISC_TRAIN_PROFILE=1 timeout 300 python profiles/train_bench.py xe 256 10 2>&1 | grep -v Warning | tail -12 | cut -c1-600
timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-200
