#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
PYTHONPATH=$PWD timeout 900 python profiles/loader_bench.py 2048 512 > $O/loader_bench.log 2>&1; echo "rc=$?"; grep -v "^gather" $O/loader_bench.log | tail -26
