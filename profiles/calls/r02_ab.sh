#!/bin/bash
O=gpurun_out/r02/ab; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python profiles/loader_bench.py 4096 512 > $O/loader_bench.log 2>&1; grep "shard(\|threads=16" $O/loader_bench.log
