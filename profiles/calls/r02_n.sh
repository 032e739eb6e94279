#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_n.log
bash profiles/ab_bench.sh 2 ab/lib_h.so ab/lib_k.so > $O/ab_n.log 2>&1; cat $O/ab_n.log
