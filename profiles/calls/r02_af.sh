#!/bin/bash
O=gpurun_out/r02/af; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 > $O/train_r02_rl_n1.json; cut -c1-140 $O/train_r02_rl_n1.json
timeout 300 python profiles/train_bench.py xe 256 5 2>/dev/null | grep '^{' | head -1 > $O/train_r02_xe_n1.json; cut -c1-140 $O/train_r02_xe_n1.json
