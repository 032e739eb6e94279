#!/bin/bash
# full GPU suite on the split-K build + attention-backward change; training timing
O=gpurun_out/r02/v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_v.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_v.log
for i in 1 2; do timeout 300 python profiles/train_bench.py xe 256 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140; done
timeout 300 python profiles/train_bench.py rl 512 3 5 2>/dev/null | grep '^{' | head -1 | cut -c1-140
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'], d.get('train'))"
