#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_s.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_s.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; }
ISC_AF_PAIR=0 run single; ISC_AF_PAIR=1 run pair; ISC_AF_PAIR=0 run single; ISC_AF_PAIR=1 run pair
ISC_AF_PAIR=1 timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('bf16 pair', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
