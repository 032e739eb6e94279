#!/bin/bash
run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step']['gemm_tc'])"; }
run base; ISC_GEMM_WIDE=0 run narrow; ISC_GEMM_WIDE=1 run wide; run base; ISC_GEMM_WIDE=0 run narrow
