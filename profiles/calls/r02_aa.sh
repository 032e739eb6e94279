#!/bin/bash
O=gpurun_out/r02/aa; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-train $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; }
ISC_GATE_FUSED=0 run unfused; ISC_GATE_FUSED=1 run fused; ISC_GATE_FUSED=0 run unfused; ISC_GATE_FUSED=1 run fused
ISC_GATE_FUSED=0 run unfused_bf16 "--precision bf16"; ISC_GATE_FUSED=1 run fused_bf16 "--precision bf16"
