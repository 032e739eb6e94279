#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loader.py -m gpu -q -x > $O/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_o.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step']['gemm_tc'])"; }
run base
for w in 0 2 3 4 5 6 8; do ISC_LSTM_WIDE1=$w run "wide1=$w"; done
for w in 0 2 3 4 5 6 8; do ISC_LSTM_WIDE2=$w run "wide2=$w"; done
run base
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2> $O/bench_o.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); e=d['e2e']; print('e2e', e['value'], e['h2d_gbs'], e['pcie_frac'], 'fp32', e['fp32_host']['value'])"
