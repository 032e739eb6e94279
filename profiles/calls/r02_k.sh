#!/bin/bash
# round-2 GPU call K: graph-replayed host path: tests + bench
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_k.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-train > $O/bench_k.log 2> $O/bench_k.err; echo "bench rc=$?"; tail -3 $O/bench_k.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_k.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','kernel_ms_per_step')}); e=d['e2e']; print({k:e[k] for k in e if k not in ('note','fp32_host')}); print({k:e['fp32_host'][k] for k in ('value','h2d_gbs','pcie_frac')})
PY
