#!/bin/bash
# round-2 GPU call J: logits row threshold (A/B by env), reference bytecode on the box, bench with the fp16-shard e2e
O=gpurun_out/r02; mkdir -p $O
ls oracle/_ref/models | head -3
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_j.log
for i in 1 2; do for T in 0 1; do ISC_LOGITS_THR=$T timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('thr=$T', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; done; done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_j.log 2> $O/bench_j.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_j.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','kernel_ms_per_step')}); e=d['e2e']; print({k:e[k] for k in e if k not in ('note','fp32_host')}); print(e['fp32_host']['value']); print(d['cpu_baseline'])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-600
