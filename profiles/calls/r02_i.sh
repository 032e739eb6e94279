#!/bin/bash
# round-2 GPU call I: best build so far: tests, bench (all blocks), training iteration profile + launch list
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_i.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_i.log 2> $O/bench_i.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_i.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','kernel_ms_per_step')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['call_frac'], d['roofline']['north_star_frac']); print(d['train']); print(d['cpu_baseline'])
PY
timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train > $O/bench_i_bf16.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/r02/bench_i_bf16.log').read().strip().splitlines()[-1]); print('bf16', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
ISC_TRAIN_PROFILE=1 timeout 300 python profiles/train_bench.py xe 256 5 > $O/train_xe_prof.log 2>&1; cat $O/train_xe_prof.log
timeout 300 python profiles/train_bench.py xe 256 2 > /dev/null 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_i.csv python profiles/train_bench.py xe 256 1 > $O/ncu_i.log 2>&1
python profiles/summarize_launches.py $O/launches_train_i.csv | head -24
