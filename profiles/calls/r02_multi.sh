#!/bin/bash
# round-2 multi-GPU call: bash profiles/calls/r02_multi.sh N [sweep]   (bench with its train block, training iterations, sweep)
N=$1; O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.log 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -2 $O/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('$O/bench_n$N.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}); e=d['e2e']; print('e2e', e['value'], e['h2d_gbs'], 'fp32_host', e['fp32_host']['value']); print('train', d['train'])
PY
timeout 600 $TR profiles/train_bench.py xe 256 5 > $O/train_xe_n$N.log 2> $O/train_xe_n$N.err; echo "xe rc=$?"; grep '^{' $O/train_xe_n$N.log | head -1 | tee profiles/train_r02_xe_n$N.json
timeout 600 $TR profiles/train_bench.py rl 512 3 5 > $O/train_rl_n$N.log 2> $O/train_rl_n$N.err; echo "rl rc=$?"; grep '^{' $O/train_rl_n$N.log | head -1 | tee profiles/train_r02_rl_n$N.json
cp profiles/train_r02_*_n$N.json $O/
if [ "$2" = "sweep" ]; then
  timeout 900 $TR profiles/sweep_bench.py bf16x3 $O/sweep_r02_n$N.json > $O/sweep_n$N.log 2>&1; echo "sweep rc=$?"; tail -9 $O/sweep_n$N.log
fi
