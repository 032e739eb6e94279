#!/bin/bash
O=gpurun_out/r02/ai; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench rc=$?"; grep '^{' $O/bench_n2.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value'])); print(json.dumps(d['train'])[:1200])"
