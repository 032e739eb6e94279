#!/bin/bash
O=gpurun_out/r02/ah; mkdir -p $O
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value'])); print(json.dumps(d['train'])[:900])"
