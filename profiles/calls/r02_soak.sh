#!/bin/bash
# repeat the GPU suite to catch flaky tests (cluster launches, prefetcher threads), then the default bench twice
O=gpurun_out/r02/soak; mkdir -p $O
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_$i.log 2>&1; echo "run $i rc=$? $(tail -1 $O/pytest_$i.log)"; done
for i in 1 2; do timeout 600 python bench.py > $O/bench_$i.log 2> $O/bench_$i.err; echo "bench $i rc=$?"; tail -1 $O/bench_$i.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['train']['xe_iteration_ms'], d['clocks'])"; done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
