#!/bin/bash
# round-2 GPU call B: tests of the 16-bit attention path + A/B against the round-1 library + bench + training iterations + tcgen05 counters
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_b.log
bash profiles/ab_bench.sh 2 ab/lib_r01.so ab/lib_a.so > $O/ab_b.log 2>&1; cat $O/ab_b.log
for P in bf16; do for L in ab/lib_r01.so ab/lib_a.so; do ISC_B200_LIB=$PWD/$L timeout 300 python bench.py --precision $P --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$P $L', round(d['value']), round(d['ms_per_step'],3), d['kernel_ms_per_step'])"; done; done > $O/ab_b_bf16.log 2>&1; cat $O/ab_b_bf16.log
python bench.py --steps 10 --warmup 3 > $O/bench_b.log 2> $O/bench_b.err; echo "bench rc=$?"; tail -c 1500 $O/bench_b.log
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_b.log 2>&1
python profiles/train_bench.py xe 256 5 > $O/train_xe_n1.json 2> $O/train_xe_n1.err; echo "xe rc=$?"; cat $O/train_xe_n1.json
python profiles/train_bench.py rl 512 3 5 > $O/train_rl_n1.json 2> $O/train_rl_n1.err; echo "rl rc=$?"; cat $O/train_rl_n1.json
M=sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tc.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,gpu__time_duration.sum
python profiles/prof_step.py bf16x3 1024 1 > $O/plain_b.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:gemm_tc -c 48 --csv --log-file $O/tc_metrics_b.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_b.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_b.log
