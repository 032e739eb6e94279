#!/bin/bash
# round-2 GPU call A: tests + bench (both arms) + training iterations + tcgen05 pipe counters of the step GEMMs
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_a.log
python bench.py --steps 10 --warmup 3 > $O/bench_a.log 2> $O/bench_a.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_a.log 2>&1
python profiles/train_bench.py xe 256 5 > $O/train_xe_n1.json 2> $O/train_xe_n1.err; echo "xe rc=$?"
python profiles/train_bench.py rl 512 3 5 > $O/train_rl_n1.json 2> $O/train_rl_n1.err; echo "rl rc=$?"
M=sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tc.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,gpu__time_duration.sum
python profiles/prof_step.py bf16x3 1024 1 > $O/plain_a.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:gemm_tc -c 48 --csv --log-file $O/tc_metrics_a.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_a.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_a.log
