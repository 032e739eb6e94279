#!/bin/bash
# round-2 evidence call (1 GPU): tests, bench, launch list, tcgen05 counters, ncu --set full of the hot kernels, training, sweep, cider
O=gpurun_out/r02/final; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.log 2>&1
timeout 300 python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-train > $O/bench_bf16.log 2>&1
timeout 300 python profiles/train_bench.py xe 256 5 2> $O/train_xe.err | grep '^{' | head -1 > $O/train_r02_xe_n1.json; cat $O/train_r02_xe_n1.json
timeout 300 python profiles/train_bench.py rl 512 3 5 2> $O/train_rl.err | grep '^{' | head -1 > $O/train_r02_rl_n1.json; cat $O/train_r02_rl_n1.json
timeout 600 python profiles/sweep_bench.py bf16x3 $O/sweep_r02_n1.json > $O/sweep.log 2>&1; tail -8 $O/sweep.log
timeout 300 python profiles/cider_bench.py 512 50 > $O/cider_r02.json 2> $O/cider.err; cat $O/cider_r02.json
M=sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tc.sum,gpu__time_duration.sum
timeout 300 python profiles/prof_step.py bf16x3 1024 1 > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_final.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_tc -c 48 --csv --log-file $O/tc_metrics.csv python profiles/prof_step.py bf16x3 1024 1 > $O/ncu2.log 2>&1
timeout 900 ncu --set full --metrics sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum --clock-control none --import-source on -k regex:'gemm_tc|attention|beam_merge|gate_mix' -s 35 -c 8 -o $O/step python profiles/prof_step.py bf16x3 1024 1 > $O/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cider_score -c 1 -o $O/cider python profiles/cider_bench.py 512 3 > $O/ncu4.log 2>&1
echo "ncu done"; ls -la $O/*.ncu-rep
