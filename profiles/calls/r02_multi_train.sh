#!/bin/bash
# training iterations only on N GPUs: bash profiles/calls/r02_multi_train.sh N
N=$1; O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR profiles/train_bench.py xe 256 5 > $O/train_xe_n$N.log 2> $O/train_xe_n$N.err; echo "xe rc=$?"; grep '^{' $O/train_xe_n$N.log | head -1 > $O/train_r02_xe_n$N.json; cut -c1-200 $O/train_r02_xe_n$N.json
timeout 600 $TR profiles/train_bench.py rl 512 3 5 > $O/train_rl_n$N.log 2> $O/train_rl_n$N.err; echo "rl rc=$?"; grep '^{' $O/train_rl_n$N.log | head -1 > $O/train_r02_rl_n$N.json; cut -c1-200 $O/train_r02_rl_n$N.json
