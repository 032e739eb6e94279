#!/bin/bash
# round-2 GPU call C: full test suite (no -x) + ncu --set full of the 16-bit attention kernel (source-level)
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_c.log
ISC_B200_LIB=$PWD/ab/lib_a.so python profiles/prof_step.py bf16x3 1024 1 > $O/plain_c.log 2>&1 &&
ISC_B200_LIB=$PWD/ab/lib_a.so ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o $O/attn16 python profiles/prof_step.py bf16x3 1024 1 > $O/ncu_c.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_c.log; ls -la $O/*.ncu-rep
