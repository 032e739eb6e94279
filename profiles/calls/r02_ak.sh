#!/bin/bash
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg1_matches or test_gemm" 2>&1 | tail -1
