#!/bin/bash
timeout 300 python profiles/shard_copy_bench.py 4096 512 2>&1 | grep -v Warning | tail -3
ISC_SHARD_MEMCPY=1 timeout 300 python profiles/shard_copy_bench.py 4096 512 2>&1 | grep -v Warning | tail -3
