#!/bin/bash
timeout 300 python profiles/loader_pipeline.py 8192 512 3 pin 2>&1 | grep -v Warning | grep "epoch [0-9]"
