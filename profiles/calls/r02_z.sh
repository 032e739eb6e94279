#!/bin/bash
timeout 300 python profiles/train_host_profile.py 256 10 2>&1 | grep -v Warning | head -60 | cut -c1-200
