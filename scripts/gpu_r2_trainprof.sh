#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_profile.py rows 2>&1 | grep -v Warning | tee gpurun_out/train_profile_rows.txt | head -60
