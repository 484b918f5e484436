#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_profile.py rows 2>&1 | grep -v Warning | grep -v _warn_once | tee gpurun_out/train_profile_rows3.txt | head -48
