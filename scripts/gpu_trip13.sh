#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sa_profile.py > gpurun_out/sa_profile.log 2>&1; cat gpurun_out/sa_profile.log
PN2_SA_TC_STAGES=2 timeout 300 python scripts/sa_profile.py 2>&1 | tail -8
