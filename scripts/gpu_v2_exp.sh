#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; env "$@" timeout 120 python scripts/sa_profile.py 2>&1 | grep -A4 "^SA1\|^SA2" | head -12; }
run PN2_SA_TC_V2=1
run PN2_SA_TC_DEBUG=1
run PN2_SA_TC_DEBUG=3
run PN2_SA_TC_V2_SLOTS=2
run PN2_SA_TC_V2_SLOTS=1
run PN2_SA_TC_V2_REGIONS=3
