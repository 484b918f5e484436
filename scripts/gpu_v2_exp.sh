#!/bin/bash
run() { env "$@" timeout 120 python scripts/sa1_iso.py 2>&1 | tail -1; }
run PN2_SA_TC_DEBUG=0
run PN2_SA_TC_DEBUG=1
run PN2_SA_TC_DEBUG=32
