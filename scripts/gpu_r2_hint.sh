#!/bin/bash
# A/B: mbarrier try_wait with / without the suspend-time hint (two builds of the library)
mkdir -p gpurun_out
export PN2_FPS_BUCKET_MIN=1000000000
for lib in "" "$PWD/situation3d_b200/libpn2_b200_nohint.so"; do
  echo "=== PN2_B200_LIB=$lib"
  export PN2_B200_LIB=$lib
  for b in 8 32; do timeout 120 python scripts/fps_sat_one.py $b | tail -1; done
  PN2_FPS_RP=6 timeout 120 python scripts/fps_sat_one.py 48 | tail -1
  timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-reference-cuda --no-sub-configs --no-e2e 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['parity']['ok']); print({k['kernel']: round(k['us'],1) for k in d['roofline_kernels']})"
done
unset PN2_B200_LIB
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -2
