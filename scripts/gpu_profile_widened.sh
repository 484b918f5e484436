#!/bin/bash
# Evidence run for the widened rows: compute-sanitizer over their tests, then one full ncu
# capture of each kernel at the bench shapes.  Summaries: python scripts/ncu_summary.py gpurun_out/prof_widened.ncu-rep
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_fused_gpu.py -q -m gpu -k "voxel_pe or projection" -x > gpurun_out/sanitize_widened.log 2>&1
tail -4 gpurun_out/sanitize_widened.log
PE_B=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"voxel_pe_kernel" -s 8 -c 2 -o gpurun_out/prof_voxel_pe -f python scripts/voxel_pe_bench.py > gpurun_out/ncu_voxel_pe.log 2>&1
tail -2 gpurun_out/ncu_voxel_pe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"projection_flags|projection_compact|project_map|project_dense" -s 8 -c 4 -o gpurun_out/prof_projection -f python scripts/projection_bench.py > gpurun_out/ncu_projection.log 2>&1
tail -2 gpurun_out/ncu_projection.log
ls -la gpurun_out | grep -E "prof_voxel_pe|prof_projection"
