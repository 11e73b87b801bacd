#!/bin/bash
# ncu captures of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/).
# Every command runs once WITHOUT ncu first and must exit 0 (B200_PROFILING.md).
set -u
export LILOC_BENCH_QUICK=1
LIO="python bench.py --steps 1 --warmup 1 --frames 80 --no-ndt --no-concurrent --no-latency --no-cpu"
NDT="python bench.py --workload relo --steps 1 --warmup 1 --no-cpu"
$LIO > gpurun_out/ncu_plain_lio.log 2>&1 && $NDT > gpurun_out/ncu_plain_ndt.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file gpurun_out/launches_r02.csv $LIO > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voxel_pipeline -s 200 -c 2 -o gpurun_out/prof_pipeline_r02 $LIO > gpurun_out/ncu_pipeline.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan2map -s 100 -c 2 -o gpurun_out/prof_s2m_r02 $LIO > gpurun_out/ncu_s2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ndt_align -s 1 -c 1 -o gpurun_out/prof_ndt_r02 $NDT > gpurun_out/ncu_ndt.log 2>&1
ls -la gpurun_out/*.ncu-rep
