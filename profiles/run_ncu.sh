#!/bin/bash
# Run under gpurun (1 GPU).  Produces the launch list and one full capture of the MF
# step kernels for the bench workload.  Usage: bash profiles/run_ncu.sh [batch] [tag]
B=${1:-65536}; TAG=${2:-r01}
mkdir -p gpurun_out
[ -n "$SKIP_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}_B${B}.csv \
    python bench.py --batch $B --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'mf_|seg_scan' -s 30 -c 6 \
    -o gpurun_out/prof_${TAG}_B${B} -f \
    python bench.py --batch $B --steps 8 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_bench_${TAG}.log 2>&1
ls -la gpurun_out
