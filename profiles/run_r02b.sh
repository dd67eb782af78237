set -x
timeout 300 python -m pytest tests/test_mf_gpu.py -m gpu -x -q -k "planned" 2>&1 | tail -15
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r02b_B524288.csv \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_r02b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'mf_user_kernel|mf_item_kernel|plan_|seg_' -s 24 -c 8 \
    -o gpurun_out/prof_r02b_B524288 -f \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_bench_r02b.log 2>&1
ls -la gpurun_out | tail -5
