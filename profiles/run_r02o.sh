timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'mf_user_kernel|mf_item_kernel' -s 8 -c 2 \
    -o gpurun_out/prof_r02o_B524288 -f \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_r02o.log 2>&1
ls -la gpurun_out/prof_r02o_B524288.ncu-rep
