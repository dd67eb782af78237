set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
bash profiles/run_variants.sh base u6 pf u10
for k in 20 20; do
python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02d_k$k.json 2> gpurun_out/r02d_k$k.err
tail -3 gpurun_out/r02d_k$k.err
python - <<E
import json
d = json.loads(open('gpurun_out/r02d_k$k.json').read().strip().splitlines()[-1])
print('K$k', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), d['roofline']['step_algorithmic'], d['roofline']['per_kernel_frac'])
E
done
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'mf_user_kernel|mf_item_kernel' -s 8 -c 2 \
    -o gpurun_out/prof_r02d_B524288 -f \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_r02d.log 2>&1
