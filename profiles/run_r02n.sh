timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r02n_B524288.csv \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_r02n.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'mf_user_kernel|mf_item_kernel|plan_|seg_' -s 24 -c 8 \
    -o gpurun_out/prof_r02n_B524288 -f \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_bench_r02n.log 2>&1
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02n_k100.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_k20.json 2>/dev/null
python - <<'E'
import json
for f in ('r02n_k100', 'r02n_k20'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    print(f, 'ms/step %.4f' % d['ms_per_step'], 'value %.1fM' % (d['value']/1e6), 'e2e %.1fM' % (d['e2e']['value']/1e6), d['roofline']['kernel_ms'], d['roofline']['step_algorithmic']['frac_of_timed_region'])
E
python profiles/bench_seq.py 2>&1 | tail -3
SEQ_DENSE=1 python profiles/bench_seq.py 2>&1 | tail -3
