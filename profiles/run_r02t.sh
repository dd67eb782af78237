timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -5
for k in 20 100; do
python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02t_k$k.json 2>/dev/null
python - <<E
import json
d = json.loads(open('gpurun_out/r02t_k$k.json').read().strip().splitlines()[-1])
print('K$k', 'ms/step %.4f' % d['ms_per_step'], 'value %.1fM' % (d['value']/1e6), 'e2e %.1fM' % (d['e2e']['value']/1e6), d['roofline']['step_algorithmic']['frac_of_timed_region'])
E
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mt19937' -c 6 --csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep mt19937 | cut -d, -f5,15- | head -6
