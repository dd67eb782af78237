set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for k in 20 20 100; do
python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02e_k$k.json 2> gpurun_out/r02e_k$k.err
tail -3 gpurun_out/r02e_k$k.err
python - <<E
import json
d = json.loads(open('gpurun_out/r02e_k$k.json').read().strip().splitlines()[-1])
print('K$k', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), {k: round(v * 1e3, 1) for k, v in d['roofline']['kernel_ms'].items()}, d['roofline']['step_algorithmic']['frac_of_timed_region'])
E
done
