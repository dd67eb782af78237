timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_k20.json 2> gpurun_out/r02k_k20.err; tail -2 gpurun_out/r02k_k20.err
python - <<'E'
import json
d = json.loads(open('gpurun_out/r02k_k20.json').read().strip().splitlines()[-1])
print('K20', 'ms/step %.4f' % d['ms_per_step'], 'value %.1fM' % (d['value']/1e6), 'e2e %.1fM' % (d['e2e']['value']/1e6), d['roofline']['kernel_ms'], d['cpu_baseline'])
E
python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | cut -c1-300
