set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_k20_$i.json 2> gpurun_out/r02a_k20_$i.err; tail -c 600 gpurun_out/r02a_k20_$i.err; done
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02a_k100.json 2> gpurun_out/r02a_k100.err
python - <<'E'
import json
for f in ['r02a_k20_1','r02a_k20_2','r02a_k100']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.1fM ms/step %.4f e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6), d['clocks'], {k: round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
E
bash profiles/run_variants.sh base bulk
