set -x
timeout 300 python -m pytest tests/test_mf_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "planned or model or fit" 2>&1 | tail -6
bash profiles/run_variants.sh base u8 u5 i8
SLB_PLAN_SAME_STREAM=1 bash profiles/run_variants.sh base
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_k20.json 2> gpurun_out/r02c_k20.err
tail -5 gpurun_out/r02c_k20.err
python - <<'E'
import json
d = json.loads(open('gpurun_out/r02c_k20.json').read().strip().splitlines()[-1])
print('K20', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6), d['roofline']['step_algorithmic'], d['roofline']['per_kernel_frac'])
E
