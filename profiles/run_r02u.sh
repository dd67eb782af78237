timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r02u_n8.err > gpurun_out/r02u_n8.json
tail -3 gpurun_out/r02u_n8.err | cut -c1-300
python - <<'E'
import json
d = json.loads(open('gpurun_out/r02u_n8.json').read().strip().splitlines()[-1])
print('N8 value %.1fM ms/step %.4f e2e %.1fM first %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['first_call_value']/1e6), d['nvlink']['hw_counters'], d['clocks'])
E
