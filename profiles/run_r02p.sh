nvidia-smi -L | wc -l
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus 4 --steps 20 --warmup 5 2>/dev/null > gpurun_out/r02p_n4_$i.json
python - <<E
import json
d = json.loads(open('gpurun_out/r02p_n4_$i.json').read().strip().splitlines()[-1])
print('N4 value %.1fM ms/step %.4f e2e %.1fM first %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['first_call_value']/1e6), d['nvlink'])
E
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29615 profiles/bench_bloom_sharded.py 2>&1 | tail -1
