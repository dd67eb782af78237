timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q --tb=line 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 profiles/bench_bloom_sharded.py 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N2 value %.1fM ms/step %.4f e2e %.1fM first %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['first_call_value']/1e6))"
