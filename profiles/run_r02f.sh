set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_n2.json 2> gpurun_out/r02f_n2.err
tail -5 gpurun_out/r02f_n2.err
python - <<'E'
import json
d = json.loads(open('gpurun_out/r02f_n2.json').read().strip().splitlines()[-1])
print('N2', 'value %.1fM' % (d['value']/1e6), 'ms/step %.4f' % d['ms_per_step'], 'e2e', d['e2e'], d['nvlink'], d['clocks'])
E
