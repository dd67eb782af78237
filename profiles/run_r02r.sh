N=$1
for i in 1 2; do
SLB_TRACE_STEP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$i bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null > gpurun_out/r02r_n${N}_$i.txt
grep trace-step gpurun_out/r02r_n${N}_$i.txt | cut -c1-300
python - <<E
import json
d = json.loads(open('gpurun_out/r02r_n${N}_$i.txt').read().strip().splitlines()[-1])
print('N$N value %.1fM ms/step %.4f e2e %.1fM first %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['first_call_value']/1e6), d['nvlink']['hw_counters'])
E
done
