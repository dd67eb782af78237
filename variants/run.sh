#!/bin/bash
# kernel breakdown of one step for each variant library
for lib in base dct m15 m14 dctm15 dctm7; do
  if [ $lib = base ]; then unset SLB_LIBRARY; else export SLB_LIBRARY=/root/repo/variants/lib_$lib.so; fi
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 100 --warmup 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms/step %.4f' % d['ms_per_step'], {k: round(v * 1e3, 1) for k, v in d['roofline']['kernel_ms'].items()})"
done
