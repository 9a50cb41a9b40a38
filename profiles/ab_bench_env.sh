#!/bin/bash
# profiles/ab_bench_env.sh "ENV=.." "ENV=.." ...: bench value (c2 x 16 frames, device resident) under each environment setting
for rep in 1 2; do
for e in "$@"; do
  env $e python bench.py --steps 40 --warmup 3 --reps 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$e', 'value', d['value'], 'ms', d['ms_per_step'], 'uf', [k['avg_us'] for k in d['kernels'] if k['kernel']=='k_uf_fused'])"
done; done
