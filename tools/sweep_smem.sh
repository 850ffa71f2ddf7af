#!/bin/bash
# developer sweep: throughput vs shared-memory budget per env (B2MJ_SMEM_TARGET_KB)
for kb in "$@"; do
  B2MJ_SMEM_TARGET_KB=$kb python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('target_kb', $kb, 'value %.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d['kernel'])
"
done
