#!/bin/bash
# developer sweep: bench under a list of "VAR=val VAR2=val" environment settings (one per argument)
for cfg in "$@"; do
  env $cfg python bench.py --steps 200 --warmup 20 --no-cpu --e2e-steps 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$cfg', 'value %.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], d['kernel'])
"
done
