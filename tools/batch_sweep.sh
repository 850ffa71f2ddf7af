# batch sweep of the C2 workload: where throughput saturates.  usage: batch_sweep.sh nenv...
for n in "$@"; do
  python bench.py --nenv $n --steps 200 --warmup 20 --no-cpu --no-parity --no-configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        k = d['kernel']
        print('nenv %7d  rollout %.4g  per-step %.4g  e2e %.4g env-steps/s   (CTA %d envs x %d CTAs, %d B smem/CTA; nefc mean %.2f)' % ($n, d['value'], d['per_step_launch']['value'], d['e2e']['value'], k['warps_per_cta'], k['ctas'], k['smem_bytes_per_cta'], d['workload_stats']['nefc_mean']))
"
done
