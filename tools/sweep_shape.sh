# per-step / rollout rates of a bench model under env-var settings; usage: sweep_shape.sh model nenv "VAR=a VAR2=b" ...
model=$1; nenv=$2; shift 2
for cfg in "$@"; do
  env $cfg python bench.py --model $model --nenv $nenv --steps 100 --warmup 10 --preroll 300 --no-cpu --no-parity --no-configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernel']
        print('$model [$cfg] rollout %.4g per-step %.4g e2e %.4g  W %d ctas %d smem/CTA %d' % (d['value'], d['per_step_launch']['value'], d['e2e']['value'], k['warps_per_cta'], k['ctas'], k['smem_bytes_per_cta']))
"
done
