# per-step / e2e rates of a bench model under forced residencies (B2MJ_ENVS_PER_SM); usage: sweep_residency.sh model nenv E...
model=$1; nenv=$2; shift 2
for E in "$@"; do
  if [ "$E" = default ]; then unset B2MJ_ENVS_PER_SM; else export B2MJ_ENVS_PER_SM=$E; fi
  python bench.py --model $model --nenv $nenv --steps 100 --warmup 10 --preroll 300 --no-cpu --no-parity --no-configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$model E=$E', 'rollout %.4g per-step %.4g e2e %.4g' % (d['value'], d['per_step_launch']['value'], d['e2e']['value']), d['kernel'])
"
done
