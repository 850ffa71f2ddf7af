# A/B of the launch-order refresh: asynchronous (default) against synchronous (B2MJ_ORDER_SYNC=1) and the round-2 classes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_ab.log
tail -3 gpurun_out/pytest_gpu_ab.log
for k in 1 2; do
  timeout 600 python bench.py --no-cpu --no-parity --steps 300 > gpurun_out/ab_async_$k.json 2> gpurun_out/ab_async_$k.err
  B2MJ_ORDER_SYNC=1 timeout 600 python bench.py --no-cpu --no-parity --steps 300 > gpurun_out/ab_sync_$k.json 2> gpurun_out/ab_sync_$k.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*sync_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('configs',{})
        print(f, 'value %.4g per_step %.4g e2e %.4g'%(d['value'],d['per_step_launch']['value'],d['e2e']['value']),
              {k:(round(v.get('per_step_launch',{}).get('value',0)), round(v.get('e2e',{}).get('value',0)) if isinstance(v.get('e2e'),dict) else None) for k,v in c.items() if isinstance(v,dict)})
    except Exception as ex: print(f,'ERR',ex)
PY
