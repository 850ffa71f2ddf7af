# A/B of the launch-order weight (measured env cycles, 64 classes) against the round-2 rows x iterations classes.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_ab.log
tail -3 gpurun_out/pytest_gpu_ab.log
for k in 1 2; do
  timeout 600 python bench.py --no-cpu --no-parity --steps 300 > gpurun_out/ab_new_$k.json 2> gpurun_out/ab_new_$k.err
  B2MJ_ORDER_LEGACY=1 timeout 600 python bench.py --no-cpu --no-parity --steps 300 > gpurun_out/ab_legacy_$k.json 2> gpurun_out/ab_legacy_$k.err
done
timeout 300 python tools/imbalance_model.py humanoid_like.xml 2048 > gpurun_out/imb_c4_new.txt 2>&1
timeout 300 python tools/imbalance_model.py bin.xml 512 > gpurun_out/imb_c5_new.txt 2>&1
timeout 300 python tools/imbalance_model.py hand_like.xml 1024 > gpurun_out/imb_c3_new.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('configs',{})
        print(f, 'value %.3g per_step %.4g e2e %.4g'%(d['value'],d['per_step_launch']['value'],d['e2e']['value']),
              {k:(round(v.get('per_step_launch',{}).get('value',0)), round(v.get('e2e',{}).get('value',0)) if isinstance(v.get('e2e'),dict) else None) for k,v in c.items() if isinstance(v,dict)})
    except Exception as ex: print(f,'ERR',ex)
PY
