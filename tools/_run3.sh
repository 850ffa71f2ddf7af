mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_m14.log 2>&1; tail -4 gpurun_out/pytest_gpu_m14.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-parity > gpurun_out/bench_m14.json 2> gpurun_out/bench_m14.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_m14.json') if x.startswith('{')]
d=json.loads(l[-1])
print('C2 rollout',round(d['value']),'per-step',round(d['per_step_launch']['value']),'e2e',round(d['e2e']['value']))
for k,v in d.get('configs',{}).items():
    print(k,'per-step',round(v['per_step_launch']['value']),'rollout',round(v.get('rollout',{}).get('value',0)),'e2e',round(v['e2e']['value']))
PY
