mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_d.log 2>&1; tail -3 gpurun_out/pytest_gpu_d.log
for v in base nopromote; do
  extra="X=1"
  [ $v = nopromote ] && extra="B2MJ_NO_PROMOTE=1"
  env $extra timeout 300 python bench.py --no-cpu --no-configs --no-parity --steps 300 --warmup 20 --e2e-steps 100 > gpurun_out/c2_$v.json 2> gpurun_out/c2_$v.err
done
timeout 300 python tools/imbalance.py > gpurun_out/imbalance3.txt 2>&1
