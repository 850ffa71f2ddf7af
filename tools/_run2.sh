set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step_host or rollout" > gpurun_out/pytest_gpu_m13.log 2>&1; tail -15 gpurun_out/pytest_gpu_m13.log
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu --no-parity --no-configs > gpurun_out/bench_m13.json 2> gpurun_out/bench_m13.err
B2MJ_NO_ZERO_COPY=1 timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu --no-parity --no-configs > gpurun_out/bench_m13_copy.json 2> gpurun_out/bench_m13_copy.err
