set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all3.log 2>&1; tail -5 gpurun_out/pytest_gpu_all3.log
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu --no-parity --no-configs > gpurun_out/bench_m12.json 2> gpurun_out/bench_m12.err
