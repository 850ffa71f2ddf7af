set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_m2.log 2>&1; tail -3 gpurun_out/pytest_gpu_m2.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-parity > gpurun_out/bench_m2.json 2> gpurun_out/bench_m2.err
timeout 300 python tools/stage_profile.py --help > gpurun_out/sp_help.txt 2>&1
