set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -m gpu -q -x -s -k "spatial" > gpurun_out/pytest_gpu_m10.log 2>&1; tail -30 gpurun_out/pytest_gpu_m10.log
