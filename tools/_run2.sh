set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py -m gpu -q -x -k "implicit or matrix or fluid" > gpurun_out/pytest_gpu_m4.log 2>&1; tail -15 gpurun_out/pytest_gpu_m4.log
