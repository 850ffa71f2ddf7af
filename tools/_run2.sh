set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_model_formats.py -m gpu -q -x -k "rangefinder or keyframe" > gpurun_out/pytest_gpu_m5.log 2>&1; tail -15 gpurun_out/pytest_gpu_m5.log
