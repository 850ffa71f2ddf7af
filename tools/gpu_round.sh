set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1
timeout 300 python tools/stage_profile.py > gpurun_out/stage_profile.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5 > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b2k_step -s 3 -c 1 -o gpurun_out/step_full -f python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
