# One GPU round: tests, bench (both arms), launch list, ncu captures, stage profile, sweeps.  Run with
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1
timeout 300 python tools/stage_profile.py > gpurun_out/stage_profile.txt 2>&1
timeout 300 python tools/imbalance.py > gpurun_out/imbalance.txt 2>&1
# launch list of the bench command (cold-cache, serialised per-launch times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5 > gpurun_out/b_ncu.log 2>&1
# full capture of one per-step launch and of the fused rollout launch (the dominant kernel, both shapes)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b2k_step -s 6 -c 1 -o gpurun_out/step_full -f python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:b2k_step -s 24 -c 1 -o gpurun_out/rollout_full -f python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5 > gpurun_out/ncu_rollout.log 2>&1
# batch sweep and the other BASELINE configs (per-GPU batch of each config)
for n in 512 1024 2048 8192 16384 65536; do timeout 300 python bench.py --no-cpu --steps 200 --warmup 20 --e2e-steps 20 --nenv $n > gpurun_out/sweep_n$n.json 2>&1; done
timeout 300 python bench.py --no-cpu --steps 200 --warmup 20 --e2e-steps 20 --model hand_like.xml --nenv 1024 > gpurun_out/cfg_c3.json 2>&1
timeout 300 python bench.py --no-cpu --steps 200 --warmup 20 --e2e-steps 20 --model humanoid_like.xml --nenv 2048 > gpurun_out/cfg_c4.json 2>&1
timeout 300 python bench.py --no-cpu --steps 100 --warmup 20 --e2e-steps 10 --model bin.xml --nenv 512 > gpurun_out/cfg_c5.json 2>&1
ls -la gpurun_out
