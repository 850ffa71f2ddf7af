# One GPU round: tests, bench (both arms), launch list, ncu captures.  Run with
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [quick]'
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
# the driver's own command lines first (short run), then the full BASELINE config
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.err; tail -3 gpurun_out/bench_driver.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>&1
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
if [ "$1" = "quick" ]; then ls -la gpurun_out; exit 0; fi
timeout 300 python tools/stage_profile.py > gpurun_out/stage_profile.txt 2>&1
timeout 300 python tools/imbalance.py > gpurun_out/imbalance.txt 2>&1
# launch list of the bench command (cold-cache, serialised per-launch times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-parity --e2e-steps 5 > gpurun_out/b_ncu.log 2>&1
# full capture of one per-step launch and of the fused rollout launch (the dominant kernel, both shapes)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b2k_step -s 6 -c 1 -o gpurun_out/step_full -f python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-parity --e2e-steps 5 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:b2k_step -s 25 -c 1 -o gpurun_out/rollout_full -f python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-parity --e2e-steps 5 > gpurun_out/ncu_rollout.log 2>&1
# optional extras of round 2c (each a few GPU-seconds to two minutes; run by hand when the budget allows):
#   bash tools/batch_sweep.sh 512 1024 2048 4096 8192 16384 65536 262144     -> profiles/r2c_batch_sweep.txt
#   bash tools/sweep_shape.sh humanoid_like.xml 2048 "X=0" "B2MJ_NO_HOTSET_TRADE=1"   (residency trade A/B)
#   bash tools/sweep_shape.sh hand_like.xml 1024 "X=0" "B2MJ_STAGE_SYNC=0"            (stage barriers A/B)
#   bash tools/ab_order.sh                                                            (launch-order refresh A/B)
#   python tools/batch_probe.py 4096 16384 65536                                      (per-env residency vs batch size)
ls -la gpurun_out
