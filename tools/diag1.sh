mkdir -p gpurun_out
B2MJ_PRINT_LAYOUT=1 timeout 200 python tools/stage_profile.py --model hand_like.xml --nenv 1024 --steps 50 --warm 300 > gpurun_out/stage_c3.txt 2>&1
B2MJ_PRINT_LAYOUT=1 timeout 200 python tools/stage_profile.py --model humanoid_like.xml --nenv 2048 --steps 50 --warm 300 > gpurun_out/stage_c4.txt 2>&1
for v in base noovl nopromote; do
  extra=""
  [ $v = noovl ] && extra="B2MJ_NO_AR_OVERLAY=1"
  [ $v = nopromote ] && extra="B2MJ_NO_PROMOTE=1"
  env $extra timeout 300 python bench.py --no-cpu --no-configs --no-parity --steps 300 --warmup 20 --e2e-steps 100 > gpurun_out/c2_$v.json 2> gpurun_out/c2_$v.err
done
