#!/bin/bash
# One gpurun call of a measurement round: GPU tests, bench lines, per-config step times, phase profiles, ncu launch list and
# --set full captures.  Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag> [tests|notests]
# Everything lands in gpurun_out/<tag>/ ; nothing printed under ncu is ever used as a bench value.
tag=${1:-r02x}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/gpu.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
  tail -3 $out/pytest_gpu.log
  cp gpurun_out/parity_report.json $out/ 2>/dev/null
fi
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > $out/bench_20_5.json 2>> $out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference.json 2>> $out/bench.err
timeout 600 python bench.py --env RoboschoolHumanoidFlagrunHarder-v1 --no-cpu-baseline --no-extra > $out/bench_flagrunharder.json 2>> $out/bench.err
timeout 600 python tools/step_time.py RoboschoolHumanoid-v1:32768 RoboschoolHopper-v1:4096 RoboschoolAnt-v1:65536 \
    RoboschoolHumanoidFlagrunHarder-v1:32768 RoboschoolHumanoidFlagrun-v1:32768 RoboschoolWalker2d-v1:65536 \
    RoboschoolHalfCheetah-v1:65536 RoboschoolHopper-v1:65536 > $out/step_time.jsonl 2> $out/step_time.err
for e in RoboschoolHumanoid-v1 RoboschoolHumanoidFlagrunHarder-v1; do
  timeout 300 python tools/phase_profile.py $e 32768 300 stagger >> $out/phases.jsonl 2>> $out/phases.err
done
# ncu: launch list of the bench command (shares), then one --set full capture of a steady-state step launch and of the policy
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2604 -c 44 --csv --log-file $out/launches.csv $B > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_static --launch-skip 1310 -c 1 -o $out/k_step_full -f $B > $out/ncu_full_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc5 --launch-skip 1310 -c 1 -o $out/k_mlp_full -f $B > $out/ncu_full_mlp.log 2>&1
ls -la $out
