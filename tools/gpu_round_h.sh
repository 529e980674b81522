#!/bin/bash
# r01h GPU visit: full GPU test-suite, bench (both arms), ncu launch list of the bench command,
# one ncu --set full capture of the time-step (CTA, dependency-ordered) kernel.
set -x
TAG=${1:-r01h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json; tail -4 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
cat gpurun_out/bench_ref_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_cta_kernel -c 1 \
    -o gpurun_out/prof_timestep_${TAG} -f python tools/profile_timestep.py > gpurun_out/prof_timestep_${TAG}.log 2>&1
tail -3 gpurun_out/prof_timestep_${TAG}.log
ls -la gpurun_out | tail -12
