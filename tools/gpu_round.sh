#!/bin/bash
# One GPU-box visit: bench (both arms), ncu launch list of the bench command, one ncu --set full
# capture of the search kernel on the bench's own records.  Outputs under gpurun_out/.
set -x
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
cat gpurun_out/bench_ref_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
REC=$(ls build/bench_triple_speed_20v_*s_35t_seed1.npz | head -1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python tools/profile_batch.py $REC 2 1 0 > gpurun_out/prof_search_${TAG}.log 2>&1
tail -5 gpurun_out/prof_search_${TAG}.log
ls -la gpurun_out
