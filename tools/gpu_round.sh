#!/bin/bash
# One GPU-box visit: one ncu --set full capture of the dominant search kernel on the bench's own records
# -> profiles/search_kernel_traffic.json (fails when the capture is not of the bench's launch), THEN the bench (both arms;
# its roofline.traffic / roofline.issue come from that capture), then the ncu launch list of the bench command.
# Outputs under gpurun_out/.
set -x
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
PDMPC_STATS_JSON=gpurun_out/prof_search_${TAG}_stats.json timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:search_tile_kernel -s 1 -c 1 -o gpurun_out/prof_search_${TAG} -f \
    python tools/profile_batch.py "build/bench_records/triple_speed_20v_35t_block00[0-6]*.npz" 2 1 0 > gpurun_out/prof_search_${TAG}.log 2>&1
tail -5 gpurun_out/prof_search_${TAG}.log
python tools/capture_to_json.py ${TAG} gpurun_out/prof_search_${TAG}.ncu-rep gpurun_out/prof_search_${TAG}_stats.json || echo "CAPTURE DOES NOT MATCH THE BENCH LAUNCH"
cp profiles/search_kernel_traffic.json gpurun_out/search_kernel_traffic_${TAG}.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
tail -c 1500 gpurun_out/bench_ref_${TAG}.json
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out
