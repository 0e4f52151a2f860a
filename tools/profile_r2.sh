#!/bin/bash
# ncu / compute-sanitizer / SASS evidence for profiles/ (round 2).  Run under gpurun on ONE GPU:
#   gpurun --timeout 1500 -- bash tools/profile_r2.sh
# The launch list shrinks the closed-loop legs to 40 ms, config 4 to 300 ms and the config-1 batch to 20 000 ms, disables
# the resident session kernel of the host-loop comparison leg and the streaming legs (a profiler serialises kernels: it
# cannot feed a resident kernel from a concurrent copy stream); no number printed by these runs is a bench value.
mkdir -p gpurun_out
NCU="ncu --target-processes application-only --clock-control none"
export GPSB_BENCH_NO_STREAM=1 GPSB_DISABLE_SESSION=1 GPSB_BENCH_NMS=40 GPSB_BENCH_LONG_MS=20000 GPSB_BENCH_C4_MS=300
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list_r2.log 2>&1
tail -1 gpurun_out/ncu_list_r2.log | cut -c1-160
unset GPSB_BENCH_NMS GPSB_BENCH_LONG_MS GPSB_BENCH_C4_MS
# the closed loop at full size: 4 satellites x 1000 ms in one launch
timeout 300 $NCU --set full --import-source on -k regex:k_track_run -s 2 -c 1 -f -o gpurun_out/prof_k_track_run_r2 \
    python tools/loop_once.py 4 > gpurun_out/ncu_k_track_run_r2.log 2>&1
# the batched open-loop correlator at full size: 400 000 cells, TMA ring, prompt arm and all three arms
for arms in 1 3; do
    timeout 300 $NCU --set full --import-source on -k regex:k_epl_batch_tma -s 4 -c 1 -f -o gpurun_out/prof_k_epl_batch_tma${arms}_r2 \
        python tools/batch_once.py 400000 $arms 0 > gpurun_out/ncu_k_epl_batch_tma$arms.log 2>&1
done
GPSB_BENCH_NMS=40 GPSB_BENCH_LONG_MS=20000 GPSB_BENCH_C4_MS=300 timeout 300 $NCU --set full --import-source on -k "regex:k_acq_dp4a" -s 3 -c 1 -f \
    -o gpurun_out/prof_k_acq_dp4a_r2 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_k_acq_dp4a_r2.log 2>&1
ls -la gpurun_out/*_r2.ncu-rep
# compute-sanitizer: memcheck + racecheck over the kernels this round touched - k_track_run (slot-phase walk, streamed and
# resident), k_epl_batch_tma, the sweep -> iq2 stream -> sweep sequence of the round-1 advisor finding, the sharded sweep,
# k_pretrack_run and k_code_rounds_run (cold start of 32 PRNs + the first tracking call)
unset GPSB_BENCH_NO_STREAM GPSB_DISABLE_SESSION
SEL="slot_walk or sweep_scratch or split_runs or streaming_run_equals or starved"
for tool in memcheck racecheck; do
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_${tool}_r2.log \
        python -m pytest tests/test_gpu_loop.py tests/test_gpu_parity.py tests/test_config4.py -x -q -k "$SEL or batch or code_rounds or pre_track" > gpurun_out/sanitizer_${tool}_pytest_r2.log 2>&1
    echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer_${tool}_pytest_r2.log)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_${tool}_r2.log | tail -2
done
