#!/bin/bash
# ncu evidence for profiles/: launch list of a (shrunk) bench run + one full capture per hot kernel.
# Run under gpurun on ONE GPU:  gpurun --timeout 1500 -- bash tools/profile_r1.sh
# The launch list shrinks the closed-loop legs to 40 ms and the config-1 batch to 20 000 ms, disables the resident
# session kernel of the host-loop comparison leg and the streaming legs (a profiler serialises kernels: it cannot feed
# a resident kernel from the host or from a concurrent copy stream); no number printed by these runs is a bench value.
mkdir -p gpurun_out
NCU="ncu --target-processes application-only --clock-control none"
GPSB_BENCH_NO_STREAM=1 GPSB_DISABLE_SESSION=1 GPSB_BENCH_NMS=40 GPSB_BENCH_LONG_MS=20000 timeout 600 $NCU --metrics gpu__time_duration.sum -c 4000 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
# the closed loop at full size: 4 satellites x 1000 ms in one launch
timeout 300 $NCU --set full --import-source on -k regex:k_track_run -s 2 -c 1 -f -o gpurun_out/prof_k_track_run_r1 \
    python tools/loop_once.py 4 > gpurun_out/ncu_k_track_run.log 2>&1
grep -E "==PROF==|Error|error" gpurun_out/ncu_k_track_run.log | tail -2 | cut -c1-200
# the batched open-loop correlator at full size: 400 000 cells, prompt arm and all three arms
for arms in 1 3; do
    timeout 300 $NCU --set full --import-source on -k regex:k_epl_batch -s 4 -c 1 -f -o gpurun_out/prof_k_epl_batch${arms}_r1 \
        python tools/batch_once.py 400000 $arms > gpurun_out/ncu_k_epl_batch$arms.log 2>&1
    grep -E "==PROF==|Error|error" gpurun_out/ncu_k_epl_batch$arms.log | tail -2 | cut -c1-200
done
for k in k_acq_dp4a k_search; do
    GPSB_BENCH_NO_STREAM=1 GPSB_DISABLE_SESSION=1 GPSB_BENCH_NMS=40 GPSB_BENCH_LONG_MS=20000 timeout 300 $NCU --set full --import-source on -k "regex:$k" -s 3 -c 1 -f \
        -o gpurun_out/prof_${k}_r1 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
    grep -E "==PROF==|Error|error" gpurun_out/ncu_$k.log | tail -2 | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
