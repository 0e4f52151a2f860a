#!/bin/bash
# ncu evidence for profiles/: launch list of a (shrunk) bench run + one full capture per hot kernel.
# Run under gpurun on ONE GPU:  gpurun --timeout 1500 -- bash tools/profile_r1.sh
# Sessions are disabled (a profiler replays kernels and cannot feed a resident one) and the closed-loop legs are
# shrunk to 40 ms; no number printed by these runs is a bench value.
export GPSB_DISABLE_SESSION=1 GPSB_BENCH_NMS=40
mkdir -p gpurun_out
NCU="ncu --target-processes application-only --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
for k in k_acq_dp4a k_epl_rt k_search "k_epl\\("; do
    n=$(echo "$k" | tr -d '\\(')
    timeout 300 $NCU --set full --import-source on -k "regex:$k" -s 3 -c 1 -f -o gpurun_out/prof_${n}_r1 \
        python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_$n.log 2>&1
    grep -E "==PROF==|Error|error" gpurun_out/ncu_$n.log | tail -2 | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
