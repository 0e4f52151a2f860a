#!/bin/bash
# k_track_run at bench size (4 satellites x 1000 ms): time per millisecond (not under a profiler), then one ncu pass for
# the instruction-cache hit rate, issued instructions and duration of the launch.  Usage: bash tools/loop_icc.sh [tag]
GPSB_LOOP_EXPERIMENT=0 python tools/rtt_probe.py 2>&1 | grep "device loop"
ncu --target-processes application-only --clock-control none -k regex:k_track_run -s 2 -c 1 \
    --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
    python tools/loop_once.py 4 2>&1 | grep -E "icc|inst_executed|time_duration|no_instruction"
