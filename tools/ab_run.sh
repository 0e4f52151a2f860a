#!/bin/bash
# On the GPU box: time k_track_run (tools/rtt_probe.py, 4 satellites x 1000 ms) for every tools/bin/ab/*.so, interleaved,
# REPS times each (STREAM=1: the streaming build, tools/e2e_breakdown.py; BATCH=1: k_epl_batch_tma, tools/batch_once.py).  The library in lib/ is put back afterwards.
cd "$(dirname "$0")/.."
lib=stm32f4_sdr_gps_b200/lib/libgpsb_cuda.so
cp $lib /tmp/libgpsb_cuda.keep
for rep in $(seq ${REPS:-2}); do
    for v in tools/bin/ab/*.so; do
        cp $v $lib
        echo -n "$(basename $v .so): "
        if [ -n "$BATCH" ]; then python tools/batch_once.py 400000 1 2>&1 | tail -1 | tr '\n' ' '; python tools/batch_once.py 400000 3 2>&1 | tail -1
        elif [ -n "$STREAM" ]; then python tools/e2e_breakdown.py 2>&1 | grep -E "resident kernel  |streaming build|track_stream" | tr '\n' ' '; echo
        else GPSB_LOOP_EXPERIMENT=0 python tools/rtt_probe.py 2>&1 | grep "device loop" | sed 's/  n_ch   4 device loop//'; fi
    done
done
cp /tmp/libgpsb_cuda.keep $lib
