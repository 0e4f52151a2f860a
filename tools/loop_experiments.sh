#!/bin/bash
# Timing experiments on k_track_run (results are WRONG by construction; only the wall time per ms is read):
# which piece is the long pole of a millisecond?  Needs a library built with -DGPSB_LOOP_EXPERIMENTS.
for e in 0 1 2 4 3 5 7; do
    echo -n "experiment $e (1 = no phase 1, 2 = no PLL/FLL, 4 = no DLL): "
    GPSB_LOOP_EXPERIMENT=$e python tools/rtt_probe.py 2>&1 | grep "n_ch   4 device loop"
done
