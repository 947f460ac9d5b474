#!/bin/bash
# usage: run_screen.sh "<dbg modes>"  -> RESULT lines in gpurun_out/screen_runs.log
python scratch/exp_screen.py 2>&1 | grep RESULT
for m in $1; do ANN_B200_SCREEN_DBG=$m ANN_B200_SCREEN=1 timeout 120 python scratch/exp_screen.py 2>&1 | grep -E "RESULT|rror"; done
