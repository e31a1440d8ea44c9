#!/bin/bash
# ncu capture of the kernel the default gprot-fit layout runs (500 walkers x 7 chunks of 286 points)
python tools/dev_sweep.py chunks small > gpurun_out/sweep_small.log 2>&1; cat gpurun_out/sweep_small.log
ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 1 -c 1 -f -o gpurun_out/prof_small python tools/dev_sweep.py chunks > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
