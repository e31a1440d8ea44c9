#!/bin/bash
# one full ncu capture of gp_lnlike_kernel on a short run (B=148: one problem per SM)
python tools/dev_perf.py 2048 148 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 1 -c 1 -f -o gpurun_out/prof_dev python tools/dev_perf.py 2048 148 > gpurun_out/ncu_dev.log 2>&1
tail -2 gpurun_out/ncu_dev.log
