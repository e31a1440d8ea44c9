#!/bin/bash
# ncu --set full captures of the likelihood kernel for several library variants (B=148, N=2048): bash tools/dev_ncu2.sh "" _r01
for v in "$@"; do
  GPROT_B200_LIB=libgprot_b200$v.so ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 1 -c 1 -f -o gpurun_out/prof_ab$v python tools/dev_perf.py 2048 148 > gpurun_out/ncu_ab$v.log 2>&1
  tail -2 gpurun_out/ncu_ab$v.log
done
