#!/bin/bash
# Round-2 evidence on one B200: tests, bench line, ncu launch list + full capture of the dominant kernel, shape / cluster
# sweeps, phase timeline, parity campaigns.  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-400 gpurun_out/r02_bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-parity"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 3 -c 1 -f -o gpurun_out/prof_r02 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/dev_sweep.py chunks small c1 c3 c4 > gpurun_out/r02_shape_sweep.log 2>&1; tail -5 gpurun_out/r02_shape_sweep.log
python tools/dev_cluster.py > gpurun_out/r02_cluster_sweep.log 2>&1
for n in 2048 1000 286; do GPROT_B200_LIB=libgprot_b200_timing.so python tools/dev_timing.py $n 1024; done > gpurun_out/r02_phase_timing.log 2>&1
python tools/parity_campaign.py 400 41 > gpurun_out/campaign41.log 2>&1
python tools/parity_campaign.py 400 11 > gpurun_out/campaign11.log 2>&1
ls -la gpurun_out | tail -30
