#!/bin/bash
# Round-2 evidence on one B200 with the final kernels: bitwise fingerprint against a baseline library when one is in the
# tree (libgprot_b200_base.so: build the baseline commit with GPROT_BUILD_VARIANT=base:), tests, bench lines, ncu launch list + full captures, shape / cluster sweeps,
# phase timelines, microbenchmarks.  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
# The parity campaigns themselves (tools/parity_campaign.py 400 41 / 11, five minutes each) are not repeated when the
# fingerprint is bitwise identical.
mkdir -p gpurun_out
if [ -f gprotation_b200/libgprot_b200_base.so ]; then
  GPROT_B200_LIB=libgprot_b200_base.so python tools/dev_fingerprint.py gpurun_out/fp.npy > /dev/null 2>&1
  { echo "tools/dev_fingerprint.py: 480 evaluations (N = 64 .. 2048, chunked and not, 1 / 2 / 4 CTAs per problem, two problems per SM, both synthesis modes),"; echo "final library against the library of commit e58e67c (the one profiles/r02_parity_campaign_seed*.json ran on):"; python tools/dev_fingerprint.py gpurun_out/fp.npy --compare; } > gpurun_out/r02_fingerprint.log 2>&1
  cat gpurun_out/r02_fingerprint.log
fi
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-300 gpurun_out/r02_bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-parity"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 3 -c 1 -f -o gpurun_out/prof_r02 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/dev_sweep.py chunks > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gp_lnlike -s 1 -c 1 -f -o gpurun_out/prof_small python tools/dev_sweep.py chunks > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
python tools/dev_sweep.py chunks small c1 c3 c4 > gpurun_out/r02_shape_sweep.log 2>&1; tail -3 gpurun_out/r02_shape_sweep.log
python tools/dev_cluster.py > gpurun_out/r02_cluster_sweep.log 2>&1
{ for a in "2048 1024" "1000 1024 1" "1000 1024 64" "286 3500 64" "128 4096 64"; do GPROT_B200_LIB=libgprot_b200_timing.so python tools/dev_timing.py $a; done; } > gpurun_out/r02_phase_timing.log 2>&1
./build/sweep_bench > gpurun_out/r02_sweep_bench.log 2>&1
./build/trsm_bench > gpurun_out/r02_trsm_bench.log 2>&1
./build/chain_bench > gpurun_out/r02_chain_bench.log 2>&1
ls -la gpurun_out | tail -25
