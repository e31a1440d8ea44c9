#!/bin/bash
# GPU box: parity tests, a short bench line, and a reduced parity campaign.  usage: bash tools/r02_check.sh [ncases]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/bench_quick.json
tail -3 gpurun_out/bench_quick.err
python tools/parity_campaign.py ${1:-150} 41 > gpurun_out/campaign.log 2>&1
head -40 gpurun_out/campaign.log
