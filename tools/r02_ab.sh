#!/bin/bash
# GPU box: bitwise fingerprints of library variants against the first one, A/B of the headline and the small-chunk
# shapes.  usage: bash tools/r02_ab.sh "" _base ...
mkdir -p gpurun_out
first=1
for v in "$@"; do
  if [ $first = 1 ]; then GPROT_B200_LIB=libgprot_b200$v.so timeout 300 python tools/dev_fingerprint.py gpurun_out/fp.npy; first=0
  else echo "fingerprint libgprot_b200$v.so vs first:"; GPROT_B200_LIB=libgprot_b200$v.so timeout 300 python tools/dev_fingerprint.py gpurun_out/fp.npy --compare; fi
done 2>&1 | tee gpurun_out/fingerprint.log
for v in "$@"; do
  echo "variant: libgprot_b200$v.so"
  for k in 1 2; do
    GPROT_B200_LIB=libgprot_b200$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-parity | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
  done
  GPROT_B200_LIB=libgprot_b200$v.so timeout 300 python tools/dev_sweep.py chunks small c1 2>&1 | grep -v parity
done 2>&1 | tee gpurun_out/ab.log
