# same-box A/B of library variants: bash tools/dev_ab.sh "" _r01 _early ...
for v in "$@"; do
  echo "variant: libgprot_b200$v.so"
  for k in 1 2; do
  GPROT_B200_LIB=libgprot_b200$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-parity | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
  done
done
