for v in "" _fullfence _endbar; do
  echo "variant: libgprot_b200$v.so"
  GPROT_B200_LIB=libgprot_b200$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
