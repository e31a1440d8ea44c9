#!/bin/bash
# Runs on the GPU box: fp64 pipe microbenchmarks + cuBLAS DGEMM reference, with clocks sampled alongside.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/mb_clocks.csv &
SMI=$!
./build/microbench > gpurun_out/microbench.log 2>&1
python - > gpurun_out/dgemm.log 2>&1 <<'PY'
import torch, time
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"cublasDgemm n={n}: {best:.3f} ms  {2*n**3/best/1e9:.2f} TFLOP/s (best of 5)")
    e0.record()
    for _ in range(10): c = a @ b
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    print(f"cublasDgemm n={n}: {ms:.3f} ms  {2*n**3/ms/1e9:.2f} TFLOP/s (10 back to back)")
# batched potrf via cusolver for context (library baseline, small batch)
n=2048; B=16
x = torch.randn(B, n, n, dtype=torch.float64, device='cuda'); K = x @ x.transpose(1,2) + n*torch.eye(n, dtype=torch.float64, device='cuda')
torch.linalg.cholesky(K); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); L = torch.linalg.cholesky(K); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"torch.linalg.cholesky batch={B} n={n}: {ms:.2f} ms => {B/ms*1e3:.1f} factorizations/s, {B*n**3/3/ms/1e9:.2f} TFLOP/s")
PY
kill $SMI
cat gpurun_out/microbench.log gpurun_out/dgemm.log
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/mb_clocks.csv')))[1:]
sm = sorted(int(r[1].split()[0]) for r in rows if 'MHz' in r[1])
print('clock samples', len(sm), 'min/median/max', sm[0], sm[len(sm)//2], sm[-1])
print('reasons seen:', set(r[4].strip() for r in rows))
PY
