#!/bin/bash
# Box probe. Run under gpurun; writes gpurun_out/probe_box.log
mkdir -p gpurun_out
{
  echo "== host"; nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2
  echo "== gpu"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv
  echo "== fp64 probe"; ./tools/probe_fp64
  echo "== cublas dgemm via torch"
  python - <<'PY'
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"dgemm n={n}: {best:.3f} ms  {2*n**3/best*1e-9:.2f} TFLOP/s")
# skinny-K shape like ours: 2048 x 2048 x 65536
m, k = 2048, 65536
a = torch.randn(m, k, dtype=torch.float64, device='cuda'); b = torch.randn(k, m, dtype=torch.float64, device='cuda')
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"dgemm {m}x{m}x{k}: {best:.3f} ms  {2*m*m*k/best*1e-9:.2f} TFLOP/s")
PY
} > gpurun_out/probe_box.log 2>&1
tail -50 gpurun_out/probe_box.log
