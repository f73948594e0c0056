"""Write-only HBM bandwidth reference points on this GPU (what a pure store stream can reach)."""
import torch, time
n = 10_452_205_568  # bytes of one fp32 info-state pass at 2^20 envs
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
ms = t(lambda: buf.zero_()); print(f"torch zero_ (memset)      {ms:.3f} ms  {n/ms/1e6:.0f} GB/s")
f = buf.view(torch.float32)
ms = t(lambda: f.fill_(1.0)); print(f"torch fill_ (store kernel) {ms:.3f} ms  {n/ms/1e6:.0f} GB/s")
src = torch.empty(n // 2, dtype=torch.uint8, device="cuda"); dst = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
ms = t(lambda: dst.copy_(src)); print(f"torch copy_ (read+write)   {ms:.3f} ms  {n/ms/1e6:.0f} GB/s")
