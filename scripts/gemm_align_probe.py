"""How much the unaligned K = 2492 of the info-state tensor costs the consumer's first GEMM (bf16)."""
import torch, time
n = 1 << 18
for k in (2492, 2496, 2560):
    x = torch.randn(n, k, device="cuda", dtype=torch.bfloat16)
    w = torch.nn.Linear(k, 1024, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): y = w(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): y = w(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"K={k}: {dt*1e3:.3f} ms  {2*n*k*1024/dt/1e12:.0f} TFLOP/s")
x = torch.randn(n, 1024, device="cuda", dtype=torch.bfloat16); w = torch.nn.Linear(1024, 1024, device="cuda", dtype=torch.bfloat16)
for _ in range(3): y = w(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): y = torch.relu(w(x))
torch.cuda.synchronize(); print(f"1024x1024 + relu: {(time.perf_counter()-t0)/10*1e3:.3f} ms")
