"""Where the time of the host-buffer (e2e) step goes. Run on the GPU box."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv

n = 1 << 20
lib = _lib.load()
env = CoupVectorEnv(n, auto_reset=True)
env.rollout(100)
out = torch.empty((n, 2492), dtype=torch.float32, device="cuda")
h_act = torch.empty(n, dtype=torch.uint8).pin_memory()
h_legal = torch.empty(n, dtype=torch.int32).pin_memory()
h_cur = torch.empty(n, dtype=torch.int8).pin_memory()
h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
h_rew = torch.empty((n, 2), dtype=torch.int8).pin_memory()
h_legal.copy_(env.legal_mask); torch.cuda.synchronize()
threads = os.cpu_count()

def sample(t=threads):
    lib.coup_host_sample_uniform(C.c_void_p(h_legal.data_ptr()), n, 1234, 0, env.step_counter, C.c_void_p(h_act.data_ptr()), t)

for t in (1, 4, 8, 16, 32):
    sample(t); t0 = time.perf_counter()
    for _ in range(20): sample(t)
    print(f"host sample threads={t}: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
for tensor in (None, out):
    for _ in range(3): sample(); env.step_host(h_act, h_legal, h_cur, h_done, h_rew, tensor_out=tensor)
    torch.cuda.synchronize(); ts = tw = 0.0; t_all = time.perf_counter()
    for _ in range(50):
        t0 = time.perf_counter(); sample(); t1 = time.perf_counter()
        env.step_host(h_act, h_legal, h_cur, h_done, h_rew, tensor_out=tensor); t2 = time.perf_counter()
        ts += t1 - t0; tw += t2 - t1
    torch.cuda.synchronize(); tot = time.perf_counter() - t_all
    print(f"tensor={'yes' if tensor is not None else 'no'}: sample {ts/50*1e3:.3f} ms, step_host call {tw/50*1e3:.3f} ms, total/step {tot/50*1e3:.3f} ms")
h_words = torch.empty(n, dtype=torch.int32).pin_memory(); h_words.copy_(env.step_word); torch.cuda.synchronize()
def sample_w(t=threads):
    lib.coup_host_sample_uniform(C.c_void_p(h_words.data_ptr()), n, 1234, 0, env.step_counter, C.c_void_p(h_act.data_ptr()), t)
for tensor in (None, out):
    for _ in range(3): sample_w(); env.step_host_packed(h_act, h_words, tensor_out=tensor)
    torch.cuda.synchronize(); ts = tw = 0.0; t_all = time.perf_counter()
    for _ in range(50):
        t0 = time.perf_counter(); sample_w(); t1 = time.perf_counter()
        env.step_host_packed(h_act, h_words, tensor_out=tensor); t2 = time.perf_counter()
        ts += t1 - t0; tw += t2 - t1
    torch.cuda.synchronize(); tot = time.perf_counter() - t_all
    print(f"packed tensor={'yes' if tensor is not None else 'no'}: sample {ts/50*1e3:.3f} ms, step_host_packed call {tw/50*1e3:.3f} ms, total/step {tot/50*1e3:.3f} ms")
# raw copy speeds
d = torch.empty(8 << 20, dtype=torch.uint8, device="cuda"); h = torch.empty(8 << 20, dtype=torch.uint8).pin_memory()
for nbytes in (1 << 20, 4 << 20, 8 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): h[:nbytes].copy_(d[:nbytes], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"D2H {nbytes>>20} MiB: {dt*1e3:.3f} ms = {nbytes/dt/1e9:.1f} GB/s")
