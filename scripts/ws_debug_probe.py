"""Where the cycles of the warp-specialised fused step go (needs a library built with
COUP_B200_NVCC_EXTRA=-DCOUP_WS_DEBUG python -m open_spiel_coup_b200.build --force)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from open_spiel_coup_b200.vector_env import CoupVectorEnv

n = 1 << 20
for dtype in (torch.float32, torch.bfloat16, torch.uint8):
    env = CoupVectorEnv(n, seed=1, auto_reset=True)
    env.rollout(100)
    out = torch.empty((n, 2492), dtype=dtype, device=env.device)
    env.rollout(5, encode_player=2, out=out)
    env.clear_stats()
    steps = 20
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); env.rollout(steps, encode_player=2, out=out); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    buf = np.zeros(32, np.uint64)
    import ctypes as C
    env._lib.coup_vec_stats(env._h, C.c_void_p(buf.ctypes.data), None)
    ctas = 148 * steps
    r_busy, r_wait, e_wait, e_busy, total, batches = [float(x) for x in buf[24:30]]
    if total == 0:
        print("library was not built with -DCOUP_WS_DEBUG"); break
    ghz = total / ctas / (ms * 1e-3) / 1e9
    us = lambda c: c / ctas / ghz / 1e3
    print("%-8s %.3f ms/step | per CTA: total %.0f us, rules busy %.0f us + waiting for a buffer %.0f us | encoder busy %.0f us + waiting for records %.0f us | %.1f batches, rules %.1f us/batch, encode %.1f us/batch" % (
        str(dtype).split(".")[-1], ms, us(total), us(r_busy), us(r_wait), us(e_busy), us(e_wait), batches / ctas,
        us(r_busy) / (batches / ctas), us(e_busy) / (batches / ctas)))
    env.close(); del out
