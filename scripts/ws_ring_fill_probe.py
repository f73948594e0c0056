import json, sys, torch
sys.path.insert(0, "/root/repo")
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv
n = 1 << 20
res = {}
for ring in (1 << 18, 0):
    env = CoupVectorEnv(n, seed=1234, auto_reset=True, finished_ring=ring)
    env.rollout(100)
    for name, dt in (("d8", torch.uint8), ("bf16", torch.bfloat16), ("d32", torch.float32)):
        buf = torch.empty((n, 2492), dtype=dt, device=env.device)
        env.rollout(5, _lib.PLAYER_CURRENT, out=buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(100, _lib.PLAYER_CURRENT, out=buf); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 100 * 1e3
        # fill of the same buffer
        for _ in range(3): buf.zero_()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): buf.zero_()
        e1.record(); torch.cuda.synchronize()
        f = e0.elapsed_time(e1) / 20 * 1e3
        res[f"{name}_ring{ring}"] = {"us": round(t, 2), "fill_us": round(f, 2), "frac_of_fill": round(f / t, 4)}
        del buf
    env.close()
print(json.dumps(res))
