"""Time of the stand-alone info-state encoder and of a zero-fill of the same bytes against the number of envs: separates the
fixed cost of a launch (ramp-up, tail) from the streaming rate.    python scripts/encode_scaling_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


rows = []
for lg in (16, 17, 18, 19, 20, 21):
    n = 1 << lg
    env = CoupVectorEnv(n, seed=3, auto_reset=True)
    env.rollout(60)
    for name, dt in (("u8", torch.uint8), ("f32", torch.float32)):
        if lg == 21 and dt == torch.float32:
            continue
        buf = torch.empty((n, 2492), dtype=dt, device=env.device)
        flat = buf.view(-1).view(torch.int32)
        t_enc = timed(lambda: env.information_state_tensor(_lib.PLAYER_CURRENT, out=buf), 30)
        t_fill = timed(lambda: flat.zero_(), 30)
        t_roll = timed(lambda: env.rollout(1, _lib.PLAYER_CURRENT, out=buf), 30)
        gb = buf.numel() * buf.element_size() / 1e9
        rows.append({"envs": n, "dtype": name, "GB": round(gb, 3), "encode_us": round(t_enc, 1), "fill_us": round(t_fill, 1),
                     "fused_step_us": round(t_roll, 1)})
        print(json.dumps(rows[-1]), flush=True)
        del buf, flat
    env.close()
