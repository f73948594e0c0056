"""Does the 80 B of state + history READ per row cost the encoder's write stream anything? The same encoder, same output,
with the rows gathered (a) from the envs in order (80 MB of DRAM reads per 2^20 rows) and (b) all from a window of 4096
envs (cache hits).    python scripts/encode_reads_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv


def timed(fn, reps=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


n = 1 << 20
env = CoupVectorEnv(n, seed=3, auto_reset=True)
env.rollout(60)
ids_all = torch.arange(n, device=env.device, dtype=torch.int32)
ids_hot = ids_all % 4096
for name, dt in (("u8", torch.uint8), ("bf16", torch.bfloat16), ("f32", torch.float32)):
    buf = torch.empty((n, 2492), dtype=dt, device=env.device)
    gb = buf.numel() * buf.element_size() / 1e9
    t_plain = timed(lambda: env.information_state_tensor(_lib.PLAYER_CURRENT, out=buf))
    t_all = timed(lambda: env.information_state_tensor_gather(ids_all, _lib.PLAYER_CURRENT, out=buf))
    t_hot = timed(lambda: env.information_state_tensor_gather(ids_hot, _lib.PLAYER_CURRENT, out=buf))
    t_fill = timed(lambda: buf.view(-1).view(torch.int32).zero_())
    print(json.dumps({"dtype": name, "GB": round(gb, 3), "encode_us": round(t_plain, 1), "gather_all_us": round(t_all, 1),
                      "gather_cached_window_us": round(t_hot, 1), "fill_us": round(t_fill, 1),
                      "TBps": {"encode": round(gb / t_plain * 1e3, 2), "cached": round(gb / t_hot * 1e3, 2), "fill": round(gb / t_fill * 1e3, 2)}}), flush=True)
    del buf
