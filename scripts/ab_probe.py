"""A/B timing of the fused rollout contracts for library builds with different -D flags (run on the GPU box):
    python scripts/ab_probe.py "" "-DCOUP_AB_X" ...        each argument = one COUP_B200_NVCC_EXTRA value"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import json, sys, torch
sys.path.insert(0, %r)
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv
n = 1 << 20
env = CoupVectorEnv(n, seed=1234, auto_reset=True, finished_ring=1 << 18)
env.rollout(100)
res = {}
for name, dt in (("d8", torch.uint8), ("bf16", torch.bfloat16), ("d32", torch.float32), ("env", None)):
    buf = None if dt is None else torch.empty((n, 2492), dtype=dt, device=env.device)
    sel = None if dt is None else _lib.PLAYER_CURRENT
    k = 640 if dt is None else 100
    env.rollout(5, sel, out=buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(k, sel, out=buf); e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / k * 1e3, 2)
    del buf
print(json.dumps(res))
''' % ROOT
for flags in sys.argv[1:] or [""]:
    env = dict(os.environ, COUP_B200_NVCC_EXTRA=flags)
    subprocess.run([sys.executable, "-m", "open_spiel_coup_b200.build", "--force"], cwd=ROOT, env=env, check=True, capture_output=True)
    out = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True)
    print("%-40s us/step %s %s" % (flags or "(default)", out.stdout.strip(), out.stderr.strip()[-200:]), flush=True)
subprocess.run([sys.executable, "-m", "open_spiel_coup_b200.build", "--force"], cwd=ROOT, check=True, capture_output=True)
