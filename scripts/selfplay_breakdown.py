"""Per-phase time of one self-play data-generation step (2^18 envs). Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.selfplay import SelfPlayDataGen

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

gen = SelfPlayDataGen(num_envs=1 << 18, seed=1)
gen.env.rollout(100)
env = gen.env
print("encode bf16 [2^18,2496]   %.3f ms" % timeit(lambda: env.information_state_tensor(_lib.PLAYER_CURRENT, out=gen.info_state)))
with torch.no_grad():
    print("policy forward            %.3f ms" % timeit(lambda: gen.policy(gen.info_state)))
    logits = gen.policy(gen.info_state)
    for i, layer in enumerate(gen.policy.net):
        x = gen.info_state if i == 0 else x_prev
        x_prev = layer(x)
        print("   layer %d %-28s %.3f ms" % (i, str(layer)[:28], timeit(lambda: layer(x))))
print("sample_policy             %.3f ms" % timeit(lambda: env.sample_policy(logits, probs_out=gen.action_probs, actions_out=gen.actions)))
print("copies (player, legal)    %.3f ms" % timeit(lambda: (gen.acting_player.copy_(env.current_player), gen.legal_before.copy_(env.legal_mask))))
env.sample_policy(logits, actions_out=gen.actions)
print("whole step                %.3f ms" % timeit(gen.step))

# dense vs move-number-bucketed first layer (selfplay.bucketed_policy_forward), same envs
from open_spiel_coup_b200.selfplay import bucketed_policy_forward, _BUCKET_MOVES
gen.run(30)                                   # states reached under the policy itself
moves = env.move_numbers()
print("move numbers under policy play: mean %.1f, bucket counts %s" % (
    float(moves.float().mean()), torch.bincount(torch.bucketize(moves, torch.tensor(_BUCKET_MOVES[:-1], device=moves.device)), minlength=9).tolist()))
print("bucketed forward (sort+encode+9 GEMMs+rest+unsort) %.3f ms" % timeit(lambda: bucketed_policy_forward(env, gen.policy, gen.info_state)))
def dense():
    env.information_state_tensor(_lib.PLAYER_CURRENT, out=gen.info_state)
    with torch.no_grad():
        return gen.policy(gen.info_state)
print("dense forward (encode+policy)                      %.3f ms" % timeit(dense))
gen.bucketed_first_layer = True
print("whole step, bucketed      %.3f ms" % timeit(gen.step))
gen.bucketed_first_layer = False
print("whole step, dense         %.3f ms" % timeit(gen.step))
