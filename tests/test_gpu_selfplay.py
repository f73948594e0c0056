"""GPU tests (-m gpu) of the on-device masked policy sampling and the self-play data generator. The sampling
kernel is floating point: it is compared with a plain PyTorch fp32 statement of the reference rule
(nfsp.py:154-167) with tolerance 1e-6 on the probabilities."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.selfplay import MLPPolicy, SelfPlayDataGen, masked_action_probs  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv  # noqa: E402


def test_sample_policy_matches_torch_reference():
    n = 8192
    env = CoupVectorEnv(n, seed=5, auto_reset=True)
    env.rollout(23)
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = torch.randn(n, 18, device="cuda", generator=g) * 4
    probs = torch.empty(n, 18, device="cuda")
    acts = env.sample_policy(logits, probs_out=probs)
    ref = masked_action_probs(logits, env.legal_mask)
    assert torch.allclose(probs, ref, rtol=1e-5, atol=1e-6)            # tolerance 1e-6 (fp32 softmax)
    assert torch.allclose(probs.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    legal = ((env.legal_mask.view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).bool()
    assert legal.gather(1, acts.long().view(-1, 1)).all(), "sampled an illegal action"
    assert (probs[~legal] == 0).all()
    # deterministic in (seed, env id, step counter); bf16 logits accepted
    assert torch.equal(acts, env.sample_policy(logits))
    acts_bf = env.sample_policy(logits.bfloat16(), probs_out=probs)
    assert torch.allclose(probs, masked_action_probs(logits.bfloat16().float(), env.legal_mask), rtol=1e-5, atol=1e-6)
    assert legal.gather(1, acts_bf.long().view(-1, 1)).all()
    env.step(acts)
    env.check_errors()


def test_sample_policy_frequencies():
    """All envs are fresh episodes (legal set {0,1,3,5,6}); with shared logits the empirical action
    frequencies must match the masked softmax (chi-square, 4 d.o.f.)."""
    n = 1 << 18
    env = CoupVectorEnv(n, seed=77)
    logits = torch.tensor([0.3, -1.0, 5.0, 1.2, 0.0, 0.7, -0.4] + [0.0] * 11, device="cuda").repeat(n, 1)
    acts = env.sample_policy(logits)
    p = masked_action_probs(logits[:1], env.legal_mask[:1])[0].cpu().numpy()
    counts = np.bincount(acts.cpu().numpy(), minlength=18)
    assert counts[[2, 4] + list(range(7, 18))].sum() == 0
    chi2 = sum((counts[a] - n * p[a]) ** 2 / (n * p[a]) for a in (0, 1, 3, 5, 6))
    assert chi2 < 23.5, chi2      # p < 1e-4 for 4 degrees of freedom


def test_selfplay_datagen_runs_and_records():
    torch.manual_seed(0)
    gen = SelfPlayDataGen(num_envs=4096, policy=MLPPolicy(), seed=9, reservoir_capacity=10000)
    for _ in range(40):
        acts = gen.step()
        legal = (gen.legal_before >> acts.to(torch.int32)) & 1
        assert legal.all()
    s = gen.env.stats()
    assert s["illegal"] == 0 and s["decision_steps"] == 40 * 4096 and s["episodes"] > 4096
    r = gen.reservoir
    assert r.size == 10000 and r.add_calls == 40 * 4096
    info, probs, mask = r.sample(256)
    assert torch.allclose(probs.sum(1), torch.ones(256, device="cuda"), atol=1e-4)
    # a stored record is a real info state: exactly one observer bit, and it equals the player to move
    f = info.float()
    assert torch.equal(f[:, 0:2].sum(1), torch.ones(256, device="cuda"))
    assert torch.equal(f[:, 0], f[:, 42])
    # stored probabilities vanish outside the stored legal mask
    bits = ((mask.view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).bool()
    assert (probs[~bits] == 0).all()
