"""CPU tests of the host-side pieces of the self-play data generator."""
import numpy as np
import pytest
import torch

from open_spiel_coup_b200.selfplay import MLPPolicy, ReservoirBuffer, masked_action_probs


def test_masked_action_probs_is_nfsp_act_rule():
    """open_spiel/python/algorithms/nfsp.py:154-167 restated with NumPy."""
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(64, 18)).astype(np.float32) * 3
    masks = rng.integers(1, 1 << 18, size=64).astype(np.int32)
    got = masked_action_probs(torch.from_numpy(logits), torch.from_numpy(masks)).numpy()
    for i in range(64):
        legal = [a for a in range(18) if (masks[i] >> a) & 1]
        sm = np.exp(logits[i] - logits[i].max())
        sm /= sm.sum()
        probs = np.zeros(18)
        probs[legal] = sm[legal]
        probs /= probs.sum()
        np.testing.assert_allclose(got[i], probs, rtol=1e-5, atol=1e-7)
        assert got[i][[a for a in range(18) if a not in legal]].sum() == 0


def test_mlp_policy_shape_follows_thesis_flagfile():
    net = MLPPolicy()
    sizes = [(m.in_features, m.out_features) for m in net.net if isinstance(m, torch.nn.Linear)]
    assert sizes == [(2492, 1024), (1024, 1024), (1024, 18)]
    assert net(torch.zeros(3, 2492)).shape == (3, 18)
    padded = MLPPolicy(padded_input_size=2496)          # GEMM-aligned input rows; pad columns are always zero
    assert padded.net[0].in_features == 2496 and padded(torch.zeros(3, 2496)).shape == (3, 18)


def test_mlp_policy_accepts_the_live_prefix():
    """Rows cut to the live prefix (1728 columns: everything beyond element 1700 is always zero) give the same logits and
    the same gradients as full rows; anything narrower is refused."""
    from open_spiel_coup_b200.selfplay import LIVE_INFO_STATE_SIZE
    torch.manual_seed(0)
    for pad in (None, 2496):
        net = MLPPolicy(hidden_sizes=(64, 32), padded_input_size=pad)
        x = torch.zeros(7, net.padded_input_size)
        x[:, :1700] = (torch.rand(7, 1700) < 0.02).float()
        full = net.net(x)
        assert torch.allclose(net(x), full, atol=1e-6)
        assert torch.allclose(net(x[:, :LIVE_INFO_STATE_SIZE].contiguous()), full, atol=1e-6)
        with torch.no_grad():
            assert torch.allclose(net(x), full, atol=1e-6)            # inference path: cut to the live prefix
        net.zero_grad()
        net(x[:, :LIVE_INFO_STATE_SIZE].contiguous()).sum().backward()
        g_live = net.net[0].weight.grad.clone()
        net.zero_grad()
        net.net(x).sum().backward()
        assert torch.allclose(g_live, net.net[0].weight.grad, atol=1e-6)
        with pytest.raises(ValueError):
            net(torch.zeros(2, 1000))


def test_reservoir_fills_then_samples_uniformly():
    cap, batches, b = 512, 60, 256
    buf = ReservoirBuffer(cap, torch.device("cpu"), info_dtype=torch.int64, seed=3)
    for k in range(batches):
        ids = torch.arange(k * b, (k + 1) * b)
        info = ids.view(-1, 1).expand(-1, 2492)
        buf.add(info, torch.zeros(b, 18), torch.ones(b, dtype=torch.int32))
        if k == 0:
            assert buf.size == b and torch.equal(buf.info_state[:b, 0], ids)       # plain fill first
    assert buf.size == cap and buf.add_calls == batches * b
    kept = buf.info_state[:, 0].double()
    total = batches * b
    # a uniform sample of 0..total-1: mean total/2, every tenth of the stream represented
    assert abs(kept.mean().item() - total / 2) < 4 * total / np.sqrt(12 * cap)
    hist = torch.histc(kept, bins=10, min=0, max=total)
    assert (hist > cap / 10 * 0.5).all()
    assert len(torch.unique(buf.info_state[:, 0])) == cap                            # no duplicates
    s_info, s_probs, s_mask = buf.sample(32)
    assert s_info.shape == (32, 2492) and s_probs.shape == (32, 18) and s_mask.shape == (32,)


def test_reservoir_valid_mask():
    buf = ReservoirBuffer(100, torch.device("cpu"), info_dtype=torch.int64)
    info = torch.arange(50).view(-1, 1).expand(-1, 2492)
    valid = torch.arange(50) % 2 == 0
    buf.add(info, torch.zeros(50, 18), torch.ones(50, dtype=torch.int32), valid=valid)
    assert buf.size == 25 and torch.equal(buf.info_state[:25, 0], torch.arange(0, 50, 2))
