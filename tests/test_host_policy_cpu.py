"""CPU test of the host-side uniform policy of the host-buffer path (coup_host_sample_uniform; no GPU involved):
every action equals a NumPy statement of the device sampler's rule -- Philox4x32-10 (Salmon et al., SC'11) with key
(env lo, env hi ^ seed lo) and counter (step lo, step hi, 0, seed hi), first output word x, k = floor(x * n_legal / 2^32),
the k-th set bit of the legal mask (coup_device.cuh: env_random, sample_action) -- for every vector path the host has
(AVX-512 / AVX2 / scalar, picked at run time; the scalar path also where the 32-bit env id wraps inside a tile)."""
import ctypes as C

import numpy as np
import pytest

from open_spiel_coup_b200 import _lib


def _philox_x(seed, env, step):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    env = env.astype(np.uint64)
    c = [np.full(env.shape, step & 0xFFFFFFFF, np.uint64), np.full(env.shape, step >> 32, np.uint64),
         np.zeros(env.shape, np.uint64), np.full(env.shape, seed >> 32, np.uint64)]
    k = [env & mask, ((env >> np.uint64(32)) ^ np.uint64(seed & 0xFFFFFFFF)) & mask]
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & mask, p0 & mask]
        k = [(k[0] + np.uint64(W0)) & mask, (k[1] + np.uint64(W1)) & mask]
    return c[0]


def _expected(words, seed, offset, step):
    n = len(words)
    x = _philox_x(seed, np.uint64(offset) + np.arange(n, dtype=np.uint64), step)
    out = np.full(n, 0xFF, np.uint8)
    for i, w in enumerate(words):
        m = int(w) & 0x3FFFF
        if m:
            bits = [a for a in range(18) if (m >> a) & 1]
            out[i] = bits[(int(x[i]) * len(bits)) >> 32]
    return out


@pytest.mark.parametrize("isa", ["native", "avx2", "scalar"])
@pytest.mark.parametrize("n,offset,threads", [(5000, 0, 1), (5000, 1 << 20, 3), (4099, (1 << 32) - 1000, 2), (17, (5 << 32) + 7, 1),
                                              (40000, (1 << 40) + 123, 8)])
def test_host_sampler_equals_numpy_statement(n, offset, threads, isa, monkeypatch):
    if isa == "native":
        monkeypatch.delenv("COUP_B200_HOST_ISA", raising=False)
    else:
        monkeypatch.setenv("COUP_B200_HOST_ISA", isa)          # read on every call: restricts the vector paths
    lib = _lib.load()
    rng = np.random.default_rng(n)
    words = rng.integers(0, 1 << 18, size=n).astype(np.uint32)
    words[::97] = 0                                            # terminal envs: no legal action -> 0xFF
    words |= rng.integers(0, 1 << 9, size=n).astype(np.uint32) << 18   # the other fields of a step word are ignored
    acts = np.zeros(n, np.uint8)
    for seed, step in ((1234, 0), ((0xABCDEF << 32) | 0x13579B, (3 << 32) | 77)):
        rc = lib.coup_host_sample_uniform(C.c_void_p(words.ctypes.data), n, seed, offset, step, C.c_void_p(acts.ctypes.data), threads)
        assert rc == 0
        np.testing.assert_array_equal(acts, _expected(words, seed, offset, step))


def test_host_sampler_rejects_null():
    lib = _lib.load()
    assert lib.coup_host_sample_uniform(None, 4, 0, 0, 0, None, 1) != 0
