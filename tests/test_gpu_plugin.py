"""GPU test (-m gpu): the reference's OWN test binary -- open_spiel/games/coup_test.cc (14 scenario tests)
and tests/basic_tests.cc (LoadGameTest, ChanceOutcomesTest, RandomSimTest x100, RandomSimTestCustomObserver
for the default and info-state observers) -- compiled UNMODIFIED against the GPU-backed "coup" plugin
(plugin/coup_b200_plugin.cc) instead of the reference's games/coup.cc. Built by `make -C plugin` where the
reference headers exist; skipped if the binary did not travel."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "plugin", "coup_test_b200")


def test_reference_coup_test_binary_passes_on_the_plugin():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/plugin/coup_test_b200 not built (needs /root/reference at build time)")
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    tail = (res.stdout + res.stderr)[-2000:]
    assert res.returncode == 0, tail
    assert "Spiel Fatal Error" not in tail
