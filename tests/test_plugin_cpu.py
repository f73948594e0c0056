"""CPU test: the plugin-backed reference test binary must REFUSE to run without a GPU (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "plugin", "coup_test_b200")


def test_plugin_binary_has_no_cpu_path():
    if os.path.isdir("/root/reference/open_spiel"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-j8"], check=True, capture_output=True)
    if not os.path.exists(BIN):
        pytest.skip("plugin test binary not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stdout + res.stderr)
