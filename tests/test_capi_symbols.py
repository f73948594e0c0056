"""CPU tests of the boundary: the CUDA library cross-compiles for sm_100a, loads, exports every symbol
include/coup_b200.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from open_spiel_coup_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "coup_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(coup_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree(lib):
    from open_spiel_coup_b200 import _lib
    declared = _declared_functions()
    assert declared == sorted(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"libcoup_b200.so does not export {name}"


def test_header_compiles_as_c():
    src = '#include "coup_b200.h"\nint main(void){coup_vec_opts o; (void)o; return COUP_STATS_LEN==32?0:1;}\n'
    exe = "/tmp/coup_b200_hdr_test"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe],
                   input=src.encode(), check=True)
    assert subprocess.run([exe]).returncode == 0


def test_binary_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "open_spiel_coup_b200", "libcoup_b200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_device_means_error_not_fallback(lib):
    from open_spiel_coup_b200 import _lib
    if lib.coup_device_count() > 0:
        pytest.skip("a CUDA device is present")
    opts = _lib.VecOpts(16, 0, 1, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.coup_vec_create(C.byref(opts), C.byref(h))
    assert rc == _lib.ERR_NO_DEVICE
    assert not h
    assert b"no CPU path" in lib.coup_last_error()


def test_product_never_touches_oracle():
    """The oracle is test tooling: nothing under open_spiel_coup_b200/ may import, include or link it."""
    pkg = os.path.join(ROOT, "open_spiel_coup_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                code = "\n".join(ln for ln in text.split("\n") if not ln.strip().startswith(("//", "#", "*", "/*")))
                assert "coup_oracle" not in code and "libcoup_ref" not in code and "from oracle" not in code \
                    and "import oracle" not in code, f"{f} references the oracle"


def test_cpp_wrapper_compiles_links_and_refuses_without_device(lib):
    """include/coup_b200.hpp (RAII layer over the C ABI) builds against the library with the host compiler."""
    src = r'''
#include "coup_b200.hpp"
#include <cstdio>
int main() {
  if (coup_device_count() == 0) {
    try { coup_b200::VectorEnv env(64); } catch (const coup_b200::Error& e) { return e.code == COUP_ERR_NO_DEVICE ? 0 : 2; }
    return 3;
  }
  coup_b200::VectorEnv env(64, 1, 0, 0, true);
  env.Rollout(10);
  auto st = env.Stats();
  return st[COUP_STAT_DECISION_STEPS] == 640 && st[COUP_STAT_ILLEGAL] == 0 ? 0 : 4;
}
'''
    exe = "/tmp/coup_b200_hpp_test"
    pkg = os.path.join(ROOT, "open_spiel_coup_b200")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-", "-o", exe,
                    "-L", pkg, "-lcoup_b200", f"-Wl,-rpath,{pkg}"], input=src.encode(), check=True)
    assert subprocess.run([exe]).returncode == 0
