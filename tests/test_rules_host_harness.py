"""CPU tests of the branch-free device rules: tests/host_harness compiles open_spiel_coup_b200/csrc/coup_device.cuh for
the host (test tooling only -- the product library has no host copy of the rules) and every state of many random,
steered and reference-recorded games is compared with the oracle and with the compiled unmodified reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness", "rules_harness.cc")
OUT_DIR = os.path.join(ROOT, "tests", "host_harness", "_build")
SO = os.path.join(OUT_DIR, "librules_host.so")
FIELDS = ("cur_player", "is_terminal", "is_chance", "move_number", "legal_mask", "rewards", "returns", "coins", "ncards")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(OUT_DIR, exist_ok=True)
    dev = os.path.join(ROOT, "open_spiel_coup_b200", "csrc", "coup_device.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(dev)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I/usr/local/cuda/include", SRC, "-o", SO])
    lib = C.CDLL(SO)
    lib.hh_random_games.restype = C.c_long
    lib.hh_random_games.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_long]
    lib.hh_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for name in ("hh_check_sample_card", "hh_check_closed_form_deal", "hh_check_history_commit"):
        getattr(lib, name).restype = C.c_long
        getattr(lib, name).argtypes = [C.c_long, C.c_uint64]
    lib.hh_check_kth_set_bit.restype = C.c_long
    lib.hh_check_kth_set_bit.argtypes = [C.c_int]
    lib.hh_check_hand_insert.restype = C.c_long
    return lib


def _games(lib, n, seed, steer):
    cap = n * 100
    actions = np.zeros(cap, np.uint8)
    offsets = np.zeros(n + 1, np.int64)
    total = lib.hh_random_games(n, seed, steer, actions.ctypes.data, offsets.ctypes.data, cap)
    assert total > 0
    return actions[:total].copy(), offsets


def _harness_trace(lib, actions, offsets):
    from oracle.bindings import TRACE_DTYPE
    n_traj = len(offsets) - 1
    out = np.zeros(len(actions) + n_traj, TRACE_DTYPE)
    hist_ok = C.c_int(1)
    a = np.ascontiguousarray(actions, np.uint8)
    off = np.ascontiguousarray(offsets, np.int64)
    bad = lib.hh_trace_batch(a.ctypes.data, off.ctypes.data, n_traj, out.ctypes.data, C.byref(hist_ok))
    return out, bad, hist_ok.value


def _compare(got, want):
    for f in FIELDS:
        same = got[f] == want[f]
        assert same.all(), (f, int(np.argmin(same.reshape(len(got), -1).all(1))))


def test_building_blocks(harness):
    assert harness.hh_check_kth_set_bit(18) == 0
    assert harness.hh_check_hand_insert() == 0
    assert harness.hh_check_sample_card(2_000_000, 7) == 0
    assert harness.hh_check_closed_form_deal(2_000_000, 11) == 0
    assert harness.hh_check_history_commit(1_000_000, 13) == 0


@pytest.mark.parametrize("steer", [0, 1], ids=["uniform", "steered-91-move"])
def test_random_games_match_oracle(harness, oracle, steer):
    n = 60000 if not steer else 3000
    actions, offsets = _games(harness, n, 1234 + steer, steer)
    lens = np.diff(offsets)
    if steer:
        assert (lens > 90).sum() > 20                      # the move-cap truncation is exercised
    else:
        assert 20.5 < lens.mean() < 21.9                   # 21.2 moves per uniform-random episode
    want, bad = oracle.trace_batch(actions, offsets)
    assert bad == 0                                        # every move the harness rules thought legal is legal
    got, bad2, hist_ok = _harness_trace(harness, actions, offsets)
    assert bad2 == 0 and hist_ok == 1
    _compare(got, want)


def test_reference_trajectories(harness, ref_trajectories):
    """The committed trajectories recorded from the compiled reference (tests/golden), record by record."""
    actions, offsets, records = ref_trajectories
    got, bad, hist_ok = _harness_trace(harness, actions, offsets)
    assert bad == 0 and hist_ok == 1
    _compare(got, records)


def test_random_games_match_compiled_reference(harness, reference):
    actions, offsets = _games(harness, 20000, 777, 0)
    a2, o2 = _games(harness, 1000, 778, 1)
    actions = np.concatenate([actions, a2])
    offsets = np.concatenate([offsets, o2[1:] + offsets[-1]])
    want, bad = reference.trace_batch(actions, offsets)
    assert bad == 0
    got, bad2, hist_ok = _harness_trace(harness, actions, offsets)
    assert bad2 == 0 and hist_ok == 1
    _compare(got, want)


def test_product_library_has_no_host_rules():
    """COUP_RULES_HOST_TEST is test tooling: the shipped library must not export or contain host versions of the rules."""
    so = os.path.join(ROOT, "open_spiel_coup_b200", "libcoup_b200.so")
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "apply_player_action" not in syms and "legal_mask_decision" not in syms
    csrc = os.path.join(ROOT, "open_spiel_coup_b200", "csrc")
    for f in [f for f in os.listdir(csrc) if f != "coup_device.cuh"]:
        assert "COUP_RULES_HOST_TEST" not in open(os.path.join(ROOT, "open_spiel_coup_b200", "csrc", f)).read()
    assert "COUP_RULES_HOST_TEST" not in open(os.path.join(ROOT, "open_spiel_coup_b200", "build.py")).read()
