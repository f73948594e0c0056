"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes bindings to (a) oracle/libcoup_oracle.so, the plain-C restatement of the reference's Coup path
(coup_oracle.c), and (b) oracle/_ref/libcoup_ref.so, the UNMODIFIED reference compiled from
/root/reference by oracle/Makefile. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(open_spiel_coup_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libcoup_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcoup_ref.so")

INFO_SIZE = 2492
OBS_SIZE = 98
NUM_ACTIONS = 18
HIST_CAP = 96

TRACE_DTYPE = np.dtype(
    [
        ("cur_player", np.int8),
        ("is_terminal", np.uint8),
        ("is_chance", np.uint8),
        ("move_number", np.uint8),
        ("legal_mask", np.uint32),
        ("rewards", np.int8, (2,)),
        ("returns", np.int8, (2,)),
        ("coins", np.uint8, (2,)),
        ("ncards", np.uint8, (2,)),
        ("hash_info", np.uint64, (2,)),
        ("hash_obs", np.uint64, (2,)),
    ],
    align=True,
)
assert TRACE_DTYPE.itemsize == 48, TRACE_DTYPE.itemsize


def build_oracle(force=False):
    """Compile the C restatement (gcc, <1 s)."""
    src = os.path.join(HERE, "coup_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "coup_oracle.h"))
    ):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref():
    """Compile the reference where /root/reference exists; returns the .so path or None."""
    if os.path.isdir("/root/reference/open_spiel"):
        subprocess.check_call(["make", "-C", HERE, "ref", "-j8"], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class OcCard(C.Structure):
    _fields_ = [("value", C.c_int32), ("state", C.c_int32)]


class OcPlayer(C.Structure):
    _fields_ = [
        ("cards", OcCard * 4),
        ("num_cards", C.c_int32),
        ("coins", C.c_int32),
        ("last_action", C.c_int32),
        ("lost_challenge", C.c_int32),
    ]


class OcState(C.Structure):
    _fields_ = [
        ("deck", C.c_int32 * 5),
        ("players", OcPlayer * 2),
        ("deal_queue", C.c_int32 * 8),
        ("deal_head", C.c_int32),
        ("deal_tail", C.c_int32),
        ("cur_player_turn", C.c_int32),
        ("cur_player_move", C.c_int32),
        ("opp_player", C.c_int32),
        ("is_turn_begin", C.c_int32),
        ("turn_number", C.c_int32),
        ("is_chance", C.c_int32),
        ("cur_rewards", C.c_int32 * 2),
        ("move_number", C.c_int32),
        ("history_len", C.c_int32),
        ("history_player", C.c_int8 * HIST_CAP),
        ("history_action", C.c_int8 * HIST_CAP),
        ("history_deal_player", C.c_int8 * HIST_CAP),
        ("error", C.c_int32),
    ]


class Oracle:
    """The C restatement (always available: compiled on demand with gcc)."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        assert L.oc_sizeof_state() == C.sizeof(OcState), (L.oc_sizeof_state(), C.sizeof(OcState))
        P = C.POINTER
        L.oc_init.argtypes = [P(OcState)]
        L.oc_is_terminal.argtypes = [P(OcState)]
        L.oc_current_player.argtypes = [P(OcState)]
        L.oc_is_chance_node.argtypes = [P(OcState)]
        L.oc_apply_action.argtypes = [P(OcState), C.c_int]
        L.oc_apply_action_checked.argtypes = [P(OcState), C.c_int]
        L.oc_legal_actions.argtypes = [P(OcState), P(C.c_int32)]
        L.oc_legal_mask.argtypes = [P(OcState)]
        L.oc_legal_mask.restype = C.c_uint32
        L.oc_chance_outcomes.argtypes = [P(OcState), P(C.c_int32), P(C.c_double)]
        L.oc_returns.argtypes = [P(OcState), P(C.c_double)]
        L.oc_rewards.argtypes = [P(OcState), P(C.c_double)]
        L.oc_information_state_tensor.argtypes = [P(OcState), C.c_int, C.c_void_p]
        L.oc_observation_tensor.argtypes = [P(OcState), C.c_int, C.c_void_p]
        L.oc_observer_tensor.argtypes = [P(OcState), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oc_tensor_hash.argtypes = [C.c_void_p, C.c_int]
        L.oc_tensor_hash.restype = C.c_uint64
        L.oc_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oc_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oc_state_from_actions.argtypes = [P(OcState), C.c_void_p, C.c_int]
        L.oc_final_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oc_final_tensors_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_replay_digest_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_batch_init.argtypes = [C.c_void_p, C.c_int]
        L.oc_batch_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_batch_info_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_batch_observation.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oc_bench.argtypes = [C.c_int, C.c_int, C.c_long, C.c_uint32, C.c_void_p]

    # -- single state -------------------------------------------------------------------------
    def new_state(self):
        s = OcState()
        self.lib.oc_init(C.byref(s))
        return s

    def state_from_actions(self, actions):
        s = OcState()
        a = _u8(actions)
        rc = self.lib.oc_state_from_actions(C.byref(s), a.ctypes.data, len(a))
        if rc != 0:
            raise ValueError(f"oracle rejected action list (rc={rc})")
        return s

    def apply(self, s, a, checked=True):
        f = self.lib.oc_apply_action_checked if checked else self.lib.oc_apply_action
        return f(C.byref(s), int(a))

    def is_terminal(self, s):
        return bool(self.lib.oc_is_terminal(C.byref(s)))

    def current_player(self, s):
        return self.lib.oc_current_player(C.byref(s))

    def legal_actions(self, s):
        out = (C.c_int32 * NUM_ACTIONS)()
        n = self.lib.oc_legal_actions(C.byref(s), out)
        if n < 0:
            raise ValueError("oracle: LegalActions() would SpielFatalError here")
        return [int(out[i]) for i in range(n)]

    def legal_mask(self, s):
        return int(self.lib.oc_legal_mask(C.byref(s)))

    def chance_outcomes(self, s):
        a = (C.c_int32 * 5)()
        p = (C.c_double * 5)()
        n = self.lib.oc_chance_outcomes(C.byref(s), a, p)
        return [(int(a[i]), float(p[i])) for i in range(n)]

    def returns(self, s):
        out = (C.c_double * 2)()
        self.lib.oc_returns(C.byref(s), out)
        return [out[0], out[1]]

    def rewards(self, s):
        out = (C.c_double * 2)()
        self.lib.oc_rewards(C.byref(s), out)
        return [out[0], out[1]]

    def info_state(self, s, player):
        out = np.empty(INFO_SIZE, np.float32)
        self.lib.oc_information_state_tensor(C.byref(s), player, out.ctypes.data)
        return out

    def observation(self, s, player):
        out = np.empty(OBS_SIZE, np.float32)
        self.lib.oc_observation_tensor(C.byref(s), player, out.ctypes.data)
        return out

    def observer_tensor(self, s, player, public_info, perfect_recall, private_info):
        out = np.empty(INFO_SIZE, np.float32)
        n = self.lib.oc_observer_tensor(
            C.byref(s), player, int(public_info), int(perfect_recall), int(private_info), out.ctypes.data
        )
        return out[:n].copy()

    def tensor_hash(self, t):
        t = np.ascontiguousarray(t, np.float32)
        return int(self.lib.oc_tensor_hash(t.ctypes.data, t.size))

    # -- traces -------------------------------------------------------------------------------
    def trace(self, actions):
        a = _u8(actions)
        out = np.zeros(len(a) + 1, TRACE_DTYPE)
        n = self.lib.oc_trace(a.ctypes.data, len(a), out.ctypes.data)
        if n < 0:
            raise ValueError(f"oracle rejected move {-n - 1}")
        return out

    def trace_batch(self, actions, offsets):
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        out = np.zeros(len(a) + n_traj, TRACE_DTYPE)
        bad = self.lib.oc_trace_batch(a.ctypes.data, off.ctypes.data, n_traj, out.ctypes.data)
        return out, bad

    def final_batch(self, actions, offsets):
        """Record of the last state of each trajectory; rejected trajectories have cur_player 127."""
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        out = np.zeros(n_traj, TRACE_DTYPE)
        bad = self.lib.oc_final_batch(a.ctypes.data, off.ctypes.data, n_traj, out.ctypes.data)
        return out, bad

    def final_tensors_batch(self, actions, offsets, info=True, obs=True):
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        ti = np.zeros((n_traj, 2, INFO_SIZE), np.float32) if info else None
        to = np.zeros((n_traj, 2, OBS_SIZE), np.float32) if obs else None
        bad = self.lib.oc_final_tensors_batch(
            a.ctypes.data, off.ctypes.data, n_traj,
            ti.ctypes.data if info else None, to.ctypes.data if obs else None)
        return ti, to, bad

    def replay_digest_batch(self, actions, offsets, threads=1):
        """(digests uint64[n], reports int32[n], rejected) -- see ref_replay_digest in ref_harness.cc."""
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        dig = np.zeros(n_traj, np.uint64)
        rep = np.zeros(n_traj, np.int32)
        bad = self.lib.oc_replay_digest_batch(a.ctypes.data, off.ctypes.data, n_traj, dig.ctypes.data, rep.ctypes.data)
        return dig, rep, bad

    # -- batches ------------------------------------------------------------------------------
    def batch_new(self, n):
        arr = (OcState * n)()
        self.lib.oc_batch_init(arr, n)
        return arr

    def batch_apply(self, arr, actions, do_mask=None):
        a = _u8(actions)
        m = None if do_mask is None else _u8(do_mask)
        return self.lib.oc_batch_apply(arr, len(arr), a.ctypes.data, None if m is None else m.ctypes.data)

    def batch_info_state(self, arr, players):
        p = np.ascontiguousarray(players, np.int8)
        out = np.empty((len(arr), INFO_SIZE), np.float32)
        self.lib.oc_batch_info_state(arr, len(arr), p.ctypes.data, out.ctypes.data)
        return out

    def batch_observation(self, arr, players):
        p = np.ascontiguousarray(players, np.int8)
        out = np.empty((len(arr), OBS_SIZE), np.float32)
        self.lib.oc_batch_observation(arr, len(arr), p.ctypes.data, out.ctypes.data)
        return out

    def bench(self, mode, threads, episodes, seed=1234):
        out = np.zeros(21, np.float64)
        self.lib.oc_bench(mode, threads, episodes, seed, out.ctypes.data)
        return _bench_dict(out, "port", threads, mode)


def _bench_dict(out, kind, threads, mode):
    secs = float(out[0])
    return {
        "kind": kind,
        "mode": mode,
        "threads": threads,
        "seconds": secs,
        "moves": int(out[1]),
        "decisions": int(out[2]),
        "chance": int(out[3]),
        "episodes": int(out[4]),
        "truncated": int(out[5]),
        "returns_hist_p0": [int(x) for x in out[6:11]],
        "legal_count_hist": [int(x) for x in out[11:19]],
        "max_moves": int(out[19]),
        "max_coins": int(out[20]),
        "moves_per_s": out[1] / secs if secs > 0 else 0.0,
        "decisions_per_s": out[2] / secs if secs > 0 else 0.0,
        "episodes_per_s": out[4] / secs if secs > 0 else 0.0,
    }


class Reference:
    """The unmodified reference (oracle/_ref/libcoup_ref.so). available() is False where it was never
    built (it needs /root/reference at build time; the built .so travels with the repo snapshot)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} not built (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        vp = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_new_state.restype = vp
        L.ref_free_state.argtypes = [vp]
        L.ref_clone_state.argtypes = [vp]
        L.ref_clone_state.restype = vp
        L.ref_apply_action.argtypes = [vp, C.c_long]
        for f in ("ref_current_player", "ref_is_terminal", "ref_is_chance", "ref_move_number"):
            getattr(L, f).argtypes = [vp]
        L.ref_legal_actions.argtypes = [vp, vp]
        L.ref_chance_outcomes.argtypes = [vp, vp, vp]
        L.ref_returns.argtypes = [vp, vp]
        L.ref_rewards.argtypes = [vp, vp]
        L.ref_information_state_tensor.argtypes = [vp, C.c_int, vp, C.c_int]
        L.ref_observation_tensor.argtypes = [vp, C.c_int, vp, C.c_int]
        L.ref_observer_tensor.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.ref_to_string.argtypes = [vp, vp, C.c_int]
        L.ref_information_state_string.argtypes = [vp, C.c_int, vp, C.c_int]
        L.ref_observation_string.argtypes = [vp, C.c_int, vp, C.c_int]
        L.ref_observer_string.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.ref_serialize.argtypes = [vp, vp, C.c_int]
        L.ref_action_to_string.argtypes = [C.c_int, C.c_long, vp, C.c_int]
        L.ref_history.argtypes = [vp, vp, vp, C.c_int]
        L.ref_get_cards.argtypes = [vp, C.c_int, vp, vp]
        L.ref_get_coins.argtypes = [vp, C.c_int]
        L.ref_get_last_action.argtypes = [vp, C.c_int]
        L.ref_tensor_hash.argtypes = [vp, C.c_int]
        L.ref_tensor_hash.restype = C.c_uint64
        L.ref_trace.argtypes = [vp, C.c_int, vp]
        L.ref_trace_batch.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.ref_replay_digest_batch.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int]
        L.ref_deserialize_state.argtypes = [C.c_char_p]
        L.ref_deserialize_state.restype = vp
        L.ref_state_from_actions.argtypes = [vp, C.c_int]
        L.ref_state_from_actions.restype = vp
        L.ref_bench.argtypes = [C.c_int, C.c_int, C.c_long, C.c_uint32, vp]
        L.ref_game_constants.argtypes = [vp]

    def last_error(self):
        return self.lib.ref_last_error().decode()

    def game_constants(self):
        out = np.zeros(11, np.float64)
        self.lib.ref_game_constants(out.ctypes.data)
        keys = [
            "NumDistinctActions", "MaxChanceOutcomes", "NumPlayers", "MinUtility", "MaxUtility",
            "InformationStateTensorSize", "ObservationTensorSize", "MaxGameLength",
            "MaxChanceNodesInHistory", "MaxMoveNumber", "UtilitySum",
        ]
        return dict(zip(keys, out.tolist()))

    def new_state(self):
        return self.lib.ref_new_state()

    def free(self, h):
        self.lib.ref_free_state(h)

    def clone(self, h):
        return self.lib.ref_clone_state(h)

    def state_from_actions(self, actions):
        a = _u8(actions)
        h = self.lib.ref_state_from_actions(a.ctypes.data, len(a))
        if not h:
            raise ValueError("reference rejected action list: " + self.last_error())
        return h

    def deserialize_state(self, text):
        h = self.lib.ref_deserialize_state(text.encode())
        if not h:
            raise ValueError("reference rejected serialized state: " + self.last_error())
        return h

    def apply(self, h, a):
        return self.lib.ref_apply_action(h, int(a))

    def is_terminal(self, h):
        return bool(self.lib.ref_is_terminal(h))

    def is_chance(self, h):
        return bool(self.lib.ref_is_chance(h))

    def current_player(self, h):
        return self.lib.ref_current_player(h)

    def move_number(self, h):
        return self.lib.ref_move_number(h)

    def legal_actions(self, h):
        out = np.zeros(32, np.int64)
        n = self.lib.ref_legal_actions(h, out.ctypes.data)
        if n < 0:
            raise ValueError(self.last_error())
        return [int(x) for x in out[:n]]

    def chance_outcomes(self, h):
        a = np.zeros(8, np.int64)
        p = np.zeros(8, np.float64)
        n = self.lib.ref_chance_outcomes(h, a.ctypes.data, p.ctypes.data)
        if n < 0:
            raise ValueError(self.last_error())
        return [(int(a[i]), float(p[i])) for i in range(n)]

    def returns(self, h):
        out = np.zeros(2, np.float64)
        self.lib.ref_returns(h, out.ctypes.data)
        return out.tolist()

    def rewards(self, h):
        out = np.zeros(2, np.float64)
        self.lib.ref_rewards(h, out.ctypes.data)
        return out.tolist()

    def info_state(self, h, player):
        out = np.full(INFO_SIZE, np.nan, np.float32)
        self.lib.ref_information_state_tensor(h, player, out.ctypes.data, INFO_SIZE)
        return out

    def observation(self, h, player):
        out = np.full(OBS_SIZE, np.nan, np.float32)
        self.lib.ref_observation_tensor(h, player, out.ctypes.data, OBS_SIZE)
        return out

    def observer_tensor(self, h, player, public_info, perfect_recall, private_info):
        out = np.full(INFO_SIZE, np.nan, np.float32)
        n = self.lib.ref_observer_tensor(
            h, player, int(public_info), int(perfect_recall), int(private_info), out.ctypes.data, INFO_SIZE
        )
        assert n >= 0
        return out[:n].copy()

    def _string(self, fn, *args):
        buf = C.create_string_buffer(1 << 16)
        n = fn(*args, buf, len(buf))
        assert n >= 0
        return buf.raw[:n].decode()

    def to_string(self, h):
        return self._string(self.lib.ref_to_string, h)

    def info_state_string(self, h, player):
        return self._string(self.lib.ref_information_state_string, h, player)

    def observation_string(self, h, player):
        return self._string(self.lib.ref_observation_string, h, player)

    def observer_string(self, h, player, public_info, perfect_recall, private_info):
        return self._string(
            self.lib.ref_observer_string, h, player, int(public_info), int(perfect_recall), int(private_info)
        )

    def serialize(self, h):
        return self._string(self.lib.ref_serialize, h)

    def action_to_string(self, player, action):
        return self._string(self.lib.ref_action_to_string, player, action)

    def history(self, h):
        a = np.zeros(160, np.int64)
        p = np.zeros(160, np.int32)
        n = self.lib.ref_history(h, a.ctypes.data, p.ctypes.data, 160)
        return a[:n].tolist(), p[:n].tolist()

    def cards(self, h, player):
        v = np.zeros(4, np.int32)
        s = np.zeros(4, np.int32)
        n = self.lib.ref_get_cards(h, player, v.ctypes.data, s.ctypes.data)
        return v[:n].tolist(), s[:n].tolist()

    def coins(self, h, player):
        return self.lib.ref_get_coins(h, player)

    def last_action(self, h, player):
        return self.lib.ref_get_last_action(h, player)

    def tensor_hash(self, t):
        t = np.ascontiguousarray(t, np.float32)
        return int(self.lib.ref_tensor_hash(t.ctypes.data, t.size))

    def trace(self, actions):
        a = _u8(actions)
        out = np.zeros(len(a) + 1, TRACE_DTYPE)
        n = self.lib.ref_trace(a.ctypes.data, len(a), out.ctypes.data)
        if n < 0:
            raise ValueError(f"reference rejected move {-n - 1}: {self.last_error()}")
        return out

    def trace_batch(self, actions, offsets, threads=8):
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        out = np.zeros(len(a) + n_traj, TRACE_DTYPE)
        bad = self.lib.ref_trace_batch(a.ctypes.data, off.ctypes.data, n_traj, out.ctypes.data, threads)
        return out, bad

    def replay_digest_batch(self, actions, offsets, threads=8):
        a = _u8(actions)
        off = np.ascontiguousarray(offsets, np.int64)
        n_traj = len(off) - 1
        dig = np.zeros(n_traj, np.uint64)
        rep = np.zeros(n_traj, np.int32)
        bad = self.lib.ref_replay_digest_batch(a.ctypes.data, off.ctypes.data, n_traj, dig.ctypes.data,
                                               rep.ctypes.data, threads)
        return dig, rep, bad

    def bench(self, mode, threads, episodes, seed=1234):
        out = np.zeros(21, np.float64)
        self.lib.ref_bench(mode, threads, episodes, seed, out.ctypes.data)
        return _bench_dict(out, "reference", threads, mode)
