"""ctypes binding of the C ABI declared in include/coup_b200.h (libcoup_b200.so).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded this module
raises, it never substitutes a Python/NumPy implementation of the rules.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcoup_b200.so")

# constants mirrored from include/coup_b200.h
NUM_PLAYERS = 2
NUM_DISTINCT_ACTIONS = 18
MAX_CHANCE_OUTCOMES = 5
MAX_GAME_LENGTH = 90
MAX_CHANCE_NODES_IN_HISTORY = 45
INFO_STATE_SIZE = 2492
LIVE_INFO_STATE_SIZE = 1728      # COUP_LIVE_INFO_STATE_SIZE: elements >= 62 + 18 * 91 = 1700 are always zero
OBSERVATION_SIZE = 98
MIN_UTILITY = -2.0
MAX_UTILITY = 2.0
CHANCE_PLAYER_ID = -1
TERMINAL_PLAYER_ID = -4
STATE_WORDS = 4
HISTORY_WORDS = 16
RECORD_WORDS = 24

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_ILLEGAL_ACTION = 0, 1, 2, 3, 4
FLAG_AUTO_RESET = 1
FLAG_PLAIN_STORE_ENCODER = 2
FLAG_NO_WARP_SPECIALISATION = 4
FLAG_BLOCKING_SYNC = 8
PLAYER_0, PLAYER_1, PLAYER_CURRENT, PLAYER_BOTH, PLAYER_FROM_RECORD = 0, 1, 2, 3, 4
DTYPE_F32, DTYPE_U8, DTYPE_BF16 = 0, 1, 2
STAT_DECISION_STEPS, STAT_CHANCE_MOVES, STAT_EPISODES, STAT_TRUNCATED = 0, 1, 2, 3
STAT_EPISODE_MOVES, STAT_ILLEGAL, STAT_RETURN_HIST, STAT_LEGAL_HIST, STATS_LEN = 4, 5, 8, 16, 32

# Every symbol include/coup_b200.h declares (tests check that the built library exports all of them).
EXPORTED_SYMBOLS = [
    "coup_last_error", "coup_device_count", "coup_vec_create", "coup_vec_destroy", "coup_vec_num_envs",
    "coup_vec_reset", "coup_vec_step", "coup_vec_new_initial_state", "coup_vec_apply_move", "coup_vec_copy_env", "coup_vec_fork", "coup_cfr_expand", "coup_cfr_children", "coup_env_new_initial_state", "coup_env_apply_action", "coup_env_clone", "coup_env_read",
    "coup_env_information_state_tensor", "coup_env_observation_tensor",
    "coup_vec_sample_uniform", "coup_vec_sample_policy", "coup_vec_rollout", "coup_vec_rollout_incremental",
    "coup_vec_legal_mask", "coup_vec_current_player", "coup_vec_done", "coup_vec_rewards",
    "coup_vec_returns", "coup_vec_step_word", "coup_vec_state", "coup_vec_history", "coup_vec_legal_actions_mask",
    "coup_vec_information_state_tensor", "coup_vec_observation_tensor",
    "coup_vec_information_state_tensor_strided", "coup_vec_rollout_strided", "coup_vec_information_state_tensor_gather", "coup_vec_step_host", "coup_vec_step_host_packed",
    "coup_host_sample_uniform", "coup_vec_stats", "coup_vec_stats_device", "coup_vec_clear_stats", "coup_vec_check_errors",
    "coup_vec_finished_ring_enable", "coup_vec_finished_ring", "coup_vec_finished_ring_ctrl", "coup_vec_finished_ring_capacity",
    "coup_vec_finished_drain", "coup_vec_finished_information_state_tensor", "coup_records_information_state_tensor",
    "coup_vec_finished_observation_tensor", "coup_records_observation_tensor", "coup_vec_observation_tensor_gather",
    "coup_vec_step_record", "coup_vec_observer_tensor", "coup_vec_step_host_packed_async", "coup_vec_host_outputs_wait", "coup_vec_information_state_tensor_prefix", "coup_cfr_level",
    "coup_vec_fork_counted", "coup_vec_pack_records", "coup_cfr_backward",
    "coup_tensor_row_hash", "coup_vec_snapshot_size", "coup_vec_snapshot", "coup_vec_restore", "coup_vec_step_counter", "coup_vec_set_step_counter",
]


class CoupError(RuntimeError):
    """Raised where the reference would raise SpielError / abort (pyspiel.cc:620-626)."""


class VecOpts(C.Structure):
    _fields_ = [
        ("num_envs", C.c_uint32),
        ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("global_env_offset", C.c_uint64),
        ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class RecorderBuffers(C.Structure):
    """coup_recorder_buffers (include/coup_b200.h)."""
    _fields_ = [
        ("d_reservoir_records", C.c_void_p), ("d_reservoir_probs", C.c_void_p), ("d_reservoir_winner", C.c_void_p),
        ("reservoir_capacity", C.c_uint64), ("reservoir_offered", C.c_uint64),
        ("d_transitions", C.c_void_p), ("replay_capacity", C.c_uint64), ("d_replay_total", C.c_void_p),
        ("d_pending", C.c_void_p),
    ]


_lib = None


def load():
    """Loads libcoup_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoupError(
            f"{LIB_PATH} is missing: build it with `python -m open_spiel_coup_b200.build` "
            "(there is no CPU fallback for the Coup environment)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, u8p = C.c_void_p, C.c_void_p
    lib.coup_last_error.restype = C.c_char_p
    lib.coup_device_count.restype = C.c_int
    lib.coup_vec_create.argtypes = [C.POINTER(VecOpts), C.POINTER(vp)]
    lib.coup_vec_destroy.argtypes = [vp]
    lib.coup_vec_num_envs.argtypes = [vp]
    lib.coup_vec_num_envs.restype = C.c_uint32
    lib.coup_vec_reset.argtypes = [vp, u8p, u8p, vp]
    lib.coup_vec_step.argtypes = [vp, u8p, u8p, vp]
    lib.coup_vec_new_initial_state.argtypes = [vp, u8p, vp]
    lib.coup_vec_apply_move.argtypes = [vp, u8p, vp]
    lib.coup_vec_copy_env.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    lib.coup_vec_fork.argtypes = [vp, vp, vp, u8p, u8p, C.c_uint32, vp]
    lib.coup_cfr_expand.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_uint64, C.c_uint64, vp, vp, vp, vp]
    lib.coup_cfr_children.argtypes = [vp, vp, C.c_uint32, vp, vp, vp]
    lib.coup_env_new_initial_state.argtypes = [vp, C.c_uint32]
    lib.coup_env_apply_action.argtypes = [vp, C.c_uint32, C.c_int]
    lib.coup_env_clone.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.coup_env_read.argtypes = [vp, C.c_uint32, vp, vp, vp]
    lib.coup_env_information_state_tensor.argtypes = [vp, C.c_uint32, C.c_int, vp, C.c_int]
    lib.coup_env_observation_tensor.argtypes = [vp, C.c_uint32, C.c_int, vp, C.c_int]
    lib.coup_vec_sample_uniform.argtypes = [vp, u8p, vp]
    lib.coup_vec_sample_policy.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    lib.coup_vec_rollout_incremental.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_vec_rollout.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    for name in ("coup_vec_legal_mask", "coup_vec_current_player", "coup_vec_done", "coup_vec_rewards",
                 "coup_vec_returns", "coup_vec_state", "coup_vec_history", "coup_vec_stats_device", "coup_vec_step_word",
                 "coup_vec_finished_ring", "coup_vec_finished_ring_ctrl"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = vp
    lib.coup_vec_legal_actions_mask.argtypes = [vp, vp, vp]
    lib.coup_vec_finished_ring_enable.argtypes = [vp, C.c_uint32]
    lib.coup_vec_finished_ring_capacity.argtypes = [vp]
    lib.coup_vec_finished_ring_capacity.restype = C.c_uint32
    lib.coup_vec_finished_drain.argtypes = [vp, vp, C.c_uint32, vp, vp, vp]
    lib.coup_vec_finished_information_state_tensor.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, C.c_uint32, vp, vp, vp]
    lib.coup_vec_finished_observation_tensor.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp, vp, vp]
    lib.coup_records_observation_tensor.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp]
    lib.coup_vec_observation_tensor_gather.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp]
    lib.coup_vec_observer_tensor.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.coup_vec_information_state_tensor_prefix.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_cfr_level.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_uint64, C.c_uint64,
                                   vp, vp, vp, vp, vp, vp, vp, vp]
    lib.coup_vec_fork_counted.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, vp]
    lib.coup_vec_pack_records.argtypes = [vp, vp, vp, vp]
    lib.coup_cfr_backward.argtypes = [vp, vp, C.c_uint32, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.coup_vec_step_record.argtypes = [vp, vp, vp, C.POINTER(RecorderBuffers), vp]
    lib.coup_records_information_state_tensor.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_vec_information_state_tensor.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    lib.coup_vec_information_state_tensor_strided.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_vec_information_state_tensor_gather.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_vec_rollout_strided.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_uint32, vp]
    lib.coup_vec_observation_tensor.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    lib.coup_vec_step_host.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, vp]
    lib.coup_vec_step_host_packed.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    lib.coup_vec_step_host_packed_async.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    lib.coup_vec_host_outputs_wait.argtypes = [vp]
    lib.coup_host_sample_uniform.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_int]
    lib.coup_vec_stats.argtypes = [vp, vp, vp]
    lib.coup_vec_clear_stats.argtypes = [vp, vp]
    lib.coup_vec_check_errors.argtypes = [vp, vp]
    lib.coup_tensor_row_hash.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp]
    lib.coup_vec_snapshot_size.argtypes = [vp]
    lib.coup_vec_snapshot_size.restype = C.c_size_t
    lib.coup_vec_snapshot.argtypes = [vp, vp, C.c_size_t, vp]
    lib.coup_vec_restore.argtypes = [vp, vp, C.c_size_t, vp]
    lib.coup_vec_step_counter.argtypes = [vp]
    lib.coup_vec_step_counter.restype = C.c_uint64
    lib.coup_vec_set_step_counter.argtypes = [vp, C.c_uint64]
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise CoupError(f"libcoup_b200 error {rc}: {load().coup_last_error().decode()}")
