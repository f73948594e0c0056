"""Batched Coup VectorEnv over the C ABI (include/coup_b200.h).

Semantics follow the reference's `rl_environment.Environment.reset/step`
(open_spiel/python/rl_environment.py:282-382) applied per env and batched like
`vector_env.SyncVectorEnv` (open_spiel/python/vector_env.py:40-78): one `step` = one player action +
every following chance node, then legal mask / current player / rewards / done (+ tensors) for the next
decision. All rules run in CUDA kernels; PyTorch is only used here to own device buffers and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F32, DTYPE_U8, FLAG_AUTO_RESET, HISTORY_WORDS, INFO_STATE_SIZE, LIVE_INFO_STATE_SIZE,
                   NUM_DISTINCT_ACTIONS, OBSERVATION_SIZE, PLAYER_0, PLAYER_1, PLAYER_BOTH, PLAYER_CURRENT,
                   RECORD_WORDS, STATE_WORDS, STATS_LEN, TERMINAL_PLAYER_ID, CoupError, check)

_TORCH_TO_DTYPE = {torch.float32: DTYPE_F32, torch.uint8: DTYPE_U8, torch.bfloat16: DTYPE_BF16}


class _DevPtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it zero-copy."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None,
        }
        self._owner = owner  # keeps the handle alive as long as a view exists


def _view(ptr, shape, typestr, device, owner):
    return torch.as_tensor(_DevPtr(ptr, shape, typestr, owner), device=device)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class CoupVectorEnv:
    """`num_envs` independent 2-player Coup games resident on one GPU."""

    num_players = 2
    num_actions = NUM_DISTINCT_ACTIONS
    info_state_size = INFO_STATE_SIZE
    observation_size = OBSERVATION_SIZE

    def __init__(self, num_envs, seed=1234, device=0, global_env_offset=0, auto_reset=False,
                 plain_store_encoder=False, warp_specialised=True, blocking_sync=False, finished_ring=0):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise CoupError("CoupVectorEnv needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
        self.num_envs = int(num_envs)
        opts = _lib.VecOpts(self.num_envs, self.device.index, seed, global_env_offset,
                            (FLAG_AUTO_RESET if auto_reset else 0)
                            | (_lib.FLAG_PLAIN_STORE_ENCODER if plain_store_encoder else 0)
                            | (0 if warp_specialised else _lib.FLAG_NO_WARP_SPECIALISATION)
                            | (_lib.FLAG_BLOCKING_SYNC if blocking_sync else 0), 0)
        self._h = C.c_void_p()
        check(self._lib.coup_vec_create(C.byref(opts), C.byref(self._h)))
        n, L, d = self.num_envs, self._lib, self.device
        self.legal_mask = _view(L.coup_vec_legal_mask(self._h), (n,), "<i4", d, self)
        self.current_player = _view(L.coup_vec_current_player(self._h), (n,), "|i1", d, self)
        self.done = _view(L.coup_vec_done(self._h), (n,), "|u1", d, self)
        self.rewards = _view(L.coup_vec_rewards(self._h), (n, 2), "|i1", d, self)
        self.returns = _view(L.coup_vec_returns(self._h), (n, 2), "|i1", d, self)
        self.state = _view(L.coup_vec_state(self._h), (n, STATE_WORDS), "<i4", d, self)
        self.history = _view(L.coup_vec_history(self._h), (n, HISTORY_WORDS), "<i4", d, self)
        self.step_word = _view(L.coup_vec_step_word(self._h), (n,), "<i4", d, self)
        self.stats_device = _view(L.coup_vec_stats_device(self._h), (STATS_LEN,), "<i8", d, self)
        self.finished_ring = self.finished_ctrl = None
        if finished_ring:
            self.enable_finished_ring(finished_ring)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.coup_vec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.num_envs

    # ---- helpers --------------------------------------------------------------------------------
    def _u8(self, t, shape, name):
        if t is None:
            return None
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t, dtype=np.uint8))
        t = t.to(device=self.device, dtype=torch.uint8).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @staticmethod
    def _row_stride(out):
        """Row stride in elements of a [rows, >=2492] output (2496 = padded, GEMM-aligned rows), or of a contiguous
        [rows, 1728] output that receives the live prefix of every row (COUP_LIVE_INFO_STATE_SIZE)."""
        live = out.dim() == 2 and out.shape[1] == LIVE_INFO_STATE_SIZE and out.stride(0) == LIVE_INFO_STATE_SIZE
        if out.dim() != 2 or out.stride(1) != 1 or (out.shape[1] < INFO_STATE_SIZE and not live):
            raise ValueError("info-state output must be [rows, >=2492] (or contiguous [rows, 1728]) with unit inner stride")
        unit = 4 * out.element_size()          # the encoders store four elements at a time (coup_b200.h, "Alignment")
        if out.data_ptr() % unit:
            raise ValueError(f"info-state output must be {unit}-byte aligned (got a view at offset {out.data_ptr() % unit})")
        return int(out.stride(0))

    def _rows(self, player):
        return self.num_envs * (2 if player == PLAYER_BOTH else 1)

    # ---- reset / step ---------------------------------------------------------------------------
    def reset(self, envs_to_reset=None, forced_deals=None):
        """vector_env.SyncVectorEnv.reset(envs_to_reset) (vector_env.py:68-78)."""
        m = self._u8(envs_to_reset, (self.num_envs,), "envs_to_reset")
        f = self._u8(forced_deals, (self.num_envs, 4), "forced_deals")
        check(self._lib.coup_vec_reset(self._h, self._ptr(m), self._ptr(f), _stream_ptr(self.device)))

    def step(self, actions, forced_chance=None):
        """One player action per env + all following chance nodes (rl_environment.py:282-322)."""
        a = self._u8(actions, (self.num_envs,), "actions")
        f = self._u8(forced_chance, (self.num_envs, 4), "forced_chance")
        check(self._lib.coup_vec_step(self._h, self._ptr(a), self._ptr(f), _stream_ptr(self.device)))

    def fork_from(self, src, parents, actions, forced_chance=None):
        """Batched `state.child(action)` (deep_cfr.py:415-497): env i of THIS handle becomes a copy of env
        parents[i] of `src` with actions[i] applied and the following chance nodes resolved, never auto-reset.
        Envs >= len(parents) keep their contents. Returns the number of children."""
        parents = parents.to(device=self.device, dtype=torch.int32).contiguous()
        count = int(parents.numel())
        a = self._u8(actions, (count,), "actions")
        f = self._u8(forced_chance, (count, 4), "forced_chance")
        check(self._lib.coup_vec_fork(self._h, src._h, self._ptr(parents) if count else None,
                                      self._ptr(a) if count else None, self._ptr(f), count, _stream_ptr(self.device)))
        return count

    def sample_uniform(self, out=None):
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        check(self._lib.coup_vec_sample_uniform(self._h, self._ptr(out), _stream_ptr(self.device)))
        return out

    def sample_policy(self, logits, probs_out=None, actions_out=None):
        """Masked softmax sampling (nfsp.py:154-167) from logits [num_envs, 18] (float32 or bfloat16)."""
        logits = logits.contiguous()
        if tuple(logits.shape) != (self.num_envs, NUM_DISTINCT_ACTIONS):
            raise ValueError(f"logits must be [{self.num_envs}, {NUM_DISTINCT_ACTIONS}]")
        if actions_out is None:
            actions_out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        check(self._lib.coup_vec_sample_policy(self._h, self._ptr(logits), _TORCH_TO_DTYPE[logits.dtype],
                                               self._ptr(probs_out), self._ptr(actions_out), _stream_ptr(self.device)))
        return actions_out

    def rollout(self, n_steps, encode_player=None, out=None, dtype=torch.float32):
        """n_steps fused (uniform-random action, step, chance, [auto-reset], encode) kernels."""
        if encode_player is None:
            check(self._lib.coup_vec_rollout(self._h, n_steps, -1, 0, None, _stream_ptr(self.device)))
            return None
        if out is None:
            out = torch.empty((self._rows(encode_player), INFO_STATE_SIZE), dtype=dtype, device=self.device)
        check(self._lib.coup_vec_rollout_strided(self._h, n_steps, encode_player, _TORCH_TO_DTYPE[out.dtype],
                                                 self._ptr(out), self._row_stride(out), _stream_ptr(self.device)))
        return out

    def rollout_incremental(self, n_steps, buf):
        """Fused random rollout that keeps `buf` ([2*num_envs, >=2492], both views of every env, previously filled by
        information_state_tensor(PLAYER_BOTH, out=buf)) up to date by rewriting only the changed elements."""
        if buf.shape[0] != 2 * self.num_envs:
            raise ValueError("incremental buffer must have 2 * num_envs rows")
        check(self._lib.coup_vec_rollout_incremental(self._h, n_steps, _TORCH_TO_DTYPE[buf.dtype], self._ptr(buf),
                                                     self._row_stride(buf), _stream_ptr(self.device)))
        return buf

    def step_host(self, h_actions, h_legal_mask=None, h_current_player=None, h_done=None, h_rewards=None,
                  tensor_out=None):
        """Host-buffer path: numpy/pinned-torch buffers in and out, tensors stay on the device."""
        def hp(x):
            if x is None:
                return None
            return C.c_void_p(x.data_ptr() if torch.is_tensor(x) else x.ctypes.data)
        dt = _TORCH_TO_DTYPE[tensor_out.dtype] if tensor_out is not None else 0
        check(self._lib.coup_vec_step_host(self._h, hp(h_actions), hp(h_legal_mask), hp(h_current_player),
                                           hp(h_done), hp(h_rewards), dt, self._ptr(tensor_out),
                                           _stream_ptr(self.device)))

    def step_host_packed(self, h_actions, h_step_words, tensor_out=None, stream=None):
        """Host-buffer path with one D2H copy: h_step_words (int32 [n], pinned) receives `step_word`."""
        dt = _TORCH_TO_DTYPE[tensor_out.dtype] if tensor_out is not None else 0
        sp = _stream_ptr(self.device) if stream is None else C.c_void_p(stream.cuda_stream)
        check(self._lib.coup_vec_step_host_packed(self._h, C.c_void_p(h_actions.data_ptr()),
                                                  C.c_void_p(h_step_words.data_ptr()), dt, self._ptr(tensor_out), sp))

    def step_host_packed_async(self, h_actions, h_step_words, tensor_out=None, stream=None):
        """`step_host_packed` without the final wait; `host_outputs_wait()` blocks until `h_step_words` is filled."""
        dt = _TORCH_TO_DTYPE[tensor_out.dtype] if tensor_out is not None else 0
        sp = _stream_ptr(self.device) if stream is None else C.c_void_p(stream.cuda_stream)
        check(self._lib.coup_vec_step_host_packed_async(self._h, C.c_void_p(h_actions.data_ptr()),
                                                        C.c_void_p(h_step_words.data_ptr()), dt, self._ptr(tensor_out), sp))

    def host_outputs_wait(self):
        check(self._lib.coup_vec_host_outputs_wait(self._h))

    # ---- observations ---------------------------------------------------------------------------
    def information_state_tensor(self, player=PLAYER_CURRENT, out=None, dtype=torch.float32):
        """CoupState::InformationStateTensor (coup.cc:1044-1049) for every env; [rows, 2492]."""
        if out is None:
            out = torch.empty((self._rows(player), INFO_STATE_SIZE), dtype=dtype, device=self.device)
        check(self._lib.coup_vec_information_state_tensor_strided(self._h, player, _TORCH_TO_DTYPE[out.dtype],
                                                                  self._ptr(out), self._row_stride(out),
                                                                  _stream_ptr(self.device)))
        return out

    def information_state_tensor_gather(self, env_ids, player=PLAYER_CURRENT, out=None, dtype=torch.float32):
        """Info-state rows of the envs listed in `env_ids` (int32/int64 device tensor) only."""
        ids = env_ids.to(device=self.device, dtype=torch.int32).contiguous()
        count = int(ids.numel())
        rows = count * (2 if player == PLAYER_BOTH else 1)
        if out is None:
            out = torch.empty((rows, INFO_STATE_SIZE), dtype=dtype, device=self.device)
        if out.shape[0] < rows:
            raise ValueError("output has too few rows")
        check(self._lib.coup_vec_information_state_tensor_gather(
            self._h, self._ptr(ids) if count else None, count, player, _TORCH_TO_DTYPE[out.dtype], self._ptr(out),
            self._row_stride(out), _stream_ptr(self.device)))
        return out[:rows]

    def observation_tensor(self, player=PLAYER_CURRENT, out=None, dtype=torch.float32):
        """CoupState::ObservationTensor (coup.cc:1051-1056) for every env; [rows, 98]."""
        if out is None:
            out = torch.empty((self._rows(player), OBSERVATION_SIZE), dtype=dtype, device=self.device)
        check(self._lib.coup_vec_observation_tensor(self._h, player, _TORCH_TO_DTYPE[out.dtype],
                                                    self._ptr(out), _stream_ptr(self.device)))
        return out

    def observer_tensor(self, player=PLAYER_CURRENT, public_info=True, perfect_recall=False, private_info=1, out=None,
                        dtype=torch.float32):
        """`game.make_observer(IIGObservationType(public_info, perfect_recall, private_info))` tensors (coup.cc:1132-1141)
        for every env; private_info 0 none / 1 single player / 2 all players. [rows, 2492 | 98 | 42]."""
        width = (INFO_STATE_SIZE if perfect_recall else OBSERVATION_SIZE) if public_info else 42
        if out is None:
            out = torch.empty((self._rows(player), width), dtype=dtype, device=self.device)
        if tuple(out.shape) != (self._rows(player), width) or not out.is_contiguous():
            raise ValueError(f"observer output must be a contiguous [{self._rows(player)}, {width}] tensor")
        check(self._lib.coup_vec_observer_tensor(self._h, player, int(bool(public_info)), int(bool(perfect_recall)),
                                                 int(private_info), _TORCH_TO_DTYPE[out.dtype], self._ptr(out),
                                                 _stream_ptr(self.device)))
        return out

    def observation_tensor_gather(self, env_ids, player=PLAYER_CURRENT, out=None, dtype=torch.float32):
        """Observation rows of the envs listed in `env_ids` only."""
        ids = env_ids.to(device=self.device, dtype=torch.int32).contiguous()
        count = int(ids.numel())
        rows = count * (2 if player == PLAYER_BOTH else 1)
        if out is None:
            out = torch.empty((rows, OBSERVATION_SIZE), dtype=dtype, device=self.device)
        if count:
            check(self._lib.coup_vec_observation_tensor_gather(self._h, self._ptr(ids), count, player,
                                                               _TORCH_TO_DTYPE[out.dtype], self._ptr(out), _stream_ptr(self.device)))
        return out[:rows]

    def legal_actions_mask(self, out=None):
        """State::LegalActionsMask (spiel.cc:371-377): uint8 [num_envs, 18]."""
        if out is None:
            out = torch.empty((self.num_envs, NUM_DISTINCT_ACTIONS), dtype=torch.uint8, device=self.device)
        check(self._lib.coup_vec_legal_actions_mask(self._h, self._ptr(out), _stream_ptr(self.device)))
        return out

    def tensor_row_hash(self, t):
        t = t.contiguous()
        rows = t.shape[0]
        out = torch.empty(rows, dtype=torch.int64, device=self.device)
        check(self._lib.coup_tensor_row_hash(self._ptr(t), _TORCH_TO_DTYPE[t.dtype], rows, t.shape[1],
                                             self._ptr(out), _stream_ptr(self.device)))
        return out

    # ---- finished episodes (auto-reset mode keeps them: include/coup_b200.h "finished episodes") ------
    def enable_finished_ring(self, capacity):
        """Device ring of `capacity` (power of two) packed records, one per episode that ends in any step/rollout
        call, appended before the env is re-dealt. 0 disables."""
        check(self._lib.coup_vec_finished_ring_enable(self._h, int(capacity)))
        self.finished_ring = self.finished_ctrl = None
        if capacity:
            self.finished_ring = _view(self._lib.coup_vec_finished_ring(self._h), (int(capacity), RECORD_WORDS), "<i4",
                                       self.device, self)
            self.finished_ctrl = _view(self._lib.coup_vec_finished_ring_ctrl(self._h), (4,), "<i8", self.device, self)

    def finished_drain(self, max_records=None):
        """Records not handed out yet, oldest first: (uint32 numpy [k, 24], records dropped so far)."""
        cap = int(self._lib.coup_vec_finished_ring_capacity(self._h))
        k = cap if max_records is None else min(int(max_records), cap)
        buf = np.empty((max(k, 1), RECORD_WORDS), np.uint32)
        count, dropped = C.c_uint32(0), C.c_uint64(0)
        check(self._lib.coup_vec_finished_drain(self._h, C.c_void_p(buf.ctypes.data), k, C.byref(count), C.byref(dropped),
                                                _stream_ptr(self.device)))
        return buf[: count.value], int(dropped.value)

    def finished_information_state_tensor(self, player=PLAYER_BOTH, out=None, dtype=torch.float32, max_episodes=None,
                                          env_ids_out=None, count_out=None):
        """Terminal info-state rows of the episodes that ended in the most recent step/rollout call, in ring order.
        Returns (rows, env_ids int32 [max_episodes], count int32 [1]); only the first `count` entries (x2 rows for
        PLAYER_BOTH) are written. Nothing is read back to the host."""
        views = 2 if player == PLAYER_BOTH else 1
        if max_episodes is None:
            max_episodes = self.num_envs if out is None else out.shape[0] // views
        if out is None:
            out = torch.empty((max_episodes * views, INFO_STATE_SIZE), dtype=dtype, device=self.device)
        if out.shape[0] < max_episodes * views:
            raise ValueError("output has too few rows")
        if env_ids_out is None:
            env_ids_out = torch.empty(max_episodes, dtype=torch.int32, device=self.device)
        if count_out is None:
            count_out = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self._lib.coup_vec_finished_information_state_tensor(
            self._h, player, _TORCH_TO_DTYPE[out.dtype], self._ptr(out), self._row_stride(out), int(max_episodes),
            self._ptr(env_ids_out), self._ptr(count_out), _stream_ptr(self.device)))
        return out, env_ids_out, count_out

    def finished_observation_tensor(self, player=PLAYER_BOTH, out=None, dtype=torch.float32, max_episodes=None,
                                    env_ids_out=None, count_out=None):
        """Terminal observation rows ([.., 98]) of the episodes that ended in the most recent step/rollout call."""
        views = 2 if player == PLAYER_BOTH else 1
        if max_episodes is None:
            max_episodes = self.num_envs if out is None else out.shape[0] // views
        if out is None:
            out = torch.empty((max_episodes * views, OBSERVATION_SIZE), dtype=dtype, device=self.device)
        if out.shape[0] < max_episodes * views or not out.is_contiguous():
            raise ValueError("output must be contiguous with max_episodes (x2) rows")
        if env_ids_out is None:
            env_ids_out = torch.empty(max_episodes, dtype=torch.int32, device=self.device)
        if count_out is None:
            count_out = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self._lib.coup_vec_finished_observation_tensor(
            self._h, player, _TORCH_TO_DTYPE[out.dtype], self._ptr(out), int(max_episodes), self._ptr(env_ids_out),
            self._ptr(count_out), _stream_ptr(self.device)))
        return out, env_ids_out, count_out

    def records_observation_tensor(self, records, indices=None, player=PLAYER_BOTH, out=None, dtype=torch.float32):
        """Decodes packed records into observation rows ([.., 98])."""
        if indices is not None:
            indices = indices.to(device=self.device, dtype=torch.int32).contiguous()
        count = int(records.shape[0] if indices is None else indices.numel())
        rows = count * (2 if player == PLAYER_BOTH else 1)
        if out is None:
            out = torch.empty((rows, OBSERVATION_SIZE), dtype=dtype, device=self.device)
        if count:
            check(self._lib.coup_records_observation_tensor(self._h, self._ptr(records), self._ptr(indices), count, player,
                                                            _TORCH_TO_DTYPE[out.dtype], self._ptr(out), _stream_ptr(self.device)))
        return out[:rows]

    def records_information_state_tensor(self, records, indices=None, player=PLAYER_BOTH, out=None, dtype=torch.float32):
        """Decodes packed records (int32 device tensor [m, 24]) into info-state rows; row i = record indices[i]."""
        if records.dim() != 2 or records.shape[1] != RECORD_WORDS or not records.is_contiguous():
            raise ValueError("records must be a contiguous [m, 24] int32 tensor")
        if indices is not None:
            indices = indices.to(device=self.device, dtype=torch.int32).contiguous()
        count = int(records.shape[0] if indices is None else indices.numel())
        rows = count * (2 if player == PLAYER_BOTH else 1)
        if out is None:
            out = torch.empty((rows, INFO_STATE_SIZE), dtype=dtype, device=self.device)
        if out.shape[0] < rows:
            raise ValueError("output has too few rows")
        if count:
            check(self._lib.coup_records_information_state_tensor(
                self._h, self._ptr(records), self._ptr(indices), count, player, _TORCH_TO_DTYPE[out.dtype], self._ptr(out),
                self._row_stride(out), _stream_ptr(self.device)))
        return out[:rows]

    # ---- statistics -----------------------------------------------------------------------------
    def stats(self):
        buf = np.zeros(STATS_LEN, np.uint64)
        check(self._lib.coup_vec_stats(self._h, C.c_void_p(buf.ctypes.data), _stream_ptr(self.device)))
        return stats_dict(buf)

    def clear_stats(self):
        check(self._lib.coup_vec_clear_stats(self._h, _stream_ptr(self.device)))

    def check_errors(self):
        check(self._lib.coup_vec_check_errors(self._h, _stream_ptr(self.device)))

    @property
    def step_counter(self):
        return int(self._lib.coup_vec_step_counter(self._h))

    @step_counter.setter
    def step_counter(self, v):
        check(self._lib.coup_vec_set_step_counter(self._h, int(v)))

    # ---- checkpoint / resume ------------------------------------------------------------------------
    def snapshot(self):
        """Whole-slab checkpoint as a numpy byte array (state, history, outputs, stats, Philox counter)."""
        size = int(self._lib.coup_vec_snapshot_size(self._h))
        buf = np.empty(size, np.uint8)
        check(self._lib.coup_vec_snapshot(self._h, C.c_void_p(buf.ctypes.data), size, _stream_ptr(self.device)))
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        check(self._lib.coup_vec_restore(self._h, C.c_void_p(buf.ctypes.data), buf.size, _stream_ptr(self.device)))

    def serialized_states(self):
        """Every env's current state in the reference's wire format (State::Serialize, spiel.cc:297-311:
        one action id per line), loadable with Game::DeserializeState / pyspiel.deserialize_game_and_state."""
        return ["\n".join(str(int(a)) for a in acts) + "\n" for acts, _ in self.trajectories()]

    # ---- trajectory export (host side, for replay through the oracle) -----------------------------
    def move_numbers(self):
        return (self.state[:, 3] & 127).to(torch.int64)

    def trajectories(self):
        """Decodes the device history into per-env action lists (chance outcomes included), i.e. the
        reference's serialisation of a state (State::Serialize, spiel.cc:297-311)."""
        hist = self.history.cpu().numpy().view(np.uint32)
        lens = self.move_numbers().cpu().numpy()
        return decode_history(hist, lens)


def decode_history(hist_words, lens):
    """uint32 [n,16] packed 5-bit codes -> list of (actions uint8[len], deal_target int8[len])."""
    n = hist_words.shape[0]
    idx = np.arange(96)
    codes = (hist_words[:, idx // 6] >> (5 * (idx % 6)).astype(np.uint32)) & 31
    out = []
    for e in range(n):
        c = codes[e, : lens[e]].astype(np.int64)
        is_chance = c >= 18
        target = np.where(is_chance, (c - 18) // 5, -1).astype(np.int8)
        action = np.where(is_chance, (c - 18) % 5, c).astype(np.uint8)
        out.append((action, target))
    return out


def decode_finished_records(records):
    """uint32 [k, 24] ring records -> dict of numpy fields + per-episode (actions, deal_target) lists."""
    r = np.asarray(records).view(np.uint32).reshape(-1, RECORD_WORDS)
    meta = r[:, 21]
    moves = (meta & 127).astype(np.int64)
    return {
        "env": r[:, 20].astype(np.int64), "moves": moves,
        "return0": ((meta >> 8) & 7).astype(np.int64) - 2, "reward0": ((meta >> 12) & 7).astype(np.int64) - 2,
        "truncated": ((meta >> 16) & 1).astype(bool),
        "step": r[:, 22].astype(np.uint64) | (r[:, 23].astype(np.uint64) << np.uint64(32)),
        "state": r[:, HISTORY_WORDS:HISTORY_WORDS + STATE_WORDS].copy(),
        "trajectories": decode_history(r[:, :HISTORY_WORDS], moves),
    }


def stats_dict(buf):
    buf = [int(x) for x in buf]
    return {
        "decision_steps": buf[_lib.STAT_DECISION_STEPS],
        "chance_moves": buf[_lib.STAT_CHANCE_MOVES],
        "episodes": buf[_lib.STAT_EPISODES],
        "truncated": buf[_lib.STAT_TRUNCATED],
        "episode_moves": buf[_lib.STAT_EPISODE_MOVES],
        "illegal": buf[_lib.STAT_ILLEGAL],
        "returns_hist_p0": buf[_lib.STAT_RETURN_HIST:_lib.STAT_RETURN_HIST + 5],
        "legal_count_hist": buf[_lib.STAT_LEGAL_HIST:_lib.STAT_LEGAL_HIST + 8],
    }


def unpack_states(state_words):
    """Decodes packed state words (uint32 [n,4], layout in include/coup_b200.h) into plain numpy fields
    for inspection and tests. hands[n,2,4] holds (value<<1 | face_up) per slot, 15 = empty."""
    w = np.asarray(state_words).view(np.uint32).reshape(-1, 4)
    out = {}
    pw = w[:, :2]
    out["hands"] = np.stack([(pw >> (4 * i)) & 15 for i in range(4)], axis=-1).astype(np.int64)
    out["num_cards"] = (out["hands"] != 15).sum(-1)
    out["coins"] = ((pw >> 16) & 31).astype(np.int64)
    last = ((pw >> 21) & 31).astype(np.int64)
    out["last_action"] = np.where(last == 31, -1, last)
    out["lost_challenge"] = ((pw >> 26) & 1).astype(np.int64)
    g = w[:, 2]
    out["deck"] = np.stack([(g >> (4 * c)) & 15 for c in range(5)], axis=-1).astype(np.int64)
    out["cur_player_turn"] = ((g >> 20) & 1).astype(np.int64)
    out["cur_player_move"] = ((g >> 21) & 1).astype(np.int64)
    out["is_turn_begin"] = ((g >> 22) & 1).astype(np.int64)
    out["is_chance"] = ((g >> 23) & 1).astype(np.int64)
    out["queued_deals"] = ((g >> 24) & 7).astype(np.int64)
    out["error"] = ((g >> 29) & 1).astype(np.int64)
    c = w[:, 3]
    out["move_number"] = (c & 127).astype(np.int64)
    out["turn_number"] = ((c >> 7) & 127).astype(np.int64)
    out["reward0"] = ((c >> 14) & 7).astype(np.int64) - 2
    return out


class BatchedTimeSteps:
    """The list of `TimeStep`s that `SyncVectorEnv.step/reset` returns (python/vector_env.py:40-78), as device tensors:
    observations["info_state"] [n, 2, 2492], observations["legal_actions_mask"] uint8 [n, 2, 18] (all zero for the
    player who is not to move, rl_environment.py:243-249), observations["current_player"] int8 [n] (-4 = terminal),
    rewards float [n, 2], discounts float [n, 2], step_type int8 [n] (0 FIRST, 1 MID, 2 LAST). `ts[i]` converts env i
    into the reference's `rl_environment.TimeStep` with Python lists, for code written against the list form."""

    def __init__(self, observations, rewards, discounts, step_type):
        self.observations, self.rewards, self.discounts, self.step_type = observations, rewards, discounts, step_type

    def __len__(self):
        return int(self.step_type.shape[0])

    def last(self):
        return self.step_type == 2

    def first(self):
        return self.step_type == 0

    def current_player(self):
        return self.observations["current_player"]

    def __getitem__(self, i):
        from .rl_environment import StepType, TimeStep
        st = StepType(int(self.step_type[i]))
        mask = self.observations["legal_actions_mask"][i].cpu().numpy()
        obs = {"info_state": [row.tolist() for row in self.observations["info_state"][i].float().cpu()],
               "legal_actions": [np.nonzero(mask[p])[0].tolist() for p in range(2)],
               "current_player": int(self.observations["current_player"][i]), "serialized_state": []}
        first = st == StepType.FIRST
        return TimeStep(observations=obs, rewards=None if first else self.rewards[i].tolist(),
                        discounts=None if first else self.discounts[i].tolist(), step_type=st)


class SyncVectorEnv:
    """`vector_env.SyncVectorEnv` (python/vector_env.py:23-78) over ONE batched device environment instead of a list of
    `rl_environment.Environment`s: same call shape -- `step(step_outputs, reset_if_done) -> (time_steps, reward, done,
    unreset_time_steps)`, `reset(envs_to_reset)` -- with batched tensors (BatchedTimeSteps) where the reference has lists.

    The slab runs in auto-reset mode with the finished-episode ring on: an env whose episode ends is re-dealt inside the
    step kernel, and its terminal time step (what `unreset_time_steps` carries) is rebuilt from the ring record. With
    `reset_if_done=False` the reference leaves a finished env alone and resets it on its NEXT step call, ignoring that
    call's action (rl_environment.py:296-297); here the env already holds the new episode, so that next call lets it sit
    the step out (action 0xFF) and reports its FIRST time step."""

    def __init__(self, num_envs, seed=1234, device=0, discount=1.0, dtype=torch.float32, env=None):
        self.env = env if env is not None else CoupVectorEnv(num_envs, seed=seed, device=device, auto_reset=True)
        n = self.env.num_envs
        ring = 1
        while ring < 2 * n:
            ring *= 2
        self.env.enable_finished_ring(ring)
        self._discount, self._dtype, dev = float(discount), dtype, self.env.device
        self._pending_first = torch.zeros(n, dtype=torch.bool, device=dev)   # finished, not yet "reset" by the caller
        self._term_rows = torch.empty((2 * n, INFO_STATE_SIZE), dtype=dtype, device=dev)
        self._term_ids = torch.empty(n, dtype=torch.int32, device=dev)
        self._term_count = torch.zeros(1, dtype=torch.int32, device=dev)

    def __len__(self):
        return self.env.num_envs

    @property
    def num_players(self):
        return 2

    def observation_spec(self):
        return dict(info_state=(INFO_STATE_SIZE,), legal_actions=(NUM_DISTINCT_ACTIONS,), current_player=(), serialized_state=())

    def _current(self, first_mask):
        """Time steps of the states the slab holds now; FIRST where `first_mask`, else MID."""
        env, n = self.env, self.env.num_envs
        info = env.information_state_tensor(PLAYER_BOTH, dtype=self._dtype).view(n, 2, INFO_STATE_SIZE)
        mover = env.current_player.to(torch.int64).clamp(min=0)
        legal = torch.zeros((n, 2, NUM_DISTINCT_ACTIONS), dtype=torch.uint8, device=env.device)
        legal[torch.arange(n, device=env.device), mover] = env.legal_actions_mask()
        rewards = env.rewards.to(torch.float32) * (~first_mask).unsqueeze(1)
        discounts = torch.full((n, 2), self._discount, device=env.device)
        step_type = torch.where(first_mask, 0, 1).to(torch.int8)
        obs = {"info_state": info, "legal_actions_mask": legal, "current_player": env.current_player.clone()}
        return BatchedTimeSteps(obs, rewards, discounts, step_type)

    def step(self, step_outputs, reset_if_done=False):
        env, n = self.env, self.env.num_envs
        if torch.is_tensor(step_outputs):
            actions = step_outputs.to(device=env.device, dtype=torch.uint8)
        else:
            actions = torch.as_tensor(np.asarray([getattr(o, "action", o) for o in step_outputs], dtype=np.uint8)).to(env.device)
        sat_out = self._pending_first.clone()
        actions = torch.where(sat_out, torch.full_like(actions, 0xFF), actions)
        env.step(actions)
        done = env.done.bool() & ~sat_out
        time_steps = self._current(first_mask=sat_out | done)
        reward = env.rewards.to(torch.float32) * (~sat_out).unsqueeze(1)
        # unreset view: the finished envs' LAST time steps, rebuilt from this call's ring records
        _, ids, count = env.finished_information_state_tensor(PLAYER_BOTH, out=self._term_rows, env_ids_out=self._term_ids,
                                                              count_out=self._term_count)
        k = int(count.item())
        un_obs = {key: t.clone() for key, t in time_steps.observations.items()}
        idx = ids[:k].long()
        un_obs["info_state"][idx] = self._term_rows[: 2 * k].view(k, 2, INFO_STATE_SIZE)
        un_obs["legal_actions_mask"][idx] = 0
        un_obs["current_player"][idx] = TERMINAL_PLAYER_ID
        un_discounts = time_steps.discounts.clone()
        un_discounts[idx] = 0.0
        un_type = torch.where(done, 2, torch.where(sat_out, 0, 1)).to(torch.int8)
        unreset = BatchedTimeSteps(un_obs, reward, un_discounts, un_type)
        if reset_if_done:
            self._pending_first.zero_()
            return time_steps, reward, done, unreset
        self._pending_first = done
        return unreset, reward, done, unreset

    def reset(self, envs_to_reset=None):
        """vector_env.py:68-78: reset the listed envs (all by default), `get_time_step()` for the others."""
        env, n = self.env, self.env.num_envs
        if envs_to_reset is None:
            mask = torch.ones(n, dtype=torch.bool, device=env.device)
        else:
            mask = torch.as_tensor(envs_to_reset).to(device=env.device).bool()
        redeal = mask & ~self._pending_first            # envs waiting for their reset already hold a fresh episode
        env.reset(redeal.to(torch.uint8))
        first = mask | self._pending_first
        self._pending_first = self._pending_first & ~mask
        ts = self._current(first_mask=first)
        # an env that is finished-and-pending but not in the mask still shows its fresh episode: the terminal step was
        # handed out by the step call that ended it
        return ts
