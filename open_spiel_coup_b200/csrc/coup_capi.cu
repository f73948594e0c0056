// coup_capi.cu -- C ABI (include/coup_b200.h) of the batched Coup environment: host runtime that owns
// the device slab of one GPU and launches the sm_100a kernels in coup_kernels.cuh. No game rule is
// evaluated on the host anywhere in this file.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / Compute, cost nothing otherwise

#include <cstdio>
#include <cstring>
#include <new>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "coup_kernels.cuh"

using namespace coup;

namespace coup_host {  // coup_host_policy.cc
void sample_uniform(const uint32_t* words, uint32_t n, uint64_t seed, uint64_t global_env_offset, uint64_t step,
                    uint8_t* actions, int threads);
}

struct coup_vec_env {
  coup_vec_opts opts;
  EnvArrays A;
  uint8_t* d_actions;   // staging for coup_vec_step_host
  uint32_t* d_scratch;  // [4]: id / illegal flag of the single-env host accessors
  float* d_row;         // [2 * 2496]: one env's info-state rows
  cudaEvent_t host_outputs_ready;
  uint64_t step_counter;
  uint64_t ring_tail;       // first ring record coup_vec_finished_drain has not handed out yet
  uint64_t ring_dropped;    // records overwritten before they were drained
  std::mutex single_env_mutex;   // serialises the coup_env_* accessors (they share d_scratch / d_row)
};

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess)                                                            \
      return fail(COUP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// Makes the handle's device current for the duration of a call and restores the caller's; every entry point that
// launches work holds one, so it also opens the call's NVTX range (named after the entry point).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev, const char* range = __builtin_FUNCTION()) {
    nvtxRangePushA(range);
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    nvtxRangePop();
  }
};

inline cudaStream_t S(void* stream) { return static_cast<cudaStream_t>(stream); }
inline unsigned blocks_for(size_t threads) { return static_cast<unsigned>((threads + kBlockThreads - 1) / kBlockThreads); }

int launch_status(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(COUP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(err));
  return COUP_OK;
}

bool valid_player_sel(int p) { return p >= COUP_PLAYER_0 && p <= COUP_PLAYER_BOTH; }
bool valid_dtype(int d) { return d == COUP_DTYPE_F32 || d == COUP_DTYPE_U8 || d == COUP_DTYPE_BF16; }

// The staged (bulk-store) encoder needs row groups that are multiples of 16 bytes: contiguous rows
// (stride 2492) or the GEMM-friendly padded stride 2496. Other strides use the plain-store encoder.
// Bulk stores also need a 16-byte aligned destination: a torch view such as out[1:] of a uint8 tensor (2492-byte rows) is
// not, and a misaligned cp.async.bulk is a sticky fault. Such outputs take the plain-store encoder, which needs only the
// alignment of its four-element store unit (16 / 8 / 4 bytes for f32 / bf16 / u8); below that the call is rejected.
// Dynamic shared memory of a kernel whose occupancy is planned around it: raises the limit AND asks for the largest
// shared-memory carveout, so that the planned number of CTAs per SM really fits. (Without the preference the driver picks the
// carveout itself: k_encode_obs<float>, 50 KB per CTA and planned for 4 per SM, ran at 80 or at 99 us from run to run.)
template <typename K>
cudaError_t set_dynamic_smem(K kernel, int bytes) {
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return err;
}

bool use_staged_encoder(const coup_vec_env* env, uint32_t stride, const void* d_out) {
  return (env->opts.flags & COUP_FLAG_PLAIN_STORE_ENCODER) == 0 &&
         (stride == COUP_INFO_STATE_SIZE || stride == 2496u || stride == COUP_LIVE_INFO_STATE_SIZE) &&
         reinterpret_cast<uintptr_t>(d_out) % 16u == 0;
}
// Full rows (optionally padded), or the live prefix of every row (COUP_LIVE_INFO_STATE_SIZE: see coup_b200.h).
bool valid_full_stride(uint32_t stride) { return stride >= COUP_INFO_STATE_SIZE && stride % 4u == 0 && stride <= 4096u; }
bool valid_stride(uint32_t stride) { return stride == COUP_LIVE_INFO_STATE_SIZE || valid_full_stride(stride); }
size_t unit_bytes(int dtype) { return dtype == COUP_DTYPE_F32 ? 16u : dtype == COUP_DTYPE_BF16 ? 8u : 4u; }
bool aligned_for(int dtype, const void* p) { return reinterpret_cast<uintptr_t>(p) % unit_bytes(dtype) == 0; }
const char* kMisaligned = "tensor output must be aligned to 16 (f32) / 8 (bf16) / 4 (u8) bytes";

// Device that owns a device pointer (for the entry points that take no handle).
int device_of(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return -1; }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged ? attr.device : -1;
}

// In front of every step/rollout launch: k_step_prologue snapshots the ring cursor and re-arms the batch counter.
void step_prologue(coup_vec_env* env, bool batch_counter, cudaStream_t st) {
  if (env->A.ring != nullptr || batch_counter)
    k_step_prologue<<<1, 32, 0, st>>>(env->A.ring_ctrl, batch_counter ? env->d_scratch + 2 : nullptr);
}

template <typename T>
int rollout_typed(coup_vec_env* env, int n_steps, int encode_player, void* d_out, uint32_t stride, cudaStream_t st) {
  const unsigned grid = blocks_for(env->A.n);
  const bool staged = encode_player >= 0 && use_staged_encoder(env, stride, d_out);
  const bool specialised = staged && (env->opts.flags & COUP_FLAG_NO_WARP_SPECIALISATION) == 0 && env->A.n >= kWsBatch;
  int sms = 0;
  if (staged) {
    cudaError_t err = set_dynamic_smem(k_rollout_tma<T>, kTmaSmemBytes);
    if (err == cudaSuccess && specialised) {
      err = set_dynamic_smem(k_rollout_ws<T>, kWsSmemBytes);
      if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, env->opts.device);
    }
    if (err != cudaSuccess) return fail(COUP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  }
  const uint32_t n_batches = env->A.n / kWsBatch, tail_base = n_batches * kWsBatch;
  if (encode_player < 0) {   // no tensor: up to 64 steps per launch, the env stays in registers in between
    for (int done = 0; done < n_steps;) {
      const int k = std::min(64, n_steps - done);
      step_prologue(env, false, st);
      k_rollout_env_multi<<<grid, kBlockThreads, 0, st>>>(env->A, env->step_counter, k);
      env->step_counter += static_cast<uint64_t>(k);
      done += k;
    }
    return launch_status("k_rollout_env_multi");
  }
  for (int i = 0; i < n_steps; ++i) {
    if (specialised) {
      // persistent, warp-specialised: one CTA per SM over the full 256-env batches, then the ragged tail (if any)
      step_prologue(env, true, st);   // also zeroes the dynamic batch counter
      k_rollout_ws<T><<<std::min<unsigned>(sms, n_batches), kWsThreads, kWsSmemBytes, st>>>(
          env->A, env->step_counter, encode_player, static_cast<T*>(d_out), stride, n_batches, env->d_scratch + 2);
      if (tail_base < env->A.n)
        k_rollout_tma<T><<<1, kTmaBlockThreads, kTmaSmemBytes, st>>>(env->A, env->step_counter, encode_player,
                                                                      static_cast<T*>(d_out), stride, tail_base);
    } else if (staged) {
      step_prologue(env, false, st);
      k_rollout_tma<T><<<grid, kTmaBlockThreads, kTmaSmemBytes, st>>>(env->A, env->step_counter, encode_player, static_cast<T*>(d_out), stride, 0u);
    } else if (encode_player >= 0) {
      step_prologue(env, false, st);
      k_rollout<T, true><<<grid, kBlockThreads, 0, st>>>(env->A, env->step_counter, encode_player, static_cast<T*>(d_out), stride);
    } else {
      k_rollout<T, false><<<grid, kBlockThreads, 0, st>>>(env->A, env->step_counter, 0, static_cast<T*>(nullptr), stride);
    }
    env->step_counter++;
  }
  return launch_status("k_rollout");
}

template <typename T, typename Src>
int encode_info_typed(coup_vec_env* env, const Src& src, uint32_t max_groups, int player, void* d_out, uint32_t stride,
                      uint32_t* d_ids_out, uint32_t* d_count_out, cudaStream_t st) {
  if (max_groups == 0) return COUP_OK;   // rows to produce (before the x2 of COUP_PLAYER_BOTH), an upper bound for ring sources
  const unsigned grid = (max_groups + 32 * kWarpsPerBlock - 1) / (32 * kWarpsPerBlock);
  if (use_staged_encoder(env, stride, d_out)) {
    cudaError_t err = set_dynamic_smem(k_encode_info_tma<T, Src>, kTmaSmemBytes);
    if (err != cudaSuccess) return fail(COUP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
    k_encode_info_tma<T, Src><<<grid, kTmaBlockThreads, kTmaSmemBytes, st>>>(src, player, static_cast<T*>(d_out), stride, d_ids_out, d_count_out);
  } else {
    k_encode_info<T, Src><<<grid, kBlockThreads, 0, st>>>(src, player, static_cast<T*>(d_out), stride, d_ids_out, d_count_out);
  }
  return launch_status("k_encode_info");
}

template <typename Src>
int encode_info_dispatch(coup_vec_env* env, const Src& src, uint32_t max_groups, int player, int dtype, void* d_out,
                         uint32_t stride, uint32_t* d_ids_out, uint32_t* d_count_out, cudaStream_t st) {
  if (!aligned_for(dtype, d_out)) return fail(COUP_ERR_INVALID_ARG, kMisaligned);
  switch (dtype) {
    case COUP_DTYPE_F32: return encode_info_typed<float>(env, src, max_groups, player, d_out, stride, d_ids_out, d_count_out, st);
    case COUP_DTYPE_U8: return encode_info_typed<uint8_t>(env, src, max_groups, player, d_out, stride, d_ids_out, d_count_out, st);
    default: return encode_info_typed<__nv_bfloat16>(env, src, max_groups, player, d_out, stride, d_ids_out, d_count_out, st);
  }
}

template <typename T, typename Src>
int encode_obs_typed(coup_vec_env* env, const Src& src, uint32_t max_groups, int player, void* d_out, uint32_t* d_ids_out,
                     uint32_t* d_count_out, cudaStream_t st) {
  if (max_groups == 0) return COUP_OK;
  const uint32_t row_len = ((player >> 8) & kVisNoPublic) ? 42u : COUP_OBSERVATION_SIZE;
  const size_t smem = static_cast<size_t>(kObsWarps) * 32 * ((player & 7) == COUP_PLAYER_BOTH ? 2 : 1) * row_len * sizeof(T);
  int sms = 0;
  auto kernel = k_encode_obs<T, Src, sizeof(T) == 4>;   // fp32 rows: poke / erase; narrower elements: composed pairs (coup_encode.cuh)
  cudaError_t err = set_dynamic_smem(kernel, static_cast<int>(smem));
  if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, env->opts.device);
  if (err != cudaSuccess) return fail(COUP_ERR_CUDA, std::string("k_encode_obs setup: ") + cudaGetErrorString(err));
  // persistent: as many CTAs as fit at once (shared memory bound, at most 8 per SM), never more than there are groups
  const unsigned per_sm = static_cast<unsigned>(std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
  const unsigned groups = (max_groups + 31) / 32;
  const unsigned grid = std::max(1u, std::min(static_cast<unsigned>(sms) * per_sm, (groups + kObsWarps - 1) / kObsWarps));
  const int use_bulk = reinterpret_cast<uintptr_t>(d_out) % 16u == 0 && (env->opts.flags & COUP_FLAG_PLAIN_STORE_ENCODER) == 0;
  kernel<<<grid, kObsThreads, smem, st>>>(src, player, static_cast<T*>(d_out), use_bulk, row_len, d_ids_out, d_count_out);
  return launch_status("k_encode_obs");
}

template <typename Src>
int encode_obs_dispatch(coup_vec_env* env, const Src& src, uint32_t max_groups, int player, int dtype, void* d_out,
                        uint32_t* d_ids_out, uint32_t* d_count_out, cudaStream_t st) {
  const size_t elem = dtype == COUP_DTYPE_F32 ? 4u : dtype == COUP_DTYPE_BF16 ? 2u : 1u;
  if (reinterpret_cast<uintptr_t>(d_out) % elem != 0) return fail(COUP_ERR_INVALID_ARG, "observation output must be aligned to its element type");
  switch (dtype) {
    case COUP_DTYPE_F32: return encode_obs_typed<float>(env, src, max_groups, player, d_out, d_ids_out, d_count_out, st);
    case COUP_DTYPE_U8: return encode_obs_typed<uint8_t>(env, src, max_groups, player, d_out, d_ids_out, d_count_out, st);
    default: return encode_obs_typed<__nv_bfloat16>(env, src, max_groups, player, d_out, d_ids_out, d_count_out, st);
  }
}

template <typename T>
int rollout_incremental_typed(coup_vec_env* env, int n_steps, void* d_buf, uint32_t stride, cudaStream_t st) {
  const unsigned grid = blocks_for(env->A.n);
  constexpr int kIncSmemBytes = kIncSmemWords * 4 + UnitLut<T>::kBytes;
  static_assert(kIncSmemBytes <= kIncSmemBytesMax, "unit table larger than reserved");
  cudaError_t err = set_dynamic_smem(k_rollout_incremental<T>, kIncSmemBytes);
  if (err != cudaSuccess) return fail(COUP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  for (int i = 0; i < n_steps; ++i) {
    step_prologue(env, false, st);
    k_rollout_incremental<T><<<grid, kBlockThreads, kIncSmemBytes, st>>>(env->A, env->step_counter, static_cast<T*>(d_buf), stride);
    env->step_counter++;
  }
  return launch_status("k_rollout_incremental");
}

}  // namespace

extern "C" {

const char* coup_last_error(void) { return g_error.c_str(); }

int coup_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int coup_vec_create(const coup_vec_opts* opts, coup_vec_env** out) {
  if (!opts || !out || opts->num_envs == 0) return fail(COUP_ERR_INVALID_ARG, "coup_vec_create: bad arguments");
  *out = nullptr;
  int ndev = coup_device_count();
  if (ndev == 0) return fail(COUP_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  if (opts->device < 0 || opts->device >= ndev) return fail(COUP_ERR_INVALID_ARG, "coup_vec_create: bad device ordinal");
  DeviceGuard guard(opts->device);
  if (!guard.ok) return fail(COUP_ERR_CUDA, "cudaSetDevice failed");
  coup_vec_env* env = new (std::nothrow) coup_vec_env();
  if (!env) return fail(COUP_ERR_INVALID_ARG, "out of host memory");
  env->opts = *opts;
  env->step_counter = 0;
  env->host_outputs_ready = nullptr;
  env->d_actions = nullptr;
  env->d_scratch = nullptr;
  env->d_row = nullptr;
  env->ring_tail = env->ring_dropped = 0;
  const size_t n = opts->num_envs;
  EnvArrays& A = env->A;
  std::memset(&A, 0, sizeof(A));
  A.n = opts->num_envs;
  A.flags = opts->flags;
  A.seed = opts->seed;
  A.global_env_offset = opts->global_env_offset;
  cudaError_t err = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(p, bytes); };
  alloc(reinterpret_cast<void**>(&A.state), n * sizeof(uint4));
  alloc(reinterpret_cast<void**>(&A.history), n * kHistoryWords * sizeof(uint32_t));
  alloc(reinterpret_cast<void**>(&A.legal), n * sizeof(uint32_t));
  alloc(reinterpret_cast<void**>(&A.cur_player), n);
  alloc(reinterpret_cast<void**>(&A.done), n);
  alloc(reinterpret_cast<void**>(&A.rewards), n * 2);
  alloc(reinterpret_cast<void**>(&A.returns), n * 2);
  alloc(reinterpret_cast<void**>(&A.step_word), n * sizeof(uint32_t));
  alloc(reinterpret_cast<void**>(&A.stats), COUP_STATS_LEN * sizeof(unsigned long long));
  alloc(reinterpret_cast<void**>(&env->d_actions), n);
  alloc(reinterpret_cast<void**>(&env->d_scratch), 4 * sizeof(uint32_t));
  alloc(reinterpret_cast<void**>(&env->d_row), 2 * 2496 * sizeof(float));
  if (err == cudaSuccess)
    err = cudaEventCreateWithFlags(&env->host_outputs_ready,
                                   cudaEventDisableTiming | ((opts->flags & COUP_FLAG_BLOCKING_SYNC) ? cudaEventBlockingSync : 0));
  if (err == cudaSuccess) err = cudaMemset(A.stats, 0, COUP_STATS_LEN * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMemset(A.history, 0, n * kHistoryWords * sizeof(uint32_t));
  if (err != cudaSuccess) {
    std::string msg = std::string("coup_vec_create: ") + cudaGetErrorString(err);
    coup_vec_destroy(env);
    return fail(COUP_ERR_CUDA, msg);
  }
  *out = env;
  // Every env starts as a freshly dealt episode.
  int rc = coup_vec_reset(env, nullptr, nullptr, nullptr);
  if (rc == COUP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(COUP_ERR_CUDA, "initial reset failed");
  if (rc != COUP_OK) { coup_vec_destroy(env); *out = nullptr; }
  return rc;
}

int coup_vec_destroy(coup_vec_env* env) {
  if (!env) return COUP_OK;
  DeviceGuard guard(env->opts.device);
  cudaFree(env->A.state); cudaFree(env->A.history); cudaFree(env->A.legal); cudaFree(env->A.cur_player);
  cudaFree(env->A.done); cudaFree(env->A.rewards); cudaFree(env->A.returns); cudaFree(env->A.stats);
  cudaFree(env->A.step_word);
  cudaFree(env->A.ring); cudaFree(env->A.ring_ctrl);
  cudaFree(env->d_actions); cudaFree(env->d_scratch); cudaFree(env->d_row);
  if (env->host_outputs_ready) cudaEventDestroy(env->host_outputs_ready);
  delete env;
  return COUP_OK;
}

uint32_t coup_vec_num_envs(const coup_vec_env* env) { return env ? env->A.n : 0; }

int coup_vec_reset(coup_vec_env* env, const uint8_t* d_reset_mask, const uint8_t* d_forced_deals, void* stream) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(env->opts.device);
  k_reset<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_reset_mask, d_forced_deals, env->step_counter);
  env->step_counter++;
  return launch_status("k_reset");
}

int coup_vec_step(coup_vec_env* env, const uint8_t* d_actions, const uint8_t* d_forced_chance, void* stream) {
  if (!env || !d_actions) return fail(COUP_ERR_INVALID_ARG, "coup_vec_step: null argument");
  DeviceGuard guard(env->opts.device);
  step_prologue(env, false, S(stream));
  k_step<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_actions, d_forced_chance, env->step_counter);
  env->step_counter++;
  return launch_status("k_step");
}

int coup_vec_step_record(coup_vec_env* env, const uint8_t* d_actions, const float* d_action_probs,
                         const coup_recorder_buffers* b, void* stream) {
  if (!env || !d_actions || !b) return fail(COUP_ERR_INVALID_ARG, "coup_vec_step_record: null argument");
  const bool reservoir = b->d_reservoir_records != nullptr, replay = b->d_transitions != nullptr;
  if (reservoir && (!b->d_reservoir_probs || !b->d_reservoir_winner || !d_action_probs || b->reservoir_capacity == 0))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_step_record: incomplete reservoir buffers");
  if (replay && (!b->d_replay_total || !b->d_pending || b->replay_capacity == 0))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_step_record: incomplete replay buffers");
  DeviceGuard guard(env->opts.device);
  RecorderArrays R{};
  if (reservoir) {
    R.res_records = b->d_reservoir_records; R.res_probs = b->d_reservoir_probs;
    R.res_winner = reinterpret_cast<unsigned long long*>(b->d_reservoir_winner);
    R.res_capacity = b->reservoir_capacity; R.res_base = b->reservoir_offered;
  }
  if (replay) {
    R.transitions = b->d_transitions; R.rb_capacity = b->replay_capacity;
    R.rb_total = reinterpret_cast<unsigned long long*>(b->d_replay_total); R.pending = b->d_pending;
  }
  step_prologue(env, false, S(stream));
  if (reservoir)
    k_reservoir_claim<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, R, d_actions, env->step_counter);
  k_step_record<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, R, d_actions, d_action_probs, env->step_counter);
  env->step_counter++;
  return launch_status("k_step_record");
}

int coup_vec_new_initial_state(coup_vec_env* env, const uint8_t* d_mask, void* stream) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(env->opts.device);
  k_single_move<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_mask, 0);
  return launch_status("k_single_move(new_initial_state)");
}

int coup_vec_apply_move(coup_vec_env* env, const uint8_t* d_moves, void* stream) {
  if (!env || !d_moves) return fail(COUP_ERR_INVALID_ARG, "coup_vec_apply_move: null argument");
  DeviceGuard guard(env->opts.device);
  k_single_move<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_moves, 1);
  return launch_status("k_single_move(apply_move)");
}

int coup_vec_copy_env(coup_vec_env* env, uint32_t src, uint32_t dst, void* stream) {
  if (!env || src >= env->A.n || dst >= env->A.n) return fail(COUP_ERR_INVALID_ARG, "coup_vec_copy_env: bad index");
  DeviceGuard guard(env->opts.device);
  k_copy_env<<<1, 32, 0, S(stream)>>>(env->A, src, dst);
  return launch_status("k_copy_env");
}

int coup_vec_fork(coup_vec_env* dst, const coup_vec_env* src, const uint32_t* d_parent, const uint8_t* d_actions,
                  const uint8_t* d_forced_chance, uint32_t count, void* stream) {
  if (!dst || !src || dst == src || (count && (!d_parent || !d_actions)))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_fork: null argument or dst == src");
  if (dst->opts.device != src->opts.device) return fail(COUP_ERR_INVALID_ARG, "coup_vec_fork: handles on different devices");
  if (count > dst->A.n) return fail(COUP_ERR_INVALID_ARG, "coup_vec_fork: count exceeds num_envs of dst");
  DeviceGuard guard(dst->opts.device);
  if (count) {
    EnvArrays D = dst->A;
    D.flags &= ~static_cast<uint32_t>(COUP_FLAG_AUTO_RESET);
    D.ring = nullptr;   // children stay terminal and readable in place: nothing to hand over
    k_fork<<<blocks_for(count), kBlockThreads, 0, S(stream)>>>(D, src->A.state, src->A.history, src->A.n, d_parent, d_actions,
                                                             d_forced_chance, count, dst->step_counter, nullptr);
  }
  dst->step_counter++;
  return launch_status("k_fork");
}

// ---- a traversal level without the host in the loop: node counts stay in device memory ---------------------------------
int coup_vec_fork_counted(coup_vec_env* dst, const coup_vec_env* src, const uint32_t* d_parent, const uint8_t* d_actions,
                          const uint32_t* d_count, uint32_t max_count, void* stream) {
  if (!dst || !src || dst == src || !d_parent || !d_actions || !d_count)
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_fork_counted: null argument or dst == src");
  if (dst->opts.device != src->opts.device) return fail(COUP_ERR_INVALID_ARG, "coup_vec_fork_counted: handles on different devices");
  max_count = std::min(max_count, dst->A.n);
  DeviceGuard guard(dst->opts.device);
  if (max_count) {
    EnvArrays D = dst->A;
    D.flags &= ~static_cast<uint32_t>(COUP_FLAG_AUTO_RESET);
    D.ring = nullptr;
    k_fork<<<blocks_for(max_count), kBlockThreads, 0, S(stream)>>>(D, src->A.state, src->A.history, src->A.n, d_parent, d_actions,
                                                                 nullptr, max_count, dst->step_counter, d_count);
  }
  dst->step_counter++;
  return launch_status("k_fork");
}

int coup_vec_information_state_tensor_prefix(coup_vec_env* env, const uint32_t* d_count, uint32_t max_count, int player, int dtype,
                                             void* d_out, uint32_t row_stride, void* stream) {
  if (!env || !d_out || !d_count || !valid_player_sel(player) || !valid_dtype(dtype) || !valid_stride(row_stride))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_information_state_tensor_prefix: bad arguments");
  DeviceGuard guard(env->opts.device);
  max_count = std::min(max_count, env->A.n);
  const SlabSource src{env->A.state, env->A.history, nullptr, max_count, d_count};
  return encode_info_dispatch(env, src, max_count, player, dtype, d_out, row_stride, nullptr, nullptr, S(stream));
}

int coup_vec_pack_records(coup_vec_env* env, const uint32_t* d_count, uint32_t* d_records_out, void* stream) {
  if (!env || !d_count || !d_records_out) return fail(COUP_ERR_INVALID_ARG, "coup_vec_pack_records: null argument");
  DeviceGuard guard(env->opts.device);
  k_pack_records<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_count, d_records_out);
  return launch_status("k_pack_records");
}

int coup_cfr_level(const float* d_advantages, const uint32_t* d_step_words, const uint32_t* d_count, uint32_t capacity,
                   int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
                   uint64_t counter, float* d_strategy_out, uint32_t* d_expand_out, uint32_t* d_offset_out,
                   uint32_t* d_parent_out, uint8_t* d_action_out, uint32_t* d_next_count, uint32_t* d_overflow, void* stream) {
  if (!d_advantages || !d_step_words || !d_count || !d_strategy_out || !d_expand_out || !d_offset_out || !d_parent_out ||
      !d_action_out || !d_next_count || !d_overflow || capacity == 0 || outcome_factor == 0 || (traverser != 0 && traverser != 1))
    return fail(COUP_ERR_INVALID_ARG, "coup_cfr_level: bad arguments");
  const int dev = device_of(d_advantages);
  if (dev < 0) return fail(COUP_ERR_INVALID_ARG, "coup_cfr_level: d_advantages is not a device pointer");
  DeviceGuard guard(dev);
  k_cfr_level<<<1, kCfrLevelThreads, 0, S(stream)>>>(d_advantages, d_step_words, d_count, capacity, traverser, external,
                                                     outcome_factor, e_outcome, expl, seed, counter, d_strategy_out, d_expand_out,
                                                     d_offset_out, d_parent_out, d_action_out, d_next_count, d_overflow);
  return launch_status("k_cfr_level");
}

int coup_cfr_backward(const uint32_t* d_step_words, const uint32_t* d_count, uint32_t capacity, int traverser,
                      const float* d_strategy, const uint32_t* d_expand, const uint32_t* d_offset, const double* d_child_value,
                      double* d_value_out, float* d_regret_out, void* stream) {
  if (!d_step_words || !d_count || !d_strategy || !d_expand || !d_offset || !d_child_value || !d_value_out || !d_regret_out ||
      capacity == 0 || (traverser != 0 && traverser != 1))
    return fail(COUP_ERR_INVALID_ARG, "coup_cfr_backward: bad arguments");
  const int dev = device_of(d_step_words);
  if (dev < 0) return fail(COUP_ERR_INVALID_ARG, "coup_cfr_backward: d_step_words is not a device pointer");
  DeviceGuard guard(dev);
  k_cfr_backward<<<blocks_for(capacity), kBlockThreads, 0, S(stream)>>>(d_step_words, d_count, capacity, traverser, d_strategy,
                                                                       d_expand, d_offset, d_child_value, d_value_out, d_regret_out);
  return launch_status("k_cfr_backward");
}

int coup_cfr_expand(const float* d_advantages, const uint32_t* d_step_words, uint32_t count, int traverser, int external,
                    uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed, uint64_t counter,
                    float* d_strategy_out, uint32_t* d_expand_out, uint32_t* d_child_count_out, void* stream) {
  if (count == 0) return COUP_OK;
  if (!d_advantages || !d_step_words || !d_strategy_out || !d_expand_out || !d_child_count_out || outcome_factor == 0 ||
      (traverser != 0 && traverser != 1))
    return fail(COUP_ERR_INVALID_ARG, "coup_cfr_expand: bad arguments");
  const int dev = device_of(d_advantages);
  if (dev < 0) return fail(COUP_ERR_INVALID_ARG, "coup_cfr_expand: d_advantages is not a device pointer");
  DeviceGuard guard(dev);
  k_cfr_expand<<<blocks_for(count), kBlockThreads, 0, S(stream)>>>(d_advantages, d_step_words, count, traverser, external,
                                                                  outcome_factor, e_outcome, expl, seed, counter,
                                                                  d_strategy_out, d_expand_out, d_child_count_out);
  return launch_status("k_cfr_expand");
}

int coup_cfr_children(const uint32_t* d_expand, const int64_t* d_offsets, uint32_t count, uint32_t* d_parent_out,
                      uint8_t* d_action_out, void* stream) {
  if (count == 0) return COUP_OK;
  if (!d_expand || !d_offsets || !d_parent_out || !d_action_out) return fail(COUP_ERR_INVALID_ARG, "coup_cfr_children: null argument");
  const int dev = device_of(d_expand);
  if (dev < 0) return fail(COUP_ERR_INVALID_ARG, "coup_cfr_children: d_expand is not a device pointer");
  DeviceGuard guard(dev);
  k_cfr_children<<<blocks_for(count), kBlockThreads, 0, S(stream)>>>(d_expand, d_offsets, count, d_parent_out, d_action_out);
  return launch_status("k_cfr_children");
}

// ---- single-env accessors with HOST buffers, in the style of rust_open_spiel.h ------------------------
static int one_move(coup_vec_env* env, uint32_t slot, uint32_t mv, int mode) {
  if (!env || slot >= env->A.n) return fail(COUP_ERR_INVALID_ARG, "coup_env_*: bad handle or slot");
  std::lock_guard<std::mutex> lock(env->single_env_mutex);
  DeviceGuard guard(env->opts.device);
  k_single_move_one<<<1, 32>>>(env->A, slot, mv, mode, env->d_scratch + 1);
  uint32_t illegal = 0;
  if (mode == 1) CUDA_TRY(cudaMemcpy(&illegal, env->d_scratch + 1, sizeof(illegal), cudaMemcpyDeviceToHost));
  else CUDA_TRY(cudaDeviceSynchronize());
  int rc = launch_status("k_single_move_one");
  if (rc != COUP_OK) return rc;
  return illegal ? fail(COUP_ERR_ILLEGAL_ACTION, "illegal action " + std::to_string(mv)) : COUP_OK;
}

int coup_env_new_initial_state(coup_vec_env* env, uint32_t slot) { return one_move(env, slot, 0, 0); }

int coup_env_apply_action(coup_vec_env* env, uint32_t slot, int action) {
  if (action < 0 || action > 255) return fail(COUP_ERR_ILLEGAL_ACTION, "illegal action " + std::to_string(action));
  return one_move(env, slot, static_cast<uint32_t>(action), 1);
}

int coup_env_clone(coup_vec_env* env, uint32_t src, uint32_t dst) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  std::lock_guard<std::mutex> lock(env->single_env_mutex);
  int rc = coup_vec_copy_env(env, src, dst, nullptr);
  if (rc != COUP_OK) return rc;
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaDeviceSynchronize());
  return COUP_OK;
}

int coup_env_read(coup_vec_env* env, uint32_t slot, uint32_t* h_state4, uint32_t* h_history16, uint32_t* h_step_word) {
  if (!env || slot >= env->A.n) return fail(COUP_ERR_INVALID_ARG, "coup_env_read: bad handle or slot");
  DeviceGuard guard(env->opts.device);
  if (h_state4) CUDA_TRY(cudaMemcpy(h_state4, env->A.state + slot, 16, cudaMemcpyDeviceToHost));
  if (h_history16) CUDA_TRY(cudaMemcpy(h_history16, env->A.history + static_cast<size_t>(slot) * kHistoryWords, 64, cudaMemcpyDeviceToHost));
  if (h_step_word) CUDA_TRY(cudaMemcpy(h_step_word, env->A.step_word + slot, 4, cudaMemcpyDeviceToHost));
  return COUP_OK;
}

int coup_env_information_state_tensor(coup_vec_env* env, uint32_t slot, int player, float* h_buf, int length) {
  if (!env || slot >= env->A.n || !h_buf || (player != 0 && player != 1) || length != COUP_INFO_STATE_SIZE)
    return fail(COUP_ERR_INVALID_ARG, "coup_env_information_state_tensor: bad arguments");
  std::lock_guard<std::mutex> lock(env->single_env_mutex);
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaMemcpy(env->d_scratch, &slot, sizeof(slot), cudaMemcpyHostToDevice));
  int rc = coup_vec_information_state_tensor_gather(env, env->d_scratch, 1, player, COUP_DTYPE_F32, env->d_row,
                                                    COUP_INFO_STATE_SIZE, nullptr);
  if (rc != COUP_OK) return rc;
  CUDA_TRY(cudaMemcpy(h_buf, env->d_row, COUP_INFO_STATE_SIZE * sizeof(float), cudaMemcpyDeviceToHost));
  return COUP_OK;
}

int coup_env_observation_tensor(coup_vec_env* env, uint32_t slot, int player, float* h_buf, int length) {
  if (!env || slot >= env->A.n || !h_buf || (player != 0 && player != 1) || length != COUP_OBSERVATION_SIZE)
    return fail(COUP_ERR_INVALID_ARG, "coup_env_observation_tensor: bad arguments");
  std::lock_guard<std::mutex> lock(env->single_env_mutex);
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaMemcpy(env->d_scratch, &slot, sizeof(slot), cudaMemcpyHostToDevice));
  int rc = coup_vec_observation_tensor_gather(env, env->d_scratch, 1, player, COUP_DTYPE_F32, env->d_row, nullptr);
  if (rc != COUP_OK) return rc;
  CUDA_TRY(cudaMemcpy(h_buf, env->d_row, COUP_OBSERVATION_SIZE * sizeof(float), cudaMemcpyDeviceToHost));
  return COUP_OK;
}

int coup_vec_sample_uniform(coup_vec_env* env, uint8_t* d_actions_out, void* stream) {
  if (!env || !d_actions_out) return fail(COUP_ERR_INVALID_ARG, "coup_vec_sample_uniform: null argument");
  DeviceGuard guard(env->opts.device);
  // Uses the x word of the Philox block the NEXT step will consume (same counter, not advanced), so
  // sample_uniform + step reproduces exactly what coup_vec_rollout does in one kernel.
  k_sample_uniform<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A, d_actions_out, env->step_counter);
  return launch_status("k_sample_uniform");
}

int coup_vec_sample_policy(coup_vec_env* env, const void* d_logits, int dtype, float* d_probs_out,
                           uint8_t* d_actions_out, void* stream) {
  if (!env || !d_logits || !d_actions_out || (dtype != COUP_DTYPE_F32 && dtype != COUP_DTYPE_BF16))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_sample_policy: bad arguments (logits must be f32 or bf16)");
  DeviceGuard guard(env->opts.device);
  const unsigned grid = blocks_for(env->A.n);
  if (dtype == COUP_DTYPE_F32)
    k_sample_policy<float><<<grid, kBlockThreads, 0, S(stream)>>>(env->A, static_cast<const float*>(d_logits), d_probs_out,
                                                                    d_actions_out, env->step_counter);
  else
    k_sample_policy<__nv_bfloat16><<<grid, kBlockThreads, 0, S(stream)>>>(
        env->A, static_cast<const __nv_bfloat16*>(d_logits), d_probs_out, d_actions_out, env->step_counter);
  return launch_status("k_sample_policy");
}

int coup_vec_rollout_strided(coup_vec_env* env, int n_steps, int encode_player, int dtype, void* d_tensor_out,
                             uint32_t row_stride, void* stream) {
  if (!env || n_steps < 0) return fail(COUP_ERR_INVALID_ARG, "coup_vec_rollout: bad arguments");
  if (encode_player >= 0 && (!valid_player_sel(encode_player) || !valid_dtype(dtype) || !d_tensor_out || !valid_stride(row_stride)))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_rollout: bad encode arguments");
  if (encode_player >= 0 && !aligned_for(dtype, d_tensor_out)) return fail(COUP_ERR_INVALID_ARG, kMisaligned);
  DeviceGuard guard(env->opts.device);
  if (encode_player < 0) return rollout_typed<float>(env, n_steps, -1, nullptr, COUP_INFO_STATE_SIZE, S(stream));
  switch (dtype) {
    case COUP_DTYPE_F32: return rollout_typed<float>(env, n_steps, encode_player, d_tensor_out, row_stride, S(stream));
    case COUP_DTYPE_U8: return rollout_typed<uint8_t>(env, n_steps, encode_player, d_tensor_out, row_stride, S(stream));
    default: return rollout_typed<__nv_bfloat16>(env, n_steps, encode_player, d_tensor_out, row_stride, S(stream));
  }
}

int coup_vec_rollout(coup_vec_env* env, int n_steps, int encode_player, int dtype, void* d_tensor_out, void* stream) {
  return coup_vec_rollout_strided(env, n_steps, encode_player, dtype, d_tensor_out, COUP_INFO_STATE_SIZE, stream);
}

int coup_vec_rollout_incremental(coup_vec_env* env, int n_steps, int dtype, void* d_buf, uint32_t row_stride, void* stream) {
  if (!env || n_steps < 0 || !d_buf || !valid_dtype(dtype) || !valid_full_stride(row_stride))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_rollout_incremental: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_buf) % 32u != 0)   // the kernel writes whole 32-byte sectors
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_rollout_incremental: the buffer must be 32-byte aligned");
  DeviceGuard guard(env->opts.device);
  switch (dtype) {
    case COUP_DTYPE_F32: return rollout_incremental_typed<float>(env, n_steps, d_buf, row_stride, S(stream));
    case COUP_DTYPE_U8: return rollout_incremental_typed<uint8_t>(env, n_steps, d_buf, row_stride, S(stream));
    default: return rollout_incremental_typed<__nv_bfloat16>(env, n_steps, d_buf, row_stride, S(stream));
  }
}

const uint32_t* coup_vec_legal_mask(const coup_vec_env* env) { return env ? env->A.legal : nullptr; }
const int8_t* coup_vec_current_player(const coup_vec_env* env) { return env ? env->A.cur_player : nullptr; }
const uint8_t* coup_vec_done(const coup_vec_env* env) { return env ? env->A.done : nullptr; }
const int8_t* coup_vec_rewards(const coup_vec_env* env) { return env ? env->A.rewards : nullptr; }
const int8_t* coup_vec_returns(const coup_vec_env* env) { return env ? env->A.returns : nullptr; }
const uint32_t* coup_vec_step_word(const coup_vec_env* env) { return env ? env->A.step_word : nullptr; }
uint32_t* coup_vec_state(coup_vec_env* env) { return env ? reinterpret_cast<uint32_t*>(env->A.state) : nullptr; }
uint32_t* coup_vec_history(coup_vec_env* env) { return env ? env->A.history : nullptr; }

int coup_vec_legal_actions_mask(coup_vec_env* env, uint8_t* d_out, void* stream) {
  if (!env || !d_out) return fail(COUP_ERR_INVALID_ARG, "coup_vec_legal_actions_mask: null argument");
  DeviceGuard guard(env->opts.device);
  k_legal_actions_mask<<<blocks_for(env->A.n), kBlockThreads, 0, S(stream)>>>(env->A.legal, d_out, env->A.n);
  return launch_status("k_legal_actions_mask");
}

int coup_vec_information_state_tensor_gather(coup_vec_env* env, const uint32_t* d_env_ids, uint32_t count, int player,
                                             int dtype, void* d_out, uint32_t row_stride, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype) || !valid_stride(row_stride) ||
      (count > 0 && !d_env_ids))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_information_state_tensor_gather: bad arguments");
  DeviceGuard guard(env->opts.device);
  const SlabSource src{env->A.state, env->A.history, d_env_ids, count, nullptr};   // count == 0 is an empty gather, not "all envs"
  return encode_info_dispatch(env, src, count, player, dtype, d_out, row_stride, nullptr, nullptr, S(stream));
}

int coup_vec_information_state_tensor_strided(coup_vec_env* env, int player, int dtype, void* d_out,
                                              uint32_t row_stride, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype) || !valid_stride(row_stride))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_information_state_tensor: bad arguments");
  DeviceGuard guard(env->opts.device);
  const SlabSource src{env->A.state, env->A.history, nullptr, env->A.n, nullptr};
  return encode_info_dispatch(env, src, env->A.n, player, dtype, d_out, row_stride, nullptr, nullptr, S(stream));
}

// ---- finished-episode ring ----------------------------------------------------------------------------------------
int coup_vec_finished_ring_enable(coup_vec_env* env, uint32_t capacity_records) {
  if (!env || (capacity_records & (capacity_records - 1)) != 0)
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_ring_enable: capacity must be 0 or a power of two");
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaDeviceSynchronize());
  cudaFree(env->A.ring); cudaFree(env->A.ring_ctrl);
  env->A.ring = nullptr; env->A.ring_ctrl = nullptr; env->A.ring_mask = 0;
  env->ring_tail = env->ring_dropped = 0;
  if (capacity_records == 0) return COUP_OK;
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->A.ring), static_cast<size_t>(capacity_records) * kRecordWords * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->A.ring_ctrl), 4 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(env->A.ring_ctrl, 0, 4 * sizeof(unsigned long long)));
  env->A.ring_mask = capacity_records - 1;
  return COUP_OK;
}

const uint32_t* coup_vec_finished_ring(const coup_vec_env* env) { return env ? env->A.ring : nullptr; }
const uint64_t* coup_vec_finished_ring_ctrl(const coup_vec_env* env) {
  return env ? reinterpret_cast<const uint64_t*>(env->A.ring_ctrl) : nullptr;
}
uint32_t coup_vec_finished_ring_capacity(const coup_vec_env* env) { return env && env->A.ring ? env->A.ring_mask + 1 : 0; }

int coup_vec_finished_drain(coup_vec_env* env, void* out_records, uint32_t max_records, uint32_t* h_count_out,
                            uint64_t* h_dropped_out, void* stream) {
  if (!env || !h_count_out || (max_records && !out_records)) return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_drain: null argument");
  *h_count_out = 0;
  if (h_dropped_out) *h_dropped_out = env->ring_dropped;
  if (!env->A.ring) return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_drain: the ring is not enabled");
  DeviceGuard guard(env->opts.device);
  cudaStream_t st = S(stream);
  unsigned long long head = 0;
  CUDA_TRY(cudaMemcpyAsync(&head, env->A.ring_ctrl, sizeof(head), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint64_t cap = static_cast<uint64_t>(env->A.ring_mask) + 1;
  if (head - env->ring_tail > cap) {   // the producer lapped the consumer: the oldest records are gone
    env->ring_dropped += head - env->ring_tail - cap;
    env->ring_tail = head - cap;
  }
  const uint64_t n = std::min<uint64_t>(head - env->ring_tail, max_records);
  const size_t rec_bytes = kRecordWords * sizeof(uint32_t);
  uint64_t copied = 0;
  while (copied < n) {                 // at most two pieces (wrap-around)
    const uint64_t pos = (env->ring_tail + copied) & env->A.ring_mask;
    const uint64_t piece = std::min<uint64_t>(n - copied, cap - pos);
    CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out_records) + copied * rec_bytes, env->A.ring + pos * kRecordWords,
                             piece * rec_bytes, cudaMemcpyDefault, st));
    copied += piece;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  env->ring_tail += n;
  *h_count_out = static_cast<uint32_t>(n);
  if (h_dropped_out) *h_dropped_out = env->ring_dropped;
  return COUP_OK;
}

int coup_vec_finished_information_state_tensor(coup_vec_env* env, int player, int dtype, void* d_out, uint32_t row_stride,
                                               uint32_t max_episodes, uint32_t* d_env_ids_out, uint32_t* d_count_out,
                                               void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype) || !valid_stride(row_stride))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_information_state_tensor: bad arguments");
  if (!env->A.ring) return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_information_state_tensor: the ring is not enabled");
  DeviceGuard guard(env->opts.device);
  const RecordSource src{env->A.ring, nullptr, env->A.ring_ctrl, env->A.ring_mask, max_episodes};
  return encode_info_dispatch(env, src, max_episodes, player, dtype, d_out, row_stride, d_env_ids_out, d_count_out, S(stream));
}

int coup_records_information_state_tensor(coup_vec_env* env, const uint32_t* d_records, const uint32_t* d_indices,
                                          uint32_t count, int player, int dtype, void* d_out, uint32_t row_stride, void* stream) {
  if (!env || !d_records || !d_out || !(valid_player_sel(player) || player == COUP_PLAYER_FROM_RECORD) || !valid_dtype(dtype) ||
      !valid_stride(row_stride))
    return fail(COUP_ERR_INVALID_ARG, "coup_records_information_state_tensor: bad arguments");
  DeviceGuard guard(env->opts.device);
  const RecordSource src{d_records, d_indices, nullptr, 0xFFFFFFFFu, count};
  return encode_info_dispatch(env, src, count, player, dtype, d_out, row_stride, nullptr, nullptr, S(stream));
}

int coup_vec_information_state_tensor(coup_vec_env* env, int player, int dtype, void* d_out, void* stream) {
  return coup_vec_information_state_tensor_strided(env, player, dtype, d_out, COUP_INFO_STATE_SIZE, stream);
}

int coup_vec_observation_tensor(coup_vec_env* env, int player, int dtype, void* d_out, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_observation_tensor: bad arguments");
  DeviceGuard guard(env->opts.device);
  const SlabSource src{env->A.state, env->A.history, nullptr, env->A.n, nullptr};
  return encode_obs_dispatch(env, src, env->A.n, player, dtype, d_out, nullptr, nullptr, S(stream));
}

int coup_vec_observer_tensor(coup_vec_env* env, int player, int public_info, int perfect_recall, int private_info, int dtype,
                             void* d_out, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype) || private_info < 0 || private_info > 2)
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_observer_tensor: bad arguments");
  DeviceGuard guard(env->opts.device);
  const uint32_t vis = (private_info == 0 ? kVisPrivateNone : private_info == 2 ? kVisPrivateAll : 0u) | (public_info ? 0u : kVisNoPublic);
  const int sel = player | static_cast<int>(vis << 8);
  const SlabSource src{env->A.state, env->A.history, nullptr, env->A.n, nullptr};
  if (public_info && perfect_recall)    // the info-state layout (2492), other visibility
    return encode_info_dispatch(env, src, env->A.n, sel, dtype, d_out, COUP_INFO_STATE_SIZE, nullptr, nullptr, S(stream));
  return encode_obs_dispatch(env, src, env->A.n, sel, dtype, d_out, nullptr, nullptr, S(stream));   // 98, or 42 without public info
}

int coup_vec_observation_tensor_gather(coup_vec_env* env, const uint32_t* d_env_ids, uint32_t count, int player, int dtype,
                                       void* d_out, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype) || (count > 0 && !d_env_ids))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_observation_tensor_gather: bad arguments");
  DeviceGuard guard(env->opts.device);
  const SlabSource src{env->A.state, env->A.history, d_env_ids, count, nullptr};
  return encode_obs_dispatch(env, src, count, player, dtype, d_out, nullptr, nullptr, S(stream));
}

int coup_vec_finished_observation_tensor(coup_vec_env* env, int player, int dtype, void* d_out, uint32_t max_episodes,
                                         uint32_t* d_env_ids_out, uint32_t* d_count_out, void* stream) {
  if (!env || !d_out || !valid_player_sel(player) || !valid_dtype(dtype))
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_observation_tensor: bad arguments");
  if (!env->A.ring) return fail(COUP_ERR_INVALID_ARG, "coup_vec_finished_observation_tensor: the ring is not enabled");
  DeviceGuard guard(env->opts.device);
  const RecordSource src{env->A.ring, nullptr, env->A.ring_ctrl, env->A.ring_mask, max_episodes};
  return encode_obs_dispatch(env, src, max_episodes, player, dtype, d_out, d_env_ids_out, d_count_out, S(stream));
}

int coup_records_observation_tensor(coup_vec_env* env, const uint32_t* d_records, const uint32_t* d_indices, uint32_t count,
                                    int player, int dtype, void* d_out, void* stream) {
  if (!env || !d_records || !d_out || !(valid_player_sel(player) || player == COUP_PLAYER_FROM_RECORD) || !valid_dtype(dtype))
    return fail(COUP_ERR_INVALID_ARG, "coup_records_observation_tensor: bad arguments");
  DeviceGuard guard(env->opts.device);
  const RecordSource src{d_records, d_indices, nullptr, 0xFFFFFFFFu, count};
  return encode_obs_dispatch(env, src, count, player, dtype, d_out, nullptr, nullptr, S(stream));
}

int coup_vec_step_host(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_legal_mask, int8_t* h_current_player,
                       uint8_t* h_done, int8_t* h_rewards, int dtype, void* d_tensor_out, void* stream) {
  if (!env || !h_actions) return fail(COUP_ERR_INVALID_ARG, "coup_vec_step_host: null argument");
  DeviceGuard guard(env->opts.device);
  cudaStream_t st = S(stream);
  const size_t n = env->A.n;
  CUDA_TRY(cudaMemcpyAsync(env->d_actions, h_actions, n, cudaMemcpyHostToDevice, st));
  int rc = coup_vec_step(env, env->d_actions, nullptr, stream);
  if (rc != COUP_OK) return rc;
  // Host outputs first, then the tensor: the call returns as soon as the host buffers are filled, while
  // the encoder is still running. The tensor is complete in stream order (its consumer is on the device).
  if (h_legal_mask) CUDA_TRY(cudaMemcpyAsync(h_legal_mask, env->A.legal, n * 4, cudaMemcpyDeviceToHost, st));
  if (h_current_player) CUDA_TRY(cudaMemcpyAsync(h_current_player, env->A.cur_player, n, cudaMemcpyDeviceToHost, st));
  if (h_done) CUDA_TRY(cudaMemcpyAsync(h_done, env->A.done, n, cudaMemcpyDeviceToHost, st));
  if (h_rewards) CUDA_TRY(cudaMemcpyAsync(h_rewards, env->A.rewards, n * 2, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(env->host_outputs_ready, st));
  if (d_tensor_out) {
    rc = coup_vec_information_state_tensor(env, COUP_PLAYER_CURRENT, dtype, d_tensor_out, stream);
    if (rc != COUP_OK) return rc;
  }
  CUDA_TRY(cudaEventSynchronize(env->host_outputs_ready));
  return COUP_OK;
}

int coup_vec_step_host_packed_async(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_step_words, int dtype,
                                    void* d_tensor_out, void* stream) {
  if (!env || !h_actions || !h_step_words) return fail(COUP_ERR_INVALID_ARG, "coup_vec_step_host_packed: null argument");
  DeviceGuard guard(env->opts.device);
  cudaStream_t st = S(stream);
  const size_t n = env->A.n;
  CUDA_TRY(cudaMemcpyAsync(env->d_actions, h_actions, n, cudaMemcpyHostToDevice, st));
  int rc = coup_vec_step(env, env->d_actions, nullptr, stream);
  if (rc != COUP_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_step_words, env->A.step_word, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(env->host_outputs_ready, st));
  if (d_tensor_out) {
    rc = coup_vec_information_state_tensor(env, COUP_PLAYER_CURRENT, dtype, d_tensor_out, stream);
    if (rc != COUP_OK) return rc;
  }
  return COUP_OK;
}

int coup_vec_host_outputs_wait(coup_vec_env* env) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaEventSynchronize(env->host_outputs_ready));
  return COUP_OK;
}

int coup_vec_step_host_packed(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_step_words, int dtype,
                              void* d_tensor_out, void* stream) {
  const int rc = coup_vec_step_host_packed_async(env, h_actions, h_step_words, dtype, d_tensor_out, stream);
  return rc != COUP_OK ? rc : coup_vec_host_outputs_wait(env);
}

int coup_host_sample_uniform(const uint32_t* h_legal_mask, uint32_t n, uint64_t seed, uint64_t global_env_offset,
                             uint64_t step, uint8_t* h_actions_out, int threads) {
  if (!h_legal_mask || !h_actions_out) return fail(COUP_ERR_INVALID_ARG, "coup_host_sample_uniform: null argument");
  coup_host::sample_uniform(h_legal_mask, n, seed, global_env_offset, step, h_actions_out, threads);
  return COUP_OK;
}

int coup_vec_stats(coup_vec_env* env, uint64_t* h_out, void* stream) {
  if (!env || !h_out) return fail(COUP_ERR_INVALID_ARG, "coup_vec_stats: null argument");
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaMemcpyAsync(h_out, env->A.stats, COUP_STATS_LEN * sizeof(uint64_t), cudaMemcpyDeviceToHost, S(stream)));
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  return COUP_OK;
}

uint64_t* coup_vec_stats_device(coup_vec_env* env) { return env ? reinterpret_cast<uint64_t*>(env->A.stats) : nullptr; }

int coup_vec_clear_stats(coup_vec_env* env, void* stream) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(env->opts.device);
  CUDA_TRY(cudaMemsetAsync(env->A.stats, 0, COUP_STATS_LEN * sizeof(uint64_t), S(stream)));
  return COUP_OK;
}

int coup_vec_check_errors(coup_vec_env* env, void* stream) {
  uint64_t stats[COUP_STATS_LEN];
  int rc = coup_vec_stats(env, stats, stream);
  if (rc != COUP_OK) return rc;
  if (stats[COUP_STAT_ILLEGAL] != 0)
    return fail(COUP_ERR_ILLEGAL_ACTION, std::to_string(stats[COUP_STAT_ILLEGAL]) + " illegal action(s) rejected");
  return COUP_OK;
}

int coup_tensor_row_hash(const void* d_tensor, int dtype, uint32_t rows, uint32_t row_len, uint64_t* d_hash_out, void* stream) {
  if (rows == 0) return COUP_OK;
  if (!d_tensor || !d_hash_out || !valid_dtype(dtype)) return fail(COUP_ERR_INVALID_ARG, "coup_tensor_row_hash: bad arguments");
  const int dev = device_of(d_tensor);
  if (dev < 0) return fail(COUP_ERR_INVALID_ARG, "coup_tensor_row_hash: d_tensor is not a device pointer");
  DeviceGuard guard(dev);
  const unsigned grid = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  switch (dtype) {
    case COUP_DTYPE_F32:
      k_row_hash<float><<<grid, kBlockThreads, 0, S(stream)>>>(static_cast<const float*>(d_tensor), rows, row_len, d_hash_out);
      break;
    case COUP_DTYPE_U8:
      k_row_hash<uint8_t><<<grid, kBlockThreads, 0, S(stream)>>>(static_cast<const uint8_t*>(d_tensor), rows, row_len, d_hash_out);
      break;
    default:
      k_row_hash<__nv_bfloat16><<<grid, kBlockThreads, 0, S(stream)>>>(static_cast<const __nv_bfloat16*>(d_tensor), rows, row_len, d_hash_out);
      break;
  }
  return launch_status("k_row_hash");
}

// ---- snapshot / restore: raw copy of the packed slab + outputs + statistics + the Philox step counter ----------
namespace {
struct SnapshotHeader {
  uint64_t magic, num_envs, seed, global_env_offset, step_counter, flags;
};
constexpr uint64_t kSnapshotMagic = 0x31304e53424f4355ull;  // "UCOBSN01"
struct Part { void* ptr; size_t bytes; };
std::vector<Part> snapshot_parts(coup_vec_env* env) {
  const size_t n = env->A.n;
  return {{env->A.state, n * 16}, {env->A.history, n * kHistoryWords * 4}, {env->A.legal, n * 4}, {env->A.cur_player, n},
          {env->A.done, n}, {env->A.rewards, n * 2}, {env->A.returns, n * 2}, {env->A.step_word, n * 4},
          {env->A.stats, COUP_STATS_LEN * 8}};
}
}  // namespace

size_t coup_vec_snapshot_size(const coup_vec_env* env) {
  if (!env) return 0;
  size_t total = sizeof(SnapshotHeader);
  for (const Part& p : snapshot_parts(const_cast<coup_vec_env*>(env))) total += p.bytes;
  return total;
}

int coup_vec_snapshot(coup_vec_env* env, void* h_buf, size_t bytes, void* stream) {
  if (!env || !h_buf || bytes < coup_vec_snapshot_size(env)) return fail(COUP_ERR_INVALID_ARG, "coup_vec_snapshot: buffer too small");
  DeviceGuard guard(env->opts.device);
  SnapshotHeader hdr{kSnapshotMagic, env->A.n, env->A.seed, env->A.global_env_offset, env->step_counter, env->A.flags};
  std::memcpy(h_buf, &hdr, sizeof(hdr));
  char* dst = static_cast<char*>(h_buf) + sizeof(hdr);
  for (const Part& p : snapshot_parts(env)) {
    CUDA_TRY(cudaMemcpyAsync(dst, p.ptr, p.bytes, cudaMemcpyDeviceToHost, S(stream)));
    dst += p.bytes;
  }
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  return COUP_OK;
}

int coup_vec_restore(coup_vec_env* env, const void* h_buf, size_t bytes, void* stream) {
  if (!env || !h_buf || bytes < coup_vec_snapshot_size(env)) return fail(COUP_ERR_INVALID_ARG, "coup_vec_restore: buffer too small");
  SnapshotHeader hdr;
  std::memcpy(&hdr, h_buf, sizeof(hdr));
  if (hdr.magic != kSnapshotMagic || hdr.num_envs != env->A.n)
    return fail(COUP_ERR_INVALID_ARG, "coup_vec_restore: not a snapshot of a handle with this many envs");
  DeviceGuard guard(env->opts.device);
  const char* src = static_cast<const char*>(h_buf) + sizeof(hdr);
  for (const Part& p : snapshot_parts(env)) {
    CUDA_TRY(cudaMemcpyAsync(p.ptr, src, p.bytes, cudaMemcpyHostToDevice, S(stream)));
    src += p.bytes;
  }
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  env->A.seed = env->opts.seed = hdr.seed;
  env->A.global_env_offset = env->opts.global_env_offset = hdr.global_env_offset;
  env->step_counter = hdr.step_counter;
  return COUP_OK;
}

uint64_t coup_vec_step_counter(const coup_vec_env* env) { return env ? env->step_counter : 0; }
int coup_vec_set_step_counter(coup_vec_env* env, uint64_t value) {
  if (!env) return fail(COUP_ERR_INVALID_ARG, "null handle");
  env->step_counter = value;
  return COUP_OK;
}

}  // extern "C"
