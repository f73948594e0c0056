// coup_b200.hpp -- header-only C++17 convenience layer over the C ABI (coup_b200.h): RAII handle, exceptions
// instead of status codes. Method names follow the reference interfaces each call stands in for
// (rl_environment.Environment / vector_env.SyncVectorEnv: reset, step; State: InformationStateTensor ...).
// Device pointers are plain `void*` / typed pointers owned by the caller; streams are cudaStream_t as void*.
#ifndef COUP_B200_HPP_
#define COUP_B200_HPP_

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
#include "coup_b200.h"
}

namespace coup_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void Check(int rc) {
  if (rc != COUP_OK) throw Error(rc, coup_last_error());
}

class VectorEnv {
 public:
  VectorEnv(uint32_t num_envs, uint64_t seed = 1234, int device = 0, uint64_t global_env_offset = 0,
            bool reset_if_done = false, bool plain_store_encoder = false) {
    coup_vec_opts opts{};
    opts.num_envs = num_envs;
    opts.device = device;
    opts.seed = seed;
    opts.global_env_offset = global_env_offset;
    opts.flags = (reset_if_done ? COUP_FLAG_AUTO_RESET : 0u) | (plain_store_encoder ? COUP_FLAG_PLAIN_STORE_ENCODER : 0u);
    Check(coup_vec_create(&opts, &env_));
  }
  ~VectorEnv() { coup_vec_destroy(env_); }
  VectorEnv(const VectorEnv&) = delete;
  VectorEnv& operator=(const VectorEnv&) = delete;
  VectorEnv(VectorEnv&& o) noexcept : env_(o.env_) { o.env_ = nullptr; }

  uint32_t size() const { return coup_vec_num_envs(env_); }
  coup_vec_env* handle() { return env_; }

  // SyncVectorEnv.reset(envs_to_reset) / Environment.step(actions)
  void Reset(const uint8_t* d_mask = nullptr, const uint8_t* d_forced_deals = nullptr, void* stream = nullptr) {
    Check(coup_vec_reset(env_, d_mask, d_forced_deals, stream));
  }
  void Step(const uint8_t* d_actions, const uint8_t* d_forced_chance = nullptr, void* stream = nullptr) {
    Check(coup_vec_step(env_, d_actions, d_forced_chance, stream));
  }
  // state.child(action) for `count` (parent, action) pairs: children land in slots 0..count-1 of THIS slab
  void ForkFrom(const VectorEnv& src, const uint32_t* d_parent, const uint8_t* d_actions, uint32_t count,
                const uint8_t* d_forced_chance = nullptr, void* stream = nullptr) {
    Check(coup_vec_fork(env_, src.env_, d_parent, d_actions, d_forced_chance, count, stream));
  }
  void SampleUniform(uint8_t* d_actions_out, void* stream = nullptr) { Check(coup_vec_sample_uniform(env_, d_actions_out, stream)); }
  void SamplePolicy(const void* d_logits, int dtype, float* d_probs_out, uint8_t* d_actions_out, void* stream = nullptr) {
    Check(coup_vec_sample_policy(env_, d_logits, dtype, d_probs_out, d_actions_out, stream));
  }
  void Rollout(int n_steps, int encode_player = -1, int dtype = COUP_DTYPE_F32, void* d_tensor_out = nullptr,
               uint32_t row_stride = COUP_INFO_STATE_SIZE, void* stream = nullptr) {
    Check(coup_vec_rollout_strided(env_, n_steps, encode_player, dtype, d_tensor_out, row_stride, stream));
  }

  // State observers, batched
  void InformationStateTensor(int player, int dtype, void* d_out, uint32_t row_stride = COUP_INFO_STATE_SIZE, void* stream = nullptr) {
    Check(coup_vec_information_state_tensor_strided(env_, player, dtype, d_out, row_stride, stream));
  }
  void ObservationTensor(int player, int dtype, void* d_out, void* stream = nullptr) {
    Check(coup_vec_observation_tensor(env_, player, dtype, d_out, stream));
  }
  void LegalActionsMask(uint8_t* d_out, void* stream = nullptr) { Check(coup_vec_legal_actions_mask(env_, d_out, stream)); }

  const uint32_t* legal_mask() const { return coup_vec_legal_mask(env_); }
  const int8_t* current_player() const { return coup_vec_current_player(env_); }
  const uint8_t* done() const { return coup_vec_done(env_); }
  const int8_t* rewards() const { return coup_vec_rewards(env_); }
  const int8_t* returns() const { return coup_vec_returns(env_); }
  const uint32_t* step_word() const { return coup_vec_step_word(env_); }

  std::array<uint64_t, COUP_STATS_LEN> Stats(void* stream = nullptr) {
    std::array<uint64_t, COUP_STATS_LEN> out{};
    Check(coup_vec_stats(env_, out.data(), stream));
    return out;
  }
  void CheckErrors(void* stream = nullptr) { Check(coup_vec_check_errors(env_, stream)); }

  std::vector<uint8_t> Snapshot(void* stream = nullptr) {
    std::vector<uint8_t> buf(coup_vec_snapshot_size(env_));
    Check(coup_vec_snapshot(env_, buf.data(), buf.size(), stream));
    return buf;
  }
  void Restore(const std::vector<uint8_t>& buf, void* stream = nullptr) { Check(coup_vec_restore(env_, buf.data(), buf.size(), stream)); }

 private:
  coup_vec_env* env_ = nullptr;
};

}  // namespace coup_b200
#endif  // COUP_B200_HPP_
