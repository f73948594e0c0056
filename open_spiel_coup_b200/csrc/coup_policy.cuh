// coup_policy.cuh -- action sampling and the dense legal mask: uniform-random legal actions, masked softmax sampling of a
// policy network's logits (python/algorithms/nfsp.py:154-167), State::LegalActionsMask for every env.
#pragma once
#include "coup_step.cuh"

namespace coup {

// ---- uniform-random legal action (same draw the fused rollout would use at this step counter) ------
__global__ void __launch_bounds__(kBlockThreads)
k_sample_uniform(EnvArrays A, uint8_t* __restrict__ actions_out, uint64_t step) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n) return;
  const Env s = load_env(A.state + e);
  uint32_t a = 0xFFu;
  if (!is_terminal(s)) {
    const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 0);
    a = sample_action(legal_mask_decision(s), rnd.x);
  }
  actions_out[e] = static_cast<uint8_t>(a);
}

// ---- masked policy sampling: the acting rule of the reference's agents (python/algorithms/nfsp.py:154-167)
// fused on the device: probs = softmax(logits); illegal -> 0; renormalise; action ~ probs. One thread per
// env; the draw is the x word of the step's Philox block (the slot coup_vec_sample_uniform uses).
template <typename T> __device__ __forceinline__ float logit_to_float(T v);
template <> __device__ __forceinline__ float logit_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float logit_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Rows of 18 elements are 72 (36) bytes apart: a lane reading its own row touches a different cache line than its
// neighbour for every element. The warp therefore moves its 32 rows -- one contiguous 2 304-byte span -- with coalesced
// loads/stores through shared memory (row pitch 19 words: conflict-free) and each lane works on its row there.
constexpr int kRowPitch = kNumActions + 1;

#ifndef COUP_POLICY_BLOCKS
#define COUP_POLICY_BLOCKS 5   // resident CTAs per SM: 56.7 -> 47.2 us per 2^20 envs with probabilities, 37.7 -> 29.7 without
#endif
template <typename T>
__global__ void __launch_bounds__(kBlockThreads, COUP_POLICY_BLOCKS)
k_sample_policy(EnvArrays A, const T* __restrict__ logits, float* __restrict__ probs_out,
                uint8_t* __restrict__ actions_out, uint64_t step) {
  __shared__ float s_rows[kWarpsPerBlock][32 * kRowPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + warp) * 32u;
  if (e0 >= A.n) return;
  const uint32_t e = e0 + lane;
  const uint32_t span = min(32u, A.n - e0) * kNumActions;
  float* rows = s_rows[warp];
  const T* src = logits + static_cast<size_t>(e0) * kNumActions;
#pragma unroll
  for (uint32_t i = 0; i < kNumActions; ++i) {
    const uint32_t j = lane + 32u * i;
    if (j < span) {
      const uint32_t r = (j * 3641u) >> 16;                    // j / 18 for j < 576
      rows[r * kRowPitch + (j - r * kNumActions)] = logit_to_float<T>(src[j]);
    }
  }
  __syncwarp();
  const uint32_t legal = e < A.n ? A.legal[e] : 0u;
  float p[kNumActions];
  float mx = -INFINITY;
#pragma unroll
  for (int a = 0; a < kNumActions; ++a) {
    p[a] = rows[lane * kRowPitch + a];
    if ((legal >> a) & 1u) mx = fmaxf(mx, p[a]);
  }
  // softmax over all actions followed by masking and renormalising == softmax over the legal ones;
  // subtracting the legal maximum keeps it finite.
  float sum = 0.f;
#pragma unroll
  for (int a = 0; a < kNumActions; ++a) {
    p[a] = ((legal >> a) & 1u) ? expf(p[a] - mx) : 0.f;
    sum += p[a];
  }
  uint32_t action = 0xFFu;
  if (legal != 0) {
    const float inv = 1.f / sum;
    const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 0);
    const float u = static_cast<float>(rnd.x >> 8) * (1.0f / 16777216.0f);  // 24-bit uniform in [0,1)
    float cdf = 0.f;
    action = 31u - __clz(legal);  // falls back to the last legal action if rounding leaves u >= cdf
    bool found = false;
#pragma unroll
    for (int a = 0; a < kNumActions; ++a) {
      p[a] *= inv;
      cdf += p[a];
      if (!found && ((legal >> a) & 1u) && u < cdf) { action = a; found = true; }
    }
  }
  if (e < A.n) actions_out[e] = static_cast<uint8_t>(action);
  if (probs_out != nullptr) {
    __syncwarp();
#pragma unroll
    for (int a = 0; a < kNumActions; ++a) rows[lane * kRowPitch + a] = legal ? p[a] : 0.f;
    __syncwarp();
    float* dst = probs_out + static_cast<size_t>(e0) * kNumActions;
#pragma unroll
    for (uint32_t i = 0; i < kNumActions; ++i) {
      const uint32_t j = lane + 32u * i;
      if (j < span) {
        const uint32_t r = (j * 3641u) >> 16;
        dst[j] = rows[r * kRowPitch + (j - r * kNumActions)];
      }
    }
  }
}

// ---- dense legal mask: uint8[n][18] (State::LegalActionsMask, spiel.cc:371-377) ----------------------
// Every lane expands its own env's mask into 18 bytes (nine byte-pair stores) in the warp's 576-byte staging span in shared
// memory; the span then leaves as 36 sixteen-byte stores.
__global__ void __launch_bounds__(kBlockThreads)
k_legal_actions_mask(const uint32_t* __restrict__ legal, uint8_t* __restrict__ out, uint32_t n) {
  __shared__ __align__(16) uint8_t s_span[kWarpsPerBlock][32 * kNumActions];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + warp) * 32u;
  if (e0 >= n) return;
  const uint32_t mine = e0 + lane < n ? legal[e0 + lane] : 0u;
  uint16_t* row = reinterpret_cast<uint16_t*>(s_span[warp] + lane * kNumActions);      // 18 * lane is even
#pragma unroll
  for (uint32_t k = 0; k < kNumActions / 2; ++k)
    row[k] = static_cast<uint16_t>(((mine >> (2u * k)) & 1u) | (((mine >> (2u * k + 1u)) & 1u) << 8));
  __syncwarp();
  const uint32_t span = min(32u, n - e0) * kNumActions;                 // bytes of this warp
  uint8_t* dst = out + static_cast<size_t>(e0) * kNumActions;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && span == 32u * kNumActions) {
    const uint4* src = reinterpret_cast<const uint4*>(s_span[warp]);
    reinterpret_cast<uint4*>(dst)[lane] = src[lane];
    if (lane < 4) reinterpret_cast<uint4*>(dst)[32 + lane] = src[32 + lane];
  } else {
    for (uint32_t i = lane; i < span; i += 32u) dst[i] = s_span[warp][i];
  }
}

}  // namespace coup
