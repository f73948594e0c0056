// coup_cfr.cuh -- one level of a sampled CFR traversal (python/algorithms/deep_cfr.py:415-525) for all nodes of the level:
// regret matching and child selection, child lists, the device-count forms that keep the level sizes off the host, the
// backward sweep and the packing of a level's nodes into 96-byte records.
#pragma once
#include "coup_step.cuh"

namespace coup {

// ---- one level of a sampled CFR traversal (python/algorithms/deep_cfr.py:415-525), thread per node ------------------
// From the advantage-network outputs of the player to move: regret matching (positive parts over the legal actions,
// normalised; if none is positive, probability one on the legal action with the largest raw advantage, :499-525),
// then which children to expand: at the traverser's nodes every legal action (external sampling, :438-441) or
// min(n_legal, k) actions drawn without replacement from expl * uniform + (1 - expl) * strategy (outcome sampling,
// :442-466; k = outcome_factor, or per node outcome_factor with probability e_outcome and 1 otherwise); at the
// opponent's nodes one action drawn from the strategy (:482-487). Sampling without replacement is the Gumbel-top-k
// order of the log-probabilities, i.e. the sequential renormalised draw of np.random.choice(replace=False).
__device__ __forceinline__ float u01(uint32_t r) { return (static_cast<float>(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ uint32_t cfr_expand_node(const float* __restrict__ adv_row, uint32_t word, uint32_t i, int traverser,
                                                    int external, uint32_t outcome_factor, float e_outcome, float expl,
                                                    uint64_t seed, uint64_t counter, float* __restrict__ strategy_row);

__global__ void __launch_bounds__(kBlockThreads)
k_cfr_expand(const float* __restrict__ advantages, const uint32_t* __restrict__ step_words, uint32_t count,
             int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
             uint64_t counter, float* __restrict__ strategy_out, uint32_t* __restrict__ expand_out,
             uint32_t* __restrict__ count_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t expand = cfr_expand_node(advantages + static_cast<size_t>(i) * 18, step_words[i], i, traverser, external,
                                          outcome_factor, e_outcome, expl, seed, counter, strategy_out + static_cast<size_t>(i) * 18);
  expand_out[i] = expand;
  count_out[i] = __popc(expand);
}

// Children of a level in parent order: child j of node i (j-th set bit of expand[i]) lands at offsets[i] + j, where
// offsets is the exclusive prefix sum of the child counts.
__global__ void __launch_bounds__(kBlockThreads)
k_cfr_children(const uint32_t* __restrict__ expand, const int64_t* __restrict__ offsets, uint32_t count,
               uint32_t* __restrict__ parent_out, uint8_t* __restrict__ action_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t bits = expand[i];
  int64_t pos = offsets[i];
  while (bits) {
    const int a = __ffs(bits) - 1;
    bits &= bits - 1;
    parent_out[pos] = i;
    action_out[pos] = static_cast<uint8_t>(a);
    ++pos;
  }
}

// ---- a whole level of a sampled CFR traversal with DEVICE-side node counts ------------------------------------------------
// The level-by-level expansion above, without the host in the loop: the number of nodes of a level lives in device memory
// (levels shrink and grow with the sampling), every kernel is launched for the capacity of the level buffers and works on
// the first *count nodes, and one single-CTA kernel per level does regret matching, child selection, the prefix sum of the
// child counts and the (parent, action) lists, and leaves the next level's count. The host only checks "is the frontier
// empty" every few levels.
constexpr int kCfrLevelThreads = 1024;

// Regret matching + child selection of ONE node (the body of k_cfr_expand as a function).
__device__ __forceinline__ uint32_t cfr_expand_node(const float* __restrict__ adv_row, uint32_t word, uint32_t i, int traverser,
                                                    int external, uint32_t outcome_factor, float e_outcome, float expl,
                                                    uint64_t seed, uint64_t counter, float* __restrict__ strategy_row) {
  const uint32_t legal = word & 0x3FFFFu;
  const int player = (word >> 18) & 1u;
  const int n_legal = __popc(legal);
  float adv[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) adv[a] = adv_row[a];
  float total = 0.f, best = -INFINITY;
  int best_a = 0;
#pragma unroll
  for (int a = 0; a < 18; ++a) {
    if ((legal >> a) & 1u) {
      total += fmaxf(adv[a], 0.f);
      if (adv[a] > best) { best = adv[a]; best_a = a; }
    }
  }
  float strat[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) {
    const bool ok = (legal >> a) & 1u;
    strat[a] = !ok ? 0.f : total > 0.f ? fmaxf(adv[a], 0.f) / total : (a == best_a ? 1.f : 0.f);
    strategy_row[a] = strat[a];
  }
  uint32_t expand = 0;
  if (n_legal > 0) {
    const uint4 r0 = env_random(seed, i, counter, 2), r1 = env_random(seed, i, counter, 3), r2 = env_random(seed, i, counter, 4);
    if (player != traverser) {
      float sum = 0.f;
#pragma unroll
      for (int a = 0; a < 18; ++a) sum += strat[a];
      const float target = u01(r0.x) * sum;
      float acc = 0.f;
      int pick = best_a;
#pragma unroll
      for (int a = 17; a >= 0; --a) if (strat[a] > 0.f) pick = a;           // fall-back: first action with mass
      bool done = false;
#pragma unroll
      for (int a = 0; a < 18; ++a) {
        if (!done && strat[a] > 0.f) { acc += strat[a]; pick = a; if (target < acc) done = true; }
      }
      expand = 1u << pick;
    } else if (external) {
      expand = legal;
    } else {
      uint32_t k = outcome_factor;
      if (e_outcome >= 0.f) k = u01(r0.y) < e_outcome ? outcome_factor : 1u;
      k = min(k, static_cast<uint32_t>(n_legal));
      float key[18];
      int slot = 0;                                                          // legal actions draw r0.z, r0.w, r1.*, r2.* in order
#pragma unroll
      for (int a = 0; a < 18; ++a) {
        key[a] = -INFINITY;
        if ((legal >> a) & 1u) {
          const uint32_t r = slot == 0 ? r0.z : slot == 1 ? r0.w : slot == 2 ? r1.x : slot == 3 ? r1.y : slot == 4 ? r1.z
                           : slot == 5 ? r1.w : slot == 6 ? r2.x : slot == 7 ? r2.y : slot == 8 ? r2.z : r2.w;
          ++slot;
          const float p = expl / n_legal + (1.f - expl) * strat[a];
          if (p > 0.f) key[a] = logf(p) - logf(-logf(u01(r)));
        }
      }
      for (uint32_t t = 0; t < k; ++t) {
        int arg = -1;
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 18; ++a) if (!((expand >> a) & 1u) && key[a] > m) { m = key[a]; arg = a; }
        if (arg < 0) break;
        expand |= 1u << arg;
      }
    }
  }
  return expand;
}

// One CTA. Nodes [0, *count) of the level: terminal nodes (bit 19 of the step word) expand nothing. Writes, per node,
// strategy [18], expand mask and the exclusive prefix `offset` of its children; per child (parent order, ascending action)
// parent index and action; *next_count = number of children, clipped to `capacity` (then *overflow is set: the level is
// truncated -- the caller sized the buffers too small).
__global__ void __launch_bounds__(kCfrLevelThreads)
k_cfr_level(const float* __restrict__ advantages, const uint32_t* __restrict__ step_words, const uint32_t* __restrict__ count_ptr,
            uint32_t capacity, int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
            uint64_t counter, float* __restrict__ strategy_out, uint32_t* __restrict__ expand_out, uint32_t* __restrict__ offset_out,
            uint32_t* __restrict__ parent_out, uint8_t* __restrict__ action_out, uint32_t* __restrict__ next_count,
            uint32_t* __restrict__ overflow) {
  __shared__ uint32_t s_warp[kCfrLevelThreads / 32];
  __shared__ uint32_t s_base;
  const uint32_t count = min(*count_ptr, capacity);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (uint32_t start = 0; start < count; start += kCfrLevelThreads) {     // uniform trip count
    const uint32_t i = start + threadIdx.x;
    uint32_t expand = 0;
    if (i < count) {
      const uint32_t word = step_words[i];
      if (((word >> 19) & 1u) == 0)
        expand = cfr_expand_node(advantages + static_cast<size_t>(i) * 18, word, i, traverser, external, outcome_factor,
                                 e_outcome, expl, seed, counter, strategy_out + static_cast<size_t>(i) * 18);
      expand_out[i] = expand;
    }
    // block-wide exclusive scan of the child counts of this chunk
    const uint32_t c = __popc(expand);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;                      // inclusive over warps
    }
    __syncthreads();
    const uint32_t base = s_base + (warp ? s_warp[warp - 1] : 0u) + incl - c;
    const uint32_t chunk_total = s_warp[kCfrLevelThreads / 32 - 1];
    if (i < count) {
      offset_out[i] = base;
      uint32_t bits = expand, pos = base;
      while (bits) {
        const int a = __ffs(bits) - 1;
        bits &= bits - 1;
        if (pos < capacity) { parent_out[pos] = i; action_out[pos] = static_cast<uint8_t>(a); }
        ++pos;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t total = s_base;
    *next_count = min(total, capacity);
    if (total > capacity) *overflow = 1u;
  }
}

// Backward pass of one level (deep_cfr.py:468-480, 492-497), thread per node: a terminal node's value is the traverser's
// return; an opponent node's value is its sampled child's; a traverser node's is cfv = sum_a strategy[a] * payoff[a] over
// its expanded children (unsampled actions count as payoff 0, as in the reference), and its sampled regrets are
// payoff[a] - cfv on the legal actions. `child_value` are the values of the next level (node offset[i] + j = j-th child).
__global__ void __launch_bounds__(kBlockThreads)
k_cfr_backward(const uint32_t* __restrict__ step_words, const uint32_t* __restrict__ count_ptr, uint32_t capacity, int traverser,
               const float* __restrict__ strategy, const uint32_t* __restrict__ expand, const uint32_t* __restrict__ offset,
               const double* __restrict__ child_value, double* __restrict__ value_out, float* __restrict__ regret_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(*count_ptr, capacity)) return;
  const uint32_t word = step_words[i];
  const double sign = traverser == 0 ? 1.0 : -1.0;
  if ((word >> 19) & 1u) {                                     // terminal: Returns()[traverser]
    value_out[i] = sign * (static_cast<int>((word >> 24) & 7u) - 2);
    return;
  }
  const uint32_t legal = word & 0x3FFFFu;
  const bool is_trav = static_cast<int>((word >> 18) & 1u) == traverser;
  double payoff[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) payoff[a] = 0.0;
  uint32_t bits = expand[i], pos = offset[i];
  double sum = 0.0;
  while (bits) {
    const int a = __ffs(bits) - 1;
    bits &= bits - 1;
    const double v = pos < capacity ? child_value[pos] : 0.0;
#pragma unroll
    for (int b = 0; b < 18; ++b) if (b == a) payoff[b] = v;
    sum += v;
    ++pos;
  }
  if (!is_trav) { value_out[i] = sum; return; }
  double cfv = 0.0;
#pragma unroll
  for (int a = 0; a < 18; ++a) if ((legal >> a) & 1u) cfv += static_cast<double>(strategy[static_cast<size_t>(i) * 18 + a]) * payoff[a];
  value_out[i] = cfv;
#pragma unroll
  for (int a = 0; a < 18; ++a)
    regret_out[static_cast<size_t>(i) * 18 + a] = ((legal >> a) & 1u) ? static_cast<float>(payoff[a] - cfv) : 0.f;
}

// The nodes [0, *count) of a slab as packed records (history, state, meta: node index, seat<<31 | step word bits 0-26).
__global__ void __launch_bounds__(kBlockThreads)
k_pack_records(EnvArrays A, const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ records) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= min(*count_ptr, A.n)) return;
  uint4* dst = reinterpret_cast<uint4*>(records + static_cast<size_t>(e) * kRecordWords);
  const uint4* h4 = reinterpret_cast<const uint4*>(A.history + static_cast<size_t>(e) * kHistoryWords);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = h4[k];
  dst[4] = A.state[e];
  const uint32_t word = A.step_word[e];
  dst[5] = make_uint4(e, (((word >> 18) & 1u) << 31) | (word & 0x7FFFFFFu), 0u, 0u);
}

}  // namespace coup
