// coup_host_policy.cc -- host-side helpers of the host-buffer path (no game rules in this file):
// a uniform-random policy over legal masks that live in HOST memory, drawing from the same
// Philox4x32-10 stream as the device sampler (coup_device.cuh: env_random, purpose 0, word x), on a
// small persistent thread pool. Compiled by the host compiler; AVX2 path selected at run time.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace coup_host {

namespace {

constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;

inline uint32_t mulhi32(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }

// First output word of Philox4x32-10 with counter (step lo, step hi, 0, seed hi) and key
// (env lo, env hi ^ seed lo): identical to coup::env_random(seed, env, step, 0).x on the device.
inline uint32_t philox_x_scalar(uint64_t seed, uint64_t env, uint64_t step) {
  uint32_t c0 = static_cast<uint32_t>(step), c1 = static_cast<uint32_t>(step >> 32), c2 = 0,
           c3 = static_cast<uint32_t>(seed >> 32);
  uint32_t k0 = static_cast<uint32_t>(env), k1 = static_cast<uint32_t>(env >> 32) ^ static_cast<uint32_t>(seed);
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(kM0, c0), lo0 = kM0 * c0, hi1 = mulhi32(kM1, c2), lo1 = kM1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += kW0; k1 += kW1;
  }
  return c0;
}

struct Mul256 {
  __m256i hi, lo;
};
__attribute__((target("avx2"))) inline Mul256 mul_hilo_256(__m256i m, __m256i c) {   // see mul_hilo_512
  const __m256i even = _mm256_mul_epu32(m, c);
  const __m256i odd = _mm256_mul_epu32(m, _mm256_srli_epi64(c, 32));
  Mul256 r;
  r.hi = _mm256_blend_epi32(_mm256_srli_epi64(even, 32), odd, 0xAA);
  r.lo = _mm256_blend_epi32(even, _mm256_slli_epi64(odd, 32), 0xAA);
  return r;
}

// Eight consecutive envs per iteration. Requires that env lo does not wrap inside [env, env+count).
__attribute__((target("avx2"))) void philox_x_avx2(uint64_t seed, uint64_t env, uint64_t step, uint32_t count, uint32_t* out) {
  const __m256i m0 = _mm256_set1_epi32(static_cast<int>(kM0)), m1 = _mm256_set1_epi32(static_cast<int>(kM1));
  const __m256i w0 = _mm256_set1_epi32(static_cast<int>(kW0)), w1 = _mm256_set1_epi32(static_cast<int>(kW1));
  const __m256i iota = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
  uint32_t i = 0;
  for (; i + 8 <= count; i += 8) {
    __m256i c0 = _mm256_set1_epi32(static_cast<int>(static_cast<uint32_t>(step)));
    __m256i c1 = _mm256_set1_epi32(static_cast<int>(static_cast<uint32_t>(step >> 32)));
    __m256i c2 = _mm256_setzero_si256();
    __m256i c3 = _mm256_set1_epi32(static_cast<int>(static_cast<uint32_t>(seed >> 32)));
    __m256i k0 = _mm256_add_epi32(_mm256_set1_epi32(static_cast<int>(static_cast<uint32_t>(env + i))), iota);
    __m256i k1 = _mm256_set1_epi32(static_cast<int>(static_cast<uint32_t>((env + i) >> 32) ^ static_cast<uint32_t>(seed)));
    for (int r = 0; r < 10; ++r) {
      const Mul256 p0 = mul_hilo_256(m0, c0), p1 = mul_hilo_256(m1, c2);
      c0 = _mm256_xor_si256(_mm256_xor_si256(p1.hi, c1), k0);
      c2 = _mm256_xor_si256(_mm256_xor_si256(p0.hi, c3), k1);
      c1 = p1.lo; c3 = p0.lo;
      k0 = _mm256_add_epi32(k0, w0);
      k1 = _mm256_add_epi32(k1, w1);
    }
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i), c0);
  }
  for (; i < count; ++i) out[i] = philox_x_scalar(seed, env + i, step);
}

// 32x32 -> 64-bit products of 16 lanes from two VPMULUDQ (even and odd lanes); hi and lo halves are both taken from
// them (VPMULLD is a slow two-uop instruction on every AVX-512 core).
struct Mul512 {
  __m512i hi, lo;
};
__attribute__((target("avx512f"))) inline Mul512 mul_hilo_512(__m512i m, __m512i c) {
  const __m512i even = _mm512_mul_epu32(m, c);
  const __m512i odd = _mm512_mul_epu32(m, _mm512_srli_epi64(c, 32));   // m is a broadcast constant: same in odd lanes
  Mul512 r;
  r.hi = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(even, 32), odd);
  r.lo = _mm512_mask_blend_epi32(0xAAAA, even, _mm512_slli_epi64(odd, 32));
  return r;
}

// 32 consecutive envs per iteration: two independent 16-lane Philox states interleaved, so that the multiplier
// pipeline is busy while the other state's round result is still in flight (same stream as the scalar version).
__attribute__((target("avx512f"))) void philox_x_avx512(uint64_t seed, uint64_t env, uint64_t step, uint32_t count, uint32_t* out) {
  const __m512i m0 = _mm512_set1_epi32(static_cast<int>(kM0)), m1 = _mm512_set1_epi32(static_cast<int>(kM1));
  const __m512i w0 = _mm512_set1_epi32(static_cast<int>(kW0)), w1 = _mm512_set1_epi32(static_cast<int>(kW1));
  const __m512i iota = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  const __m512i s0 = _mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>(step)));
  const __m512i s1 = _mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>(step >> 32)));
  const __m512i s3 = _mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>(seed >> 32)));
  uint32_t i = 0;
  for (; i + 32 <= count; i += 32) {
    __m512i a0 = s0, a1 = s1, a2 = _mm512_setzero_si512(), a3 = s3;
    __m512i b0 = s0, b1 = s1, b2 = _mm512_setzero_si512(), b3 = s3;
    __m512i ka0 = _mm512_add_epi32(_mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>(env + i))), iota);
    __m512i kb0 = _mm512_add_epi32(_mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>(env + i + 16))), iota);
    __m512i ka1 = _mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>((env + i) >> 32) ^ static_cast<uint32_t>(seed)));
    __m512i kb1 = _mm512_set1_epi32(static_cast<int>(static_cast<uint32_t>((env + i + 16) >> 32) ^ static_cast<uint32_t>(seed)));
#pragma GCC unroll 10
    for (int r = 0; r < 10; ++r) {
      const Mul512 pa0 = mul_hilo_512(m0, a0), pa1 = mul_hilo_512(m1, a2);
      const Mul512 pb0 = mul_hilo_512(m0, b0), pb1 = mul_hilo_512(m1, b2);
      a0 = _mm512_xor_si512(_mm512_xor_si512(pa1.hi, a1), ka0);
      a2 = _mm512_xor_si512(_mm512_xor_si512(pa0.hi, a3), ka1);
      a1 = pa1.lo; a3 = pa0.lo;
      b0 = _mm512_xor_si512(_mm512_xor_si512(pb1.hi, b1), kb0);
      b2 = _mm512_xor_si512(_mm512_xor_si512(pb0.hi, b3), kb1);
      b1 = pb1.lo; b3 = pb0.lo;
      ka0 = _mm512_add_epi32(ka0, w0); ka1 = _mm512_add_epi32(ka1, w1);
      kb0 = _mm512_add_epi32(kb0, w0); kb1 = _mm512_add_epi32(kb1, w1);
    }
    _mm512_storeu_si512(out + i, a0);
    _mm512_storeu_si512(out + i + 16, b0);
  }
  for (; i < count; ++i) out[i] = philox_x_scalar(seed, env + i, step);
}

// k-th legal action with BMI2: PDEP deposits a single bit at the position of the k-th set bit of the mask.
__attribute__((target("bmi2,popcnt"))) void select_bmi2(const uint32_t* x, uint32_t count, const uint32_t* words, uint8_t* actions) {
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t m = words[i] & 0x3FFFFu;
    const uint32_t k = mulhi32(x[i], static_cast<uint32_t>(__builtin_popcount(m)));
    actions[i] = m ? static_cast<uint8_t>(__builtin_ctz(_pdep_u32(1u << k, m))) : 0xFF;
  }
}

struct KthBitTable {
  uint8_t pos[64][6];
  constexpr KthBitTable() : pos() {
    for (int m = 0; m < 64; ++m) {
      int k = 0;
      for (int b = 0; b < 6; ++b)
        if ((m >> b) & 1) pos[m][k++] = static_cast<uint8_t>(b);
      for (; k < 6; ++k) pos[m][k] = 0;
    }
  }
};
constexpr KthBitTable kKthBit{};
struct PopTable {
  uint8_t n[64];
  constexpr PopTable() : n() {
    for (int m = 0; m < 64; ++m) n[m] = static_cast<uint8_t>((m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1) + ((m >> 4) & 1) + ((m >> 5) & 1));
  }
};
constexpr PopTable kPop{};

struct CpuFeatures {
  bool avx2, avx512, bmi2;
};

void sample_tile(uint64_t seed, uint64_t first_env, uint64_t step, uint32_t count, const uint32_t* words, uint8_t* actions,
                 CpuFeatures cpu) {
  uint32_t x[2048];
  const bool wraps = static_cast<uint32_t>(first_env) > static_cast<uint32_t>(first_env + count);
  if (cpu.avx512 && !wraps) philox_x_avx512(seed, first_env, step, count, x);
  else if (cpu.avx2 && !wraps) philox_x_avx2(seed, first_env, step, count, x);
  else for (uint32_t i = 0; i < count; ++i) x[i] = philox_x_scalar(seed, first_env + i, step);
  if (cpu.bmi2) { select_bmi2(x, count, words, actions); return; }
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t m = words[i] & 0x3FFFFu;  // legal mask, or the low 18 bits of a packed step word
    const uint32_t c0 = kPop.n[m & 63u], c1 = kPop.n[(m >> 6) & 63u], c2 = kPop.n[m >> 12];
    uint32_t k = mulhi32(x[i], c0 + c1 + c2);  // k-th (0-based) legal action, no data-dependent branch
    const uint32_t ge0 = 0u - static_cast<uint32_t>(k >= c0), ge1 = 0u - static_cast<uint32_t>(k >= c0 + c1);
    const uint32_t chunk = (ge0 & 1u) + (ge1 & 1u);
    k -= (c0 & ge0) + (c1 & ge1);
    const uint32_t bits = (m >> (6u * chunk)) & 63u;
    actions[i] = m ? static_cast<uint8_t>(6u * chunk + kKthBit.pos[bits][k]) : 0xFF;
  }
}

// Minimal persistent pool: workers sleep on a condition variable between parallel_for calls.
class HostPool {
 public:
  static HostPool& instance() {
    static HostPool pool;
    return pool;
  }
  void parallel_for(uint32_t items, int threads, const std::function<void(uint32_t)>& fn) {
    threads = std::max(1, std::min(threads, 256));
    if (threads == 1 || items < 2) {
      for (uint32_t i = 0; i < items; ++i) fn(i);
      return;
    }
    std::lock_guard<std::mutex> run_lock(run_mutex_);  // one parallel_for at a time
    while (static_cast<int>(workers_.size()) < threads - 1) {
      const int id = static_cast<int>(workers_.size());
      workers_.emplace_back([this, id] { worker(id); });
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      items_ = items;
      next_.store(0, std::memory_order_relaxed);
      active_ = threads - 1;
      pending_ = active_;
      ++generation_;
    }
    cv_.notify_all();
    for (uint32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < items;) fn(i);  // the caller works too
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() = default;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  void worker(int id) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    while (true) {
      cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      if (id >= active_) continue;
      const std::function<void(uint32_t)>* f = fn_;
      const uint32_t items = items_;
      lk.unlock();
      for (uint32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < items;) (*f)(i);
      lk.lock();
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  std::mutex run_mutex_, m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(uint32_t)>* fn_ = nullptr;
  std::atomic<uint32_t> next_{0};
  uint32_t items_ = 0;
  int active_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

}  // namespace

void sample_uniform(const uint32_t* words, uint32_t n, uint64_t seed, uint64_t global_env_offset, uint64_t step,
                    uint8_t* actions, int threads) {
  static const CpuFeatures detected = {static_cast<bool>(__builtin_cpu_supports("avx2")), static_cast<bool>(__builtin_cpu_supports("avx512f")),
                                       static_cast<bool>(__builtin_cpu_supports("bmi2")) && static_cast<bool>(__builtin_cpu_supports("popcnt"))};
  CpuFeatures cpu = detected;
  // COUP_B200_HOST_ISA=avx2|scalar restricts the vector paths (tests run every path the host has; all give the same stream)
  if (const char* isa = std::getenv("COUP_B200_HOST_ISA")) {
    const std::string want(isa);
    if (want == "avx2") cpu.avx512 = false;
    if (want == "scalar") cpu.avx512 = cpu.avx2 = cpu.bmi2 = false;
  }
  constexpr uint32_t kTile = 2048;
  const uint32_t tiles = (n + kTile - 1) / kTile;
  // A tile is ~2 us of work: waking a worker pays off from about four tiles per thread, and beyond 16 threads the wake-up
  // and join cost more than they save (measured on the 32-core GPU box: 32 threads were slower than 16).
  threads = std::max(1, std::min({threads, 16, static_cast<int>(tiles / 4)}));
  HostPool::instance().parallel_for(tiles, threads, [=](uint32_t t) {
    const uint32_t lo = t * kTile, cnt = std::min(kTile, n - lo);
    sample_tile(seed, global_env_offset + lo, step, cnt, words + lo, actions + lo, cpu);
  });
}

}  // namespace coup_host
