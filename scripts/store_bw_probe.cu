// store_bw_probe.cu -- what a pure write stream of the info-state shape can reach on this GPU, for the store
// strategies the encoder could use. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw_probe store_bw_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kRowBytes = 9968;                 // one fp32 info-state row
constexpr int kRowUnits = kRowBytes / 16;       // 623

__global__ void k_fill(uint4* out, size_t n_units) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (; i < n_units; i += stride) out[i] = z;
}

// warp per 32 consecutive rows, lane-strided 16-byte stores row by row (the plain-store encoder's pattern)
__global__ void k_rows_lsu(uint4* out, uint32_t n_rows, int row_units) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (r0 >= n_rows) return;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int r = 0; r < 32; ++r) {
    uint4* row = out + (r0 + r) * row_units;
#pragma unroll 4
    for (int q = lane; q < row_units; q += 32) row[q] = z;
  }
}

// same but the warp treats its 32 rows as ONE contiguous span (no per-row loop restart)
__global__ void k_span_lsu(uint4* out, uint32_t n_rows, int row_units) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (r0 >= n_rows) return;
  const uint4 z = make_uint4(0, 0, 0, 0);
  uint4* p = out + r0 * row_units;
  const int total = 32 * row_units;
#pragma unroll 8
  for (int q = lane; q < total; q += 32) p[q] = z;
}

// torch-like: no loops; block b owns a contiguous chunk of U * blockDim * 16 bytes, thread t writes U units strided by blockDim
template <int U, int HINT>
__global__ void k_chunk(uint4* out, size_t n_units) {
  const size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const size_t i = base + (size_t)u * blockDim.x;
    if (i < n_units) {
      if (HINT == 0) out[i] = z;
      else if (HINT == 1) __stcs(out + i, z);
      else if (HINT == 2) __stwt(out + i, z);
      else asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(out + i), "r"(0) : "memory");
    }
  }
}

// copy-out of a per-warp staging buffer with LDS.128 + STG.128 (what a non-TMA staged encoder would do)
__global__ void k_rows_stage_lsu(uint4* out, uint32_t n_rows, int row_units) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* stage = reinterpret_cast<uint4*>(smem) + (size_t)warp * 624;
  for (int q = lane; q < 624; q += 32) stage[q] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const size_t r0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (r0 >= n_rows) return;
  for (int r = 0; r < 32; ++r) {
    uint4* row = out + (r0 + r) * row_units;
#pragma unroll 4
    for (int q = lane; q < row_units; q += 32) row[q] = stage[q];
  }
}

__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(saddr), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// warp per 32 rows, one bulk store per row from a zero staging buffer; wait_each: wait for the read after every op
template <bool kWaitEach>
__global__ void k_rows_tma(unsigned char* out, uint32_t n_rows, int row_bytes, int stage_bytes_per_warp) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + (size_t)warp * stage_bytes_per_warp;
  for (int q = lane; q < stage_bytes_per_warp / 16; q += 32) reinterpret_cast<uint4*>(stage)[q] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t r0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (r0 >= n_rows) return;
  if (lane == 0) {
    for (int r = 0; r < 32; ++r) {
      bulk_store(out + (r0 + r) * (size_t)row_bytes, stage, row_bytes);
      if (kWaitEach) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// block of W warps owns 32*W consecutive rows; at step r warp w writes row base + r*W + w (adjacent rows in flight)
__global__ void k_rows_tma_interleaved(unsigned char* out, uint32_t n_rows, int row_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  unsigned char* stage = smem + (size_t)warp * 9984;
  for (int q = lane; q < 624; q += 32) reinterpret_cast<uint4*>(stage)[q] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t base = (size_t)blockIdx.x * 32 * W;
  if (lane == 0) {
    for (int r = 0; r < 32; ++r) {
      bulk_store(out + (base + (size_t)r * W + warp) * row_bytes, stage, row_bytes);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// persistent: every warp takes the next row from a global counter (all resident warps write adjacent rows)
__global__ void k_rows_tma_sweep(unsigned char* out, uint32_t n_rows, int row_bytes, unsigned int* counter, int rows_per_grab) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage = smem + (size_t)warp * 9984;
  for (int q = lane; q < 624; q += 32) reinterpret_cast<uint4*>(stage)[q] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    while (true) {
      const unsigned int r0 = atomicAdd(counter, rows_per_grab);
      if (r0 >= n_rows) break;
      for (int k = 0; k < rows_per_grab; ++k) {
        bulk_store(out + (size_t)(r0 + k) * row_bytes, stage, row_bytes);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// block sweeps its own contiguous region (rows_per_block rows) with 4 KB block-wide steps of 16-byte stores
__global__ void k_blockspan_lsu(uint4* out, uint32_t n_rows, int rows_per_block, int unroll_dummy) {
  const size_t base = (size_t)blockIdx.x * rows_per_block * kRowUnits;
  const int total = rows_per_block * kRowUnits;
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 4
  for (int u = threadIdx.x; u < total; u += blockDim.x) out[base + u] = z;
}

template <typename F>
float time_ms(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const uint32_t n_rows = 1u << 20;
  const size_t bytes_max = (size_t)n_rows * 9984;
  unsigned char* out;
  cudaMalloc(&out, bytes_max);
  auto report = [&](const char* name, float ms, size_t bytes) { printf("%-44s %8.3f ms  %7.0f GB/s\n", name, ms, bytes / ms / 1e6); };
  const size_t bytes = (size_t)n_rows * kRowBytes;
  report("grid-stride fill, 148x8 blocks x 256", time_ms([&] { k_fill<<<148 * 8, 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("grid-stride fill, 4096 blocks x 256", time_ms([&] { k_fill<<<4096, 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("rows LSU (per-row loop), 8 warps/block", time_ms([&] { k_rows_lsu<<<n_rows / 256, 256>>>((uint4*)out, n_rows, kRowUnits); }), bytes);
  report("rows LSU padded 9984 B rows", time_ms([&] { k_rows_lsu<<<n_rows / 256, 256>>>((uint4*)out, n_rows, 624); }), (size_t)n_rows * 9984);
  report("span LSU (32 rows contiguous)", time_ms([&] { k_span_lsu<<<n_rows / 256, 256>>>((uint4*)out, n_rows, kRowUnits); }), bytes);
  report("chunk U=4 plain  (torch-like)", time_ms([&] { k_chunk<4, 0><<<(unsigned)((bytes / 16 + 1023) / 1024), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=1 plain", time_ms([&] { k_chunk<1, 0><<<(unsigned)((bytes / 16 + 255) / 256), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=8 plain", time_ms([&] { k_chunk<8, 0><<<(unsigned)((bytes / 16 + 2047) / 2048), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=4 __stcs", time_ms([&] { k_chunk<4, 1><<<(unsigned)((bytes / 16 + 1023) / 1024), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=4 __stwt", time_ms([&] { k_chunk<4, 2><<<(unsigned)((bytes / 16 + 1023) / 1024), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=4 L1::no_allocate", time_ms([&] { k_chunk<4, 3><<<(unsigned)((bytes / 16 + 1023) / 1024), 256>>>((uint4*)out, bytes / 16); }), bytes);
  report("chunk U=4 plain, 128 threads", time_ms([&] { k_chunk<4, 0><<<(unsigned)((bytes / 16 + 511) / 512), 128>>>((uint4*)out, bytes / 16); }), bytes);
  cudaFuncSetAttribute(k_rows_stage_lsu, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9984);
  report("rows staged LDS+STG, 8 warps/block", time_ms([&] { k_rows_stage_lsu<<<n_rows / 256, 256, 8 * 9984>>>((uint4*)out, n_rows, kRowUnits); }), bytes);
  for (int warps : {8}) {
    const int smem = warps * 9984;
    cudaFuncSetAttribute(k_rows_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_rows_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    char name[96];
    snprintf(name, sizeof name, "rows TMA bulk 9968 B, wait each, %d warps/block", warps);
    report(name, time_ms([&] { k_rows_tma<true><<<n_rows / (32 * warps), 32 * warps, smem>>>(out, n_rows, 9968, 9984); }), bytes);
    snprintf(name, sizeof name, "rows TMA bulk 9968 B, no wait,   %d warps/block", warps);
    report(name, time_ms([&] { k_rows_tma<false><<<n_rows / (32 * warps), 32 * warps, smem>>>(out, n_rows, 9968, 9984); }), bytes);
    snprintf(name, sizeof name, "rows TMA bulk 9984 B aligned, wait each, %d w", warps);
    report(name, time_ms([&] { k_rows_tma<true><<<n_rows / (32 * warps), 32 * warps, smem>>>(out, n_rows, 9984, 9984); }), (size_t)n_rows * 9984);
  }
  cudaFuncSetAttribute(k_rows_tma_interleaved, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 9984);
  report("rows TMA interleaved in block, 8 warps", time_ms([&] { k_rows_tma_interleaved<<<n_rows / 256, 256, 8 * 9984>>>(out, n_rows, 9968); }), bytes);
  report("rows TMA interleaved in block, 16 warps", time_ms([&] { k_rows_tma_interleaved<<<n_rows / 512, 512, 16 * 9984>>>(out, n_rows, 9968); }), bytes);
  report("block-span LSU, 256 rows/block, 256 thr", time_ms([&] { k_blockspan_lsu<<<n_rows / 256, 256>>>((uint4*)out, n_rows, 256, 0); }), bytes);
  report("block-span LSU, 256 rows/block, 512 thr", time_ms([&] { k_blockspan_lsu<<<n_rows / 256, 512>>>((uint4*)out, n_rows, 256, 0); }), bytes);
  report("block-span LSU, 64 rows/block, 256 thr", time_ms([&] { k_blockspan_lsu<<<n_rows / 64, 256>>>((uint4*)out, n_rows, 64, 0); }), bytes);
  report("block-span LSU, 32 rows/block, 256 thr", time_ms([&] { k_blockspan_lsu<<<n_rows / 32, 256>>>((uint4*)out, n_rows, 32, 0); }), bytes);
  report("block-span LSU, 8 rows/block, 256 thr", time_ms([&] { k_blockspan_lsu<<<n_rows / 8, 256>>>((uint4*)out, n_rows, 8, 0); }), bytes);
  unsigned int* counter;
  cudaMalloc(&counter, 4);
  cudaFuncSetAttribute(k_rows_tma_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9984);
  for (int grab : {1, 4, 32}) {
    char name[96];
    snprintf(name, sizeof name, "rows TMA global sweep, grab %d rows, 296 blocks x 8w", grab);
    report(name, time_ms([&] { cudaMemsetAsync(counter, 0, 4); k_rows_tma_sweep<<<296, 256, 8 * 9984>>>(out, n_rows, 9968, counter, grab); }), bytes);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
