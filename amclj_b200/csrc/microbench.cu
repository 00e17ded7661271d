// microbench.cu — measured ceilings for the ray caster's inner operation: a
// chain of DEPENDENT distance-field lookups, from shared memory (nibble-packed table, the
// lgfloor placement) and from an L2-resident byte table (the large-map placement), with the
// least arithmetic that still makes the next address depend on the loaded value.  bench.py
// reports the sensor kernel's probe rate against these.
#include "internal.h"

namespace amclj {

constexpr int MB_THREADS = 1024;

__global__ void __launch_bounds__(MB_THREADS, 1)
smem_lookup_kernel(const uint8_t *__restrict__ table, uint32_t table_bytes, int iters,
                   unsigned long long *sink) {
  extern __shared__ uint4 smem_dyn[];
  uint8_t *s = reinterpret_cast<uint8_t *>(smem_dyn);
  for (uint32_t i = threadIdx.x * 16u; i < table_bytes; i += blockDim.x * 16u)
    *reinterpret_cast<uint4 *>(s + i) = *reinterpret_cast<const uint4 *>(table + i);
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
  const uint32_t mask = table_bytes * 8u - 1u; /* index in units of 4 per nibble, power of two */
  uint32_t idx = (blockIdx.x * 7919u + threadIdx.x * 104729u) & mask, acc = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      uint32_t byte;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(byte) : "r"(base + (idx >> 3)));
      uint32_t d = (byte >> (idx & 4u)) & 15u;
      acc += d;
      idx = (idx + d * 2724u + 1324u) & mask; /* next probe depends on this one */
    }
  }
  if (acc == 0xffffffffu) *sink = acc;
}

__global__ void __launch_bounds__(MB_THREADS, 1)
l2_lookup_kernel(const uint8_t *__restrict__ table, uint32_t mask, int iters,
                 unsigned long long *sink) {
  uint32_t idx = (blockIdx.x * 7919u + threadIdx.x * 104729u) * 97u & mask, acc = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      uint32_t d = __ldg(table + idx);
      acc += d;
      idx = (idx + d * 70001u + 12347u) & mask;
    }
  }
  if (acc == 0xffffffffu) *sink = acc;
}

__global__ void fill_random_kernel(uint8_t *p, size_t n, uint32_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, 0x4d42u, seed, 0u, r);
  p[i] = (uint8_t)(r[0] | 0x11u); /* no zero nibble: every lookup moves on */
}

}  // namespace amclj

using namespace amclj;

extern "C" int32_t amclj_microbench(amclj_ctx *ctx, int32_t which, double *lookups_per_s) {
  if (!ctx || !lookups_per_s) return AMCLJ_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const bool smem = which == 0;
  const size_t bytes = smem ? (128u << 10) : (which == 2 ? (64u << 10) : (64u << 20));
  uint8_t *tab = nullptr;
  if (cudaMalloc(&tab, bytes) != cudaSuccess) return AMCLJ_ERR_NOMEM;
  fill_random_kernel<<<(unsigned)((bytes + 255) / 256), 256, 0, ctx->stream>>>(tab, bytes, 12345u);
  const int iters = smem || which == 2 ? 2048 : 256;
  const int blocks = ctx->sm_count;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaError_t err = cudaSuccess;
  if (smem) err = cudaFuncSetAttribute(smem_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  float best = 1e30f;
  for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) { /* first repetition warms up */
    cudaEventRecord(e0, ctx->stream);
    if (smem)
      smem_lookup_kernel<<<blocks, MB_THREADS, bytes, ctx->stream>>>(tab, (uint32_t)bytes, iters, ctx->stats + 7);
    else
      l2_lookup_kernel<<<blocks, MB_THREADS, 0, ctx->stream>>>(tab, (uint32_t)bytes - 1u, iters, ctx->stats + 7);
    cudaEventRecord(e1, ctx->stream);
    err = cudaStreamSynchronize(ctx->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(tab);
  if (err != cudaSuccess) {
    ctx->err = std::string("microbench: ") + cudaGetErrorString(err);
    return AMCLJ_ERR_CUDA;
  }
  *lookups_per_s = (double)blocks * MB_THREADS * iters * 8.0 / (best * 1e-3);
  return AMCLJ_OK;
}
