// pipe_rates.cu — measured issue rates of the instruction classes the look-up kernel is made of, on the GPU
// this runs on (B200: sm_100a).  One CTA of 24 warps per SM (the look-up kernel's shape), every thread runs
// ILP independent chains of one instruction class; prints warp-instructions per clock per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates tools/pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096, THREADS = 768;

#define CHAIN_KERNEL(NAME, DECL, BODY)                                                         \
  __global__ void __launch_bounds__(THREADS, 1) NAME(uint32_t *out, uint32_t seed) {           \
    DECL;                                                                                      \
    for (int it = 0; it < ITERS; it++) {                                                       \
      _Pragma("unroll") for (int u = 0; u < ILP; u++) { BODY; }                                \
    }                                                                                          \
    uint32_t acc = 0;                                                                          \
    _Pragma("unroll") for (int u = 0; u < ILP; u++) acc ^= sink(v[u]);                         \
    if (acc == 0x12345678u) out[threadIdx.x] = acc;                                            \
  }

__device__ __forceinline__ uint32_t sink(uint32_t x) { return x; }
__device__ __forceinline__ uint32_t sink(float x) { return __float_as_uint(x); }
__device__ __forceinline__ uint32_t sink(unsigned long long x) { return (uint32_t)x ^ (uint32_t)(x >> 32); }

#define DECL_F float v[ILP]; float a = __uint_as_float(seed | 0x3f800000u), b = a * 0.5f; \
  _Pragma("unroll") for (int u = 0; u < ILP; u++) v[u] = a + (float)(threadIdx.x + u)
#define DECL_I uint32_t v[ILP]; uint32_t a = seed | 1u, b = seed * 3u + 7u; \
  _Pragma("unroll") for (int u = 0; u < ILP; u++) v[u] = a + threadIdx.x * 17u + u
#define DECL_P unsigned long long v[ILP]; unsigned long long a = ((unsigned long long)(seed | 0x3f800000u) << 32) | (seed | 0x3f800000u), b = a; \
  _Pragma("unroll") for (int u = 0; u < ILP; u++) v[u] = a + (unsigned long long)(threadIdx.x + u)

CHAIN_KERNEL(k_ffma, DECL_F, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[u]) : "f"(a), "f"(b)))
CHAIN_KERNEL(k_fmul, DECL_F, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[u]) : "f"(a)))
CHAIN_KERNEL(k_fadd, DECL_F, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[u]) : "f"(a)))
CHAIN_KERNEL(k_ffma2, DECL_P, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[u]) : "l"(a), "l"(b)))
CHAIN_KERNEL(k_fmul2, DECL_P, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[u]) : "l"(a)))
CHAIN_KERNEL(k_fadd2, DECL_P, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[u]) : "l"(a)))
CHAIN_KERNEL(k_iadd, DECL_I, asm volatile("add.u32 %0, %0, %1;" : "+r"(v[u]) : "r"(a)))
CHAIN_KERNEL(k_lop3, DECL_I, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_shf, DECL_I, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_shr, DECL_I, asm volatile("shr.s32 %0, %0, %1;" : "+r"(v[u]) : "r"(b & 3u)))
CHAIN_KERNEL(k_imad, DECL_I, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_imin, DECL_I, asm volatile("min.u32 %0, %0, %1;" : "+r"(v[u]) : "r"(a)))
CHAIN_KERNEL(k_prmt, DECL_I, asm volatile("prmt.b32 %0, %0, %1, 0x4442;" : "+r"(v[u]) : "r"(a)))
CHAIN_KERNEL(k_abs, DECL_I, asm volatile("abs.s32 %0, %0;" : "+r"(v[u])))
CHAIN_KERNEL(k_setp_sel, DECL_I, asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_fsetp_sel, DECL_F, asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(v[u]) : "f"(a), "f"(b)))
CHAIN_KERNEL(k_ex2, DECL_F, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[u])))
// mixtures: one fma-class and one alu-class instruction per chain step
CHAIN_KERNEL(k_ffma_lop3, DECL_I, asm volatile("{.reg .f32 t; mov.b32 t, %0; fma.rn.f32 t, t, t, t; mov.b32 %0, t; lop3.b32 %0, %0, %1, %2, 0x96;}" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_imad_lop3, DECL_I, asm volatile("mad.lo.u32 %0, %0, %1, %2; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_iadd_lop3, DECL_I, asm volatile("add.u32 %0, %0, %1; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[u]) : "r"(a), "r"(b)))
CHAIN_KERNEL(k_ffma2_lop3, DECL_P, asm volatile("{.reg .b32 lo, hi; fma.rn.f32x2 %0, %0, %1, %2; mov.b64 {lo, hi}, %0; lop3.b32 lo, lo, hi, hi, 0x96; mov.b64 %0, {lo, hi};}" : "+l"(v[u]) : "l"(a), "l"(b)))
CHAIN_KERNEL(k_ex2_ffma, DECL_F, asm volatile("ex2.approx.ftz.f32 %0, %0; fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[u]) : "f"(a), "f"(b)))

template <typename K>
static void run(const char *name, K kern, int per_step, int sms, double clk_hz) {
  uint32_t *out;
  cudaMalloc(&out, THREADS * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    kern<<<sms, THREADS>>>(out, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double warp_inst = (double)sms * (THREADS / 32) * ITERS * ILP * per_step;
  const double per_clk_smsp = warp_inst / (best * 1e-3) / clk_hz / (sms * 4.0);
  printf("%-14s %8.3f ms  %6.3f warp-inst/clk/SMSP (%d per chain step)\n", name, best, per_clk_smsp, per_step);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = clk_khz * 1e3;
  printf("%s, %d SMs, %.0f MHz (attribute; rates assume the SM runs at it)\n", p.name, p.multiProcessorCount, clk / 1e6);
  const int sms = p.multiProcessorCount;
  run("FFMA", k_ffma, 1, sms, clk);
  run("FMUL", k_fmul, 1, sms, clk);
  run("FADD", k_fadd, 1, sms, clk);
  run("FFMA2", k_ffma2, 1, sms, clk);
  run("FMUL2", k_fmul2, 1, sms, clk);
  run("FADD2", k_fadd2, 1, sms, clk);
  run("IADD", k_iadd, 1, sms, clk);
  run("LOP3", k_lop3, 1, sms, clk);
  run("SHF", k_shf, 1, sms, clk);
  run("SHR", k_shr, 1, sms, clk);
  run("IMAD", k_imad, 1, sms, clk);
  run("IMIN", k_imin, 1, sms, clk);
  run("PRMT", k_prmt, 1, sms, clk);
  run("IABS", k_abs, 1, sms, clk);
  run("ISETP+SEL", k_setp_sel, 2, sms, clk);
  run("FSETP+FSEL", k_fsetp_sel, 2, sms, clk);
  run("MUFU.EX2", k_ex2, 1, sms, clk);
  run("FFMA+LOP3", k_ffma_lop3, 2, sms, clk);
  run("IMAD+LOP3", k_imad_lop3, 2, sms, clk);
  run("IADD+LOP3", k_iadd_lop3, 2, sms, clk);
  run("FFMA2+LOP3", k_ffma2_lop3, 2, sms, clk);
  run("EX2+FFMA", k_ex2_ffma, 2, sms, clk);
  return 0;
}
