// Microbenchmark: issue-rate / latency of tcgen05.mma (kind::f16, M=128, cta_group::1, SS operands) on B200 as a
// function of N, of the accumulator dependency pattern and of the A-operand addressing pattern used by conv_tc.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
// Prints clocks per MMA for each variant (one CTA per SM, all SMs busy, smem contents irrelevant).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: same A/B every MMA, one accumulator (dependent chain)
// mode 1: same A/B, 8 accumulators round robin (independent)
// mode 2: conv-like: taps = shifted A rows (row pitch 128 B, SBO = 10 rows), 4 k-steps inside a tap, one accumulator
// mode 3: GEMM-like: A advances 32 B per MMA inside 128 B rows (4 k-steps) then next 16 KB tile, one accumulator
// mode 4: like 2 but k outer / tap inner
template <int N>
__global__ void __launch_bounds__(128, 1) bench(int mode, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < 36; ++j) {
        uint32_t a = a0, b = b0, sbo = 1024, d = tm;
        if (mode == 1) d = tm + (j & 7) * (N <= 64 ? N : 0);
        if (mode == 2) { const int t = j / 4, k = j % 4; a = a0 + ((t / 3) * 10 + t % 3) * 128 + k * 32; b = b0 + k * 32; sbo = 1280; }
        if (mode == 3) { a = a0 + (j % 4) * 32 + ((j / 4) % 2) * 16384; b = b0 + (j % 4) * 32; }
        if (mode == 4) { const int k = j / 9, t = j % 9; a = a0 + ((t / 3) * 10 + t % 3) * 128 + k * 32; b = b0 + k * 32; sbo = 1280; }
        umma(d, make_desc(a, sbo, 2), make_desc(b, 1024, 2), idesc, (it | j) ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int N>
void run(int mode, const char* name) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 200, smem = 100 * 1024;
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N><<<148, 128, smem>>>(mode, 10, d);
  bench<N><<<148, 128, smem>>>(mode, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d %-34s %7.1f clk/MMA (ideal %3d)  %s\n", N, name, (double)h / (iters * 36.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const char* names[5] = {"same operands, 1 accumulator", "same operands, 8 accumulators", "conv taps (shifted rows), k inner",
                          "gemm k-advance, 1 accumulator", "conv taps, k outer"};
  for (int mode = 0; mode < 5; ++mode) {
    run<32>(mode, names[mode]);
    run<64>(mode, names[mode]);
    run<128>(mode, names[mode]);
    run<256>(mode, names[mode]);
  }
  return 0;
}
