// Microbenchmark 2: can the ~105-clock fixed cost of a small-N tcgen05.mma be overlapped?
//   (a) 1 vs 2 CTAs resident per SM, each issuing its own MMA stream;
//   (b) 1 vs 2 issuing threads (different warps) inside one CTA, each with its own accumulator;
//   (c) M = 64 instead of 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench2 tools/mma_bench2.cu && ./mma_bench2
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

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int M, int N, int ELECT>
__global__ void __launch_bounds__(128) bench(int issuers, int iters, int tmem_cols, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 40 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (ELECT == 2 && warp < issuers) {
    // warp-converged issue loop: every lane computes the (uniform) descriptors, one elected lane issues
    const uint32_t a0 = smem_u32(smem) + warp * 64, b0 = smem_u32(smem + 20 * 1024);
    const uint32_t d = tm + warp * N;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 36; ++j) {
        const int t = j / 4, k = j % 4;
        const uint64_t ad = make_desc(a0 + ((t / 3) * 10 + t % 3) * 128 + (k & 1) * 32, 1280, 2), bd = make_desc(b0 + k * 32, 1024, 2);
        if (elect_one()) umma(d, ad, bd, idesc, (it | j) ? 1u : 0u);
      }
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0 && lane == 0) *out = t1 - t0;
  } else if (ELECT != 2 && warp < issuers && (ELECT ? elect_one() : lane == 0)) {
    const uint32_t a0 = smem_u32(smem) + warp * 64, b0 = smem_u32(smem + 20 * 1024);
    const uint32_t d = tm + warp * N;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int j = 0; j < 36; ++j) {
        const int t = j / 4, k = j % 4;
        umma(d, make_desc(a0 + ((t / 3) * 10 + t % 3) * 128 + (k & 1) * 32, 1280, 2), make_desc(b0 + k * 32, 1024, 2), idesc, (it | j) ? 1u : 0u);
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(tmem_cols) : "memory");
}

template <int M, int N, int ELECT = 0>
void run(int ctas_per_sm, int issuers) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 200, smem = 64 * 1024;     // 64 KB + alignment: three CTAs fit per SM, we launch 1 or 2 per SM
  const int tmem_cols = 256;
  cudaFuncSetAttribute(bench<M, N, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<M, N, ELECT><<<148 * ctas_per_sm, 128, smem>>>(issuers, 10, tmem_cols, d);
  bench<M, N, ELECT><<<148 * ctas_per_sm, 128, smem>>>(issuers, iters, tmem_cols, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double clk = (double)h / (iters * 36.0);
  printf("%s M=%3d N=%3d  CTAs/SM=%d issuers/CTA=%d : %7.1f clk per MMA per issuer -> %6.0f MAC/clk/SM (peak 4096)  %s\n", ELECT == 2 ? "warpcv" : (ELECT ? "elect " : "lane0 "), M, N, ctas_per_sm,
         issuers, clk, (double)M * N * 16 * ctas_per_sm * issuers / clk, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<128, 32, 2>(1, 1); run<128, 64, 2>(1, 1); run<128, 64, 2>(1, 2); run<128, 128, 2>(1, 1); run<128, 32, 2>(1, 2);
  run<128, 32, 1>(1, 1); run<128, 64, 1>(1, 1);
  run<128, 32>(1, 1); run<128, 32>(2, 1); run<128, 32>(1, 2); run<128, 32>(2, 2);
  run<128, 64>(1, 1); run<128, 64>(2, 1); run<128, 64>(1, 2); run<128, 64>(2, 2);
  run<128, 128>(1, 1); run<128, 128>(2, 1); run<128, 128>(1, 2);
  run<64, 64>(1, 1); run<64, 64>(2, 1); run<64, 128>(1, 1); run<64, 256>(1, 1); run<64, 256>(2, 1);
  run<128, 256>(1, 1); run<128, 256>(2, 1);
  return 0;
}
