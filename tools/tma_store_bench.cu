// Microbenchmark: how fast can one SM push data to HBM through cp.async.bulk.tensor stores?
// Every CTA (one per SM) has W warps; each warp owns two staging buffers and loops: wait until the store that last read the
// buffer is done -> (optional) rewrite the buffer with st.shared -> fence.proxy.async -> one TMA tensor store -> commit.
// Variants: box rows x bytes, row pitch of the global tensor (contiguous tile vs scattered rows), warps per CTA, and a plain
// st.global.v4 loop for comparison.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_store_bench tools/tma_store_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tensor: [rows_total][cols] fp32, row pitch `pitch` bytes.  A warp's k-th store covers rows [r0, r0 + BR) x cols [c0, c0 + BC).
template <int BR, int BC /*floats*/>
__global__ void __launch_bounds__(512) store_bench(const __grid_constant__ CUtensorMap map, int warps, int iters, int rewrite, int col_tiles,
                                                   long long rows_total) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  constexpr int BYTES = BR * BC * 4;
  uint8_t* stg = smem + warp * 2 * BYTES;
  for (int i = lane; i < 2 * BYTES / 16; i += 32) ((uint4*)stg)[i] = make_uint4(i, warp, blockIdx.x, 7);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  // this warp's row band: global warp id w owns rows [w * BR * iters_rows ...): walk column tiles first (contiguous in a row)
  const long long gw = (long long)blockIdx.x * warps + warp;
  int buf = 0;
  for (int it = 0; it < iters; ++it) {
    const long long tile = gw * iters + it;
    const int ct = (int)(tile % col_tiles);
    const long long rt = tile / col_tiles;
    const long long r0 = (rt * BR) % rows_total;
    uint8_t* s = stg + buf * BYTES;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    if (rewrite) {
      for (int i = lane; i < BYTES / 16; i += 32) ((uint4*)s)[i] = make_uint4(it, i, warp, 3);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    }
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map), "r"(smem_u32(s)), "r"(ct * BC),
                   "r"((int)r0) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512) stg_bench(float4* out, int warps, int iters, long long n4) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const long long gw = (long long)blockIdx.x * warps + warp;
  for (int it = 0; it < iters; ++it) {
    // each warp writes 4 KB contiguous per trip: 8 x (32 lanes x 16 B)
    const long long base = ((gw * iters + it) * 256) % n4;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[base + k * 32 + lane] = make_float4(it, k, lane, 1.f);
  }
}

int main(int argc, char** argv) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_encodeTiled encode = (PFN_encodeTiled)fp;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  if (argc > 1) sms = atoi(argv[1]);   // number of CTAs (one per SM): few CTAs take HBM out of the picture -> per-SM engine limits
  const size_t bytes = (size_t)6 << 30;
  float* buf = nullptr;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, int br, int bc, long long cols, long long pitch_bytes, int warps, int rewrite, CUtensorMapSwizzle sw) {
    const long long rows_total = (long long)(bytes / pitch_bytes) / br * br;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows_total};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br}, es[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("%-44s encode failed %d\n", name, (int)cr); return; }
    const int iters = 2000;
    const int smem = warps * 2 * br * bc * 4 + 1024;
    const int col_tiles = (int)(cols / bc);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (br == 32 && bc == 32) { cudaFuncSetAttribute(store_bench<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); store_bench<32, 32><<<sms, 512, smem>>>(map, warps, iters, rewrite, col_tiles, rows_total); }
      else if (br == 64 && bc == 32) { cudaFuncSetAttribute(store_bench<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); store_bench<64, 32><<<sms, 512, smem>>>(map, warps, iters, rewrite, col_tiles, rows_total); }
      else if (br == 32 && bc == 64) { cudaFuncSetAttribute(store_bench<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); store_bench<32, 64><<<sms, 512, smem>>>(map, warps, iters, rewrite, col_tiles, rows_total); }
      else if (br == 128 && bc == 32) { cudaFuncSetAttribute(store_bench<128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); store_bench<128, 32><<<sms, 512, smem>>>(map, warps, iters, rewrite, col_tiles, rows_total); }
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("%-44s error %s\n", name, cudaGetErrorString(e)); return; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double total = (double)sms * warps * iters * br * bc * 4;
    printf("%-58s %7.0f GB/s  %6.1f GB/s per SM  (%d warps, rewrite %d)\n", name, total / best / 1e6, total / best / 1e6 / sms, warps, rewrite);
  };
  printf("SMs %d\n", sms);
  for (int warps : {4, 8}) {
    run("box 32 rows x 128 B, contiguous tile (pitch 1 KB)", 32, 32, 256, 1024, warps, 0, CU_TENSOR_MAP_SWIZZLE_128B);
    run("box 32 rows x 128 B, contiguous tile (pitch 1 KB)", 32, 32, 256, 1024, warps, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    run("box 32 rows x 128 B, scattered rows (pitch 88 KB)", 32, 32, 20736, 88 * 1024 + 512, warps, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    run("box 64 rows x 128 B, contiguous tile (pitch 1 KB)", 64, 32, 256, 1024, warps, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    run("box 128 rows x 128 B, pitch 128 B (fully contiguous)", 128, 32, 32, 128, warps, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    run("box 32 rows x 256 B (no swizzle), pitch 1 KB", 32, 64, 256, 1024, warps, 1, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      stg_bench<<<sms, 512>>>((float4*)buf, warps, iters, (long long)(bytes / 16) - 4096);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double total = (double)sms * warps * iters * 4096;
    printf("%-58s %7.0f GB/s  %6.1f GB/s per SM  (%d warps)\n", "st.global.v4, 4 KB contiguous per warp trip", total / best / 1e6, total / best / 1e6 / sms, warps);
  }
  return 0;
}
