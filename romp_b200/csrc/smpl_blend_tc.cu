// SMPL shape + pose blend on the tensor cores (lbs, simple_romp/romp/smpl.py:151-170):
//     v_posed[n, (v,c)] = v_template[(v,c)] + sum_k [betas | R[1:]-I](n, k) * [shapedirs ; posedirs](k, (v,c)),   K = 10 + 207
// = the (B*6890) x 207 contraction the north star names, as ONE tcgen05 GEMM  [persons x K'] x [K' x 20670]  per batch.
//
// Precision: the result feeds a 1e-4 tolerance on vertices of magnitude ~1; a single fp16 (or bf16) product leaves
// 9.5e-5 (7e-4) max error on the synthetic pack, a two-term split 6e-5 - not safe.  So both operands are split
// x = hi + lo (fp16 each, 22 significant bits) and three products are accumulated in fp32:  hi*hi + lo*hi + hi*lo
// (measured 4e-7).  The split is folded into K (3 x 224 = 672 per output), but each operand is stored and moved once:
// B' = [B_hi | B_lo] (448 halfs per vertex coordinate, built once at smpl_create), A' = [f_hi | f_lo] (448 halfs per person,
// written by smpl_pose_kernel).  A streamed B_hi chunk is multiplied by the resident f_hi AND f_lo chunks while it sits
// in shared memory (4 MMAs), a B_lo chunk by f_hi (2 MMAs).
//
// Kernel (CTA pairs, cta_group::2, one UMMA = 256 persons x 256 coordinates x 16):
//   * the pair owns a 256-person tile: each CTA keeps ITS 128 persons' A' rows resident in shared memory (112 KB, 14 chunks of
//     32 halfs, SWIZZLE_64B) and walks the 81 coordinate tiles of 256 (a slice of them when there are fewer person tiles than SM pairs);
//   * B' is streamed: per (coordinate tile, chunk) each CTA TMA-loads its 128 of the 256 rows (8 KB) into a 6-stage ring; both
//     loads complete on the leader's barrier (conv_tc_2cta.cu protocol); 14 chunks and 42 MMAs per tile, N = 256 = full rate.
//     The ring (48 KB in flight per CTA) is what bounds the kernel: the first version streamed B_hi twice (21 chunks per tile,
//     61 GB/s per SM at full MMA rate against ~32 GB/s that 48 KB in flight sustain - ncu: tensor pipe 48 %);
//   * D: 2 accumulators x 256 fp32 columns in TMEM; each CTA's 8 epilogue warps drain every tile (warp = 32 persons x 128 columns) in
//     chunks of 32 columns, the TMEM load of the next chunk in flight: + v_template, staged in shared memory (128 B rows, SWIZZLE_128B), one TMA store per chunk into the
//     coordinate-tile-major v_posed buffer [81][capacity][256] that the skinning kernel reads (a tile = one contiguous block).
// Bytes: B' re-streamed per person pair: 18.6 MB x (N/256) pairs through L2 (4.8 GB at N = 65,536);
// HBM: v_posed written once (82.9 KB/person) and read once by the skinning kernel.
#include "conv_tc.cuh"
#include "tc_device.cuh"

namespace b200romp {

constexpr int kBlK = 448;                 // B': 2 x 224 (217 features zero-padded to 224): [B_hi | B_lo]
constexpr int kBlChunks = kBlK / 32;      // 14 chunks of 32 halfs = 64 B rows
constexpr int kBlAK = 448;                // A': [f_hi | f_lo]; B_hi chunk c (< 7) pairs with A' chunks c and 7 + c, B_lo chunk c (>= 7) with A' chunk c - 7
constexpr int kBlAChunks = kBlAK / 32;    // 14
constexpr int kBlCols = 20736;            // 20670 vertex coordinates padded to 81 x 256
constexpr int kBlColTiles = kBlCols / 256;
constexpr int kBlAChunk = 128 * 64;       // one resident A' chunk: 128 persons x 64 B
constexpr int kBlABytes = kBlAChunks * kBlAChunk;         // 114,688
constexpr int kBlBStage = 128 * 64;       // this CTA's 128 rows of one B' chunk
constexpr int kBlStages = 6;              // 8 KB each: ~1 us of TMA latency at 2 MMAs (256 clk) per chunk needs a deep ring
constexpr int kBlStgBytes = 32 * 128;     // epilogue staging: 32 persons x 32 fp32 columns per warp buffer
constexpr int kBlThreads = 320;           // warp 0 producer, warp 1 MMA, warps 2-9 epilogue: (person quarter, column half) of every tile
constexpr uint32_t kBlIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);   // F32 acc, F16 x F16, N 256, M 256

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct BlendMaps {
  CUtensorMap a, b, out;
};

__global__ void __launch_bounds__(kBlThreads, 1)
smpl_blend_tc_kernel(const __grid_constant__ BlendMaps maps, const float* __restrict__ v_template /*[20736], zero padded*/, int n_host,
                     const int* __restrict__ d_count, int col_splits, float* __restrict__ v_posed /*[81][cap][256]*/, int cap,
                     int debug /*B200ROMP_SMPL_DEBUG: 1 no stores, 2 no staging writes, 4 no MMAs, 8 st.global through a staging transpose instead of TMA stores*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                          // 14 x [128 persons x 64 B]
  uint8_t* sB = sA + kBlABytes;                                // kBlStages x [128 rows x 64 B]
  uint8_t* sStg = sB + kBlStages * kBlBStage;                  // 8 warps x 2 buffers x 4 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(sStg + 8 * 2 * kBlStgBytes);
  uint64_t* empty = full + kBlStages;
  uint64_t* a_full = empty + kBlStages;      // this CTA's A' tile landed
  uint64_t* a_peer = a_full + 1;             // (leader) the peer's A' tile landed
  uint64_t* a_free = a_peer + 1;             // MMAs of the current person tile retired: A' may be overwritten
  uint64_t* tmem_full = a_free + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBlStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(a_full, 1); mbar_init(a_peer, 1); mbar_init(a_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 16); }   // 8 epilogue warps of each CTA of the pair
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int N = d_count ? min(n_host, *d_count) : n_host;
  // work item = (256-person tile, slice of the 81 coordinate tiles): with few persons (cfg2: ~370 per batch) the coordinate
  // tiles are split over the pairs as well, so every SM pair has work; with many (cfg5) col_splits = 1
  const int person_tiles = ((N + 255) / 256) * col_splits;     // work items
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int cols_per = (kBlColTiles + col_splits - 1) / col_splits;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0, pt_i = 0;
      uint32_t phase = 0;
      for (int pt = pair; pt < person_tiles; pt += npairs, ++pt_i) {
        if (pt_i > 0) mbar_wait(a_free, (pt_i - 1) & 1);       // previous person tile's MMAs no longer read A'
        mbar_arrive_expect_tx(a_full, kBlABytes);
        const int row0 = (pt / col_splits) * 256 + (int)rank * 128;
        const int j0 = (pt % col_splits) * cols_per, j1 = min(kBlColTiles, j0 + cols_per);
        for (int c = 0; c < kBlAChunks; ++c) tma_load_2d(sA + c * kBlAChunk, &maps.a, a_full, c * 32, row0);
        for (int j = j0; j < j1; ++j)
          for (int c = 0; c < kBlChunks; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * kBlBStage);
            tma_load_2d_2cta(sB + stage * kBlBStage, &maps.b, &full[stage], c * 32, j * 256 + (int)rank * 128);
            if (++stage == kBlStages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (rank != 0) {
      // peer: report "A' tile resident" to the leader for every person tile
      if (elect_one()) {
        int pt_i = 0;
        for (int pt = pair; pt < person_tiles; pt += npairs, ++pt_i) {
          mbar_wait(a_full, pt_i & 1);
          mbar_arrive_cluster(a_peer, 0);
        }
      }
    } else if (elect_one()) {
      // ===================== MMA issuer (leader) =====================
      int stage = 0, it = 0, pt_i = 0;
      uint32_t phase = 0;
      const uint32_t a_base = smem_u32(sA);
      for (int pt = pair; pt < person_tiles; pt += npairs, ++pt_i) {
        mbar_wait(a_full, pt_i & 1);
        mbar_wait(a_peer, pt_i & 1);
        tc_fence_after();
        const int j0 = (pt % col_splits) * cols_per, j1 = min(kBlColTiles, j0 + cols_per);
        for (int j = j0; j < j1; ++j, ++it) {
          const int acc = it & 1;
          mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tile = tmem_base + (uint32_t)(acc * 256);
          for (int c = 0; c < kBlChunks; ++c) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sB + stage * kBlBStage);
            constexpr int kHalf = kBlChunks / 2;                   // 7 chunks of B_hi, then 7 of B_lo
            const int na = c < kHalf ? 2 : 1;                      // B_hi: x f_hi and x f_lo;  B_lo: x f_hi
            for (int a = 0; a < na; ++a) {
              const int ac = c < kHalf ? c + a * kHalf : c - kHalf;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint64_t adesc = make_smem_desc(a_base + ac * kBlAChunk + k * 32, 8 * 64, 4);
                const uint64_t bdesc = make_smem_desc(b_base + k * 32, 8 * 64, 4);
                if (!(debug & 4)) umma_f16_2cta(d_tile, adesc, bdesc, kBlIdesc, (c | a | k) ? 1u : 0u);
              }
            }
            umma_commit_2cta(&empty[stage]);
            if (++stage == kBlStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_2cta(&tmem_full[acc]);
        }
        umma_commit_2cta(a_free);              // (arrives in both CTAs once every MMA of this person tile has retired)
      }
    }
  } else {
    // ===================== epilogue: warp q drains persons [32q, 32q+32) of this CTA, 32 columns at a time =====================
    // All eight warps drain EVERY tile: warp (q, h) owns persons [32q, 32q + 32) x columns [128h, 128h + 128) = 4 chunks of 32
    // columns.  (First version: two groups of four warps alternating tiles, 8 chunks each - an accumulator was then busy
    // for drain (11 us) + MMA (2.8 us) in series, 6.9 us per tile, with each group idle a third of the time waiting for its
    // MMAs; ncu: 34 % of the samples on the tmem_full wait, 27 % on the first use of a tcgen05.ld result.)  The TMEM load of
    // chunk c + 1 is in flight while chunk c is converted, staged and stored.
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg0 = sStg + (warp - 2) * 2 * kBlStgBytes;
    int it = 0, buf = 0;
    for (int pt = pair; pt < person_tiles; pt += npairs) {
      const int prow = (pt / col_splits) * 256 + (int)rank * 128 + q * 32;
      const int j0 = (pt % col_splits) * cols_per, j1 = min(kBlColTiles, j0 + cols_per);
      for (int j = j0; j < j1; ++j, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* stg = stg0 + buf * kBlStgBytes;
          const int col = half * 128 + c * 32;
          const float4* vt4 = reinterpret_cast<const float4*>(v_template + j * 256 + col);
          float4 tv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) tv[i] = __ldg(vt4 + i);
          const bool tma_out = (debug & 8) == 0;       // default: TMA tensor stores (1.71 ms at cfg5); bit 3: plain st.global (1.83 ms)
          if (tma_out && lane == 0) bulk_wait_read(1);  // the store that last read this buffer (two chunks ago) is done
          __syncwarp();                               // (st.global path: every lane has read the buffer's previous contents)
          tmem_ld_wait();                             // chunk c has arrived
          if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
          else tc_fence_before();                     // accumulator fully read (handed back below)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t = tv[i];
            const uint32_t* rr = r[c & 1];
            const float4 v = make_float4(__uint_as_float(rr[4 * i + 0]) + t.x, __uint_as_float(rr[4 * i + 1]) + t.y,
                                         __uint_as_float(rr[4 * i + 2]) + t.z, __uint_as_float(rr[4 * i + 3]) + t.w);
            if (!(debug & 2)) *reinterpret_cast<float4*>(stg + lane * 128 + ((i ^ (lane & 7)) * 16)) = v;
          }
          if (tma_out) fence_proxy_async();
          __syncwarp();
          if (!tma_out && !(debug & 1)) {
            // transpose through the staging tile: thread = (row k*4 + lane/8, 16 B slot lane%8) -> every st.global.v4 of the warp
            // writes four complete 128 B lines.  Measured alternative to the TMA store (B200ROMP_SMPL_DEBUG=8): 1.83 vs 1.71 ms -
            // the ~1.5 us per chunk do not come from the store mechanism (tools/tma_store_bench.cu: both paths reach the same
            // 62 GB/s per SM, 5-6 TB/s on the full chip)
            const int slot = lane & 7;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int row = k * 4 + (lane >> 3);
              const float4 v = *reinterpret_cast<const float4*>(stg + row * 128 + ((slot ^ (row & 7)) * 16));
              if (prow + row < cap)
                *reinterpret_cast<float4*>(v_posed + ((size_t)j * cap + prow + row) * 256 + col + slot * 4) = v;
            }
          }
          if (lane == 0) {
            if (tma_out && !(debug & 1)) {
              tma_store_3d(&maps.out, stg, col, prow, j);   // v_posed [81 tiles][capacity][256]: persons beyond the capacity are clipped
              bulk_commit_group();
            }
            if (c == 3) mbar_arrive_cluster(&tmem_empty[acc], 0);
          }
          buf ^= 1;
        }
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int kBlendSmem = kBlABytes + kBlStages * kBlBStage + 8 * 2 * kBlStgBytes + 256 + 1024;

// a_rows: fp16 [cap] rows of 448 inside the per-person scratch records of smpl.cu (`row_stride_bytes` apart);  v_posed: fp32
// [81][capacity][256], coordinate-tile major
int smpl_blend_tc_launch(const void* a_rows, int row_stride_bytes, int capacity, const void* b_rows /*fp16 [20736][448]*/,
                         float* v_posed, const float* v_template_pad, int n, const int* d_count, int sm_count, cudaStream_t stream) {
  const int a_row_stride_bytes = row_stride_bytes;
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) { set_error("smpl_blend_tc: cuTensorMapEncodeTiled is unavailable"); return B200ROMP_ECUDA; }
  static bool attr_set = false;
  if (!attr_set) {
    B2R_CUDA_OK(cudaFuncSetAttribute(smpl_blend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlendSmem));
    attr_set = true;
  }
  BlendMaps m;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)kBlAK, (cuuint64_t)capacity};
    const cuuint64_t gstr[1] = {(cuuint64_t)a_row_stride_bytes};
    const cuuint32_t box[2] = {32, 128};
    if (encode(&m.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a_rows), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("smpl_blend_tc: tensor map (A') failed"); return B200ROMP_ECUDA;
    }
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)kBlK, (cuuint64_t)kBlCols};
    const cuuint64_t gstr[1] = {(cuuint64_t)kBlK * 2};
    const cuuint32_t box[2] = {32, 128};
    if (encode(&m.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(b_rows), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("smpl_blend_tc: tensor map (B') failed"); return B200ROMP_ECUDA;
    }
  }
  {
    const cuuint64_t gdim[3] = {256, (cuuint64_t)capacity, (cuuint64_t)kBlColTiles};
    const cuuint64_t gstr[2] = {256 * 4, (cuuint64_t)capacity * 256 * 4};
    const cuuint32_t box[3] = {32, 32, 1}, estr3[3] = {1, 1, 1};
    if (encode(&m.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v_posed, gdim, gstr, box, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("smpl_blend_tc: tensor map (v_posed) failed"); return B200ROMP_ECUDA;
    }
  }
  // `n` is the host-side upper bound of the person count (the device count may be smaller: surplus work items exit at once)
  const int ptiles = (n + 255) / 256, max_pairs = std::max(1, sm_count / 2);
  const int col_splits = std::max(1, std::min(kBlColTiles, max_pairs / std::max(1, ptiles)));
  const int pairs = std::max(1, std::min(max_pairs, ptiles * col_splits));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kBlThreads);
  cfg.dynamicSmemBytes = kBlendSmem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  static const int debug = [] { const char* e = getenv("B200ROMP_SMPL_DEBUG"); return e ? atoi(e) : 0; }();   // profiling experiments only
  B2R_CUDA_OK(cudaLaunchKernelEx(&cfg, smpl_blend_tc_kernel, m, v_template_pad, n, d_count, col_splits, v_posed, capacity, debug));
  return B200ROMP_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Linear-blend skinning on the tensor cores (lbs, simple_romp/romp/smpl.py:176-186):
//     T[n, v, e] = sum_j W[v, j] * A[n, j, e]   (e = 12 entries of the 3x4 transform),   vert = T [v_posed ; 1]
// As FFMA this is 300 FMA per (vertex, person) - 3.75 ms at the fp32 peak for cfg5, measured 12 ms.  As a GEMM it is tiny-K:
//     D[v (M = 128), (p, e) (N = 12 x 16 persons = 192)] = W'[v, k'] x A'[(p, e), k'],   k' = 96 = 3 x (24 joints padded to 32)
// with the same 3-term fp16 split as the blend: W' = [W_hi | W_lo | W_hi] (built at smpl_create), A' = [A_hi | A_hi | A_lo]
// (written per person by smpl_pose_kernel).  6 MMAs per 128 x 16 tile; the epilogue (thread = vertex) reads its 12 transform
// entries per person from TMEM, v_posed from the blend's output and writes the vertex: the kernel is bound by that HBM
// stream (82.9 KB read + 82.7 KB written per person).
constexpr int kSkP = 16;                   // persons per tile
constexpr int kSkN = 12 * kSkP;            // 192 accumulator columns
constexpr int kSkChunk = 128 * 64;         // resident W' chunk: 128 vertices x 64 B
constexpr int kSkBChunk = kSkN * 64;       // streamed A' chunk: 192 rows x 64 B
constexpr int kSkStage = 3 * kSkBChunk;    // 36,864
constexpr int kSkStages = 4;
constexpr int kSkThreads = 320;           // warp 0 producer, warp 1 MMA, warps 2-9 epilogue: two groups of four, one accumulator each
constexpr uint32_t kSkIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kSkN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tma_load_3d_plain(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct SkinMaps {
  CUtensorMap w, a;
};

__global__ void __launch_bounds__(kSkThreads, 1)
smpl_skin_tc_kernel(const __grid_constant__ SkinMaps maps, const float* __restrict__ v_posed /*[81][cap][256]*/, int cap, int n_host,
                    const int* __restrict__ d_count, float* __restrict__ verts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                  // 3 x [128 vertices x 64 B]
  uint8_t* sB = sW + 3 * kSkChunk;                     // kSkStages x 3 x [192 rows x 64 B]
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kSkStages * kSkStage);
  uint64_t* empty = full + kSkStages;
  uint64_t* w_full = empty + kSkStages;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSkStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int N = d_count ? min(n_host, *d_count) : n_host;
  const int ptiles = (N + kSkP - 1) / kSkP;
  const int vt = blockIdx.x;                           // vertex tile

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, 3 * kSkChunk);
      for (int c = 0; c < 3; ++c) tma_load_2d(sW + c * kSkChunk, &maps.w, w_full, c * 32, vt * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = blockIdx.y; pt < ptiles; pt += gridDim.y) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], kSkStage);
        for (int c = 0; c < 3; ++c) tma_load_3d_plain(sB + stage * kSkStage + c * kSkBChunk, &maps.a, &full[stage], c * 32, 0, pt * kSkP);
        if (++stage == kSkStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      int stage = 0, it = 0;
      uint32_t phase = 0;
      const uint32_t w_base = smem_u32(sW);
      for (int pt = blockIdx.y; pt < ptiles; pt += gridDim.y, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t b_base = smem_u32(sB + stage * kSkStage);
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * 256);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t adesc = make_smem_desc(w_base + c * kSkChunk + k * 32, 8 * 64, 4);
            const uint64_t bdesc = make_smem_desc(b_base + c * kSkBChunk + k * 32, 8 * 64, 4);
            umma_f16(d_tile, adesc, bdesc, kSkIdesc, (c | k) ? 1u : 0u);
          }
        umma_commit(&empty[stage]);
        umma_commit(&tmem_full[acc]);
        if (++stage == kSkStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: thread = vertex, loops over the tile's 16 persons =====================
    // two groups of four warps alternate tiles (group g owns accumulator g): while one group waits for / drains its
    // accumulator, the other group's 48 v_posed loads per thread are in flight
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;
    const int v = vt * 128 + q * 32 + lane;
    const bool v_ok = v < 6890;
    int it = group;
    // v_posed of this vertex for all 16 persons of a tile: 48 independent loads.  They do not depend on the MMAs, so they are
    // issued one tile AHEAD (double-buffered in registers): while this group drains tile t its loads of tile t + 2 are in
    // flight, and the other group's likewise - the ~1 us HBM latency is off the critical path (first version, loads per
    // person: 17.8 ms; loads per tile before the accumulator wait: 2.64 ms at 4.2 TB/s, latency-bound)
    float vpx[kSkP], vpy[kSkP], vpz[kSkP], nx[kSkP], ny[kSkP], nz[kSkP];
    // the three coordinates of this vertex inside v_posed [81][cap][256]: (tile, column) of coordinate 3v + c
    const int cc = 3 * (v_ok ? v : 0);
    const size_t o0 = (size_t)(cc >> 8) * cap * 256 + (cc & 255), o1 = (size_t)((cc + 1) >> 8) * cap * 256 + ((cc + 1) & 255),
                 o2 = (size_t)((cc + 2) >> 8) * cap * 256 + ((cc + 2) & 255);
    auto load_tile = [&](int pt, float (&x)[kSkP], float (&y)[kSkP], float (&z)[kSkP]) {
      const int np = min(kSkP, N - pt * kSkP);
#pragma unroll
      for (int p = 0; p < kSkP; ++p) {
        const bool ok = v_ok && p < np;
        const float* vp = v_posed + (size_t)(pt * kSkP + (ok ? p : 0)) * 256;
        x[p] = ok ? __ldg(vp + o0) : 0.f; y[p] = ok ? __ldg(vp + o1) : 0.f; z[p] = ok ? __ldg(vp + o2) : 0.f;
      }
    };
    const int pt0 = blockIdx.y + group * gridDim.y;
    if (pt0 < ptiles) load_tile(pt0, nx, ny, nz);
    for (int pt = pt0; pt < ptiles; pt += 2 * gridDim.y, it += 2) {
      const int acc = it & 1;
      const int np = min(kSkP, N - pt * kSkP);
#pragma unroll
      for (int p = 0; p < kSkP; ++p) { vpx[p] = nx[p]; vpy[p] = ny[p]; vpz[p] = nz[p]; }
      if (pt + 2 * (int)gridDim.y < ptiles) load_tile(pt + 2 * gridDim.y, nx, ny, nz);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll
      for (int p = 0; p < kSkP; ++p) {
        uint32_t r[16];
        tmem_ld16(taddr + p * 12, r);                 // 12 transform entries (+4 columns of the next person, ignored)
        tmem_ld_wait();
        if (p + 1 == kSkP) {                          // accumulator fully read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (p < np && v_ok) {
          float* o = verts + ((size_t)(pt * kSkP + p) * 6890 + v) * 3;
          const float px = vpx[p], py = vpy[p], pz = vpz[p];
          o[0] = __uint_as_float(r[0]) * px + __uint_as_float(r[1]) * py + __uint_as_float(r[2]) * pz + __uint_as_float(r[3]);
          o[1] = __uint_as_float(r[4]) * px + __uint_as_float(r[5]) * py + __uint_as_float(r[6]) * pz + __uint_as_float(r[7]);
          o[2] = __uint_as_float(r[8]) * px + __uint_as_float(r[9]) * py + __uint_as_float(r[10]) * pz + __uint_as_float(r[11]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr int kSkinSmem = 3 * kSkChunk + kSkStages * kSkStage + 256 + 1024;

// w_rows: fp16 [6912][96] = [W_hi | W_lo | W_hi] per vertex (24 joints padded to 32);  a' rows: fp16, 12 rows of 96 per person at
// byte offset a_off inside the per-person scratch record (record stride row_stride_bytes);  v_posed: [81][capacity][256] floats (smpl.cu)
int smpl_skin_tc_launch(const void* w_rows, const void* ws_base, int a_off_bytes, const float* v_posed, int row_stride_bytes, int capacity,
                        int n, const int* d_count, int sm_count, float* verts, cudaStream_t stream) {
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) { set_error("smpl_skin_tc: cuTensorMapEncodeTiled is unavailable"); return B200ROMP_ECUDA; }
  static bool attr_set = false;
  if (!attr_set) {
    B2R_CUDA_OK(cudaFuncSetAttribute(smpl_skin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkinSmem));
    attr_set = true;
  }
  SkinMaps m;
  {
    const cuuint64_t gdim[2] = {96, 6912};
    const cuuint64_t gstr[1] = {96 * 2};
    const cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    if (encode(&m.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_rows), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("smpl_skin_tc: tensor map (W') failed"); return B200ROMP_ECUDA;
    }
  }
  {
    const cuuint64_t gdim[3] = {96, 12, (cuuint64_t)capacity};
    const cuuint64_t gstr[2] = {96 * 2, (cuuint64_t)row_stride_bytes};
    const cuuint32_t box[3] = {32, 12, (cuuint32_t)kSkP}, estr[3] = {1, 1, 1};
    void* base = const_cast<char*>(static_cast<const char*>(ws_base) + a_off_bytes);
    if (encode(&m.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("smpl_skin_tc: tensor map (A') failed"); return B200ROMP_ECUDA;
    }
  }
  const int ptiles = (n + kSkP - 1) / kSkP;
  const int gy = std::max(1, std::min(ptiles, (2 * sm_count + 53) / 54));
  dim3 grid(54, gy);
  smpl_skin_tc_kernel<<<grid, kSkThreads, kSkinSmem, stream>>>(m, v_posed, capacity, n, d_count, verts);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

}  // namespace b200romp
