// Stem conv on tcgen05: backbone.conv1 = Conv2d(3, 64, 3, stride 2, pad 1) + BN + ReLU on raw uint8 frames with the
// input normalisation x/255*2-1 folded in (simple_romp/romp/model.py:384-387).
//
// GEMM view: M = 128 output pixels (one 16x8 tile), N = 64, K = 27 taps*channels padded to 32 (two UMMA_K=16 steps).
// The A tile cannot come straight from TMA (3-channel u8, stride 2).  A 3-D tensor map over the frames viewed as
// [N][H][W*3 bytes] brings the raw 33-row x 80-byte window of a tile into shared memory (zero fill outside the image,
// two windows in flight per producer warp = 8 per SM, so DRAM latency is hidden); the producer threads then convert:
// thread = output pixel reads its 27 bytes from the window and writes one 64-byte K-major row in the SWIZZLE_64B pattern
// the MMA descriptor expects.  (A first version gathered the 27 bytes from global memory: 382 us, latency-bound.)
// Exactness: the operand is stored as (x - 127.5) in bf16 - exact for every integer 0..255 (8 significant bits) - and the
// weights carry the factor 2/255, so  sum w*(2/255)*(x-127.5) = sum w*(x/255*2-1)  and zero padding stays zero; the only
// rounding is the bf16 rounding of the scaled weights (same as every other layer on this engine).
// Everything after the MMA is the shared TMA epilogue (tc_device.cuh): bias + ReLU + bf16, staged, tensor store.
// Warp roles (416 threads): warps 0, 1, 11, 12 producers - each builds whole tiles in its private stage, so four gathers
// (one DRAM latency each) are in flight per SM; warp 2 TMEM allocator + MMA issuer; warps 3-10 the shared epilogue.
#include "conv_tc.cuh"
#include "tc_device.cuh"

namespace b200romp {

namespace {
constexpr int kStemStages = 4;                   // = producer warps, stage w is private to producer w
constexpr int kStemThreads = kTcThreads + 64;    // two extra producer warps after the epilogue warps
constexpr int kStemK = 32;                       // 27 padded to two UMMA_K steps
constexpr int kStemRowB = kStemK * 2;            // 64 B rows -> SWIZZLE_64B
constexpr int kStemABytes = 128 * kStemRowB;     // 8 KB per stage
constexpr int kStemNT = 64;
constexpr int kStemBBytes = kStemNT * kStemRowB; // 4 KB weight image
constexpr int kStemAcc = AccCfg<1>::ACC;
constexpr int kStemEpi = tc_epi_with_nbuf(kTmaEpiOut, 2);   // TMA epilogue, two staging tiles per warp
constexpr int kRawRowB = 80;                     // bytes [6*x0 - 16, 6*x0 + 64) of each input row: 16 B aligned, covers ix = 2*x0-1 .. 2*x0+15
constexpr int kRawRows = 33;                     // iy = 2*y0 - 1 .. 2*y0 + 31
constexpr int kRawBytes = kRawRows * kRawRowB;   // 2640 B per window (TMA box)
constexpr int kRawStage = 2688;                  // padded to 128 B
constexpr int kRawDepth = 2;                     // windows in flight per producer warp
constexpr int kStemRawStride = kRawStage;
constexpr uint32_t kStemIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kStemNT >> 3) << 17) | ((128u >> 4) << 24);
}  // namespace

__global__ void __launch_bounds__(kStemThreads, 1)
conv_stem_tc_kernel(const __grid_constant__ CUtensorMap raw_map, const __grid_constant__ TcEpiMaps epi_maps, const ConvParams p,
                    const uint8_t* __restrict__ wpack, int tiles_x, int tiles_y, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + kStemBBytes;
  uint8_t* epi_smem = sA + kStemStages * kStemABytes;
  uint8_t* raw_smem = epi_smem + tc_epi_total_bytes(kStemEpi, kStemNT);
  uint64_t* full = reinterpret_cast<uint64_t*>(raw_smem + kStemStages * kRawDepth * kRawStage);
  uint64_t* empty = full + kStemStages;
  uint64_t* b_full = empty + kStemStages;
  uint64_t* tmem_full = b_full + 1;
  uint64_t* tmem_empty = tmem_full + kStemAcc;
  uint64_t* res_bar = tmem_empty + kStemAcc;
  uint64_t* raw_full = res_bar + 3 * kEpiWarps;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(raw_full + kStemStages * kRawDepth);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 2) + 15) & ~(uintptr_t)15);   // 16 B: ld.shared.v4

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStemStages; ++i) {
      mbar_init(&full[i], 1);                    // its producer warp
      mbar_init(&empty[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < kStemAcc; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < kStemStages * kRawDepth; ++i) mbar_init(&raw_full[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= kFirstEpiWarp * 32 && threadIdx.x < kFirstEpiWarp * 32 + kStemNT)
    s_bias[threadIdx.x - kFirstEpiWarp * 32] = p.bias[threadIdx.x - kFirstEpiWarp * 32];
  if (warp == 2) tmem_alloc(tmem_ptr, tc_tmem_cols(kStemAcc * kStemNT));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_frame = tiles_x * tiles_y;
  pdl_trigger();

  if (warp < 2 || warp >= kFirstEpiWarp + kEpiWarps) {
    // ===================== im2col producers =====================
    const int pw = warp < 2 ? warp : warp - (kFirstEpiWarp + kEpiWarps) + 2;   // 0..3 = private stage
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(b_full, kStemBBytes);
      bulk_copy_g2s(sB, wpack, kStemBBytes, b_full);
    }
    pdl_wait();                                  // frames may be produced by a predecessor kernel / copy
    const uint64_t pol = l2_policy_stream(p.debug);
    uint8_t* a = sA + pw * kStemABytes;
    uint8_t* raw0 = raw_smem + pw * kRawDepth * kStemRawStride;
    uint64_t* rbar = raw_full + pw * kRawDepth;
    auto load_window = [&](int tile, int slot) {       // lane 0: raw input window of `tile` -> raw slot
      const int n = tile / per_frame, rem = tile % per_frame;
      const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
      mbar_arrive_expect_tx(&rbar[slot], kRawBytes);
      tma_load_3d(raw0 + slot * kStemRawStride, &raw_map, &rbar[slot], 6 * x0 - 16, 2 * y0 - 1, n, pol);
    };
    const int tstep = kStemStages * gridDim.x;
    const int first = blockIdx.x + pw * gridDim.x;
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < kRawDepth; ++d)
        if (first + d * tstep < num_tiles) load_window(first + d * tstep, d);
    }
    uint32_t phase = 0;
    int t_local = 0;
    for (int tile = first; tile < num_tiles; tile += tstep, ++t_local) {
      const int rem = tile % per_frame;
      const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
      const int slot = t_local % kRawDepth;
      const uint8_t* raw = raw0 + slot * kStemRawStride;
      mbar_wait(&rbar[slot], (uint32_t)(t_local / kRawDepth) & 1u);
      mbar_wait(&empty[pw], phase ^ 1);
#pragma unroll 2
      for (int j = 0; j < 4; ++j) {
        const int m = j * 32 + lane;                       // A row = TMEM lane = pixel (m >> 3, m & 7) of the tile
        const int ty = m >> 3, px = m & 7;
        const int oy = y0 + ty, ox = x0 + px;
        uint32_t w[16];                                    // 32 bf16, k = r*9 + s*3 + c
        float v[32];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = 2 * oy - 1 + r;
          const bool row_ok = iy >= 0 && iy < p.Hin;
          const uint8_t* src = raw + (2 * ty + r) * kRawRowB + 13 + 6 * px;    // byte of (iy, ix = 2*ox-1, c = 0)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const int ix = 2 * ox - 1 + s2;
            const bool ok = row_ok && ix >= 0 && ix < p.Win;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[r * 9 + s2 * 3 + c] = ok ? (float)src[s2 * 3 + c] - 127.5f : 0.f;
          }
        }
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
          w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)                     // 16 B chunk q4 of row m, Swizzle<2,4,3>
          *reinterpret_cast<uint4*>(a + m * kStemRowB + ((q4 ^ ((m >> 1) & 3)) * 16)) = make_uint4(w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
      }
      fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full[pw]);
        if (tile + kRawDepth * tstep < num_tiles) load_window(tile + kRawDepth * tstep, slot);   // window slot is free again
      }
      phase ^= 1;
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(b_full, 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(sB);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & (kStemAcc - 1);
        mbar_wait(&tmem_empty[acc], ((it / kStemAcc) & 1) ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * kStemABytes);
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * kStemNT);
#pragma unroll
        for (int k = 0; k < kStemK / 16; ++k)
          umma_bf16(d_tile, make_smem_desc(a_base + k * 32, 8 * kStemRowB, 4), make_smem_desc(b_base + k * 32, 8 * kStemRowB, 4),
                    kStemIdesc, k ? 1u : 0u);
        umma_commit(&empty[stage]);
        umma_commit(&tmem_full[acc]);
        if (++stage == kStemStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    tc_epilogue_loop_tma<kStemNT>(p, epi_maps, kStemEpi, epi_smem, res_bar, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x,
                                  per_frame, num_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, tc_tmem_cols(kStemAcc * kStemNT));
  }
}

// ------------------------------------------------------------------------------------------------
bool tc_stem_supported(const ConvParams& p, int ksize, int stride) {
  return ksize == 3 && stride == 2 && p.cin == 3 && p.in_C == 3 && p.in_c_off == 0 && p.cout == 64 && p.in_dtype == B200ROMP_U8 &&
         p.input_norm && p.out_dtype == B200ROMP_BF16 && !p.out_nchw && p.up == 1 && p.res == nullptr && p.relu &&
         p.pow_channel < 0 && p.out_C == 64 && p.out_c_off == 0 && p.Hout % 16 == 0 && p.Wout % 8 == 0 &&
         p.Hout * 2 == p.Hin && p.Wout * 2 == p.Win && (p.Win * 3) % 16 == 0;
}

int tc_stem_prepare(const ConvParams& p, const float* w_oihw, int sm_count, bool out_final, TcConvPlan* plan,
                    std::vector<void*>* allocs) {
  plan->ksplit = 1;
  if (!tc_epi_prepare(p, kStemNT, out_final, plan)) {
    set_error("conv_stem_tc: output tensor must be an internal bf16 NHWC tensor");
    return B200ROMP_EINVAL;
  }
  // weights as a 64 x 32 K-major matrix, k = r*9 + s*3 + c, scaled by 2/255 (the normalisation), zero padded
  std::vector<float> wk((size_t)64 * kStemK, 0.f);
  for (int co = 0; co < 64; ++co)
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s)
          wk[(size_t)co * kStemK + r * 9 + s * 3 + c] = w_oihw[(((size_t)co * 3 + c) * 3 + r) * 3 + s] * (2.0f / 255.0f);
  int rc = tc_pack_weights(wk.data(), kStemK, 64, 1, kStemNT, &plan->d_wpack, allocs, kStemK * 2, 2);
  if (rc) return rc;
  plan->kind = 33;
  plan->cin = 3; plan->cout = 64; plan->nt = kStemNT; plan->stages = kStemStages;
  plan->grid_x = sm_count; plan->grid_y = 1;
  plan->smem_bytes = kStemBBytes + kStemStages * kStemABytes + tc_epi_total_bytes(kStemEpi, kStemNT) +
                     kStemStages * kRawDepth * kRawStage + 2048;
  B2R_CUDA_OK(cudaFuncSetAttribute(conv_stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
  return B200ROMP_OK;
}

int tc_stem_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream) {
  TcEpiMaps em;
  memcpy(&em, plan.tmap_epi, sizeof(em));
  // raw-window tensor map over the (caller-owned, re-bindable) u8 frames: [N][H][W*3 bytes], zero fill outside
  PFN_encodeTiled encode = tc_get_encode();
  B2R_REQUIRE(encode && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.Win * 3) % 16 == 0,
              "conv_stem_tc: frames must be 16 B aligned with a row pitch that is a multiple of 16 B");
  CUtensorMap raw;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)p.Win * 3, (cuuint64_t)p.Hin, (cuuint64_t)p.B};
    const cuuint64_t gstr[2] = {(cuuint64_t)p.Win * 3, (cuuint64_t)p.Win * 3 * p.Hin};
    const cuuint32_t box[3] = {(cuuint32_t)kRawRowB, (cuuint32_t)kRawRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&raw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(p.in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2R_REQUIRE(cr == CUDA_SUCCESS, "conv_stem_tc: cuTensorMapEncodeTiled (frames) failed with %d", (int)cr);
  }
  const int tiles_x = p.Wout / 8, tiles_y = p.Hout / 16;
  const int num_tiles = tiles_x * tiles_y * p.B;
  dim3 grid(std::min(plan.grid_x, num_tiles), 1);
  B2R_CUDA_OK(tc_launch(conv_stem_tc_kernel, grid, kStemThreads, plan.smem_bytes, stream, raw, em, p,
                        reinterpret_cast<const uint8_t*>(plan.d_wpack), tiles_x, tiles_y, num_tiles));
  return B200ROMP_OK;
}

}  // namespace b200romp
