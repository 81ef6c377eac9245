// tcgen05 GEMM for BEV's bird's-eye-view Conv1d stack (simple_romp/bev/model.py:24-45,179-182: three BasicBlock_1D =
// six Conv1d(k=3, pad=1)+BN1d+ReLU, 2560 -> 512 -> 512 | 512 -> 512 | 512 -> 128 -> 128, over the 128 image columns).
// Round 1 ran these on the fp32 SIMT engine: 27.6 % of the BEV step (profiles/r01_launches_final_bev.md).
//
// The activation is the NHWC "image" [B, 1, W = 128, C] (channels innermost = K-major rows), ksize code 13 = 1x3.
//   GEMM M = 128 positions (one row of one frame; UMMA_M = 128), N = NT output channels, K = 3 taps x Cin.
//   A: one TMA load per (tile, 64-channel chunk) of the 130-position halo row (x0 = -1, out-of-bounds = zero padding);
//      the 3 taps are 3 shared-memory descriptors shifted by one 128 B row each (same trick as conv_tc.cu).
//   B: Cin = 2560 makes the weights of one N tile 983 KB - not resident: each pipeline stage carries its own
//      [3 taps][NT rows x 128 B] weight slab (24 KB for NT = 64) next to the A row (17 KB); weights are packed
//      chunk-major so the slab of a stage is one contiguous bulk copy.
//   D: fp32 in TMEM, two accumulators; epilogue thread = position: + bias, (+ residual), ReLU, 16 B stores.
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocation), 2-5 = epilogue (TMEM lane quarter = warp % 4).
#include "conv_tc.cuh"
#include "tc_device.cuh"

namespace b200romp {

constexpr int k1dThreads = 192;
constexpr int k1dARows = 130;
constexpr int k1dABytes = (k1dARows * 128 + 1023) / 1024 * 1024;   // 17408

template <int NT>
struct C1dCfg {
  static constexpr int BBYTES = 3 * NT * 128;
  static constexpr int STAGE_BYTES = k1dABytes + BBYTES;
  static constexpr int TMEM_COLS = tc_tmem_cols(2 * NT);
  static constexpr uint32_t IDESC = tc_idesc(2, 128, NT);
};

template <int NT>
__global__ void __launch_bounds__(k1dThreads, 1)
conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvParams p, const uint8_t* __restrict__ wpack, int kch, int tiles_w,
                 int num_tiles, int stages) {
  using Cfg = C1dCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * Cfg::STAGE_BYTES);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NT) s_bias[threadIdx.x - 64] = p.bias[blockIdx.y * NT + threadIdx.x - 64];
  if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();

  if (warp == 0) {
    if (elect_one()) {
      pdl_wait();
      const uint64_t pol = l2_policy_stream(0);
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* wsrc = wpack + (size_t)blockIdx.y * kch * Cfg::BBYTES;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row = tile / tiles_w, x0 = (tile % tiles_w) * 128;
        for (int c = 0; c < kch; ++c) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = smem + (size_t)stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], k1dARows * 128 + Cfg::BBYTES);
          tma_load_3d(dst, &tmap, &full[stage], p.in_c_off + c * 64, x0 - 1, row, pol);
          bulk_copy_g2s(dst + k1dABytes, wsrc + (size_t)c * Cfg::BBYTES, Cfg::BBYTES, &full[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * NT);
        int mma_i = 0;
        for (int c = 0; c < kch; ++c) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + k1dABytes;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(a_base + t * 128 + k * 32, 1024, 2);
              const uint64_t bdesc = make_smem_desc(b_base + t * NT * 128 + k * 32, 1024, 2);
              umma_bf16(d_tile, adesc, bdesc, Cfg::IDESC, mma_i > 0 ? 1u : 0u);
              ++mma_i;
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue: thread = one position of the tile =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int co0 = blockIdx.y * NT;
    pdl_wait();
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int row = tile / tiles_w, x = (tile % tiles_w) * 128 + m;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
      const size_t pix = (size_t)row * p.Wout + x;
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
        if (co0 + c0 < p.cout) tc_store32(p, pix, pix, co0 + c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
bool tc_conv1d_supported(const ConvParams& p) {
  if (p.in_dtype != B200ROMP_BF16 || p.input_norm || p.out_nchw || p.pow_channel >= 0 || p.up != 1) return false;
  if (p.cin % 64 != 0 || p.cout % 32 != 0 || p.Wout % 128 != 0 || p.Wout != p.Win || p.Hout != p.Hin) return false;
  if (p.in_C % 8 != 0 || p.in_c_off % 8 != 0 || p.out_C % 8 != 0 || p.out_c_off % 8 != 0) return false;
  if (p.res != nullptr && (p.res_C % 8 != 0 || p.res_c_off % 8 != 0 || p.res_broadcast)) return false;
  return true;
}

static int conv1d_encode(const ConvParams& p, TcConvPlan* plan) {
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) { set_error("conv1d_tc: cuTensorMapEncodeTiled is unavailable"); return B200ROMP_ECUDA; }
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)p.in_C, (cuuint64_t)p.Win, (cuuint64_t)p.B * p.Hin};
  const cuuint64_t gstr[2] = {(cuuint64_t)p.in_C * 2, (cuuint64_t)p.Win * p.in_C * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)k1dARows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error("conv1d_tc: cuTensorMapEncodeTiled failed with %d", (int)cr); return B200ROMP_ECUDA; }
  memcpy(plan->tmap_in, &tm, sizeof(tm));
  plan->encoded_in = p.in;
  plan->encoded_batch = p.B;
  return B200ROMP_OK;
}

template <int NT>
static int conv1d_inst(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
  auto kern = conv1d_tc_kernel<NT>;
  if (attr) {
    B2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return B200ROMP_OK;
  }
  CUtensorMap tm;
  memcpy(&tm, plan.tmap_in, sizeof(tm));
  const int tiles_w = p.Wout / 128, num_tiles = tiles_w * p.Hout * p.B;
  dim3 grid(std::min(plan.grid_x, num_tiles), plan.grid_y);
  B2R_CUDA_OK(tc_launch(kern, grid, k1dThreads, plan.smem_bytes, stream, tm, p, reinterpret_cast<const uint8_t*>(plan.d_wpack), plan.cin / 64,
                        tiles_w, num_tiles, plan.stages));
  return B200ROMP_OK;
}

int tc_conv1d_prepare(const ConvParams& p, const float* w_oi3, int sm_count, TcConvPlan* plan, std::vector<void*>* allocs) {
  const int nt = (p.cout % 64 == 0) ? 64 : 32;
  const int kch = p.cin / 64, ntiles = p.cout / nt;
  const int stage_bytes = k1dABytes + 3 * nt * 128;
  plan->kind = 13; plan->eb = 2; plan->ksplit = 1; plan->tma_epi = 0;
  plan->cin = p.cin; plan->cout = p.cout; plan->nt = nt;
  plan->stages = std::min(6, (227 * 1024 - 2048) / stage_bytes);
  plan->grid_y = ntiles;
  plan->grid_x = std::max(1, sm_count / ntiles);
  plan->smem_bytes = plan->stages * stage_bytes + 2048;
  // weights: [ntile][chunk][tap][nt rows x 128 B], SWIZZLE_128B inside each (tap) tile; w_oi3 = [cout][cin][3]
  std::vector<__nv_bfloat16> img((size_t)ntiles * kch * 3 * nt * 64, __float2bfloat16_rn(0.f));
  for (int j = 0; j < ntiles; ++j)
    for (int c = 0; c < kch; ++c)
      for (int t = 0; t < 3; ++t) {
        __nv_bfloat16* tile = img.data() + ((((size_t)j * kch + c) * 3 + t) * nt) * 64;
        for (int n = 0; n < nt; ++n)
          for (int k = 0; k < 64; ++k) {
            const float w = w_oi3[((size_t)(j * nt + n) * p.cin + c * 64 + k) * 3 + t];
            const size_t byte = (size_t)n * 128 + (size_t)(((k / 8) ^ (n & 7)) * 16) + (k % 8) * 2;
            tile[byte / 2] = __float2bfloat16_rn(w);
          }
      }
  B2R_CUDA_OK(cudaMalloc(&plan->d_wpack, img.size() * sizeof(__nv_bfloat16)));
  allocs->push_back(plan->d_wpack);
  B2R_CUDA_OK(cudaMemcpy(plan->d_wpack, img.data(), img.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  int rc = conv1d_encode(p, plan);
  if (rc) return rc;
  return nt == 64 ? conv1d_inst<64>(*plan, p, nullptr, true) : conv1d_inst<32>(*plan, p, nullptr, true);
}

// the input of the bird's-eye graph is an EXTERNAL tensor (assembled by b200romp_bev_bv_input): re-encode the tensor map
// when the bound pointer (or the batch) differs from the one the map was built for
int tc_conv1d_launch(TcConvPlan& plan, const ConvParams& p, cudaStream_t stream) {
  if (plan.encoded_in != p.in || plan.encoded_batch != p.B) {
    int rc = conv1d_encode(p, &plan);
    if (rc) return rc;
  }
  return plan.nt == 64 ? conv1d_inst<64>(plan, p, stream, false) : conv1d_inst<32>(plan, p, stream, false);
}

}  // namespace b200romp
