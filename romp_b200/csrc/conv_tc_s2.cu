// tcgen05 implicit-GEMM for the 3x3 stride-2 convolutions (HRNet transitions, fuse down-sampling chains, stem
// conv2, fused head-in conv; simple_romp/romp/model.py:201-218,275-285,341-343,449-457).
//
// A stride-2 3x3 conv on X[H,W,C] is a 2x2-tap stride-1 conv on the space-to-depth view X'[H/2,W/2,(ph,pw,C)]:
// input row 2*oy+r-1 is (hh,ph) = (oy-1,1), (oy,0), (oy,1) for r = 0,1,2 (same for columns).  The NHWC tensor is
// addressed by a 5-D TMA tensor map (dims: [pw*C+c], ww, ph, hh, n) so no data is rearranged in HBM.  Per
// (tile, 64-channel chunk) the producer issues 4 TMA loads - one per input parity (ph,pw) - of
// (17|16)x(9|8)-pixel sub-tiles; the 9 taps are shifted shared-memory descriptors into those 4 sub-tiles:
//   tap (r,s) -> sub-tile (ph = r!=1, pw = s!=1), start row (r==2)*BW + (s==2), stride-byte-offset = BW rows.
// Every input element is fetched once per tile (561 pixel rows instead of 9 x 128).  Everything else (resident
// weights, TMEM accumulator ring, epilogue) is shared with conv_tc.cu.
#include "tc_device.cuh"

namespace b200romp {

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(pol)
      : "memory");
}

// bytes per pixel row of a stage: half rows (64 B) once a full row would leave a single pipeline stage next to the resident
// weights (bf16 Cin = 256: 147 KB of weights + 74 KB per 128 B-row stage; measured 328 us single-stage for 256->64 @128^2)
__host__ __device__ constexpr int s2_row_bytes(int cin, int eb) { return cin * eb >= 512 ? 64 : (cin * eb < 128 ? cin * eb : 128); }

template <int CIN, int NT, int KSPLIT, int EB>
struct S2Cfg {
  static constexpr int ROWB = s2_row_bytes(CIN, EB);
  static constexpr int CW = ROWB / EB;
  static constexpr int KCH = CIN / CW;
  static constexpr int KSTEPS = ROWB / 32;
  static constexpr int LAYOUT = ROWB == 128 ? 2 : 4;
  // sub-tile i = (ph ? 0 : 2) + (pw ? 0 : 1):  (1,1) 17x9, (1,0) 17x8, (0,1) 16x9, (0,0) 16x8
  __host__ __device__ static constexpr int BH(int i) { return i < 2 ? 17 : 16; }
  __host__ __device__ static constexpr int BW(int i) { return (i & 1) ? 8 : 9; }
  __host__ __device__ static constexpr int SUB_BYTES(int i) { return (BH(i) * BW(i) * ROWB + 1023) / 1024 * 1024; }
  __host__ __device__ static constexpr int SUB_OFF(int i) { return i == 0 ? 0 : SUB_OFF(i - 1) + SUB_BYTES(i - 1); }
  static constexpr int STAGE_BYTES = SUB_OFF(3) + SUB_BYTES(3);
  static constexpr int STAGE_PAYLOAD = (17 * 9 + 17 * 8 + 16 * 9 + 16 * 8) * ROWB;
  static constexpr int BTILE = NT * ROWB;
  static constexpr int B_BYTES = 9 * KCH * BTILE;
  static constexpr int ACC = AccCfg<KSPLIT>::ACC;
  static constexpr int TMEM_COLS = tc_tmem_cols(ACC * KSPLIT * NT);
  static constexpr uint32_t IDESC = tc_idesc(EB, 128, NT);
};

struct S2Maps {
  CUtensorMap m[4];
};

template <int CIN, int NT, int KSPLIT, int EB>
__global__ void __launch_bounds__(tc_threads(EB), 1)
conv_tc_s2_kernel(const __grid_constant__ S2Maps maps, const __grid_constant__ TcEpiMaps epi_maps, const ConvParams p,
                  const uint8_t* __restrict__ wpack, int tiles_x, int tiles_y, int num_tiles, int stages, int tma_epi) {
  using Cfg = S2Cfg<CIN, NT, KSPLIT, EB>;
  constexpr int kAccStages = Cfg::ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + Cfg::B_BYTES;
  uint8_t* epi_smem = sA + (size_t)stages * Cfg::STAGE_BYTES;       // TMA-epilogue staging tiles, if any
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + (EB == 4 ? tc_epi_f32_total_bytes(tma_epi) : tc_epi_total_bytes(tma_epi, NT)));
  uint64_t* empty = full + stages;
  uint64_t* b_full = empty + stages;
  uint64_t* tmem_full = b_full + 1;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint64_t* res_bar = tmem_empty + kAccStages;
  uint64_t* landed = res_bar + 3 * kEpiWarps;        // EB = 4: "TMA tiles landed", consumed by the TF32 rounding warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(landed + (EB == 4 ? stages : 0));
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 2) + 15) & ~(uintptr_t)15);   // 16 B: ld.shared.v4

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], EB == 4 ? kCvtWarps : 1);
      mbar_init(&empty[i], 1);
      if (EB == 4) mbar_init(&landed[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= kFirstEpiWarp * 32 && threadIdx.x < kFirstEpiWarp * 32 + NT)
    s_bias[threadIdx.x - kFirstEpiWarp * 32] = p.bias[blockIdx.y * NT + threadIdx.x - kFirstEpiWarp * 32];
  if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();
  const int per_frame = tiles_x * tiles_y;
  const int nrings = tc_num_rings(stages);   // MMA-issuing warps in use = private stage rings

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, Cfg::B_BYTES);
      const uint8_t* wsrc = wpack + (size_t)blockIdx.y * Cfg::B_BYTES;
      for (int i = 0; i < 9 * Cfg::KCH; ++i)
        bulk_copy_g2s(sB + (size_t)i * Cfg::BTILE, wsrc + (size_t)i * Cfg::BTILE, Cfg::BTILE, b_full);
      pdl_wait();                             // weights are constants; activations must wait for the predecessor grids
      const uint64_t pol = l2_policy_stream(p.debug);
      int stage = 0, stage_other = 0;                  // one private stage ring per MMA warp (see conv_tc.cu)
      uint32_t phase = 0, phase_other = 0;   // (stage, phase) of the current tile's ring / of the other ring
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_frame, rem = tile % per_frame;
        const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;      // output coordinates
        const int ring = nrings == 2 ? (it & 1) : 0, rbase = tc_ring_base(stages, ring), rsize = tc_ring_size(stages, ring);
        for (int c = 0; c < Cfg::KCH; ++c) {
          const int sidx = rbase + stage;
          mbar_wait(&empty[sidx], phase ^ 1);
          uint64_t* land = EB == 4 ? &landed[sidx] : &full[sidx];
          mbar_arrive_expect_tx(land, Cfg::STAGE_PAYLOAD);
          uint8_t* dst = sA + (size_t)sidx * Cfg::STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ph = i < 2 ? 1 : 0, pw = (i & 1) ? 0 : 1;
            tma_load_5d(dst + Cfg::SUB_OFF(i), &maps.m[i], land, pw * p.in_C + p.in_c_off + c * Cfg::CW, x0 - pw, ph, y0 - ph, n, pol);
          }
          if (++stage == rsize) { stage = 0; phase ^= 1; }
        }
        if (nrings == 2) { const int ts = stage; stage = stage_other; stage_other = ts; const uint32_t tp = phase; phase = phase_other; phase_other = tp; }
      }
    }
  } else if (warp <= kMmaWarps) {
    if (warp <= nrings && elect_one()) {      // two MMA-issuing warps alternate tiles (see conv_tc.cu)
      mbar_wait(b_full, 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(sB);
      const int rbase = tc_ring_base(stages, warp - 1), rsize = tc_ring_size(stages, warp - 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = warp - 1;
      for (int tile = blockIdx.x + it * gridDim.x; tile < num_tiles; tile += nrings * gridDim.x, it += nrings) {
        const int acc = it & (kAccStages - 1);
        mbar_wait(&tmem_empty[acc], ((it / kAccStages) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * KSPLIT * NT);
        int mma_i = 0;
        for (int c = 0; c < Cfg::KCH; ++c) {
          const int sidx = rbase + stage;
          mbar_wait(&full[sidx], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + (size_t)sidx * Cfg::STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, s = t % 3;
            const int sub = (r != 1 ? 0 : 2) + (s != 1 ? 0 : 1);
            const int bw = Cfg::BW(sub);
            const uint32_t a_tap = a_base + (uint32_t)(Cfg::SUB_OFF(sub) + ((r == 2 ? bw : 0) + (s == 2 ? 1 : 0)) * Cfg::ROWB);
            const uint32_t b_tap = b_base + (uint32_t)((t * Cfg::KCH + c) * Cfg::BTILE);
#pragma unroll
            for (int k = 0; k < Cfg::KSTEPS; ++k) {
              const uint64_t adesc = make_smem_desc(a_tap + k * 32, bw * Cfg::ROWB, Cfg::LAYOUT);
              const uint64_t bdesc = make_smem_desc(b_tap + k * 32, 8 * Cfg::ROWB, Cfg::LAYOUT);
              umma_any<EB, false>(d_tile + (uint32_t)((mma_i % KSPLIT) * NT), adesc, bdesc, Cfg::IDESC, mma_i >= KSPLIT ? 1u : 0u);
              ++mma_i;
            }
          }
          umma_commit(&empty[sidx]);
          if (++stage == rsize) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (EB == 4 && warp >= kFirstCvtWarp) {
    // ===================== TF32 rounding warps: landed -> round in place (cvt.rna.tf32) -> full =====================
    const int cw = warp - kFirstCvtWarp, lane = threadIdx.x & 31;
    int stage = 0, stage_other = 0;
    uint32_t phase = 0, phase_other = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ring = nrings == 2 ? (it & 1) : 0, rbase = tc_ring_base(stages, ring), rsize = tc_ring_size(stages, ring);
      for (int c = 0; c < Cfg::KCH; ++c) {
        const int sidx = rbase + stage;
        mbar_wait(&landed[sidx], phase);
        tf32_round_smem(sA + (size_t)sidx * Cfg::STAGE_BYTES, Cfg::STAGE_BYTES, cw, lane);   // 4 sub-tiles (+ alignment gaps)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[sidx]);
        if (++stage == rsize) { stage = 0; phase ^= 1; }
      }
      if (nrings == 2) { const int ts = stage; stage = stage_other; stage_other = ts; const uint32_t tp = phase; phase = phase_other; phase_other = tp; }
    }
  } else if (EB == 4 && tma_epi) {
    tc_epilogue_loop_tma_f32<NT>(p, epi_maps, tma_epi, epi_smem, res_bar, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame,
                                 num_tiles);
  } else if (EB == 2 && KSPLIT == 1 && NT == 32 && (tma_epi & kEpiCoalesced)) {   // (instantiated for NT = 32 only: register pressure)
    tc_epilogue_loop_coalesced<NT>(p, tma_epi, epi_smem, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame, num_tiles);
  } else if (EB == 2 && KSPLIT == 1 && tma_epi) {
    tc_epilogue_loop_tma<NT>(p, epi_maps, tma_epi, epi_smem, res_bar, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame,
                             num_tiles);
  } else {
    tc_epilogue_loop<NT, KSPLIT>(p, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame, num_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
bool tc_s2_supported(const ConvParams& p) {
  if ((p.in_dtype != B200ROMP_BF16 && p.in_dtype != B200ROMP_F32) || p.input_norm || p.out_nchw || p.pow_channel >= 0) return false;
  const int eb = p.in_dtype == B200ROMP_F32 ? 4 : 2;
  if (eb == 4 && (p.out_dtype != B200ROMP_F32 || (p.res != nullptr && p.res_dtype != B200ROMP_F32))) return false;
  if (p.cin != 32 && p.cin != 64 && p.cin != 128 && p.cin != 256) return false;
  if (eb == 4 && p.cin == 256) return false;                   // fp32 weights of 256 channels do not fit: the graph builder splits K
  if (p.in_C % 8 != 0 || p.in_c_off % 8 != 0) return false;    // channel slice [in_c_off, in_c_off + cin) of the merged (pw, c) dim
  if (p.cout % 32 != 0 || p.Hin % 2 != 0 || p.Win % 2 != 0) return false;
  if (p.Hout % 16 != 0 || p.Wout % 8 != 0 || p.Hout * 2 != p.Hin || p.Wout * 2 != p.Win) return false;
  if (p.out_C % 8 != 0 || p.out_c_off % 8 != 0) return false;
  if (p.res != nullptr && (p.res_C % 8 != 0 || p.res_c_off % 8 != 0)) return false;
  if ((reinterpret_cast<uintptr_t>(p.in) & 15) != 0) return false;
  return true;
}

template <int CIN, int NT, int KSPLIT, int EB>
static int s2_inst(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
  auto kern = conv_tc_s2_kernel<CIN, NT, KSPLIT, EB>;
  if (attr) {
    B2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* plans of one instantiation differ (TMA epilogue staging) */));
    return B200ROMP_OK;
  }
  S2Maps maps;
  memcpy(&maps, plan.tmap_s2, sizeof(maps));
  TcEpiMaps em;
  memcpy(&em, plan.tmap_epi, sizeof(em));
  const int tiles_x = p.Wout / 8, tiles_y = p.Hout / 16;
  const int num_tiles = tiles_x * tiles_y * p.B;
  dim3 grid(std::min(plan.grid_x, num_tiles), plan.grid_y);
  B2R_CUDA_OK(tc_launch(kern, grid, tc_threads(EB), plan.smem_bytes, stream, maps, em, p, reinterpret_cast<const uint8_t*>(plan.d_wpack), tiles_x,
                        tiles_y, num_tiles, plan.stages, KSPLIT == 1 ? plan.tma_epi : 0));
  return B200ROMP_OK;
}

static int s2_dispatch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
#define B2R_S2(C, N)                                                                 \
  if (plan.cin == C && plan.nt == N && plan.eb == 2) {                             \
    if (plan.ksplit == 1) return s2_inst<C, N, 1, 2>(plan, p, stream, attr);       \
    return s2_inst<C, N, tc_ksplit(9 * (C / 16), N), 2>(plan, p, stream, attr);    \
  }
#define B2R_S2F(C, N) \
  if (plan.cin == C && plan.nt == N && plan.eb == 4) return s2_inst<C, N, 1, 4>(plan, p, stream, attr);
  B2R_S2(32, 32) B2R_S2(32, 64) B2R_S2(64, 64) B2R_S2(128, 32) B2R_S2(256, 32)
  B2R_S2F(32, 32) B2R_S2F(32, 64) B2R_S2F(64, 32) B2R_S2F(128, 32)
#undef B2R_S2
#undef B2R_S2F
  set_error("conv_tc_s2: no instantiation for cin%d nt%d", plan.cin, plan.nt);
  return B200ROMP_EINVAL;
}

int tc_s2_prepare(const ConvParams& p, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan, std::vector<void*>* allocs) {
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) {
    set_error("conv_tc_s2: cuTensorMapEncodeTiled is unavailable");
    return B200ROMP_ECUDA;
  }
  const int eb = p.in_dtype == B200ROMP_F32 ? 4 : 2;
  plan->eb = eb;
  const int rowb = s2_row_bytes(p.cin, eb), cw = rowb / eb, kch = p.cin / cw;
  auto sub_bytes = [&](int i) { return ((i < 2 ? 17 : 16) * ((i & 1) ? 8 : 9) * rowb + 1023) / 1024 * 1024; };
  const int stage_bytes = sub_bytes(0) + sub_bytes(1) + sub_bytes(2) + sub_bytes(3);
  const int budget = 227 * 1024 - 2048;
  auto bbytes = [&](int n) { return 9 * kch * n * rowb; };
  int nt = (p.cout % 64 == 0) ? 64 : 32;
  if (nt == 64 && bbytes(64) + 2 * stage_bytes > budget && bbytes(32) + 2 * stage_bytes <= budget) nt = 32;
  if (nt == 64 && bbytes(64) + stage_bytes > budget) nt = 32;
  if (bbytes(nt) + stage_bytes > budget) {
    set_error("conv_tc_s2: cin%d does not fit shared memory", p.cin);
    return B200ROMP_EINVAL;
  }
  { const char* e = getenv("B200ROMP_TC_KSPLIT"); plan->ksplit = (e && e[0] == '4' && eb == 2) ? 0 : 1; }   /* K-split measured slower: off by default */
  // TMA epilogue (bf16 tensors) only where it does not cost the second pipeline stage
  int epi_bytes = 0;
  plan->tma_epi = 0;
  if (eb == 2 && tc_epi_prepare(p, nt, ptrs_final, plan)) {
    if (tc_epi_want_coalesced(nt)) plan->tma_epi |= kEpiCoalesced;
    epi_bytes = tc_epi_total_bytes(plan->tma_epi, nt);
    const int without = (budget - bbytes(nt)) / stage_bytes, with = (budget - bbytes(nt) - epi_bytes) / stage_bytes;
    if (with < 1 || (with < 2 && without >= 2)) { plan->tma_epi = 0; epi_bytes = 0; }
  }
  plan->stages = std::min(4, (budget - bbytes(nt) - epi_bytes) / stage_bytes);
  plan->kind = 32;
  plan->cin = p.cin; plan->cout = p.cout; plan->nt = nt;
  plan->grid_y = p.cout / nt;
  plan->grid_x = std::max(1, sm_count / plan->grid_y);
  plan->smem_bytes = bbytes(nt) + plan->stages * stage_bytes + epi_bytes + 2048;
  int rc = tc_pack_weights(w_oihw, p.cin, p.cout, 9, nt, &plan->d_wpack, allocs, rowb, eb);
  if (rc) return rc;
  const cuuint64_t C = (cuuint64_t)p.in_C;
  const cuuint64_t gdim[5] = {2 * C, (cuuint64_t)p.Win / 2, 2, (cuuint64_t)p.Hin / 2, (cuuint64_t)p.B};
  const cuuint64_t E = (cuuint64_t)eb;
  const cuuint64_t gstr[4] = {2 * C * E, (cuuint64_t)p.Win * C * E, 2 * (cuuint64_t)p.Win * C * E, (cuuint64_t)p.Hin * p.Win * C * E};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    const cuuint32_t box[5] = {(cuuint32_t)cw, (cuuint32_t)((i & 1) ? 8 : 9), 1, (cuuint32_t)(i < 2 ? 17 : 16), 1};
    CUtensorMap tm;
    CUresult cr = encode(&tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(p.in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("conv_tc_s2: cuTensorMapEncodeTiled failed with %d", (int)cr);
      return B200ROMP_ECUDA;
    }
    memcpy(plan->tmap_s2[i], &tm, sizeof(tm));
  }
  return s2_dispatch(*plan, p, nullptr, true);
}

int tc_s2_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream) { return s2_dispatch(plan, p, stream, false); }

}  // namespace b200romp
