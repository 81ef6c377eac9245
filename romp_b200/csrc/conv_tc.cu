// tcgen05 implicit-GEMM convolution engine for sm_100a (the dominant kernel of the hot path).
//
// Replaces, per layer, cuDNN Conv2d + separate BatchNorm/ReLU/add/Upsample kernels of the reference
// (simple_romp/romp/model.py:49-123,185-244) for 3x3-stride-1 and 1x1 convolutions with
// Cin in {32,64,128,256} on NHWC bf16 activations (fp32 accumulate in TMEM).
//
// Mapping (im2col-free):
//   GEMM M = 128 output pixels = one 16x8 spatial tile of one frame          (UMMA_M = 128, cta_group::1)
//   GEMM N = NT output channels (32 or 64) per CTA, weights RESIDENT in shared memory for the CTA's
//            lifetime (persistent CTAs loop over pixel tiles) - streaming them would cost 64 B/clk/SM
//   GEMM K = taps x Cin, walked as (64-channel chunk) x (tap) x (16-channel UMMA_K step)
//   A operand: ONE TMA load per (tile, chunk) brings the (16+2)x(8+2) halo tile, channels innermost,
//            hardware-swizzled (128B, or 64B for Cin=32).  The 9 taps are 9 *shifted shared-memory
//            descriptors* into that same halo tile: start address + (r*10+s) pixel rows, stride-byte-offset
//            = 10 pixel rows (one image row of the halo), so the input is read from L2 once, not 9 times.
//            Zero padding comes from TMA out-of-bounds fill (negative start coordinates).
//   D: fp32 accumulators in TMEM, ring of 8 (8 x NT columns) so epilogues overlap the MMAs of later tiles.
//   Epilogue (2 groups x 4 warps, one TMEM lane quarter per warp, groups alternate tiles), tcgen05.ld 32x32b.x32 -> +bias:
//            * TMA epilogue (bf16 NHWC outputs at conv resolution): per-warp staging tile in the TMA swizzle pattern,
//              residual by TMA load, result by TMA tensor store (tc_device.cuh: tc_epilogue_loop_tma);
//            * direct epilogue (fp32 / upsampled / NCHW-map outputs): residual / ReLU / nearest-upsample replication /
//              dtype conversion with batched 16-byte global accesses (tc_epilogue_loop).
// Warp roles (352 threads): warp0 = TMA producer, warps 1-2 = MMA issuers (one elected thread each, alternating tiles, a
//            private stage ring each; warp1 also allocates TMEM), warps 3-10 = epilogue.
// Pipelines: per-ring smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), persistent tile loop; launched
//            with programmatic dependent launch (prologue overlaps the previous conv's tail).
// The 3x3 stride-1 convs with bf16 outputs normally run on the CTA-pair variant of this kernel (conv_tc_2cta.cu).
#include <mutex>

#include "conv_tc.cuh"
#include "tc_device.cuh"

namespace b200romp {

thread_local int g_tc_pdl_override = -1;

template <int KS, int CIN, int NT, bool PER_TAP, int KSPLIT, int EB>
struct TcCfg {
  static constexpr int TAPS = KS * KS;
  static constexpr int PAD = KS / 2;
  // bytes per pixel row in smem (64 or 128 = one swizzle span); channels per chunk = ROWB / EB.  The weight-heavy 3x3 layers
  // (Cin >= 128: 147 KB resident bf16 weights) use half rows: 4-5 stages of 12 KB instead of 2 of 23 KB, so loads run
  // ahead of the MMAs
  static constexpr int ROWB = tc_row_bytes(KS, CIN, EB);
  static constexpr int CW = ROWB / EB;
  static constexpr int KCH = CIN / CW;
  static constexpr int KSTEPS = ROWB / 32;             // UMMA K steps (32 B = 16 bf16 / 8 tf32) per row
  static constexpr int LAYOUT = ROWB == 128 ? 2 : 4;   // SWIZZLE_128B / SWIZZLE_64B
  static constexpr int HW_ = PER_TAP ? 8 : 8 + 2 * PAD;
  static constexpr int HH = PER_TAP ? 16 : 16 + 2 * PAD;
  static constexpr int LOADS_PER_CHUNK = PER_TAP ? TAPS : 1;
  static constexpr int STAGE_PAYLOAD = HH * HW_ * ROWB;
  static constexpr int STAGE_BYTES = (STAGE_PAYLOAD + 1023) / 1024 * 1024;
  static constexpr int BTILE = NT * ROWB;
  static constexpr int B_BYTES = TAPS * KCH * BTILE;
  static constexpr int ACC = AccCfg<KSPLIT>::ACC;
  static constexpr int TMEM_COLS = tc_tmem_cols(ACC * KSPLIT * NT);
  static constexpr uint32_t IDESC = tc_idesc(EB, 128, NT);
};

template <int KS, int CIN, int NT, bool PER_TAP, int KSPLIT, int EB>
__global__ void __launch_bounds__(tc_threads(EB), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcEpiMaps epi_maps, const ConvParams p,
               const uint8_t* __restrict__ wpack, int tiles_x, int tiles_y, int num_tiles, int stages, int tma_epi) {
  using Cfg = TcCfg<KS, CIN, NT, PER_TAP, KSPLIT, EB>;
  constexpr int kAccStages = Cfg::ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + Cfg::B_BYTES;
  uint8_t* epi_smem = sA + (size_t)stages * Cfg::STAGE_BYTES;       // TMA-epilogue staging tiles (1024 B aligned), if any
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + (EB == 4 ? tc_epi_f32_total_bytes(tma_epi) : tc_epi_total_bytes(tma_epi, NT)));
  uint64_t* empty = full + stages;
  uint64_t* b_full = empty + stages;
  uint64_t* tmem_full = b_full + 1;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint64_t* res_bar = tmem_empty + kAccStages;
  uint64_t* landed = res_bar + 3 * kEpiWarps;        // EB = 4: "TMA tile landed", consumed by the TF32 rounding warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(landed + (EB == 4 ? stages : 0));
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 2) + 15) & ~(uintptr_t)15);   // 16 B: ld.shared.v4

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], EB == 4 ? kCvtWarps : 1);   // EB = 4: the rounding warps hand the converted tile to the MMA warp
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
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, Cfg::B_BYTES);
      const uint8_t* wsrc = wpack + (size_t)blockIdx.y * Cfg::B_BYTES;
      for (int i = 0; i < Cfg::TAPS * Cfg::KCH; ++i)
        bulk_copy_g2s(sB + (size_t)i * Cfg::BTILE, wsrc + (size_t)i * Cfg::BTILE, Cfg::BTILE, b_full);
      pdl_wait();                             // weights are constants; activations must wait for the predecessor grids
      const uint64_t pol = l2_policy_stream(p.debug);
      // Each MMA warp owns a private stage ring (ring r = stages [ring_base(r), ring_base(r) + ring_size(r))): mbarrier
      // waits only see the phase parity, so a ring must have exactly one in-order consumer (TMA completions of
      // different stages arrive out of order, a shared ring would alias phases).  Tile i of this CTA goes to ring i & 1.
      int stage = 0, stage_other = 0;        
      uint32_t phase = 0, phase_other = 0;   // (stage, phase) of the current tile's ring / of the other ring
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_frame, rem = tile % per_frame;
        const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
        const int ring = nrings == 2 ? (it & 1) : 0, rbase = tc_ring_base(stages, ring), rsize = tc_ring_size(stages, ring);
        for (int c = 0; c < Cfg::KCH; ++c) {
          for (int l = 0; l < Cfg::LOADS_PER_CHUNK; ++l) {
            const int sidx = rbase + stage;
            mbar_wait(&empty[sidx], phase ^ 1);
            uint64_t* land = EB == 4 ? &landed[sidx] : &full[sidx];
            if (p.debug & 4) {
              mbar_arrive(land);
            } else {
              mbar_arrive_expect_tx(land, Cfg::STAGE_PAYLOAD);
              const int dy = PER_TAP ? l / KS - Cfg::PAD : -Cfg::PAD;
              const int dx = PER_TAP ? l % KS - Cfg::PAD : -Cfg::PAD;
              tma_load_4d(sA + (size_t)sidx * Cfg::STAGE_BYTES, &tmap, land, c * Cfg::CW, x0 + dx, y0 + dy, n, pol);
            }
            if (++stage == rsize) { stage = 0; phase ^= 1; }
          }
        }
        if (nrings == 2) { const int ts = stage; stage = stage_other; stage_other = ts; const uint32_t tp = phase; phase = phase_other; phase_other = tp; }
      }
    }
  } else if (warp <= kMmaWarps) {
    // ===================== MMA issuers: warp w takes the CTA's tiles w-1, w-1+kMmaWarps, ... (one elected thread each) =====================
    if (warp <= nrings && elect_one()) {
      mbar_wait(b_full, 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(sB);
      const int rbase = tc_ring_base(stages, warp - 1), rsize = tc_ring_size(stages, warp - 1);   // this warp's private ring
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
          for (int l = 0; l < Cfg::LOADS_PER_CHUNK; ++l) {
            const int sidx = rbase + stage;
            mbar_wait(&full[sidx], phase);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA + (size_t)sidx * Cfg::STAGE_BYTES);
            const int t_lo = PER_TAP ? l : 0, t_hi = PER_TAP ? l + 1 : Cfg::TAPS;
            for (int t = t_lo; t < t_hi; ++t) {
              const int r = t / KS, s = t % KS;
              const uint32_t a_tap = PER_TAP ? a_base : a_base + (uint32_t)((r * Cfg::HW_ + s) * Cfg::ROWB);
              const uint32_t b_tap = b_base + (uint32_t)((t * Cfg::KCH + c) * Cfg::BTILE);
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k) {
                const uint64_t adesc = make_smem_desc(a_tap + k * 32, Cfg::HW_ * Cfg::ROWB, Cfg::LAYOUT);
                const uint64_t bdesc = make_smem_desc(b_tap + k * 32, 8 * Cfg::ROWB, Cfg::LAYOUT);
                if (!(p.debug & 2)) umma_any<EB, false>(d_tile + (uint32_t)((mma_i % KSPLIT) * NT), adesc, bdesc, Cfg::IDESC, mma_i >= KSPLIT ? 1u : 0u);
                ++mma_i;
              }
            }
            umma_commit(&empty[sidx]);           // smem stage reusable once these MMAs retire
            if (++stage == rsize) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full[acc]);            // accumulator complete -> epilogue
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
      for (int c = 0; c < Cfg::KCH * Cfg::LOADS_PER_CHUNK; ++c) {
        const int sidx = rbase + stage;
        mbar_wait(&landed[sidx], phase);
        tf32_round_smem(sA + (size_t)sidx * Cfg::STAGE_BYTES, Cfg::STAGE_PAYLOAD, cw, lane);
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
// host side
// ------------------------------------------------------------------------------------------------
PFN_encodeTiled tc_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

int tc_pack_weights(const float* w_oihw, int cin, int cout, int taps, int nt, void** d_out, std::vector<void*>* allocs, int rowb, int eb) {
  // shared-memory image [n-tile][tap][chunk][NT rows x rowb bytes] with the TMA/UMMA XOR swizzle; elements bf16 (eb = 2)
  // or fp32 rounded to TF32 (eb = 4)
  const int cw = rowb / eb, kch = cin / cw, ntiles = (cout + nt - 1) / nt, per16 = 16 / eb;
  std::vector<uint8_t> img((size_t)ntiles * taps * kch * nt * rowb, 0);
  for (int j = 0; j < ntiles; ++j)
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < kch; ++c) {
        uint8_t* tile = img.data() + (((size_t)j * taps + t) * kch + c) * nt * rowb;
        for (int n = 0; n < nt; ++n)
          for (int k = 0; k < cw; ++k) {
            const int co = j * nt + n, ci = c * cw + k;
            if (co >= cout) continue;                                     // zero rows pad cout up to a multiple of NT
            const float w = w_oihw[((size_t)co * cin + ci) * taps + t];
            const int chunk16 = k / per16;
            const int phase = rowb == 128 ? (n & 7) : ((n >> 1) & 3);     // Swizzle<3,4,3> / Swizzle<2,4,3>
            const size_t byte = (size_t)n * rowb + (size_t)((chunk16 ^ phase) * 16) + (k % per16) * eb;
            if (eb == 2) { const __nv_bfloat16 b = __float2bfloat16_rn(w); memcpy(tile + byte, &b, 2); }
            else { const float f = tc_round_tf32_host(w); memcpy(tile + byte, &f, 4); }
          }
      }
  B2R_CUDA_OK(cudaMalloc(d_out, img.size()));
  allocs->push_back(*d_out);
  B2R_CUDA_OK(cudaMemcpy(*d_out, img.data(), img.size(), cudaMemcpyHostToDevice));
  return B200ROMP_OK;
}

std::string TcConvPlan::describe() const {
  char buf[128];
  snprintf(buf, sizeof(buf), " [tc%s k%d v%d nt%d grid %dx%d smem %d stages %d epi%d%s]", eb == 4 ? "-tf32" : "", kind / 10, kind % 10, nt, grid_x, grid_y, smem_bytes, stages, tma_epi,
           kmask != 0xFFFFFFFFu ? " pixel-pairs" : "");
  return buf;
}

bool tc_conv_supported(const ConvParams& p, int ksize, int stride) {
  if (stride == 2 && ksize == 3) return tc_s2_supported(p);
  if (stride != 1 || (ksize != 1 && ksize != 3)) return false;
  if ((p.in_dtype != B200ROMP_BF16 && p.in_dtype != B200ROMP_F32) || p.input_norm) return false;   // F32 = the TF32 engine
  if (p.in_dtype == B200ROMP_F32 && (p.out_dtype != B200ROMP_F32 || (p.res != nullptr && p.res_dtype != B200ROMP_F32))) return false;
  if (p.cin != 32 && p.cin != 64 && p.cin != 128 && p.cin != 256) return false;
  if (p.Hout % 16 != 0 || p.Wout % 8 != 0) return false;
  if (p.in_C % 8 != 0 || p.in_c_off % 8 != 0) return false;             // TMA: 16 B aligned base and strides
  if (p.out_nchw) {                                                      // map outputs: any cout (padded to 32), scalar stores
    if (p.out_dtype != B200ROMP_F32 || p.up != 1 || p.res != nullptr) return false;
    return (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
  }
  if (p.pow_channel >= 0 || p.cout % 32 != 0) return false;
  if (p.out_C % 8 != 0 || p.out_c_off % 8 != 0) return false;           // 16 B vector epilogue
  if (p.res != nullptr && (p.res_C % 8 != 0 || p.res_c_off % 8 != 0)) return false;
  if ((reinterpret_cast<uintptr_t>(p.in) & 15) != 0) return false;
  return true;
}

template <int KS, int CIN, int NT, bool PER_TAP, int KSPLIT, int EB>
static int launch_inst(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool set_attr_only) {
  auto kern = conv_tc_kernel<KS, CIN, NT, PER_TAP, KSPLIT, EB>;
  if (set_attr_only) {
    B2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* plans of one instantiation differ (TMA epilogue staging) */));
    return B200ROMP_OK;
  }
  CUtensorMap tm;
  memcpy(&tm, plan.tmap_in, sizeof(tm));
  TcEpiMaps em;
  memcpy(&em, plan.tmap_epi, sizeof(em));
  const int tiles_x = p.Wout / 8, tiles_y = p.Hout / 16;
  const int num_tiles = tiles_x * tiles_y * p.B;
  dim3 grid(std::min(plan.grid_x, num_tiles), plan.grid_y);
  B2R_CUDA_OK(tc_launch(kern, grid, tc_threads(EB), plan.smem_bytes, stream, tm, em, p, reinterpret_cast<const uint8_t*>(plan.d_wpack), tiles_x,
                        tiles_y, num_tiles, plan.stages, KSPLIT == 1 ? plan.tma_epi : 0));
  return B200ROMP_OK;
}

template <bool PER_TAP>
static int dispatch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
  const int ks = plan.kind / 10;
#define B2R_CASE(K, C, N)                                                                            \
  if (ks == K && plan.cin == C && plan.nt == N && plan.eb == 2) {                                   \
    if (plan.ksplit == 1) return launch_inst<K, C, N, PER_TAP, 1, 2>(plan, p, stream, attr);         \
    return launch_inst<K, C, N, PER_TAP, tc_ksplit(K * K * (C / 16), N), 2>(plan, p, stream, attr);   \
  }
#define B2R_CASE4(K, C, N) \
  if (ks == K && plan.cin == C && plan.nt == N && plan.eb == 4) return launch_inst<K, C, N, PER_TAP, 1, 4>(plan, p, stream, attr);
  B2R_CASE(3, 32, 32) B2R_CASE(3, 32, 64) B2R_CASE(3, 64, 64) B2R_CASE(3, 128, 64) B2R_CASE(3, 256, 32)
  B2R_CASE(1, 64, 32) B2R_CASE(1, 64, 64) B2R_CASE(1, 128, 32) B2R_CASE(1, 128, 64) B2R_CASE(1, 256, 32)
  B2R_CASE(1, 256, 64) B2R_CASE(1, 32, 64) B2R_CASE(1, 32, 32)
  // TF32 engine (fp32 tensors): every 1x1 conv; 3x3 only as the fall-back of the CTA-pair engine (odd tile counts)
  B2R_CASE4(1, 32, 32) B2R_CASE4(1, 32, 64) B2R_CASE4(1, 64, 32) B2R_CASE4(1, 64, 64) B2R_CASE4(1, 128, 32) B2R_CASE4(1, 128, 64)
  B2R_CASE4(1, 256, 32) B2R_CASE4(1, 256, 64) B2R_CASE4(3, 32, 32) B2R_CASE4(3, 64, 32) B2R_CASE4(3, 64, 64)
#undef B2R_CASE4
#undef B2R_CASE
  set_error("conv_tc: no instantiation for k%d cin%d nt%d", ks, plan.cin, plan.nt);
  return B200ROMP_EINVAL;
}

// TMA epilogue eligibility + tensor maps over the output / residual tensors: dims (C, W, H, N), box (NT, 8, 4, 1) = the
// 32 pixels one epilogue warp owns, swizzle matching tc_epi_chunk().  Returns 0 when the direct epilogue must be used.
int tc_epi_prepare(const ConvParams& p, int nt, bool ptrs_final, TcConvPlan* plan) {
  plan->tma_epi = 0;
  const char* e = getenv("B200ROMP_TC_NO_TMA_EPI");
  if (e && e[0] == '1') return 0;
  if (!ptrs_final || plan->ksplit != 1 || p.out_dtype != B200ROMP_BF16 || p.out_nchw || p.up != 1 || p.pow_channel >= 0) return 0;
  if (p.cout % nt != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return 0;
  const bool res_tma = p.res != nullptr;
  if (res_tma && (p.res_dtype != B200ROMP_BF16 || p.res_broadcast || (reinterpret_cast<uintptr_t>(p.res) & 15) != 0)) return 0;
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) return 0;
  auto make = [&](const void* base, int C, unsigned char* dst) -> bool {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)p.Wout, (cuuint64_t)p.Hout, (cuuint64_t)p.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)p.Wout * C * 2, (cuuint64_t)p.Hout * p.Wout * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)nt, 8, 4, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, nt == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return false;
    memcpy(dst, &tm, sizeof(tm));
    return true;
  };
  if (!make(p.out, p.out_C, plan->tmap_epi[0])) return 0;
  if (res_tma) {
    if (!make(p.res, p.res_C, plan->tmap_epi[1])) return 0;
  } else {
    memcpy(plan->tmap_epi[1], plan->tmap_epi[0], 128);
  }
  plan->tma_epi = kTmaEpiOut | (res_tma ? kTmaEpiRes : 0);
  return plan->tma_epi;
}

int tc_epi_prepare_f32(const ConvParams& p, int nt, bool ptrs_final, int avail, int stage_bytes, TcConvPlan* plan) {
  plan->tma_epi = 0;
  const char* e = getenv("B200ROMP_TC_NO_TMA_EPI");
  if (e && e[0] == '1') return 0;
  if (!ptrs_final || p.out_dtype != B200ROMP_F32 || p.out_nchw || p.up != 1 || p.pow_channel >= 0) return 0;
  if (p.cout % nt != 0 || nt % 32 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || (p.out_C % 4) != 0 || (p.out_c_off % 4) != 0) return 0;
  const bool res_tma = p.res != nullptr;
  if (res_tma && (p.res_dtype != B200ROMP_F32 || p.res_broadcast || (reinterpret_cast<uintptr_t>(p.res) & 15) != 0 || (p.res_C % 4) != 0 ||
                  (p.res_c_off % 4) != 0))
    return 0;
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) return 0;
  auto make = [&](const void* base, int C, unsigned char* dst) -> bool {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)p.Wout, (cuuint64_t)p.Hout, (cuuint64_t)p.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)p.Wout * C * 4, (cuuint64_t)p.Hout * p.Wout * C * 4};
    const cuuint32_t box[4] = {32, 8, 4, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    memcpy(dst, &tm, sizeof(tm));
    return true;
  };
  if (!make(p.out, p.out_C, plan->tmap_epi[0])) return 0;
  if (res_tma) {
    if (!make(p.res, p.res_C, plan->tmap_epi[1])) return 0;
  } else {
    memcpy(plan->tmap_epi[1], plan->tmap_epi[0], 128);
  }
  const int flags = kTmaEpiOut | (res_tma ? kTmaEpiRes : 0);
  // staging depth: 3 buffers with a residual (compute / prefetch / store in flight), 2 without, while >= 4 pipeline stages
  // remain, then while >= 3 remain; never a single buffer: a plan that cannot keep 3 stages next to two buffers per warp uses
  // the direct epilogue (measured: 256->256 @16x16 with one buffer and 3 stages 118 us against 88 us direct with 6 stages;
  // 64->64 @64x64 with two buffers and 3 stages 43-50 us against 68 us direct)
  const int want = res_tma ? 3 : 2;
  for (int min_stages = 4; min_stages >= 3; --min_stages)
    for (int nb = want; nb >= 2; --nb) {
      const int bytes = kEpiWarps * nb * kF32ChunkBytes;
      if ((avail - bytes) / stage_bytes >= min_stages) {
        plan->tma_epi = tc_epi_with_nbuf(flags, nb);
        return bytes;
      }
    }
  return 0;
}

int tc_conv_prepare(const ConvParams& p, int ksize, int stride, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan,
                    std::vector<void*>* allocs) {
  if (stride == 2) return tc_s2_prepare(p, w_oihw, sm_count, ptrs_final, plan, allocs);
  {
    const int r2 = tc2_try_prepare(p, ksize, stride, w_oihw, sm_count, ptrs_final, plan, allocs);   // CTA pairs where they apply
    if (r2 != 0) return r2 < 0 ? r2 : B200ROMP_OK;
  }
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) {
    set_error("conv_tc: cuTensorMapEncodeTiled is unavailable");
    return B200ROMP_ECUDA;
  }
  const bool per_tap = false;   // the per-tap TMA variant (PER_TAP=true) was only the bring-up fallback
  const int eb = p.in_dtype == B200ROMP_F32 ? 4 : 2;
  plan->eb = eb;
  { const char* e = getenv("B200ROMP_TC_KSPLIT"); plan->ksplit = (e && e[0] == '4' && eb == 2) ? 0 : 1; }   /* K-split measured slower: off by default */
  const int taps = ksize * ksize;
  const int rowb = tc_row_bytes(ksize, p.cin, eb), cw = rowb / eb, kch = p.cin / cw;
  // N tile: weights must stay resident next to >= 2 pipeline stages
  int nt = (p.cout % 64 == 0) ? 64 : 32;
  const int hh = per_tap ? 16 : 16 + 2 * (ksize / 2), hw = per_tap ? 8 : 8 + 2 * (ksize / 2);
  const int stage_bytes = (hh * hw * rowb + 1023) / 1024 * 1024;
  const int budget = 227 * 1024 - 1024 /*align slack*/ - 1024 /*barriers + bias*/;
  auto bbytes = [&](int n) { return taps * kch * n * rowb; };
  if (bbytes(nt) + 3 * stage_bytes > budget && nt == 64) nt = 32;
  if (bbytes(nt) + 2 * stage_bytes > budget) {
    set_error("conv_tc: k%d cin%d eb%d does not fit shared memory", ksize, p.cin, eb);
    return B200ROMP_EINVAL;
  }
  // TMA epilogue (bf16 tensors only): needs kEpiWarps staging tiles next to >= 2 stages
  int epi_bytes = 0;
  plan->tma_epi = 0;
  if (eb == 2 && tc_epi_prepare(p, nt, ptrs_final, plan)) {
    const int nb = tc_epi_pick_nbuf(plan->tma_epi, nt, budget - bbytes(nt), stage_bytes);
    if (nb == 0) plan->tma_epi = 0;
    else {
      plan->tma_epi = tc_epi_with_nbuf(plan->tma_epi, nb) | (tc_epi_want_coalesced(nt) ? kEpiCoalesced : 0);
      epi_bytes = tc_epi_total_bytes(plan->tma_epi, nt);
    }
  } else if (eb == 4) {
    epi_bytes = tc_epi_prepare_f32(p, nt, ptrs_final, budget - bbytes(nt), stage_bytes, plan);
  }
  int stages = (budget - bbytes(nt) - epi_bytes) / stage_bytes;
  stages = std::min(stages, per_tap ? 12 : 8);   // split into two rings (one per MMA warp)
  plan->kind = ksize * 10 + (per_tap ? 1 : 0);
  plan->cin = p.cin; plan->cout = p.cout; plan->nt = nt; plan->stages = stages;
  plan->grid_y = (p.cout + nt - 1) / nt;
  plan->grid_x = std::max(1, sm_count / plan->grid_y);
  plan->smem_bytes = bbytes(nt) + stages * stage_bytes + epi_bytes + 1024 + 1024;
  int rcw = tc_pack_weights(w_oihw, p.cin, p.cout, taps, nt, &plan->d_wpack, allocs, rowb, eb);
  if (rcw) return rcw;
  // ---- tensor map over the NHWC input: dims (C slice, W, H, N), halo box, OOB -> zeros
  CUtensorMap tm;
  const cuuint64_t gdim[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)p.in_C * eb, (cuuint64_t)p.Win * p.in_C * eb, (cuuint64_t)p.Hin * p.Win * p.in_C * eb};
  const cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)hw, (cuuint32_t)hh, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = const_cast<uint8_t*>(static_cast<const uint8_t*>(p.in) + (size_t)p.in_c_off * eb);
  CUresult cr = encode(&tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled failed with %d", (int)cr);
    return B200ROMP_ECUDA;
  }
  memcpy(plan->tmap_in, &tm, sizeof(tm));
  return dispatch<false>(*plan, p, nullptr, true);
}

int tc_conv_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream) {
  if (plan.kind == 33) return tc_stem_launch(plan, p, stream);
  if (plan.kind == 34) return tc2_launch(plan, p, stream);
  if (plan.kind % 10 == 2) return tc_s2_launch(plan, p, stream);
  return dispatch<false>(plan, p, stream, false);
}

}  // namespace b200romp
