// tcgen05 3x3 stride-1 implicit-GEMM conv on CTA PAIRS (cta_group::2): one tcgen05.mma covers M = 256 pixels - the 16x8
// tiles of the two CTAs of a 2-CTA cluster - so the per-instruction issue cost (the limiter of the N <= 64 layers, see
// DESIGN 4.1) is paid once per two SMs, and each CTA keeps only HALF of the weight slab (its N/2 rows of every B tile),
// which frees shared memory for pipeline stages on the Cin >= 128 layers.
//
// Everything else is conv_tc.cu's design: halo tile by TMA + 9 shifted descriptors, persistent tile loop, two MMA warps
// with private stage rings, TMEM accumulator ring, the TMA epilogue.  Pair protocol (PTX ISA cta_group::2 / CUTLASS sm100):
//   * CTA rank 0 of the cluster is the leader: only its MMA warps issue.  Tile it of CTA r is blockIdx.x + it*gridDim.x,
//     so the pair (2c, 2c+1) always works on two adjacent tiles.
//   * A: each CTA TMA-loads its own halo tile into its own smem with the .cta_group::2 form whose mbarrier operand has
//     the peer bit cleared: both loads complete_tx on the LEADER's full[stage]; the leader's producer posts
//     expect_tx(2 x payload).  Each producer paces itself on its own empty[stage].
//   * B: CTA r holds rows [r*N/2, (r+1)*N/2) of every (tap, chunk) weight tile; the peer tells the leader when its
//     slab has landed by a remote arrive on b_peer.
//   * tcgen05.commit.cta_group::2 with multicast mask 0b11 arrives on empty[stage] / tmem_full[acc] of BOTH CTAs.
//   * D: rows 0-127 of the 256-row accumulator live in the leader's TMEM, rows 128-255 in the peer's, same columns; each
//     CTA's epilogue drains its own TMEM and arrives (remotely for the peer) on the leader's tmem_empty[acc] (count 8).
#include <mutex>

#include "conv_tc.cuh"
#include "tc_device.cuh"

namespace b200romp {

template <int CIN, int NT, int EB>
struct Tc2Cfg {
  static constexpr int KS = 3, TAPS = 9, PAD = 1;
  static constexpr int ROWB = tc_row_bytes(3, CIN, EB);          // bytes per pixel row of a stage (one swizzle span)
  static constexpr int CW = ROWB / EB;                           // channels per K chunk
  static constexpr int KCH = CIN / CW;
  static constexpr int KSTEPS = ROWB / 32;                       // UMMA K steps (32 B) per row
  static constexpr int LAYOUT = ROWB == 128 ? 2 : 4;
  static constexpr int HW_ = 10, HH = 18;
  static constexpr int STAGE_PAYLOAD = HH * HW_ * ROWB;
  static constexpr int STAGE_BYTES = (STAGE_PAYLOAD + 1023) / 1024 * 1024;
  static constexpr int BTILE = (NT / 2) * ROWB;               // this CTA's half of one (tap, chunk) weight tile
  static constexpr int B_BYTES = TAPS * KCH * BTILE;
  static constexpr int ACC = AccCfg<1>::ACC;
  static constexpr int TMEM_COLS = tc_tmem_cols(ACC * NT);
  static constexpr uint32_t IDESC = tc_idesc(EB, 256, NT);
};

template <int CIN, int NT, int EB>
__global__ void __launch_bounds__(tc_threads(EB), 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcEpiMaps epi_maps, const ConvParams p,
                const uint8_t* __restrict__ wpack, int tiles_x, int tiles_y, int num_tiles, int stages, int tma_epi,
                unsigned kmask) {
  using Cfg = Tc2Cfg<CIN, NT, EB>;
  constexpr int kAccStages = Cfg::ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + (Cfg::B_BYTES + 1023) / 1024 * 1024;
  uint8_t* epi_smem = sA + (size_t)stages * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + (EB == 4 ? tc_epi_f32_total_bytes(tma_epi) : tc_epi_total_bytes(tma_epi, NT)));
  uint64_t* empty = full + stages;
  uint64_t* b_full = empty + stages;
  uint64_t* b_peer = b_full + 1;
  uint64_t* tmem_full = b_peer + 1;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint64_t* res_bar = tmem_empty + kAccStages;
  uint64_t* landed = res_bar + 3 * kEpiWarps;        // EB = 4: "TMA tile landed" (local), consumed by the TF32 rounding warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(landed + (EB == 4 ? stages : 0));
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 2) + 15) & ~(uintptr_t)15);   // 16 B: ld.shared.v4

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  if (threadIdx.x == 0) {
    tc_stamp(p, 0);                                 // kernel entry
    for (int i = 0; i < stages; ++i) {
      // EB = 2: the two TMA loads of a pair complete_tx on the leader's full[]; EB = 4: the rounding warps of both
      // CTAs arrive on it once their CTA's tile is converted
      mbar_init(&full[i], EB == 4 ? 2 * kCvtWarps : 1);
      mbar_init(&empty[i], 1);
      if (EB == 4) mbar_init(&landed[i], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_peer, 1);
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);                 // 4 epilogue warps of each CTA of the pair
    }
    for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= kFirstEpiWarp * 32 && threadIdx.x < kFirstEpiWarp * 32 + NT)
    s_bias[threadIdx.x - kFirstEpiWarp * 32] = p.bias[blockIdx.y * NT + threadIdx.x - kFirstEpiWarp * 32];
  if (warp == 1) tmem_alloc_2cta(tmem_ptr, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                               // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_frame = tiles_x * tiles_y;
  const int nrings = tc_num_rings(stages);
  pdl_trigger();
  if (threadIdx.x == 0) tc_stamp(p, 1);             // prologue (barriers, TMEM, cluster sync) done

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, Cfg::B_BYTES);
      const uint8_t* wsrc = wpack + ((size_t)blockIdx.y * 2 + rank) * Cfg::B_BYTES;
      for (int i = 0; i < Cfg::TAPS * Cfg::KCH; ++i)
        bulk_copy_g2s(sB + (size_t)i * Cfg::BTILE, wsrc + (size_t)i * Cfg::BTILE, Cfg::BTILE, b_full);
      pdl_wait();
      tc_stamp(p, 2);                               // predecessor grid complete
      const uint64_t pol = l2_policy_stream(p.debug);
      int stage = 0, stage_other = 0;
      uint32_t phase = 0, phase_other = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_frame, rem = tile % per_frame;
        const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
        const int ring = nrings == 2 ? (it & 1) : 0, rbase = tc_ring_base(stages, ring), rsize = tc_ring_size(stages, ring);
        for (int c = 0; c < Cfg::KCH; ++c) {
          const int sidx = rbase + stage;
          mbar_wait(&empty[sidx], phase ^ 1);
          if (EB == 4) {
            mbar_arrive_expect_tx(&landed[sidx], Cfg::STAGE_PAYLOAD);
            tma_load_4d(sA + (size_t)sidx * Cfg::STAGE_BYTES, &tmap, &landed[sidx], c * Cfg::CW, x0 - 1, y0 - 1, n, pol);
          } else if (p.debug & 4) {                 // profiling experiment: no activation loads
            if (rank == 0) mbar_arrive(&full[sidx]);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full[sidx], 2 * Cfg::STAGE_PAYLOAD);      // own tile + the peer's
            tma_load_4d_2cta(sA + (size_t)sidx * Cfg::STAGE_BYTES, &tmap, &full[sidx], c * Cfg::CW, x0 - 1, y0 - 1, n, pol);
          }
          if (++stage == rsize) { stage = 0; phase ^= 1; }
        }
        if (nrings == 2) { const int ts = stage; stage = stage_other; stage_other = ts; const uint32_t tp = phase; phase = phase_other; phase_other = tp; }
      }
    }
  } else if (warp <= kMmaWarps) {
    if (rank != 0) {
      // peer: report "weight slab resident" to the leader, nothing else to do
      if (warp == 1 && elect_one()) {
        mbar_wait(b_full, 0);
        mbar_arrive_cluster(b_peer, 0);
      }
    } else if (warp <= nrings && elect_one()) {
      // ===================== MMA issuers (leader CTA only) =====================
      mbar_wait(b_full, 0);
      mbar_wait(b_peer, 0);
      tc_fence_after();
      if (warp == 1) tc_stamp(p, 3);                // weight slabs of both CTAs resident
      const uint32_t b_base = smem_u32(sB);
      const int rbase = tc_ring_base(stages, warp - 1), rsize = tc_ring_size(stages, warp - 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = warp - 1;
      for (int tile = blockIdx.x + it * gridDim.x; tile < num_tiles; tile += nrings * gridDim.x, it += nrings) {
        const int acc = it & (kAccStages - 1);
        mbar_wait(&tmem_empty[acc], ((it / kAccStages) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * NT);
        int mma_i = 0;
        for (int c = 0; c < Cfg::KCH; ++c) {
          const int sidx = rbase + stage;
          mbar_wait(&full[sidx], phase);
          tc_fence_after();
          if (warp == 1 && it == 0 && c == 0) tc_stamp(p, 4);   // first activation tile landed
          const uint32_t a_base = smem_u32(sA + (size_t)sidx * Cfg::STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, s = t % 3;
            const uint32_t a_tap = a_base + (uint32_t)((r * Cfg::HW_ + s) * Cfg::ROWB);
            const uint32_t b_tap = b_base + (uint32_t)((t * Cfg::KCH + c) * Cfg::BTILE);
#pragma unroll
            for (int k = 0; k < Cfg::KSTEPS; ++k) {
              // channel half of this K step: which of the two chunks (KCH = 2) or which half of the row (KCH = 1)
              const int half = Cfg::KCH == 2 ? c : (Cfg::KCH == 1 ? k / (Cfg::KSTEPS / 2) : 0);
              if (Cfg::KCH <= 2 && !((kmask >> (t * 2 + half)) & 1u)) continue;   // folded conv: this weight block is all zero
              if (p.debug & 2) continue;                                          // profiling experiment: no MMAs
              const uint64_t adesc = make_smem_desc(a_tap + k * 32, Cfg::HW_ * Cfg::ROWB, Cfg::LAYOUT);
              const uint64_t bdesc = make_smem_desc(b_tap + k * 32, 8 * Cfg::ROWB, Cfg::LAYOUT);
              umma_any<EB, true>(d_tile, adesc, bdesc, Cfg::IDESC, mma_i > 0 ? 1u : 0u);
              ++mma_i;
            }
          }
          umma_commit_2cta(&empty[sidx]);          // both CTAs' stage buffers are reusable once these MMAs retire
          if (++stage == rsize) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tmem_full[acc]);         // both halves of the accumulator complete -> both epilogues
      }
      if (warp == 1) tc_stamp(p, 5);                // last MMA of this ring issued
    }
  } else if (EB == 4 && warp >= kFirstCvtWarp) {
    // ===================== TF32 rounding warps (both CTAs): landed -> round in place -> full (on the leader) =====================
    const int cw = warp - kFirstCvtWarp, lane = threadIdx.x & 31;
    int stage = 0, stage_other = 0;
    uint32_t phase = 0, phase_other = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ring = nrings == 2 ? (it & 1) : 0, rbase = tc_ring_base(stages, ring), rsize = tc_ring_size(stages, ring);
      for (int c = 0; c < Cfg::KCH; ++c) {
        const int sidx = rbase + stage;
        mbar_wait(&landed[sidx], phase);
        tf32_round_smem(sA + (size_t)sidx * Cfg::STAGE_BYTES, Cfg::STAGE_PAYLOAD, cw, lane);
        fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&full[sidx], 0);
        if (++stage == rsize) { stage = 0; phase ^= 1; }
      }
      if (nrings == 2) { const int ts = stage; stage = stage_other; stage_other = ts; const uint32_t tp = phase; phase = phase_other; phase_other = tp; }
    }
  } else if (EB == 4 && tma_epi) {
    tc_epilogue_loop_tma_f32<NT, true>(p, epi_maps, tma_epi, epi_smem, res_bar, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x,
                                       per_frame, num_tiles);
  } else if (EB == 4 || tma_epi == 0) {
    tc_epilogue_loop<NT, 1, true>(p, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame, num_tiles);
  } else {
    if (EB == 2 && NT == 32 && (tma_epi & kEpiCoalesced))
      tc_epilogue_loop_coalesced<NT, true>(p, tma_epi, epi_smem, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x, per_frame, num_tiles);
    else if (EB == 2)
      tc_epilogue_loop_tma<NT, true>(p, epi_maps, tma_epi, epi_smem, res_bar, tmem_base, tmem_full, tmem_empty, s_bias, tiles_x,
                                     per_frame, num_tiles);
  }
  if (threadIdx.x == kFirstEpiWarp * 32) tc_stamp(p, 6);   // first epilogue warp finished its last tile
  tc_fence_before();
  cluster_sync_all();                               // no CTA leaves while its peer may still address its barriers / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    if ((threadIdx.x & 31) == 0) tc_stamp(p, 7);    // kernel exit
  }
}

// ------------------------------------------------------------------------------------------------
// weights: [ntile][half][tap][chunk][NT/2 rows x ROWB], swizzled by the row index inside the half tile; elements are
// bf16 (eb = 2) or fp32 values rounded to TF32 (eb = 4, ties away from zero like cvt.rna.tf32.f32)
float tc_round_tf32_host(float w) {
  uint32_t u;
  memcpy(&u, &w, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static int pack_weights_2cta(const float* w_oihw, int cin, int cout, int nt, int rowb, int eb, void** d_out, std::vector<void*>* allocs) {
  const int cw = rowb / eb, kch = cin / cw, ntiles = cout / nt, hn = nt / 2, per16 = 16 / eb;
  std::vector<uint8_t> img((size_t)ntiles * 2 * 9 * kch * hn * rowb, 0);
  for (int j = 0; j < ntiles; ++j)
    for (int h = 0; h < 2; ++h)
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < kch; ++c) {
          uint8_t* tile = img.data() + ((((size_t)j * 2 + h) * 9 + t) * kch + c) * hn * rowb;
          for (int n = 0; n < hn; ++n)
            for (int k = 0; k < cw; ++k) {
              const int co = j * nt + h * hn + n, ci = c * cw + k;
              const float w = w_oihw[((size_t)co * cin + ci) * 9 + t];
              const int chunk16 = k / per16;
              const int phase = rowb == 128 ? (n & 7) : ((n >> 1) & 3);
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

template <int CIN, int NT, int EB>
static int tc2_inst(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
  auto kern = conv_tc2_kernel<CIN, NT, EB>;
  if (attr) {
    B2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return B200ROMP_OK;
  }
  CUtensorMap tm;
  memcpy(&tm, plan.tmap_in, sizeof(tm));
  TcEpiMaps em;
  memcpy(&em, plan.tmap_epi, sizeof(em));
  const int tiles_x = p.Wout / 8, tiles_y = p.Hout / 16;
  const int num_tiles = tiles_x * tiles_y * p.B;
  dim3 grid(std::min(plan.grid_x, num_tiles) & ~1, plan.grid_y);
  static const bool pdl_default = [] { const char* e = getenv("B200ROMP_NO_PDL"); return !(e && e[0] == '1'); }();
  const bool pdl = pdl_default && g_tc_pdl_override != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(tc_threads(EB));
  cfg.dynamicSmemBytes = (size_t)plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 2 : 1;
  B2R_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm, em, p, reinterpret_cast<const uint8_t*>(plan.d_wpack), tiles_x, tiles_y, num_tiles,
                                 plan.stages, plan.tma_epi, plan.kmask));
  return B200ROMP_OK;
}

static int tc2_dispatch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream, bool attr) {
#define B2R_C2(C, N, E) \
  if (plan.cin == C && plan.nt == N && plan.eb == E) return tc2_inst<C, N, E>(plan, p, stream, attr);
  B2R_C2(32, 32, 2) B2R_C2(64, 64, 2) B2R_C2(128, 64, 2) B2R_C2(256, 64, 2) B2R_C2(256, 32, 2) B2R_C2(32, 64, 2)
  B2R_C2(32, 32, 4) B2R_C2(64, 64, 4) B2R_C2(128, 64, 4) B2R_C2(256, 32, 4) B2R_C2(32, 64, 4)
#undef B2R_C2
  set_error("conv_tc_2cta: no instantiation for cin%d nt%d eb%d", plan.cin, plan.nt, plan.eb);
  return B200ROMP_EINVAL;
}

// Decides whether the CTA-pair engine applies and, if so, fills the plan (kind 34).  Returns 1 = taken, 0 = not
// applicable (caller continues with the single-CTA plan), < 0 = error.
int tc2_try_prepare(const ConvParams& p, int ksize, int stride, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan,
                    std::vector<void*>* allocs) {
  static const bool off = [] { const char* e = getenv("B200ROMP_TC_NO_2CTA"); return e && e[0] == '1'; }();
  const int eb = p.in_dtype == B200ROMP_F32 ? 4 : 2;
  if (off || ksize != 3 || stride != 1 || (eb == 2 && !ptrs_final)) return 0;
  if (p.cin != 32 && p.cin != 64 && p.cin != 128 && p.cin != 256) return 0;
  if (p.cout % 32 != 0) return 0;
  if (eb == 4 && (p.out_nchw || p.pow_channel >= 0)) return 0;          // map outputs stay on the single-CTA kernel
  const int tiles = (p.Wout / 8) * (p.Hout / 16);
  if (tiles % 2 != 0) return 0;                                        // pairs must never split across the tail
  PFN_encodeTiled encode = tc_get_encode();
  if (!encode) return 0;
  const int rowb = tc_row_bytes(3, p.cin, eb), cw = rowb / eb, kch = p.cin / cw;
  int nt = (p.cout % 64 == 0) ? 64 : 32;
  const int stage_bytes = (18 * 10 * rowb + 1023) / 1024 * 1024;
  const int budget = 227 * 1024 - 1024 - 1024;
  auto bbytes = [&](int n) { return (9 * kch * (n / 2) * rowb + 1023) / 1024 * 1024; };
  plan->ksplit = 1;
  plan->eb = eb;
  int epi_bytes = 0;
  if (eb == 2) {
    if (!tc_epi_prepare(p, nt, ptrs_final, plan)) return 0;             // the bf16 pair engine only has the TMA epilogue
    const int nb = tc_epi_pick_nbuf(plan->tma_epi, nt, budget - bbytes(nt), stage_bytes);
    if (nb == 0) return 0;
    plan->tma_epi = tc_epi_with_nbuf(plan->tma_epi, nb) | (tc_epi_want_coalesced(nt) ? kEpiCoalesced : 0);
    epi_bytes = tc_epi_total_bytes(plan->tma_epi, nt);
  } else {
    if (nt == 64 && bbytes(64) + 4 * stage_bytes > budget) nt = 32;     // fp32 weights are twice the bytes: keep >= 4 stages
    if (bbytes(nt) + 2 * stage_bytes > budget) return 0;
    // fp32 tensors: TMA epilogue where shared memory allows (tc_epilogue_loop_tma_f32), else the direct epilogue
    epi_bytes = tc_epi_prepare_f32(p, nt, ptrs_final, budget - bbytes(nt), stage_bytes, plan);
  }
  int stages = std::min(8, (budget - bbytes(nt) - epi_bytes) / stage_bytes);
  plan->kind = 34;
  plan->cin = p.cin; plan->cout = p.cout; plan->nt = nt; plan->stages = stages;
  plan->grid_y = p.cout / nt;
  plan->grid_x = std::max(2, (sm_count / plan->grid_y) & ~1);
  plan->smem_bytes = bbytes(nt) + stages * stage_bytes + epi_bytes + 1024 + 1024;
  int rc = pack_weights_2cta(w_oihw, p.cin, p.cout, nt, rowb, eb, &plan->d_wpack, allocs);
  if (rc) return rc;
  CUtensorMap tm;
  const cuuint64_t gdim[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)p.in_C * eb, (cuuint64_t)p.Win * p.in_C * eb, (cuuint64_t)p.Hin * p.Win * p.in_C * eb};
  const cuuint32_t box[4] = {(cuuint32_t)cw, 10, 18, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const void* base = reinterpret_cast<const uint8_t*>(p.in) + (size_t)p.in_c_off * eb;
  CUresult cr = encode(&tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base),
                       gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv_tc_2cta: cuTensorMapEncodeTiled failed with %d", (int)cr);
    return B200ROMP_ECUDA;
  }
  memcpy(plan->tmap_in, &tm, sizeof(tm));
  rc = tc2_dispatch(*plan, p, nullptr, true);
  return rc ? rc : 1;
}

int tc2_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream) { return tc2_dispatch(plan, p, stream, false); }

}  // namespace b200romp
