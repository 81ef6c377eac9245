// Device-side building blocks shared by the tcgen05 conv kernels (conv_tc.cu, conv_tc_s2.cu): PTX wrappers for
// mbarrier / TMA / tcgen05, the shared-memory matrix descriptor, and the TMEM -> global epilogue.
#pragma once
#include <cuda.h>

#include <vector>

#include "conv_tc.cuh"

namespace b200romp {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box - trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b200romp conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Programmatic dependent launch: the next conv of the stream may start its prologue (barrier init, TMEM allocation,
// weight-slab copy - all independent of activations) while this grid drains; pdl_wait() then blocks until every
// predecessor grid has completed and its writes are visible.  Both are no-ops for a launch without the attribute.
// timeline probe (diagnostics): one %globaltimer stamp per (CTA < 4, slot); p.stamps is null in normal runs
__device__ __forceinline__ void tc_stamp(const ConvParams& p, int slot) {
  if (p.stamps != nullptr && blockIdx.x < 4 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.stamps[blockIdx.x * 16 + slot] = t;
  }
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// L2 eviction policy of activation loads (experiment switch).  Idea: the batch-64 tensors are 33-537 MB against 126 MB of
// L2, so a layer's input stream evicts the output it is producing; loads tagged evict_first should leave L2 first and let
// the fresh outputs survive for the next layer.  Measured (per-op profile, B=64): 32->32@128x128 +res 47.3 -> 45.3 us, but
// 1x1 64->256 +res 209 -> 256 us and 64->64@64x64 +res 25.4 -> 27.0 us - net slightly negative, so the default stays
// evict_normal; B200ROMP_TC_DEBUG bit 3 (value 8) turns the evict_first hint on.
__device__ __forceinline__ uint64_t l2_policy_stream(int debug) {
  uint64_t pol;
  if (debug & 8) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
// tensor store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read(int pending) {   // at most `pending` (0..2) store groups may still be reading
  if (pending <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (pending == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cta_group::2 (CTA pair) forms.  Conventions follow the PTX ISA / CUTLASS sm100 wrappers: the even CTA of the
// pair (cluster rank 0) is the leader that issues MMAs; TMA loads of both CTAs credit the LEADER's mbarrier (peer bit of
// the shared::cluster address cleared); commits are multicast to the same barrier offset in both CTAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta_rank) {   // arrive on `bar` of CTA `cta_rank`
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta_rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- element-size generic forms: EB = 2 -> bf16 operands (kind::f16), EB = 4 -> fp32 storage consumed as TF32 (kind::tf32).
// All shared-memory geometry of the conv kernels is expressed in BYTES (64 / 128 B swizzle rows, 32 B per UMMA K step =
// 16 bf16 or 8 tf32 elements), so the TF32 engine is the same pipeline with twice the bytes per channel.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int EB, bool CTA2>
__device__ __forceinline__ void umma_any(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (EB == 2) { if (CTA2) umma_bf16_2cta(d_tmem, adesc, bdesc, idesc, accumulate); else umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate); }
  else         { if (CTA2) umma_tf32_2cta(d_tmem, adesc, bdesc, idesc, accumulate); else umma_tf32(d_tmem, adesc, bdesc, idesc, accumulate); }
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 | a/b format (1 = BF16, 2 = TF32) | N >> 3 | M >> 4
__host__ __device__ constexpr uint32_t tc_idesc(int eb, int m, int n) {
  return (1u << 4) | ((eb == 2 ? 1u : 2u) << 7) | ((eb == 2 ? 1u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// bytes of one shared-memory pixel row (= swizzle span) of a conv's A / B tiles.  The weight-heavy 3x3 layers
// (Cin >= 128) use half rows: more, smaller pipeline stages next to the resident weight slab.
__host__ __device__ constexpr int tc_row_bytes(int ksize, int cin, int eb) {
  return (ksize == 3 && cin >= 128) ? 64 : (cin * eb < 128 ? cin * eb : 128);
}

// ---- TF32 conversion stage (EB = 4 kernels) -------------------------------------------------------
// tcgen05.mma kind::tf32 reads fp32 containers and IGNORES the low 13 mantissa bits (truncation).  Measured on the CPU
// oracle: truncating the activations of the ~100-layer stack biases the maps (mean |err| 1.2e-3 vs 4.8e-4
// for round-to-nearest), and rounding when a tensor is STORED (so that the truncation becomes exact) doubles the error
// because the fp32 skip connections get rounded too.  So the TF32 engine does what a cvt.rna.tf32.f32 in front of
// mma.sync does in the reference's cuDNN/cuBLAS TF32 kernels: kCvtWarps extra warps round every landed A tile in place
// (shared memory, ties away from zero) between the TMA completion and the MMA issue.  Tensors in HBM stay exact fp32.
constexpr int kCvtWarps = 4;
__device__ __forceinline__ uint32_t tf32_rna(uint32_t x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void tf32_round_smem(uint8_t* base, int bytes, int cvt_warp, int lane) {
  const uint32_t s0 = smem_u32(base);
  for (int off = (cvt_warp * 32 + lane) * 16; off < bytes; off += kCvtWarps * 32 * 16) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(s0 + off));
    a = tf32_rna(a); b = tf32_rna(b); c = tf32_rna(c); d = tf32_rna(d);
    asm volatile("st.shared.v4.b32 [%4], {%0, %1, %2, %3};" ::"r"(a), "r"(b), "r"(c), "r"(d), "r"(s0 + off) : "memory");
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, swizzled (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 |
//   [49,52) base offset = 0 (pattern anchored at 1024 B) | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}

// 32 consecutive output channels of one full-resolution pixel: residual add, ReLU, dtype conversion, all with
// 16-byte vector accesses; every load is issued before the first use so one thread keeps 4-8 requests in flight.
__device__ __forceinline__ void tc_store32(const ConvParams& p, size_t pix, size_t rpix, int co, const float (&acc)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = acc[j];
  if (p.res != nullptr) {
    if (p.res_dtype == B200ROMP_BF16) {
      const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + rpix * p.res_C + p.res_c_off + co);
      uint4 t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = r[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[i * 8 + 2 * k] += __low2float(h[k]);
          v[i * 8 + 2 * k + 1] += __high2float(h[k]);
        }
      }
    } else {
      const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + rpix * p.res_C + p.res_c_off + co);
      float4 t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = r[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i * 4 + 0] += t[i].x; v[i * 4 + 1] += t[i].y; v[i * 4 + 2] += t[i].z; v[i * 4 + 3] += t[i].w;
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.out_dtype == B200ROMP_BF16) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_C + p.out_c_off + co);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i * 8 + 0], v[i * 8 + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i * 8 + 2], v[i * 8 + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i * 8 + 4], v[i * 8 + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[i * 8 + 6], v[i * 8 + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      o[i] = pk;
    }
  } else {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_C + p.out_c_off + co);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = make_float4(v[i * 4 + 0], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
  }
}

// TMEM accumulator ring: ACC tiles in flight per CTA (8 x NT <= 512 columns for NT <= 64).  The ring depth bounds the
// throughput by (tiles in flight) / (MMA -> commit -> epilogue wake-up -> tcgen05.ld -> tmem_empty arrive -> MMA wake-up
// round trip); with 4 stages the 32->32@128x128 pair kernel was measured waiting on exactly this loop (MMA thread spinning
// on tmem_empty while the epilogue spun on tmem_full).  (KSPLIT > 1 = the abandoned K-split experiment, 2 tiles.)
template <int KSPLIT>
struct AccCfg {
  static constexpr int ACC = KSPLIT == 1 ? 8 : 2;
};
// accumulators per tile: as many as TMEM allows (2 tiles x KSPLIT x NT <= 512 columns), never more than the MMAs of a tile
__host__ __device__ constexpr int tc_ksplit(int mmas_per_tile, int nt) {
  return (nt == 32 ? 8 : 4) < mmas_per_tile ? (nt == 32 ? 8 : 4) : (mmas_per_tile >= 4 ? 4 : (mmas_per_tile >= 2 ? 2 : 1));
}
__host__ __device__ constexpr int tc_tmem_cols(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }
constexpr int kEpiWarps = 8;         // two groups of 4 warps, alternating tiles
constexpr int kMmaWarps = 2;         // two MMA-issuing warps alternate tiles: one thread sustains only ~1 tcgen05.mma / 50-75 clk
constexpr int kFirstEpiWarp = 1 + kMmaWarps;
// the smem stages are split into one private ring per MMA warp: ring 0 gets the larger half
// (a single stage cannot be split: then only the first MMA warp works and owns it)
// With fewer than 4 stages (weight-heavy layers: Cin >= 128) a split would leave one stage per ring, i.e. no overlap of a
// ring's TMA load with its MMAs (measured: 128->128@32x32 29 us); then one MMA warp owns all stages.
__host__ __device__ constexpr int tc_chunk_width(int ksize, int cin) { return (ksize == 3 && cin >= 128) ? 32 : (cin < 64 ? cin : 64); }
__host__ __device__ constexpr int tc_num_rings(int stages) { return stages >= 4 ? kMmaWarps : 1; }
__host__ __device__ constexpr int tc_ring_size(int stages, int ring) { return tc_num_rings(stages) == 1 ? stages : (stages + 1 - ring) / 2; }
__host__ __device__ constexpr int tc_ring_base(int stages, int ring) { return ring ? (stages + 1) / 2 : 0; }
constexpr int kTcThreads = (kFirstEpiWarp + kEpiWarps) * 32;
constexpr int kFirstCvtWarp = kFirstEpiWarp + kEpiWarps;                 // EB = 4 kernels only: TF32 rounding warps
__host__ __device__ constexpr int tc_threads(int eb) { return kTcThreads + (eb == 4 ? kCvtWarps * 32 : 0); }

// Epilogue of a persistent tile loop: 2 groups x 4 warps (warps 2..9), group g takes the CTA's tiles g, g+2, ...
// Each warp owns the TMEM lane quarter (warp id mod 4); thread = one pixel of the 16x8 tile.
template <int NT, int KSPLIT, bool CTA2 = false>
__device__ __forceinline__ void tc_epilogue_loop(const ConvParams& p, uint32_t tmem_base, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, const float* s_bias, int tiles_x, int per_frame,
                                                 int num_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = (warp - kFirstEpiWarp) >> 2;
  const int q = warp & 3;
  const int m = q * 32 + lane;
  const int co0 = blockIdx.y * NT;
  const int up = p.up;
  const int Hf = p.Hout * up, Wf = p.Wout * up;
  pdl_wait();                                 // residual reads / output writes must follow the predecessor grids
  int it = group;
  for (int tile = blockIdx.x + group * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
    constexpr int ACC = AccCfg<KSPLIT>::ACC;
    const int acc = it & (ACC - 1);
    const int n = tile / per_frame, rem = tile % per_frame;
    const int oy = (rem / tiles_x) * 16 + (m >> 3), ox = (rem % tiles_x) * 8 + (m & 7);
    mbar_wait(&tmem_full[acc], (it / ACC) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * KSPLIT * NT);
#pragma unroll
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = s_bias[c0 + j];
#pragma unroll
      for (int ks = 0; ks < KSPLIT; ++ks) {      // partial sums of the K-split accumulators
        uint32_t r[32];
        tmem_ld32(taddr + ks * NT + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
      }
      if (p.out_nchw) {
        // map outputs (ROMP head: [B,C,H,W] fp32, main.py:112-113): per channel a warp writes 4 x 32 B row segments
        float* o = reinterpret_cast<float*>(p.out) + (((size_t)n * p.out_C + p.out_c_off + co0 + c0) * p.Hout + oy) * p.Wout + ox;
        const size_t cstride = (size_t)p.Hout * p.Wout;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = co0 + c0 + j;
          if (c < p.cout) {
            float x = p.relu ? fmaxf(v[j], 0.f) : v[j];
            if (c == p.pow_channel) x = powf(1.1f, x);
            o[j * cstride] = x;
          }
        }
        continue;
      }
      for (int dy = 0; dy < up; ++dy) {
        for (int dx = 0; dx < up; ++dx) {
          const int fy = oy * up + dy, fx = ox * up + dx;
          const size_t pix = ((size_t)n * Hf + fy) * Wf + fx;
          const size_t rpix = ((size_t)(p.res_broadcast ? 0 : n) * Hf + fy) * Wf + fx;
          if (!(p.debug & 1)) tc_store32(p, pix, rpix, co0 + c0, v);
          else if (v[0] == 123.456f) reinterpret_cast<float*>(p.out)[0] = v[1];   // keep the TMEM loads alive
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CTA2) mbar_arrive_cluster(&tmem_empty[acc], 0);     // the pair's leader issues the MMAs of both CTAs
      else mbar_arrive(&tmem_empty[acc]);
    }
  }
}

// ---- TMA epilogue ---------------------------------------------------------------------------------
// The direct epilogue above has thread = pixel, so one warp-wide 16 B access touches 32 different 128 B lines (32 L1
// wavefronts); measured, that traffic costs as much as the MMAs of a tile.  For the common case (bf16 NHWC output at the
// conv resolution, optional bf16 residual of the same shape) each epilogue warp instead owns an NT x 32-pixel staging
// tile in shared memory (TMA swizzle pattern, conflict-free for row-per-thread 16 B accesses): the residual box arrives
// by a TMA tensor load issued before the accumulator is awaited, every thread finishes its pixel in place, and one
// TMA tensor store writes the box (4 rows x 8 pixels x NT channels) with full-line transactions.
struct TcEpiMaps {
  CUtensorMap out, res;
};
constexpr int kTmaEpiOut = 1, kTmaEpiRes = 2, kTmaEpiBufShift = 2;   // bits 2-3: staging tiles per warp - 1
__host__ __device__ constexpr int tc_epi_nbuf(int tma_epi) { return 1 + ((tma_epi >> kTmaEpiBufShift) & 3); }
__host__ __device__ constexpr int tc_epi_with_nbuf(int tma_epi, int nbuf) { return (tma_epi & 3) | ((nbuf - 1) << kTmaEpiBufShift); }
__host__ __device__ constexpr int tc_epi_stage_bytes(int nt) { return 32 * nt * 2; }   // per epilogue warp and buffer
__host__ __device__ constexpr int tc_epi_total_bytes(int tma_epi, int nt) {
  return tma_epi ? kEpiWarps * tc_epi_nbuf(tma_epi) * tc_epi_stage_bytes(nt) : 0;
}

template <int NT>
__device__ __forceinline__ uint4* tc_epi_chunk(uint8_t* stg, int row, int chunk) {
  // 16 B chunk `chunk` of pixel row `row`: SWIZZLE_128B (NT = 64, 128 B rows) or SWIZZLE_64B (NT = 32, 64 B rows)
  const int sw = NT == 64 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
  return reinterpret_cast<uint4*>(stg + row * (NT * 2) + sw * 16);
}

template <int NT, bool CTA2 = false>
__device__ __forceinline__ void tc_epilogue_loop_tma(const ConvParams& p, const TcEpiMaps& maps, int tma_epi, uint8_t* epi_smem,
                                                     uint64_t* res_bar, uint32_t tmem_base, uint64_t* tmem_full,
                                                     uint64_t* tmem_empty, const float* s_bias, int tiles_x, int per_frame,
                                                     int num_tiles) {
  static_assert(NT == 32 || NT == 64, "staging layout");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - kFirstEpiWarp;
  const int group = ew >> 2;
  const int q = warp & 3;
  const int co0 = blockIdx.y * NT;
  const bool dbg_quiet = (p.debug & 1) != 0;   // B200ROMP_TC_DEBUG bit 0: epilogue without global traffic (profiling experiments)
  const bool has_res = (tma_epi & kTmaEpiRes) != 0 && !dbg_quiet;
  // staging: nbuf (1..3) tiles per warp, used round-robin.  A tile may only be rewritten once the TMA store that last read
  // it has finished reading; with a single tile that wait (TMA queue latency, ~1 us behind the producer's loads) sits
  // on every tile's critical path - measured as the limiter of the 32->32@128x128 layers.  Without a residual, nbuf = 2
  // lets one store stay in flight; with a residual the NEXT tile's residual is prefetched one tile ahead into the tile
  // after the current one, so nbuf = 3 keeps both the prefetch and one store off the critical path.
  const int nbuf = tc_epi_nbuf(tma_epi);
  uint8_t* stg0 = epi_smem + ew * nbuf * tc_epi_stage_bytes(NT);
  uint64_t* rbar = &res_bar[ew * 3];
  constexpr int ACC = AccCfg<1>::ACC;
  pdl_wait();                                 // residual reads / output writes must follow the predecessor grids
  const uint64_t res_pol = l2_policy_stream(p.debug);
  auto load_res = [&](int tile, int buf) {    // lane 0 only
    const int n = tile / per_frame, rem = tile % per_frame;
    mbar_arrive_expect_tx(&rbar[buf], tc_epi_stage_bytes(NT));
    tma_load_4d(stg0 + buf * tc_epi_stage_bytes(NT), &maps.res, &rbar[buf], p.res_c_off + co0, (rem % tiles_x) * 8,
                (rem / tiles_x) * 16 + q * 4, n, res_pol);
  };
  const int first_tile = blockIdx.x + group * gridDim.x;
  if (nbuf >= 2 && has_res && lane == 0 && first_tile < num_tiles) load_res(first_tile, 0);
  const bool relu = p.relu != 0;
  const uint32_t bias_saddr = smem_u32(s_bias);
  // (frame, tile-in-frame) advance incrementally: two integer divisions per tile were ~25 % of this loop's instructions
  const int tstep = 2 * (int)gridDim.x, step_n = tstep / per_frame, step_rem = tstep % per_frame;
  const int tx_shift = (tiles_x & (tiles_x - 1)) == 0 ? __ffs(tiles_x) - 1 : -1;
  int n = first_tile / per_frame, rem = first_tile % per_frame;
  int it = group, t_local = 0, buf = 0;
  uint32_t rphase = 0;
  for (int tile = first_tile; tile < num_tiles; tile += tstep, it += 2, ++t_local) {
    const int acc = it & (ACC - 1);
    const int ty = tx_shift >= 0 ? (rem >> tx_shift) : rem / tiles_x, tx = rem - ty * tiles_x;
    const int y0 = ty * 16 + q * 4, x0 = tx * 8;                               // this warp's 4 x 8 pixel box
    uint8_t* stg = stg0 + buf * tc_epi_stage_bytes(NT);
    const int buf_next = buf + 1 == nbuf ? 0 : buf + 1;
    if (lane == 0) {
      if (has_res) {
        if (nbuf == 1) {
          bulk_wait_read(0);
          load_res(tile, 0);
        } else {
          bulk_wait_read(nbuf - 2);             // the store that last read staging tile buf_next is done
          if (tile + 2 * (int)gridDim.x < num_tiles) load_res(tile + 2 * gridDim.x, buf_next);
        }
      } else {
        bulk_wait_read(nbuf - 1);               // the store that last read staging tile buf is done
      }
    }
    __syncwarp();
    mbar_wait(&tmem_full[acc], (it / ACC) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
    if (has_res) mbar_wait(&rbar[buf], rphase);
#pragma unroll
    for (int c0 = 0; c0 < NT; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
      if (c0 + 32 == NT) tc_fence_before();   // accumulator fully read (released below, after the store is issued)
      float v[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {        // bias: 8 x ld.shared.v4 (the generic-pointer form compiled to slow generic loads)
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_saddr + (c0 + 4 * j4) * 4));
        v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + b.x;
        v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + b.y;
        v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + b.z;
        v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + b.w;
      }
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 t = *tc_epi_chunk<NT>(stg, lane, c0 / 8 + i);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[i * 8 + 2 * k] += __low2float(h[k]);
            v[i * 8 + 2 * k + 1] += __high2float(h[k]);
          }
        }
      }
      // round to bf16 first, ReLU on the packed pairs: max(round(x), 0) == round(max(x, 0)) (rounding is monotonic, 0 exact)
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i * 8 + 0], v[i * 8 + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i * 8 + 2], v[i * 8 + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i * 8 + 4], v[i * 8 + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[i * 8 + 6], v[i * 8 + 7]);
        if (relu) { h0 = __hmax2(h0, zero2); h1 = __hmax2(h1, zero2); h2 = __hmax2(h2, zero2); h3 = __hmax2(h3, zero2); }
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *tc_epi_chunk<NT>(stg, lane, c0 / 8 + i) = pk;
      }
    }
    fence_proxy_async();                      // generic-proxy writes -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      if (!dbg_quiet) {
        tma_store_4d(&maps.out, stg, p.out_c_off + co0, x0, y0, n);
        bulk_commit_group();
      }
      // Hand the TMEM stage back.  Done after the proxy fence on purpose: fence.proxy.async compiles to MEMBAR.ALL.CTA,
      // which would otherwise wait for this (for the pair's peer CTA: remote) arrive to be performed on every tile.
      if (CTA2) mbar_arrive_cluster(&tmem_empty[acc], 0);     // the pair's leader issues the MMAs of both CTAs
      else mbar_arrive(&tmem_empty[acc]);
    }
    if (buf_next == 0) rphase ^= 1;           // every staging tile (and its residual barrier) was used once more
    buf = buf_next;
    n += step_n; rem += step_rem;
    if (rem >= per_frame) { rem -= per_frame; ++n; }
  }
  if (lane == 0) bulk_wait0();                // all stores performed before the CTA's shared memory goes away
  __syncwarp();
}

// ---- TMA epilogue for fp32 tensors (TF32 engine) ---------------------------------------------------------
// Same idea as tc_epilogue_loop_tma, fp32 elements: the unit of work is one 32-channel CHUNK of one tile (32 pixels x
// 32 channels x 4 B = 4 KB per epilogue warp, 128 B rows in the SWIZZLE_128B pattern, conflict-free for row-per-thread
// 16 B accesses).  A warp walks the linear sequence of its chunks (tile-major) through `nbuf` staging buffers:
//   chunk j: [lane 0] wait until the store that last read buffer (j+1) % nbuf is done, prefetch the residual box of chunk
//            j+1 into it (TMA load);  wait for the accumulator (first chunk of a tile) and for this chunk's residual;
//            tcgen05.ld 32 columns -> + bias (+ residual) -> ReLU -> fp32 in place;  fence, one TMA tensor store.
// Why: the direct fp32 epilogue (thread = pixel, 8 x 16 B stores per 128 B line, 32 lines per warp-wide access) costs
// ~8x its payload in L1/shared-memory wavefronts, and these kernels are shared-memory-bandwidth bound.
constexpr int kF32ChunkBytes = 32 * 32 * 4;
__host__ __device__ constexpr int tc_epi_f32_total_bytes(int tma_epi) { return tma_epi ? kEpiWarps * tc_epi_nbuf(tma_epi) * kF32ChunkBytes : 0; }

template <int NT, bool CTA2 = false>
__device__ __forceinline__ void tc_epilogue_loop_tma_f32(const ConvParams& p, const TcEpiMaps& maps, int tma_epi, uint8_t* epi_smem,
                                                         uint64_t* res_bar, uint32_t tmem_base, uint64_t* tmem_full,
                                                         uint64_t* tmem_empty, const float* s_bias, int tiles_x, int per_frame,
                                                         int num_tiles) {
  static_assert(NT % 32 == 0, "chunks of 32 channels");
  constexpr int NCH = NT / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - kFirstEpiWarp;
  const int group = ew >> 2;
  const int q = warp & 3;
  const int co0 = blockIdx.y * NT;
  const bool has_res = (tma_epi & kTmaEpiRes) != 0;
  const int nbuf = tc_epi_nbuf(tma_epi);
  uint8_t* stg0 = epi_smem + ew * nbuf * kF32ChunkBytes;
  uint64_t* rbar = &res_bar[ew * 3];
  constexpr int ACC = AccCfg<1>::ACC;
  pdl_wait();
  const uint64_t res_pol = l2_policy_stream(p.debug);
  const bool relu = p.relu != 0;
  const uint32_t bias_saddr = smem_u32(s_bias);
  const int tstep = 2 * (int)gridDim.x;
  const int first_tile = blockIdx.x + group * gridDim.x;
  // coordinates of chunk j of this warp: tile = first_tile + (j / NCH) * tstep, channels (j % NCH) * 32
  auto load_res = [&](int tile, int c, int buf) {   // lane 0 only
    const int n = tile / per_frame, rem = tile % per_frame;
    mbar_arrive_expect_tx(&rbar[buf], kF32ChunkBytes);
    tma_load_4d(stg0 + buf * kF32ChunkBytes, &maps.res, &rbar[buf], p.res_c_off + co0 + c * 32, (rem % tiles_x) * 8,
                (rem / tiles_x) * 16 + q * 4, n, res_pol);
  };
  if (has_res && lane == 0 && first_tile < num_tiles) load_res(first_tile, 0, 0);
  int it = group, buf = 0;
  uint32_t rphase = 0;                          // parity of the residual barriers: flips each time the buffer ring wraps
  for (int tile = first_tile; tile < num_tiles; tile += tstep, it += 2) {
    const int acc = it & (ACC - 1);
    const int n = tile / per_frame, rem = tile % per_frame;
    const int y0 = (rem / tiles_x) * 16 + q * 4, x0 = (rem % tiles_x) * 8;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint8_t* stg = stg0 + buf * kF32ChunkBytes;
      const int buf_next = buf + 1 == nbuf ? 0 : buf + 1;
      if (lane == 0) {
        if (has_res) {
          if (nbuf == 1) {
            if (!(tile == first_tile && c == 0)) { bulk_wait_read(0); load_res(tile, c, 0); }
          } else {
            bulk_wait_read(nbuf - 2);             // the store that last read buffer buf_next is done
            const int nt_tile = c + 1 < NCH ? tile : tile + tstep, nc = c + 1 < NCH ? c + 1 : 0;
            if (nt_tile < num_tiles) load_res(nt_tile, nc, buf_next);
          }
        } else {
          bulk_wait_read(nbuf - 1);               // the store that last read this buffer is done
        }
      }
      __syncwarp();
      if (c == 0) {
        mbar_wait(&tmem_full[acc], (it / ACC) & 1);
        tc_fence_after();
      }
      if (has_res) mbar_wait(&rbar[buf], rphase);
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      if (c + 1 == NCH) tc_fence_before();        // accumulator fully read (released below)
#pragma unroll
      for (int i = 0; i < 8; ++i) {               // 8 x 16 B = the 32 channels of this thread's pixel
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_saddr + (c * 32 + 4 * i) * 4));
        float4 v = make_float4(__uint_as_float(r[4 * i + 0]) + b.x, __uint_as_float(r[4 * i + 1]) + b.y,
                               __uint_as_float(r[4 * i + 2]) + b.z, __uint_as_float(r[4 * i + 3]) + b.w);
        float4* slot = reinterpret_cast<float4*>(stg + lane * 128 + ((i ^ (lane & 7)) * 16));
        if (has_res) {
          const float4 t = *slot;
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *slot = v;
      }
      fence_proxy_async();                        // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&maps.out, stg, p.out_c_off + co0 + c * 32, x0, y0, n);
        bulk_commit_group();
        if (c + 1 == NCH) {
          if (CTA2) mbar_arrive_cluster(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (buf_next == 0) rphase ^= 1;
      buf = buf_next;
    }
  }
  if (lane == 0) bulk_wait0();                    // all stores performed before the CTA's shared memory goes away
  __syncwarp();
}

// launch with the programmatic-stream-serialization attribute (see pdl_trigger / pdl_wait); B200ROMP_NO_PDL=1 disables it
template <typename... KArgs, typename... Args>
static inline cudaError_t tc_launch(void (*kern)(KArgs...), dim3 grid, int threads, int smem_bytes, cudaStream_t stream, Args&&... args) {
  static const bool pdl_default = [] { const char* e = getenv("B200ROMP_NO_PDL"); return !(e && e[0] == '1'); }();
  const bool pdl = pdl_default && g_tc_pdl_override != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- coalesced direct epilogue ----------------------------------------------------------------------
// Same in-place staging tile as the TMA epilogue, but the global traffic is done by the warp itself: the staging tile is
// read back "transposed" (NT/8 consecutive lanes = the NT channels of one pixel, so one warp-wide 16 B access covers
// 512 contiguous bytes when the tensor has NT channels) and written with plain st.global.v4; the residual comes in by
// the mirror-image ld.global.v4, software-pipelined one tile ahead in registers.  Written to test whether the TMA unit's
// per-row cost (64 B box rows for NT = 32) is what makes the 32->32@128x128 layers slow: it is not - this path measures
// 46.5 us against 41.0 / 46.9 us (without / with residual) for the TMA epilogue - so it stays an experiment switch
// (B200ROMP_EPI_COALESCED=1); parity-tested like the default path.
constexpr int kEpiCoalesced = 16;   // bit in plan->tma_epi: use tc_epilogue_loop_coalesced instead of the TMA epilogue
template <int NT, bool CTA2 = false>
__device__ __forceinline__ void tc_epilogue_loop_coalesced(const ConvParams& p, int tma_epi, uint8_t* epi_smem, uint32_t tmem_base,
                                                           uint64_t* tmem_full, uint64_t* tmem_empty, const float* s_bias,
                                                           int tiles_x, int per_frame, int num_tiles) {
  static_assert(NT == 32 || NT == 64, "staging layout");
  constexpr int CH = NT / 8;                  // 16 B chunks per pixel = lanes per pixel in the transposed pattern
  constexpr int PPI = 32 / CH;                // pixels per warp-wide access
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - kFirstEpiWarp;
  const int group = ew >> 2;
  const int q = warp & 3;
  const int co0 = blockIdx.y * NT;
  const bool dbg_quiet = (p.debug & 1) != 0;
  const bool has_res = p.res != nullptr && !dbg_quiet;
  const bool relu = p.relu != 0;
  uint8_t* stg = epi_smem + ew * tc_epi_nbuf(tma_epi) * tc_epi_stage_bytes(NT);
  const uint32_t bias_saddr = smem_u32(s_bias);
  constexpr int ACC = AccCfg<1>::ACC;
  const int tp = lane / CH, tc = lane % CH;   // transposed role of this lane: pixel-in-group, chunk
  pdl_wait();
  const int tstep = 2 * (int)gridDim.x, step_n = tstep / per_frame, step_rem = tstep % per_frame;
  const int tx_shift = (tiles_x & (tiles_x - 1)) == 0 ? __ffs(tiles_x) - 1 : -1;
  const int first_tile = blockIdx.x + group * gridDim.x;
  int n = first_tile / per_frame, rem = first_tile % per_frame;
  // element offset (in 16 B units) of transposed access i of the tile whose box starts at (n, y0, x0)
  auto goff = [&](const int C_total, const int c_off, int nn, int y0, int x0, int i) -> size_t {
    const int pp = i * PPI + tp;              // pixel 0..31 of the warp's 4 x 8 box
    const size_t pix = ((size_t)nn * p.Hout + y0 + (pp >> 3)) * p.Wout + x0 + (pp & 7);
    return (pix * C_total + c_off + co0) / 8 + tc;
  };
  auto box_of = [&](int r, int& y0, int& x0) {
    const int ty = tx_shift >= 0 ? (r >> tx_shift) : r / tiles_x, tx = r - ty * tiles_x;
    y0 = ty * 16 + q * 4; x0 = tx * 8;
  };
  const uint4* res16 = reinterpret_cast<const uint4*>(p.res);
  uint4* out16 = reinterpret_cast<uint4*>(p.out);
  uint4 rnext[CH];
  if (has_res && first_tile < num_tiles) {
    int y0, x0;
    box_of(rem, y0, x0);
#pragma unroll
    for (int i = 0; i < CH; ++i) rnext[i] = res16[goff(p.res_C, p.res_c_off, n, y0, x0, i)];
  }
  int it = group;
  for (int tile = first_tile; tile < num_tiles; tile += tstep, it += 2) {
    const int acc = it & (ACC - 1);
    int y0, x0;
    box_of(rem, y0, x0);
    int n2 = n + step_n, rem2 = rem + step_rem;
    if (rem2 >= per_frame) { rem2 -= per_frame; ++n2; }
    if (has_res) {
      // this tile's residual (fetched during the previous tile) -> staging; then prefetch the next tile's
#pragma unroll
      for (int i = 0; i < CH; ++i) *tc_epi_chunk<NT>(stg, i * PPI + tp, tc) = rnext[i];
      if (tile + tstep < num_tiles) {
        int y1, x1;
        box_of(rem2, y1, x1);
#pragma unroll
        for (int i = 0; i < CH; ++i) rnext[i] = res16[goff(p.res_C, p.res_c_off, n2, y1, x1, i)];
      }
      __syncwarp();
    }
    mbar_wait(&tmem_full[acc], (it / ACC) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
#pragma unroll
    for (int c0 = 0; c0 < NT; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
      if (c0 + 32 == NT) {                    // accumulator fully read: hand the TMEM stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      float v[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_saddr + (c0 + 4 * j4) * 4));
        v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + b.x;
        v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + b.y;
        v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + b.z;
        v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + b.w;
      }
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 t = *tc_epi_chunk<NT>(stg, lane, c0 / 8 + i);
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[i * 8 + 2 * k] += __uint_as_float(w[k] << 16);
            v[i * 8 + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
          }
        }
      }
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i * 8 + 0], v[i * 8 + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i * 8 + 2], v[i * 8 + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i * 8 + 4], v[i * 8 + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[i * 8 + 6], v[i * 8 + 7]);
        if (relu) { h0 = __hmax2(h0, zero2); h1 = __hmax2(h1, zero2); h2 = __hmax2(h2, zero2); h3 = __hmax2(h3, zero2); }
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *tc_epi_chunk<NT>(stg, lane, c0 / 8 + i) = pk;
      }
    }
    __syncwarp();
    if (!dbg_quiet) {
#pragma unroll
      for (int i = 0; i < CH; ++i) out16[goff(p.out_C, p.out_c_off, n, y0, x0, i)] = *tc_epi_chunk<NT>(stg, i * PPI + tp, tc);
    }
    __syncwarp();                             // staging tile is rewritten by the next tile's residual / outputs
    n = n2; rem = rem2;
  }
}

// staging tiles per epilogue warp for a plan with `avail` bytes left for staging + pipeline stages: as many as useful
// (3 with a residual, 2 without) while keeping >= 6 stages, else >= 4, else >= 2; 0 = the TMA epilogue does not fit
// B200ROMP_EPI_COALESCED=1 selects the coalesced direct epilogue instead of the TMA epilogue
inline bool tc_epi_want_coalesced(int nt) {
  static const int mode = [] { const char* e = getenv("B200ROMP_EPI_COALESCED"); return e ? atoi(e) : -1; }();
  return mode > 0 && nt == 32;   // measured equal to (NT = 32) or slower than the TMA epilogue: kept as an experiment switch
}
inline int tc_epi_pick_nbuf(int tma_epi, int nt, int avail, int stage_bytes) {
  if (tc_epi_want_coalesced(nt)) return (avail - tc_epi_total_bytes(tc_epi_with_nbuf(tma_epi, 1), nt)) / stage_bytes >= 2 ? 1 : 0;
  const int maxb = 3;
  const int wants[3] = {6, 4, 2};
  for (int w = 0; w < 3; ++w)
    for (int nb = maxb; nb >= 1; --nb)
      if ((avail - tc_epi_total_bytes(tc_epi_with_nbuf(tma_epi, nb), nt)) / stage_bytes >= wants[w]) return nb;
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled tc_get_encode();
// decides whether the TMA epilogue applies (sets plan->tma_epi and the out/res tensor maps); 0 = direct epilogue
int tc_epi_prepare(const ConvParams& p, int nt, bool ptrs_final, TcConvPlan* plan);
// fp32 variant (TF32 engine): maps with box (32 ch, 8, 4, 1) fp32 SWIZZLE_128B; picks the staging depth for `avail` bytes next to
// pipeline stages of `stage_bytes`; returns the staging bytes (0 = direct epilogue)
int tc_epi_prepare_f32(const ConvParams& p, int nt, bool ptrs_final, int avail, int stage_bytes, TcConvPlan* plan);
// bf16 weight slab in shared-memory-image order [ntile][tap][chunk][NT rows x ROWB] with the TMA/UMMA XOR swizzle
int tc_pack_weights(const float* w_oihw, int cin, int cout, int taps, int nt, void** d_out, std::vector<void*>* allocs, int rowb, int eb);
float tc_round_tf32_host(float w);   // fp32 -> TF32, ties away from zero (cvt.rna.tf32.f32)
// stride-2 3x3 engine (conv_tc_s2.cu)
bool tc_s2_supported(const ConvParams& p);
int tc_s2_prepare(const ConvParams& p, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan, std::vector<void*>* allocs);
int tc_s2_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream);

}  // namespace b200romp
