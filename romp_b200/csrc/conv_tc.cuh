// tcgen05 implicit-GEMM convolution engine (interface).  Implementation: conv_tc.cu
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

namespace b200romp {

struct TcConvPlan {
  int kind = 0;                 // kernel family, 0 = none
  int cin = 0, cout = 0, nt = 0;  // channels, output channels per CTA
  int eb = 2;                   // operand element bytes: 2 = bf16 (kind::f16), 4 = fp32 storage consumed as TF32 (kind::tf32)
  int grid_x = 0, grid_y = 0, stages = 0;
  int ksplit = 0;               // 1 = single accumulator per tile (bring-up mode), 0 = K-split accumulators
  int smem_bytes = 0;
  void* d_wpack = nullptr;      // weights pre-arranged as the shared-memory image (bf16, swizzled)
  alignas(64) unsigned char tmap_in[128];   // CUtensorMap for the NHWC input tensor
  alignas(64) unsigned char tmap_s2[4][128];  // stride-2: one map per input parity (ph,pw)
  alignas(64) unsigned char tmap_epi[2][128]; // TMA epilogue: output / residual tensor (box NT x 8 x 4 x 1)
  int tma_epi = 0;              // bit 0: output through a TMA store, bit 1: residual through a TMA load
  // CTA-pair engine: bit (tap * 2 + half) = 0 skips the MMAs of that channel half of that tap (all-zero weights of a
  // pixel-pair folded conv, see net.cu fold_pixel_pairs)
  unsigned kmask = 0xFFFFFFFFu;
  const void* encoded_in = nullptr;   // conv1d engine: input pointer / batch the tensor map was encoded for (external inputs)
  int encoded_batch = 0;
  std::string describe() const;
};

// true when (shape, dtypes, flags) can run on the tcgen05 engine
bool tc_conv_supported(const ConvParams& p, int ksize, int stride);
// packs weights, builds tensor maps, picks the tiling; device allocations are appended to `allocs`
// ptrs_final: the output / residual device pointers in `p` are the ones every launch will use (internal tensors)
int tc_conv_prepare(const ConvParams& p, int ksize, int stride, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan,
                    std::vector<void*>* allocs);
int tc_conv_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream);
// CTA-pair engine for 3x3 stride-1 convs (conv_tc_2cta.cu): 1 = plan filled, 0 = not applicable, < 0 = error
int tc2_try_prepare(const ConvParams& p, int ksize, int stride, const float* w_oihw, int sm_count, bool ptrs_final, TcConvPlan* plan,
                    std::vector<void*>* allocs);
int tc2_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream);
// Conv1d (ksize code 13 = 1x3 along W) engine with streamed weights (conv1d_tc.cu): BEV's bird's-eye-view stack
bool tc_conv1d_supported(const ConvParams& p);
int tc_conv1d_prepare(const ConvParams& p, const float* w_oi3, int sm_count, TcConvPlan* plan, std::vector<void*>* allocs);
int tc_conv1d_launch(TcConvPlan& plan, const ConvParams& p, cudaStream_t stream);
// stem engine (conv_stem_tc.cu): 3->64 3x3 stride-2 conv on raw u8 frames with the input normalisation folded in
bool tc_stem_supported(const ConvParams& p, int ksize, int stride);
int tc_stem_prepare(const ConvParams& p, const float* w_oihw, int sm_count, bool out_final, TcConvPlan* plan,
                    std::vector<void*>* allocs);
int tc_stem_launch(const TcConvPlan& plan, const ConvParams& p, cudaStream_t stream);

}  // namespace b200romp
