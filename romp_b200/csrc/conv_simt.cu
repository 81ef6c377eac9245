// Generic fused direct convolution on CUDA cores (fp32 accumulate).
//
// Role: (1) the fp32 "parity" engine - bit-for-bit the reference's arithmetic type (model.py convs run
// fp32), used to prove graph/BN-fold/fuse semantics against the oracle; (2) the engine for the few
// layers whose shapes do not map onto tcgen05 tiles (3-channel stem, 142/3/1-channel head outputs).
// The dominant layers run on the tcgen05 implicit-GEMM engine in conv_tc.cu.
//
// Replaces per layer: nn.Conv2d + nn.BatchNorm2d (folded) + ReLU + residual add + nn.Upsample(nearest)
// (simple_romp/romp/model.py:49-83,85-123,185-244,338-343,449-466) and the input normalisation of
// model.py:384.
//
// Tiling: one CTA = 8x8 output pixels x 64 output channels, 256 threads, each thread a 4(pixel) x
// 4(channel) register tile; input halo tile and weight slab staged in shared memory in chunks of 8
// input channels.  HBM/L2 traffic per CTA: input halo once per 64-channel slab, weights once per tile.
#include "common.cuh"

namespace b200romp {

enum { IN_F32 = 0, IN_BF16 = 1, IN_U8 = 2 };

template <int KS, int STRIDE>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
  constexpr int TS = 8;
  constexpr int IT = (TS - 1) * STRIDE + KS;
  constexpr int KC = 8;
  constexpr int PAD = KS / 2;
  __shared__ float s_in[IT * IT][KC + 1];
  __shared__ __align__(16) float s_w[KS * KS][KC][64];

  const int tid = threadIdx.x;
  const int tilesX = (p.Wout + TS - 1) / TS;
  const int oy0 = (blockIdx.x / tilesX) * TS, ox0 = (blockIdx.x % tilesX) * TS;
  const int co0 = blockIdx.y * 64;
  const int n = blockIdx.z;
  const int tc = tid & 15, tp = tid >> 4;
  const int prow = tp >> 1, pcol = (tp & 1) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int iy0 = oy0 * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;
  for (int c0 = 0; c0 < p.cin; c0 += KC) {
    for (int idx = tid; idx < IT * IT * KC; idx += 256) {
      const int ci = idx % KC, pix = idx / KC;
      const int gy = iy0 + pix / IT, gx = ix0 + pix % IT;
      float v = 0.f;
      if (gy >= 0 && gy < p.Hin && gx >= 0 && gx < p.Win && c0 + ci < p.cin) {
        const size_t gi = (((size_t)n * p.Hin + gy) * p.Win + gx) * p.in_C + p.in_c_off + c0 + ci;
        if (p.in_dtype == IN_F32) v = reinterpret_cast<const float*>(p.in)[gi];
        else if (p.in_dtype == IN_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in)[gi]);
        else v = (float)reinterpret_cast<const unsigned char*>(p.in)[gi];
        if (p.input_norm) v = (v / 255.f) * 2.f - 1.f;   // model.py:384, same op order
      }
      s_in[pix][ci] = v;
    }
    for (int idx = tid; idx < KS * KS * KC * 64; idx += 256) {
      const int co = idx & 63, ci = (idx >> 6) % KC, tap = idx / (64 * KC);
      float v = 0.f;
      if (c0 + ci < p.cin) v = p.w[((size_t)tap * p.cin + c0 + ci) * p.coutPad + co0 + co];
      s_w[tap][ci][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
        for (int ci = 0; ci < KC; ++ci) {
          const float4 w4 = *reinterpret_cast<const float4*>(&s_w[ky * KS + kx][ci][tc * 4]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = s_in[(prow * STRIDE + ky) * IT + (pcol + i) * STRIDE + kx][ci];
            acc[i][0] = fmaf(a, w4.x, acc[i][0]);
            acc[i][1] = fmaf(a, w4.y, acc[i][1]);
            acc[i][2] = fmaf(a, w4.z, acc[i][2]);
            acc[i][3] = fmaf(a, w4.w, acc[i][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  const int co = co0 + tc * 4;
  if (co >= p.cout) return;
  float b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = p.bias[co + j];   // bias is padded to coutPad
  const int oy = oy0 + prow;
  if (oy >= p.Hout) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + pcol + i;
    if (ox >= p.Wout) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + b[j];
    conv_epilogue_store<4>(p, n, oy, ox, co, v);
  }
}

int launch_conv_simt(const ConvParams& p, int ksize, int stride, cudaStream_t stream) {
  dim3 grid(((p.Hout + 7) / 8) * ((p.Wout + 7) / 8), (p.cout + 63) / 64, p.B);
  dim3 block(256);
  if (ksize == 3 && stride == 1) conv_simt_kernel<3, 1><<<grid, block, 0, stream>>>(p);
  else if (ksize == 3 && stride == 2) conv_simt_kernel<3, 2><<<grid, block, 0, stream>>>(p);
  else if (ksize == 1 && stride == 1) conv_simt_kernel<1, 1><<<grid, block, 0, stream>>>(p);
  else if (ksize == 1 && stride == 2) conv_simt_kernel<1, 2><<<grid, block, 0, stream>>>(p);
  else {
    set_error("conv_simt: unsupported ksize=%d stride=%d", ksize, stride);
    return B200ROMP_EINVAL;
  }
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

}  // namespace b200romp
