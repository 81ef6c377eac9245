// The three layer types of the ResNet-50 backbone variant (romp/lib/models/resnet_50.py:19-62,93-120) that HRNet does not
// have: the 7x7 stride-2 stem conv, MaxPool2d(3, 2, 1) and ConvTranspose2d(4, 2, 1) (+ folded BatchNorm + ReLU).
// cfg1 (BASELINE.json configs[0]) is the reference's single-image plumbing case, so these are plain fp32 CUDA-core kernels
// (thread = one output pixel x 4 output channels); the 1x1 / 3x3 bottleneck convs run on the engines of conv_simt.cu /
// conv_tc*.cu like every other layer.
#include "common.cuh"

namespace b200romp {

// generic direct conv, any odd ksize (used for 7x7): weights in the SIMT packing [tap][cin][coutPad]
__global__ void __launch_bounds__(256) conv_generic_kernel(const ConvParams p, int ksize, int stride) {
  const int cq = (p.cout + 3) / 4;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t total = (size_t)p.B * p.Hout * p.Wout * cq;
  if (idx >= total) return;
  const int co = (int)(idx % cq) * 4;
  const int ox = (int)((idx / cq) % p.Wout), oy = (int)((idx / ((size_t)cq * p.Wout)) % p.Hout), n = (int)(idx / ((size_t)cq * p.Wout * p.Hout));
  const int pad = ksize / 2;
  float acc[4] = {p.bias[co], p.bias[co + 1], p.bias[co + 2], p.bias[co + 3]};     // bias is padded to coutPad (multiple of 64)
  for (int ky = 0; ky < ksize; ++ky) {
    const int iy = oy * stride - pad + ky;
    if (iy < 0 || iy >= p.Hin) continue;
    for (int kx = 0; kx < ksize; ++kx) {
      const int ix = ox * stride - pad + kx;
      if (ix < 0 || ix >= p.Win) continue;
      const size_t gi = (((size_t)n * p.Hin + iy) * p.Win + ix) * p.in_C + p.in_c_off;
      const float* w = p.w + (size_t)(ky * ksize + kx) * p.cin * p.coutPad + co;
      for (int ci = 0; ci < p.cin; ++ci) {
        float v;
        if (p.in_dtype == B200ROMP_F32) v = reinterpret_cast<const float*>(p.in)[gi + ci];
        else if (p.in_dtype == B200ROMP_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in)[gi + ci]);
        else v = (float)reinterpret_cast<const unsigned char*>(p.in)[gi + ci];
        if (p.input_norm) v = (v / 255.f) * 2.f - 1.f;
        const float4 w4 = *reinterpret_cast<const float4*>(w + (size_t)ci * p.coutPad);
        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]); acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
      }
    }
  }
  conv_epilogue_store<4>(p, n, oy, ox, co, acc);
}

// ConvTranspose2d(kernel 4, stride 2, padding 1): out[oy][ox] = sum over the two (ky, kx) taps per axis whose parity matches:
// iy = (oy + 1 - ky) / 2 when (oy + 1 - ky) is even and 0 <= iy < Hin.  Weights packed [ky*4+kx][cin][coutPad].
__global__ void __launch_bounds__(256) deconv4x4s2_kernel(const ConvParams p) {
  const int cq = (p.cout + 3) / 4;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t total = (size_t)p.B * p.Hout * p.Wout * cq;
  if (idx >= total) return;
  const int co = (int)(idx % cq) * 4;
  const int ox = (int)((idx / cq) % p.Wout), oy = (int)((idx / ((size_t)cq * p.Wout)) % p.Hout), n = (int)(idx / ((size_t)cq * p.Wout * p.Hout));
  float acc[4] = {p.bias[co], p.bias[co + 1], p.bias[co + 2], p.bias[co + 3]};
  for (int ky = (oy + 1) & 1; ky < 4; ky += 2) {
    const int iy = (oy + 1 - ky) / 2;
    if (oy + 1 - ky < 0 || iy >= p.Hin) continue;
    for (int kx = (ox + 1) & 1; kx < 4; kx += 2) {
      const int ix = (ox + 1 - kx) / 2;
      if (ox + 1 - kx < 0 || ix >= p.Win) continue;
      const size_t gi = (((size_t)n * p.Hin + iy) * p.Win + ix) * p.in_C + p.in_c_off;
      const float* w = p.w + (size_t)(ky * 4 + kx) * p.cin * p.coutPad + co;
      for (int ci = 0; ci < p.cin; ++ci) {
        const float v = p.in_dtype == B200ROMP_F32 ? reinterpret_cast<const float*>(p.in)[gi + ci]
                                                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in)[gi + ci]);
        const float4 w4 = *reinterpret_cast<const float4*>(w + (size_t)ci * p.coutPad);
        acc[0] = fmaf(v, w4.x, acc[0]); acc[1] = fmaf(v, w4.y, acc[1]); acc[2] = fmaf(v, w4.z, acc[2]); acc[3] = fmaf(v, w4.w, acc[3]);
      }
    }
  }
  conv_epilogue_store<4>(p, n, oy, ox, co, acc);
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC (implicit -inf padding), thread = one output element
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const void* __restrict__ in, void* __restrict__ out, int dtype, int B, int Hin, int Win,
                                                           int Hout, int Wout, int C) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (size_t)B * Hout * Wout * C) return;
  const int c = (int)(idx % C), ox = (int)((idx / C) % Wout), oy = (int)((idx / ((size_t)C * Wout)) % Hout), n = (int)(idx / ((size_t)C * Wout * Hout));
  float m = -INFINITY;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= Hin) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= Win) continue;
      m = fmaxf(m, load_as_float(in, (((size_t)n * Hin + iy) * Win + ix) * C + c, dtype));
    }
  }
  store_from_float(out, idx, dtype, m);
}

int launch_conv_generic(const ConvParams& p, int ksize, int stride, cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.Hout * p.Wout * ((p.cout + 3) / 4);
  conv_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, ksize, stride);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

int launch_deconv4x4s2(const ConvParams& p, cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.Hout * p.Wout * ((p.cout + 3) / 4);
  deconv4x4s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

int launch_maxpool3x3s2(const void* in, void* out, int dtype, int B, int Hin, int Win, int C, cudaStream_t stream) {
  const int Hout = (Hin + 2 - 3) / 2 + 1, Wout = (Win + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * Hout * Wout * C;
  maxpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, dtype, B, Hin, Win, Hout, Wout, C);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

}  // namespace b200romp
