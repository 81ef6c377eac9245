// Seam S3: SMPL forward (fp32), fused so that vertices are written to HBM exactly once.
//
// Replaces simple_romp/romp/smpl.py: SMPL.forward :62-108, lbs :111-188, batch_rodrigues :191-222,
// transform_mat :224-234, batch_rigid_transform :236-290, VertexJointSelector.forward :24-35.
// The reference materialises v_shaped, pose_offsets, T[N,6890,4,4] (441 KB/person written and re-read)
// and runs the kinematic chain as 23 serial launches; here:
//
//   K1 smpl_pose_kernel   (1 warp / person): Rodrigues, joint regression, kinematic chain by tree level,
//        relative transforms A[24][3x4], pose feature; writes 24 posed joints.   ~2 KB/person scratch.
//   K2 smpl_verts_kernel  (CTA = 128 vertices x 32 persons): v_posed = v_t + [betas|pose_feature] x
//        [shapedirs;posedirs] (one K=10+207 contraction), T = W x A, vert = T [v_posed;1]; verts written once.
//        Algorithmic HBM bytes/person: 328 B in + 82,680 B verts + 852 B joints.
//   K3 smpl_joints_kernel (CTA / person): 21 picked vertices + 9 + 17 regressed joints from CSR rows of
//        the (sparse in real SMPL, any density accepted) regressors; optional root alignment.
//   K4 smpl_root_align_kernel: verts -= root (only when root_align, smpl.py:102-106).
//
// J = J_regressor x (v_t + S b) is evaluated as (J_regressor v_t) + (J_regressor S) b with the two
// products precomputed in fp64 at create time (linear identity; differences ~1e-7, tolerance 1e-4).
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

namespace b200romp {

constexpr int kV = 6890;
constexpr int kJ = 24;
// workspace = `cap` per-person records of kWsHead floats, then v_posed.  Record: [0,224) features | [224,512) A[24][12] |
// [512,848) 448 fp16 (hi | lo split of the features, the tensor-core blend's left operand) | [848,1424) skinning operand.
// v_posed (written by the blend GEMM, read by the skinning kernel) is stored COORDINATE-TILE major, [81][cap][256] floats: a
// 256-person x 256-coordinate tile of the GEMM is one contiguous 256 KB block.  (Person-major rows of 20,736 floats made
// every 4 KB TMA store 32 scattered 128 B segments 88 KB apart: 2.8 TB/s of HBM writes.)
constexpr int kBlendCols = 20736;
constexpr int kSkinOff = 848;            // [848, 1424): skinning operand A' = 12 rows x 96 fp16 ([A_hi | A_hi | A_lo] per transform entry)
constexpr int kWsHead = 1424;
constexpr int kWsFloats = kWsHead + kBlendCols;      // floats per person of capacity (b200romp_smpl_workspace_floats)
constexpr int kFeatOff = 0, kAOff = 224, kAqOff = 512;
constexpr int kMaxBetas = 16;

struct SmplDev {
  int n_betas, K;               // K = n_betas + 207
  signed char parents[kJ];      // kinematic tree of THIS handle (by value: BEV holds SMPL-A and SMIL handles side by side)
  signed char depth[kJ];
  const float* v_template;      // [6890*3]
  const float* blend;           // [K][20670]  rows: shapedirs^T then posedirs
  const float* weights;         // [6890][24]
  const float* J_template;      // [24*3]
  const float* J_shape;         // [24*3][n_betas]
  const int* extra_idx;         // [21]
  const int* csr_rowptr;        // [27]
  const int* csr_col;
  const float* csr_val;
  const __half* blend_q;        // [20736][448] fp16: B' = [B_hi | B_lo] per vertex coordinate (smpl_blend_tc.cu)
  const __half* skin_w;         // [6912][96] fp16: W' = [W_hi | W_lo | W_hi] per vertex, 24 joints padded to 32
  const float* vt_pad;          // [20736] v_template, zero padded
};

// smpl_blend_tc.cu
int smpl_blend_tc_launch(const void* a_rows, int a_row_stride_bytes, int capacity, const void* b_rows, float* v_posed, const float* v_template_pad,
                         int n, const int* d_count, int sm_count, cudaStream_t stream);
int smpl_skin_tc_launch(const void* w_rows, const void* ws_base, int a_off_bytes, const float* v_posed, int row_stride_bytes, int capacity, int n,
                        const int* d_count, int sm_count, float* verts, cudaStream_t stream);

__device__ __forceinline__ int person_count(int n, const int* d_count) {
  return d_count ? min(n, *d_count) : n;
}

__global__ void __launch_bounds__(128) smpl_pose_kernel(SmplDev m, const float* __restrict__ betas, int betas_stride,
                                                        const float* __restrict__ thetas, int n_host,
                                                        const int* __restrict__ d_count, float* __restrict__ ws,
                                                        float* __restrict__ joints) {
  __shared__ float s_J[4][kJ][3];
  __shared__ float s_G[4][kJ][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  const int N = person_count(n_host, d_count);
  if (n >= N) return;    // whole warp exits together; only __syncwarp is used below
  float* w = ws + (size_t)n * kWsHead;
  const float* be = betas + (size_t)n * betas_stride;
  float R[9], Jl[3] = {0.f, 0.f, 0.f};
  if (lane < m.n_betas) w[kFeatOff + lane] = be[lane];
  if (lane < kJ) {
    // batch_rodrigues, smpl.py:206-221
    const float rx0 = thetas[(size_t)n * 72 + lane * 3 + 0];
    const float ry0 = thetas[(size_t)n * 72 + lane * 3 + 1];
    const float rz0 = thetas[(size_t)n * 72 + lane * 3 + 2];
    const float ex = rx0 + 1e-8f, ey = ry0 + 1e-8f, ez = rz0 + 1e-8f;
    const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    const float rx = rx0 / angle, ry = ry0 / angle, rz = rz0 / angle;
    const float s = sinf(angle), c1 = 1.f - cosf(angle);
    const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float kk = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) kk += K[a * 3 + c] * K[c * 3 + b];
        R[a * 3 + b] = (a == b ? 1.f : 0.f) + s * K[a * 3 + b] + c1 * kk;
      }
    // J = J_regressor (v_template + shapedirs betas), smpl.py:153-156
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = m.J_template[lane * 3 + k];
      for (int l = 0; l < m.n_betas; ++l) acc = fmaf(m.J_shape[(lane * 3 + k) * m.n_betas + l], be[l], acc);
      Jl[k] = acc;
      s_J[warp][lane][k] = acc;
    }
    if (lane >= 1) {   // pose_feature = (R[1:] - I).view(207), smpl.py:165
#pragma unroll
      for (int e = 0; e < 9; ++e) w[kFeatOff + m.n_betas + (lane - 1) * 9 + e] = R[e] - ((e % 4 == 0) ? 1.f : 0.f);
    }
  }
  __syncwarp();
  {
    // left operand of the tensor-core blend: features split into fp16 hi + lo, laid out [hi(224) | lo(224)]
    __half* aq = reinterpret_cast<__half*>(w + kAqOff);
    for (int k = lane; k < 224; k += 32) {
      const float f = k < m.K ? w[kFeatOff + k] : 0.f;
      const __half hi = __float2half_rn(f);
      const __half lo = __float2half_rn(f - __half2float(hi));
      aq[k] = hi; aq[224 + k] = lo;
    }
  }
  // kinematic chain by tree depth (batch_rigid_transform, smpl.py:260-277): G_i = G_parent * [R_i | J_i - J_parent]
  float G[12];
  for (int level = 0; level < 9; ++level) {
    if (lane < kJ && m.depth[lane] == level) {
      if (level == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          G[a * 4 + 0] = R[a * 3 + 0]; G[a * 4 + 1] = R[a * 3 + 1]; G[a * 4 + 2] = R[a * 3 + 2];
          G[a * 4 + 3] = Jl[a];
        }
      } else {
        const int p = m.parents[lane];
        const float t0 = Jl[0] - s_J[warp][p][0], t1 = Jl[1] - s_J[warp][p][1], t2 = Jl[2] - s_J[warp][p][2];
        float P[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) P[e] = s_G[warp][p][e];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int b = 0; b < 3; ++b)
            G[a * 4 + b] = P[a * 4 + 0] * R[0 * 3 + b] + P[a * 4 + 1] * R[1 * 3 + b] + P[a * 4 + 2] * R[2 * 3 + b];
          G[a * 4 + 3] = P[a * 4 + 0] * t0 + P[a * 4 + 1] * t1 + P[a * 4 + 2] * t2 + P[a * 4 + 3];
        }
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) s_G[warp][lane][e] = G[e];
    }
    __syncwarp();
  }
  if (lane < kJ) {
    // posed joints (:280) and A = G - [0 | G [J;0]] (:285-288)
    float* jo = joints + ((size_t)n * 71 + lane) * 3;
    jo[0] = G[3]; jo[1] = G[7]; jo[2] = G[11];
    float* A = w + kAOff + lane * 12;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      A[a * 4 + 0] = G[a * 4 + 0]; A[a * 4 + 1] = G[a * 4 + 1]; A[a * 4 + 2] = G[a * 4 + 2];
      A[a * 4 + 3] = G[a * 4 + 3] - (G[a * 4 + 0] * Jl[0] + G[a * 4 + 1] * Jl[1] + G[a * 4 + 2] * Jl[2]);
    }
  }
  __syncwarp();
  {
    // right operand of the tensor-core skinning: row e (transform entry) = [A_hi[0..31] | A_hi | A_lo] over the joints (24, zero padded)
    __half* sk = reinterpret_cast<__half*>(w + kSkinOff);
    for (int i = lane; i < 12 * 32; i += 32) {
      const int e = i >> 5, j = i & 31;
      const float a = j < kJ ? w[kAOff + j * 12 + e] : 0.f;
      const __half hi = __float2half_rn(a);
      const __half lo = __float2half_rn(a - __half2float(hi));
      sk[e * 96 + j] = hi; sk[e * 96 + 32 + j] = hi; sk[e * 96 + 64 + j] = lo;
    }
  }
}

constexpr int kVT = 128;   // vertices per CTA
constexpr int kPT = 32;    // persons per CTA (two halves of 16)

__global__ void __launch_bounds__(256) smpl_verts_kernel(SmplDev m, int n_host, const int* __restrict__ d_count,
                                                         const float* __restrict__ ws, float* __restrict__ verts) {
  extern __shared__ __align__(16) float smem[];
  float* s_feat = smem;                       // [K][32]
  float* s_A = smem + 224 * kPT;              // [32][288]
  const int N = person_count(n_host, d_count);
  const int n0 = blockIdx.y * kPT;
  if (n0 >= N) return;
  const int tid = threadIdx.x;
  const int K = m.K;
  for (int i = tid; i < K * kPT; i += 256) {
    const int q = i % kPT, p = i / kPT;
    s_feat[p * kPT + q] = (n0 + q < N) ? ws[(size_t)(n0 + q) * kWsHead + kFeatOff + p] : 0.f;
  }
  for (int i = tid; i < kPT * 288; i += 256) {
    const int q = i / 288, e = i % 288;
    s_A[i] = (n0 + q < N) ? ws[(size_t)(n0 + q) * kWsHead + kAOff + e] : 0.f;
  }
  __syncthreads();
  const int half = tid >> 7;
  const int v = blockIdx.x * kVT + (tid & 127);
  if (v >= kV) return;
  float acc[16][3];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
  const float* bl = m.blend + (size_t)3 * v;
#pragma unroll 4
  for (int p = 0; p < K; ++p) {
    const float b0 = bl[(size_t)p * (3 * kV) + 0], b1 = bl[(size_t)p * (3 * kV) + 1], b2 = bl[(size_t)p * (3 * kV) + 2];
    const float4* f4 = reinterpret_cast<const float4*>(s_feat + p * kPT + half * 16);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 f = f4[g];
      acc[g * 4 + 0][0] = fmaf(f.x, b0, acc[g * 4 + 0][0]); acc[g * 4 + 0][1] = fmaf(f.x, b1, acc[g * 4 + 0][1]); acc[g * 4 + 0][2] = fmaf(f.x, b2, acc[g * 4 + 0][2]);
      acc[g * 4 + 1][0] = fmaf(f.y, b0, acc[g * 4 + 1][0]); acc[g * 4 + 1][1] = fmaf(f.y, b1, acc[g * 4 + 1][1]); acc[g * 4 + 1][2] = fmaf(f.y, b2, acc[g * 4 + 1][2]);
      acc[g * 4 + 2][0] = fmaf(f.z, b0, acc[g * 4 + 2][0]); acc[g * 4 + 2][1] = fmaf(f.z, b1, acc[g * 4 + 2][1]); acc[g * 4 + 2][2] = fmaf(f.z, b2, acc[g * 4 + 2][2]);
      acc[g * 4 + 3][0] = fmaf(f.w, b0, acc[g * 4 + 3][0]); acc[g * 4 + 3][1] = fmaf(f.w, b1, acc[g * 4 + 3][1]); acc[g * 4 + 3][2] = fmaf(f.w, b2, acc[g * 4 + 3][2]);
    }
  }
  const float vt0 = m.v_template[3 * v + 0], vt1 = m.v_template[3 * v + 1], vt2 = m.v_template[3 * v + 2];
  float wj[kJ];
  {
    const float4* w4 = reinterpret_cast<const float4*>(m.weights + (size_t)v * kJ);
#pragma unroll
    for (int g = 0; g < 6; ++g) {
      const float4 t = w4[g];
      wj[g * 4 + 0] = t.x; wj[g * 4 + 1] = t.y; wj[g * 4 + 2] = t.z; wj[g * 4 + 3] = t.w;
    }
  }
#pragma unroll 2
  for (int q = 0; q < 16; ++q) {
    const int n = n0 + half * 16 + q;
    if (n >= N) break;
    const float4* A4 = reinterpret_cast<const float4*>(s_A + (half * 16 + q) * 288);
    float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const float4 a0 = A4[j * 3 + 0], a1 = A4[j * 3 + 1], a2 = A4[j * 3 + 2];
      const float w = wj[j];
      T0.x = fmaf(w, a0.x, T0.x); T0.y = fmaf(w, a0.y, T0.y); T0.z = fmaf(w, a0.z, T0.z); T0.w = fmaf(w, a0.w, T0.w);
      T1.x = fmaf(w, a1.x, T1.x); T1.y = fmaf(w, a1.y, T1.y); T1.z = fmaf(w, a1.z, T1.z); T1.w = fmaf(w, a1.w, T1.w);
      T2.x = fmaf(w, a2.x, T2.x); T2.y = fmaf(w, a2.y, T2.y); T2.z = fmaf(w, a2.z, T2.z); T2.w = fmaf(w, a2.w, T2.w);
    }
    const float px = vt0 + acc[q][0], py = vt1 + acc[q][1], pz = vt2 + acc[q][2];
    float* o = verts + ((size_t)n * kV + v) * 3;
    o[0] = T0.x * px + T0.y * py + T0.z * pz + T0.w;
    o[1] = T1.x * px + T1.y * py + T1.z * pz + T1.w;
    o[2] = T2.x * px + T2.y * py + T2.z * pz + T2.w;
  }
}

// Skinning on v_posed produced by the tensor-core blend (smpl_blend_tc.cu): T = W x A, vert = T [v_posed;1] (smpl.py:176-186).
// CTA = 128 vertices x 32 persons (two halves of 16), W row in registers, A broadcast from shared memory; verts written once.
__global__ void __launch_bounds__(256) smpl_skin_kernel(SmplDev m, int n_host, const int* __restrict__ d_count,
                                                        const float* __restrict__ ws, float* __restrict__ verts) {
  __shared__ __align__(16) float s_A[kPT * 288];
  const int N = person_count(n_host, d_count);
  const int n0 = blockIdx.y * kPT;
  if (n0 >= N) return;
  const int tid = threadIdx.x;
  for (int i = tid; i < kPT * 288; i += 256) {
    const int q = i / 288, e = i % 288;
    s_A[i] = (n0 + q < N) ? ws[(size_t)(n0 + q) * kWsHead + kAOff + e] : 0.f;
  }
  __syncthreads();
  const int half = tid >> 7;
  const int v = blockIdx.x * kVT + (tid & 127);
  if (v >= kV) return;
  float wj[kJ];
  {
    const float4* w4 = reinterpret_cast<const float4*>(m.weights + (size_t)v * kJ);
#pragma unroll
    for (int g = 0; g < 6; ++g) {
      const float4 t = w4[g];
      wj[g * 4 + 0] = t.x; wj[g * 4 + 1] = t.y; wj[g * 4 + 2] = t.z; wj[g * 4 + 3] = t.w;
    }
  }
#pragma unroll 2
  for (int q = 0; q < 16; ++q) {
    const int n = n0 + half * 16 + q;
    if (n >= N) break;
    const float* vp = ws + (size_t)n_host * kWsHead;          // v_posed [81][n_host][256]
    const int c0 = 3 * v, c1 = c0 + 1, c2 = c0 + 2;
    const float px = vp[((size_t)(c0 >> 8) * n_host + n) * 256 + (c0 & 255)], py = vp[((size_t)(c1 >> 8) * n_host + n) * 256 + (c1 & 255)],
                pz = vp[((size_t)(c2 >> 8) * n_host + n) * 256 + (c2 & 255)];
    const float4* A4 = reinterpret_cast<const float4*>(s_A + (half * 16 + q) * 288);
    float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const float4 a0 = A4[j * 3 + 0], a1 = A4[j * 3 + 1], a2 = A4[j * 3 + 2];
      const float w = wj[j];
      T0.x = fmaf(w, a0.x, T0.x); T0.y = fmaf(w, a0.y, T0.y); T0.z = fmaf(w, a0.z, T0.z); T0.w = fmaf(w, a0.w, T0.w);
      T1.x = fmaf(w, a1.x, T1.x); T1.y = fmaf(w, a1.y, T1.y); T1.z = fmaf(w, a1.z, T1.z); T1.w = fmaf(w, a1.w, T1.w);
      T2.x = fmaf(w, a2.x, T2.x); T2.y = fmaf(w, a2.y, T2.y); T2.z = fmaf(w, a2.z, T2.z); T2.w = fmaf(w, a2.w, T2.w);
    }
    float* o = verts + ((size_t)n * kV + v) * 3;
    o[0] = T0.x * px + T0.y * py + T0.z * pz + T0.w;
    o[1] = T1.x * px + T1.y * py + T1.z * pz + T1.w;
    o[2] = T2.x * px + T2.y * py + T2.z * pz + T2.w;
  }
}

__global__ void __launch_bounds__(256) smpl_joints_kernel(SmplDev m, int n_host, const int* __restrict__ d_count,
                                                          const float* __restrict__ verts, int root_align,
                                                          float* __restrict__ joints, float* __restrict__ ws) {
  __shared__ float s_root[3];
  const int n = blockIdx.x;
  const int N = person_count(n_host, d_count);
  if (n >= N) return;
  const float* vp = verts + (size_t)n * kV * 3;
  float* jo = joints + (size_t)n * 71 * 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 21 * 3) {   // index_select of 21 vertices, smpl.py:25
    const int j = tid / 3, k = tid % 3;
    jo[(24 + j) * 3 + k] = vp[(size_t)m.extra_idx[j] * 3 + k];
  }
  for (int row = warp; row < 26; row += 8) {   // einsum('bik,ji->bjk') for the 9 + 17 regressors, smpl.py:26-27
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int e = m.csr_rowptr[row] + lane; e < m.csr_rowptr[row + 1]; e += 32) {
      const float w = m.csr_val[e];
      const float* q = vp + (size_t)m.csr_col[e] * 3;
      a0 = fmaf(w, q[0], a0); a1 = fmaf(w, q[1], a1); a2 = fmaf(w, q[2], a2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
      jo[(45 + row) * 3 + 0] = a0; jo[(45 + row) * 3 + 1] = a1; jo[(45 + row) * 3 + 2] = a2;
    }
  }
  if (!root_align) return;
  __syncthreads();   // joints 45,46 written by warps 0,1 of this CTA (global writes visible after the barrier)
  if (tid < 3) {
    const float r = (jo[45 * 3 + tid] + jo[46 * 3 + tid]) / 2.f;   // joints54[:,[45,46]].mean(1), smpl.py:104
    s_root[tid] = r;
    ws[(size_t)n * kWsHead + kFeatOff + tid] = r;               // feature slots are dead by now
  }
  __syncthreads();
  for (int i = tid; i < 71 * 3; i += 256) jo[i] -= s_root[i % 3];
}

__global__ void __launch_bounds__(256) smpl_root_align_kernel(int n_host, const int* __restrict__ d_count,
                                                              const float* __restrict__ ws, float* __restrict__ verts) {
  const int n = blockIdx.x;
  if (n >= person_count(n_host, d_count)) return;
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= kV * 3) return;
  verts[(size_t)n * kV * 3 + i] -= ws[(size_t)n * kWsHead + kFeatOff + (i % 3)];
}

}  // namespace b200romp

using namespace b200romp;

struct b200romp_smpl {
  int device = 0, sm_count = 148;
  SmplDev dev;
  std::vector<void*> allocs;
  int smem_verts = 0;
};

template <typename T>
static int upload(b200romp_smpl* s, const std::vector<T>& host, const T** out) {
  void* d = nullptr;
  B2R_CUDA_OK(cudaMalloc(&d, host.size() * sizeof(T)));
  s->allocs.push_back(d);
  B2R_CUDA_OK(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(d);
  return B200ROMP_OK;
}

extern "C" {

int b200romp_smpl_workspace_floats(void) { return kWsFloats; }

b200romp_smpl* b200romp_smpl_create(int device, int n_betas, const float* v_template, const float* shapedirs,
                                    const float* posedirs, const float* J_regressor, const float* weights,
                                    const long long* parents, const long long* extra_joints_index,
                                    const float* J_regressor_extra9, const float* J_regressor_h36m17) {
  if (!v_template || !shapedirs || !posedirs || !J_regressor || !weights || !parents || !extra_joints_index ||
      !J_regressor_extra9 || !J_regressor_h36m17 || n_betas < 1 || n_betas > kMaxBetas) {
    set_error("smpl_create: bad arguments");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("smpl_create: cudaSetDevice(%d) failed (no CPU fallback)", device);
    return nullptr;
  }
  int par[kJ], depth[kJ];
  for (int j = 0; j < kJ; ++j) {
    par[j] = (int)parents[j];
    if (j == 0) par[j] = -1;
    if (j > 0 && (par[j] < 0 || par[j] >= j)) {
      set_error("smpl_create: kintree_table must satisfy 0 <= parent[j] < j");
      return nullptr;
    }
    depth[j] = j == 0 ? 0 : depth[par[j]] + 1;
    if (depth[j] > 8) {
      set_error("smpl_create: kinematic tree deeper than 9 levels");
      return nullptr;
    }
  }
  b200romp_smpl* s = new b200romp_smpl();
  s->device = device;
  const int K = n_betas + 207;
  bool ok = true;
  for (int j = 0; j < kJ; ++j) {
    s->dev.parents[j] = (signed char)par[j];
    s->dev.depth[j] = (signed char)depth[j];
  }
  // blend matrix: rows 0..n_betas-1 = shapedirs[:, :, l] flattened, then posedirs
  std::vector<float> blend((size_t)K * 3 * kV);
  for (int l = 0; l < n_betas; ++l)
    for (int i = 0; i < 3 * kV; ++i) blend[(size_t)l * 3 * kV + i] = shapedirs[(size_t)i * n_betas + l];
  std::copy(posedirs, posedirs + (size_t)207 * 3 * kV, blend.begin() + (size_t)n_betas * 3 * kV);
  std::vector<float> Jt(kJ * 3), Js((size_t)kJ * 3 * n_betas);
  for (int j = 0; j < kJ; ++j)
    for (int k = 0; k < 3; ++k) {
      double a = 0.0;
      for (int v = 0; v < kV; ++v) a += (double)J_regressor[(size_t)j * kV + v] * v_template[v * 3 + k];
      Jt[j * 3 + k] = (float)a;
      for (int l = 0; l < n_betas; ++l) {
        double b = 0.0;
        for (int v = 0; v < kV; ++v)
          b += (double)J_regressor[(size_t)j * kV + v] * shapedirs[((size_t)v * 3 + k) * n_betas + l];
        Js[((size_t)j * 3 + k) * n_betas + l] = (float)b;
      }
    }
  std::vector<int> rowptr(27, 0), cols;
  std::vector<float> vals;
  for (int r = 0; r < 26; ++r) {
    const float* row = r < 9 ? J_regressor_extra9 + (size_t)r * kV : J_regressor_h36m17 + (size_t)(r - 9) * kV;
    for (int v = 0; v < kV; ++v)
      if (row[v] != 0.f) {
        cols.push_back(v);
        vals.push_back(row[v]);
      }
    rowptr[r + 1] = (int)cols.size();
  }
  if (cols.empty()) { cols.push_back(0); vals.push_back(0.f); }
  std::vector<int> eidx(21);
  for (int i = 0; i < 21; ++i) {
    eidx[i] = (int)extra_joints_index[i];
    if (eidx[i] < 0 || eidx[i] >= kV) ok = false;
  }
  std::vector<float> vt(v_template, v_template + 3 * kV), w(weights, weights + (size_t)kV * kJ);
  // tensor-core blend operands: B'[(v,c)][k'] = [B_hi | B_hi | B_lo] (fp16 hi/lo split of the blend matrix), zero padded
  std::vector<__half> bq((size_t)kBlendCols * 448, __float2half(0.f));
  std::vector<float> vtp(kBlendCols, 0.f);
  for (int r = 0; r < 3 * kV; ++r) {
    vtp[r] = v_template[r];
    for (int k = 0; k < K; ++k) {
      const float b = blend[(size_t)k * 3 * kV + r];
      const __half hi = __float2half_rn(b);
      const __half lo = __float2half_rn(b - __half2float(hi));
      bq[(size_t)r * 448 + k] = hi; bq[(size_t)r * 448 + 224 + k] = lo;
    }
  }
  std::vector<__half> wq((size_t)6912 * 96, __float2half(0.f));
  for (int v = 0; v < kV; ++v)
    for (int j = 0; j < kJ; ++j) {
      const float x = weights[(size_t)v * kJ + j];
      const __half hi = __float2half_rn(x);
      const __half lo = __float2half_rn(x - __half2float(hi));
      wq[(size_t)v * 96 + j] = hi; wq[(size_t)v * 96 + 32 + j] = lo; wq[(size_t)v * 96 + 64 + j] = hi;
    }
  { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->sm_count = prop.multiProcessorCount; }
  SmplDev& d = s->dev;
  d.n_betas = n_betas; d.K = K;
  ok = ok && upload(s, vt, &d.v_template) == 0 && upload(s, blend, &d.blend) == 0 && upload(s, w, &d.weights) == 0 &&
       upload(s, Jt, &d.J_template) == 0 && upload(s, Js, &d.J_shape) == 0 && upload(s, eidx, &d.extra_idx) == 0 &&
       upload(s, rowptr, &d.csr_rowptr) == 0 && upload(s, cols, &d.csr_col) == 0 && upload(s, vals, &d.csr_val) == 0 &&
       upload(s, bq, &d.blend_q) == 0 && upload(s, vtp, &d.vt_pad) == 0 && upload(s, wq, &d.skin_w) == 0;
  s->smem_verts = (224 * kPT + kPT * 288) * (int)sizeof(float);
  ok = ok && cudaFuncSetAttribute(smpl_verts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_verts) == cudaSuccess;
  if (!ok) {
    set_error("smpl_create: upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
    b200romp_smpl_destroy(s);
    return nullptr;
  }
  return s;
}

void b200romp_smpl_destroy(b200romp_smpl* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (void* p : s->allocs) cudaFree(p);
  delete s;
}

int b200romp_smpl_forward(b200romp_smpl* s, const float* betas, int betas_stride, const float* thetas, int n,
                          const int* d_count, int root_align, float* workspace, float* verts, float* joints,
                          b200romp_stream stream_) {
  B2R_REQUIRE(s && betas && thetas && workspace && verts && joints, "smpl_forward: null pointer");
  B2R_REQUIRE(n > 0 && betas_stride >= s->dev.n_betas, "smpl_forward: n must be > 0 and betas_stride >= n_betas");
  cudaStream_t stream = (cudaStream_t)stream_;
  B2R_CUDA_OK(cudaSetDevice(s->device));
  smpl_pose_kernel<<<(n + 3) / 4, 128, 0, stream>>>(s->dev, betas, betas_stride, thetas, n, d_count, workspace, joints);
  B2R_CUDA_OK(cudaGetLastError());
  dim3 grid((kV + kVT - 1) / kVT, (n + kPT - 1) / kPT);
  static const bool simt_blend = [] { const char* e = getenv("B200ROMP_SMPL_SIMT"); return e && e[0] == '1'; }();
  if (simt_blend) {       // round-1 formulation: fp32 FFMA blend + skinning in one SIMT kernel (kept for A/B measurements)
    smpl_verts_kernel<<<grid, 256, s->smem_verts, stream>>>(s->dev, n, d_count, workspace, verts);
  } else {                // shape + pose blend as a tcgen05 GEMM (3-term fp16 split), then skinning on its v_posed
    float* v_posed = workspace + (size_t)n * kWsHead;            // [81][n][256], after the n records (n = the caller's capacity bound)
    int rc = smpl_blend_tc_launch(reinterpret_cast<const char*>(workspace) + kAqOff * sizeof(float), kWsHead * (int)sizeof(float), n, s->dev.blend_q,
                                  v_posed, s->dev.vt_pad, n, d_count, s->sm_count, stream);
    if (rc) return rc;
    static const bool simt_skin = [] { const char* e = getenv("B200ROMP_SMPL_SKIN_SIMT"); return e && e[0] == '1'; }();
    if (simt_skin) {      // FFMA skinning on the blend's v_posed (300 FMA per vertex and person; A/B switch)
      smpl_skin_kernel<<<grid, 256, 0, stream>>>(s->dev, n, d_count, workspace, verts);
    } else {              // skinning as a tiny-K tcgen05 GEMM (smpl_blend_tc.cu: smpl_skin_tc_kernel)
      rc = smpl_skin_tc_launch(s->dev.skin_w, workspace, kSkinOff * (int)sizeof(float), v_posed, kWsHead * (int)sizeof(float), n, n, d_count,
                               s->sm_count, verts, stream);
      if (rc) return rc;
    }
  }
  B2R_CUDA_OK(cudaGetLastError());
  smpl_joints_kernel<<<n, 256, 0, stream>>>(s->dev, n, d_count, verts, root_align, joints, workspace);
  B2R_CUDA_OK(cudaGetLastError());
  if (root_align) {
    dim3 g2(n, (kV * 3 + 255) / 256);
    smpl_root_align_kernel<<<g2, 256, 0, stream>>>(n, d_count, workspace, verts);
    B2R_CUDA_OK(cudaGetLastError());
  }
  return B200ROMP_OK;
}

}  // extern "C"
