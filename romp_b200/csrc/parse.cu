// Seam S2: center-map parsing + parameter sampling + 6D->axis-angle, all on the device.
//
// Replaces (simple_romp/romp): CenterMap.parse_centermap post_parser.py:27-47 (nms :50-54 with
// MaxPool2d(5,1,2) :24; two torch.topk calls; gathers; `torch.where(mask)` host sync),
// parameter_sampling :128-133 (which materialises a full [B,4096,145] transpose), pack_params_dict
// :66-79, rot6D_to_angular utils.py:471-475 (rot6d_to_rotmat :477-491, rotation_matrix_to_quaternion
// :606-682, quaternion_to_angle_axis :554-604) and the center_preds/center_confs of :144-145.
//
// Kernel 1 (one CTA per frame): 5x5 max-pool NMS in shared memory, then a full bitonic sort of the
//   <=4096 (score, index) pairs (score desc, index asc - a deterministic version of torch.topk's order),
//   top-64 > thresh kept.  HBM traffic: S*S*4 B per frame read once.
// Kernel 2 (one CTA per frame): prefix over the per-frame counts (B is small), gather of the P-vector
//   at each kept cell straight from the NCHW map (no transpose), rot6d->axis-angle, index outputs.
#include <math_constants.h>

#include "common.cuh"
#include "rot6d.cuh"

namespace b200romp {

constexpr int kMaxPerson = 64;    // CenterMap.max_person, post_parser.py:11
constexpr int kSortN = 4096;

__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(1024) parse_frame_kernel(const float* __restrict__ center_maps, int S, float thresh,
                                                            int* __restrict__ counts, int* __restrict__ cand_idx,
                                                            float* __restrict__ cand_score) {
  __shared__ float s_map[kSortN];
  __shared__ float s_key[kSortN];
  __shared__ int s_idx[kSortN];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int cells = S * S;
  const float* cm = center_maps + (size_t)b * cells;
  for (int i = tid; i < kSortN; i += 1024) s_map[i] = i < cells ? cm[i] : -CUDART_INF_F;
  __syncthreads();
  for (int i = tid; i < kSortN; i += 1024) {
    float key = -CUDART_INF_F;
    if (i < cells) {
      const int y = i / S, x = i % S;
      const float c = s_map[i];
      float m = -CUDART_INF_F;   // MaxPool2d pads with -inf
#pragma unroll
      for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < S && xx >= 0 && xx < S) m = fmaxf(m, s_map[yy * S + xx]);
        }
      }
      // det * float(maxpool(det) == det): maxima keep their value, everything else becomes 0 -> never > thresh>=0
      if (m == c && c > thresh) key = c;
    }
    s_key[i] = key;
    s_idx[i] = i;
  }
  __syncthreads();
  // bitonic sort, "before" order ascending in position
  for (int k = 2; k <= kSortN; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < kSortN / 2; t += 1024) {
        const int lo = ((t / j) * 2 * j) + (t % j);
        const int hi = lo + j;
        const bool up = ((lo & k) == 0);
        const float ka = s_key[lo], kb = s_key[hi];
        const int ia = s_idx[lo], ib = s_idx[hi];
        const bool in_order = before(ka, ia, kb, ib);
        if (in_order != up) {
          s_key[lo] = kb; s_key[hi] = ka;
          s_idx[lo] = ib; s_idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (tid < kMaxPerson) {
    const bool keep = s_key[tid] > thresh;
    cand_idx[b * kMaxPerson + tid] = keep ? s_idx[tid] : -1;
    cand_score[b * kMaxPerson + tid] = s_key[tid];
  }
  if (tid == 0) {
    int c = 0;
    while (c < kMaxPerson && s_key[c] > thresh) ++c;
    counts[b] = c;
  }
}

__global__ void __launch_bounds__(256) parse_gather_kernel(
    const float* __restrict__ center_maps, const float* __restrict__ params_maps, int B, int S, int n_betas,
    int capacity, const int* __restrict__ counts, const int* __restrict__ cand_idx, const float* __restrict__ cand_score,
    int* __restrict__ d_count, long long* __restrict__ batch_ids, long long* __restrict__ flat_inds,
    float* __restrict__ center_confs, float* __restrict__ params_pred, float* __restrict__ cam,
    float* __restrict__ thetas, float* __restrict__ betas, long long* __restrict__ center_preds) {
  __shared__ int s_part[256];
  __shared__ float s_row[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int P = 3 + 22 * 6 + n_betas;
  int acc = 0;
  for (int i = tid; i < b; i += 256) acc += counts[i];
  s_part[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) s_part[tid] += s_part[tid + s];
    __syncthreads();
  }
  const int offset = s_part[0];
  const int cnt = counts[b];
  if (b == B - 1 && tid == 0) *d_count = min(offset + cnt, capacity);
  const int cells = S * S;
  for (int k = 0; k < cnt; ++k) {
    const int n = offset + k;
    if (n >= capacity) break;
    const int flat = cand_idx[b * kMaxPerson + k];
    __syncthreads();
    for (int c = tid; c < P; c += 256) {
      const float v = params_maps[((size_t)b * P + c) * cells + flat];
      params_pred[(size_t)n * P + c] = v;
      s_row[c] = v;
    }
    __syncthreads();
    if (tid < 22) {
      float aa[3];
      rot6d_to_aa(&s_row[3 + tid * 6], aa);
      thetas[(size_t)n * 72 + tid * 3 + 0] = aa[0];
      thetas[(size_t)n * 72 + tid * 3 + 1] = aa[1];
      thetas[(size_t)n * 72 + tid * 3 + 2] = aa[2];
    } else if (tid < 28) {
      thetas[(size_t)n * 72 + 66 + (tid - 22)] = 0.f;            // post_parser.py:76
    } else if (tid >= 32 && tid < 35) {
      cam[(size_t)n * 3 + (tid - 32)] = s_row[tid - 32];
    } else if (tid >= 64 && tid < 64 + n_betas) {
      betas[(size_t)n * n_betas + (tid - 64)] = s_row[135 + (tid - 64)];
    } else if (tid == 96) {
      batch_ids[n] = b;
      flat_inds[n] = flat;
      center_confs[n] = center_maps[(size_t)b * cells + flat];   // post_parser.py:145
      center_preds[(size_t)n * 2 + 0] = (long long)(flat % S) * 512 / S;   // :144 (x, y) * 512 // 64
      center_preds[(size_t)n * 2 + 1] = (long long)(flat / S) * 512 / S;
    }
  }
  (void)cand_score;
}

}  // namespace b200romp

using namespace b200romp;

extern "C" {

// bytes of caller-provided device scratch: per-frame candidate counts, indices and scores
long long b200romp_parse_workspace_bytes(int batch) {
  return (long long)batch * (sizeof(int) + kMaxPerson * (sizeof(int) + sizeof(float)));
}

int b200romp_parse(const float* center_maps, const float* params_maps, int batch, int map_size, int n_betas,
                   float thresh, int capacity, int* d_count, long long* batch_ids, long long* flat_inds,
                   float* center_confs, float* params_pred, float* cam, float* thetas, float* betas,
                   long long* center_preds, void* workspace, b200romp_stream stream_) {
  B2R_REQUIRE(center_maps && params_maps && d_count && batch_ids && flat_inds && center_confs && params_pred && cam &&
                  thetas && betas && center_preds && workspace, "parse: null pointer");
  B2R_REQUIRE(batch > 0 && capacity > 0, "parse: batch and capacity must be positive");
  B2R_REQUIRE(map_size > 0 && map_size * map_size <= kSortN, "parse: map_size %d unsupported (max 64)", map_size);
  B2R_REQUIRE(n_betas >= 1 && n_betas <= 32, "parse: n_betas out of range");
  B2R_REQUIRE(thresh >= 0.f, "parse: thresh must be >= 0");
  cudaStream_t stream = (cudaStream_t)stream_;
  int* counts = reinterpret_cast<int*>(workspace);
  int* cand_idx = counts + batch;
  float* cand_score = reinterpret_cast<float*>(cand_idx + (size_t)batch * kMaxPerson);
  parse_frame_kernel<<<batch, 1024, 0, stream>>>(center_maps, map_size, thresh, counts, cand_idx, cand_score);
  B2R_CUDA_OK(cudaGetLastError());
  parse_gather_kernel<<<batch, 256, 0, stream>>>(center_maps, params_maps, batch, map_size, n_betas, capacity,
                                                 counts, cand_idx, cand_score, d_count,
                                                 batch_ids, flat_inds, center_confs, params_pred, cam, thetas, betas,
                                                 center_preds);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}

}  // extern "C"
