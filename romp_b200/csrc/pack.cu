// Device-side packing of per-person outputs into fixed-width records for the ONE collective of the frame-sharded path
// ("one padded all_gather of a packed per-person record"; the reference's nn.DataParallel bookkeeping it
// replaces is romp/lib/maps_utils/result_parser.py:59-64,123-124).
//
// Record buffer = [1 header row | capacity person rows] x row_bytes.  Header (int32): {magic, count, user0, user1, row_bytes};
// the person count is read from device memory, so packing + all-gather need no host synchronisation.
#include "common.cuh"

namespace b200romp {

constexpr int kMaxSeg = 16;
struct PackArgs {
  const uint32_t* src[kMaxSeg];
  int words[kMaxSeg];       // 4-byte words per person in segment s
  int nseg;
};

__global__ void pack_rows_kernel(PackArgs a, const int* __restrict__ d_count, int count_host, int capacity, int user0, int user1,
                                 uint32_t* __restrict__ dst, int row_words) {
  const int n = min(d_count ? *d_count : count_host, capacity);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 8) {
    const int hdr[8] = {0x0B200B20, n, user0, user1, row_words * 4, 0, 0, 0};
    dst[threadIdx.x] = (uint32_t)hdr[threadIdx.x];
  }
  for (int row = blockIdx.x; row < n; row += gridDim.x) {
    uint32_t* d = dst + (size_t)(row + 1) * row_words;
    int off = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const uint32_t* src = a.src[s] + (size_t)row * a.words[s];
      for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < a.words[s]; i += gridDim.y * blockDim.x) d[off + i] = src[i];
      off += a.words[s];
    }
  }
}

}  // namespace b200romp

using namespace b200romp;

extern "C" int b200romp_pack_rows(const void* const* srcs, const int* seg_bytes, int nseg, const int* d_count, int count_host,
                                  int capacity, int user0, int user1, void* dst, int dst_row_bytes, b200romp_stream stream) {
  B2R_REQUIRE(srcs && seg_bytes && dst && nseg > 0 && nseg <= kMaxSeg && capacity >= 0, "pack_rows: bad arguments");
  PackArgs a;
  memset(&a, 0, sizeof(a));
  int total = 0;
  for (int s = 0; s < nseg; ++s) {
    B2R_REQUIRE(seg_bytes[s] > 0 && seg_bytes[s] % 4 == 0 && (reinterpret_cast<uintptr_t>(srcs[s]) & 3) == 0, "pack_rows: segment %d must be 4-byte granular", s);
    a.src[s] = static_cast<const uint32_t*>(srcs[s]);
    a.words[s] = seg_bytes[s] / 4;
    total += seg_bytes[s];
  }
  a.nseg = nseg;
  B2R_REQUIRE(dst_row_bytes % 4 == 0 && dst_row_bytes >= total && dst_row_bytes >= 32, "pack_rows: row of %d bytes cannot hold %d", dst_row_bytes, total);
  const int rows = std::max(1, std::min(capacity, 1024));
  dim3 grid(rows, total > 16384 ? 4 : 1);
  pack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, d_count, count_host, capacity, user0, user1, static_cast<uint32_t*>(dst), dst_row_bytes / 4);
  B2R_CUDA_OK(cudaGetLastError());
  return B200ROMP_OK;
}
