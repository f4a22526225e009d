// Mask compaction, mask validation and row gather: the HBM-/latency-bound integer and byte work around the
// attention kernel.
//
//   csa_compact_rows   per-frame boolean mask row -> ascending key index list (== torch.nonzero of the row),
//                      replacing the dense (T*N)^2 mask of StoryDiffusion/utils/gradio_utils.py:260-286 and its
//                      slicing at StoryDiffusion/Comic_Generation.py:105-114.
//   csa_validate_mask  every row of a frame block equals the block's first row (premise of the compaction).
//   csa_gather_rows    dst[i] = src[base + idx[i]]  (sampled K/V rows made contiguous for the NVLink exchange).
#include "csa_internal.h"

namespace csa {

constexpr int kCompactThreads = 1024;
constexpr int kBytesPerThread = 16;  // one 16-byte load per thread per sweep

// value of the (virtual) mask row r at column j, see header for block_n / limit_cols
__device__ __forceinline__ bool mask_value(uint8_t byte, int r, int j, int block_n, int limit_cols) {
  if (block_n <= 0) return byte != 0;
  const bool own = (j >= r * block_n) && (j < (r + 1) * block_n);
  return own || (byte != 0 && j < limit_cols);
}

// One CTA per row.  Sweep the row in chunks of 1024 threads x 16 bytes; within a chunk: per-thread popcount ->
// warp inclusive scan (shuffles) -> cross-warp scan in shared memory -> ordered scatter of the indices.
__global__ void __launch_bounds__(kCompactThreads) compact_rows_kernel(const uint8_t* __restrict__ mask,
                                                                       int64_t row_stride, int n_cols, int block_n,
                                                                       int limit_cols, int32_t* __restrict__ idx,
                                                                       int64_t idx_stride,
                                                                       int32_t* __restrict__ counts) {
  __shared__ int warp_tot[kCompactThreads / 32];   // inclusive totals per warp
  __shared__ int warp_excl[kCompactThreads / 32];  // exclusive prefix of those
  __shared__ int chunk_total;
  const int r = blockIdx.x;
  const uint8_t* row = mask + static_cast<int64_t>(r) * row_stride;
  int32_t* out = idx + static_cast<int64_t>(r) * idx_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  int base = 0;  // keys found in earlier chunks (uniform across the CTA)

  for (int c0 = 0; c0 < n_cols; c0 += kCompactThreads * kBytesPerThread) {
    const int j0 = c0 + threadIdx.x * kBytesPerThread;
    uint32_t bits = 0;  // bit b set <=> column j0+b is attended
    if (j0 < n_cols) {
      uint8_t b[kBytesPerThread];
      if (vec_ok && j0 + kBytesPerThread <= n_cols) {
        *reinterpret_cast<uint4*>(b) = __ldg(reinterpret_cast<const uint4*>(row + j0));
      } else {
#pragma unroll
        for (int i = 0; i < kBytesPerThread; ++i) b[i] = (j0 + i < n_cols) ? row[j0 + i] : 0;
      }
#pragma unroll
      for (int i = 0; i < kBytesPerThread; ++i) {
        const int j = j0 + i;
        if (j < n_cols && mask_value(b[i], r, j, block_n, limit_cols)) bits |= (1u << i);
      }
    }
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int v = warp_tot[lane];
      int s = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += y;
      }
      warp_excl[lane] = s - v;
      if (lane == 31) chunk_total = s;
    }
    __syncthreads();
    int pos = base + warp_excl[warp] + (incl - cnt);
    while (bits) {
      const int i = __ffs(bits) - 1;
      bits &= bits - 1;
      out[pos++] = j0 + i;
    }
    base += chunk_total;
    __syncthreads();  // warp_tot / warp_excl / chunk_total are rewritten by the next chunk
  }
  if (threadIdx.x == 0) counts[r] = base;
}

// grid.x = number of row blocks, grid.y splits the rows of a block; every row is compared 16 bytes at a time
// with the first row of its block.
__global__ void __launch_bounds__(256) validate_mask_kernel(const uint8_t* __restrict__ mask, int64_t row_stride,
                                                            int n_rows, int n_cols, int block_n,
                                                            int32_t* __restrict__ n_bad) {
  const int blk = blockIdx.x;
  const int r_begin = blk * block_n;
  const int r_end = min(n_rows, r_begin + block_n);
  const uint8_t* ref = mask + static_cast<int64_t>(r_begin) * row_stride;
  // 16-byte compares when base and stride allow it, byte compares otherwise (odd-sized toy masks)
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(mask) | static_cast<uintptr_t>(row_stride)) & 15) == 0;
  const int words = vec_ok ? n_cols / 16 : 0;
  int bad = 0;
  for (int r = r_begin + 1 + blockIdx.y; r < r_end; r += gridDim.y) {
    const uint8_t* row = mask + static_cast<int64_t>(r) * row_stride;
    for (int wd = threadIdx.x; wd < words; wd += blockDim.x) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(ref) + wd);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(row) + wd);
      bad += (a.x != b.x) | (a.y != b.y) | (a.z != b.z) | (a.w != b.w);
    }
    for (int j = words * 16 + threadIdx.x; j < n_cols; j += blockDim.x) bad += (ref[j] != row[j]);
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

// one warp per destination row, 16 bytes per lane per step
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ src, int64_t src_ld,
                                                          int row_base, const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ count, int count_adjust,
                                                          int max_rows, uint8_t* __restrict__ dst, int64_t dst_ld,
                                                          int row_bytes) {
  int n = max_rows;
  if (count != nullptr) {
    const int c = __ldg(count) + count_adjust;
    n = c < n ? c : n;
  }
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = row_bytes >> 4;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
    const int r = row_base + __ldg(idx + i);
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r) * src_ld);
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(i) * dst_ld);
    for (int v = lane; v < vecs; v += 32) d[v] = __ldg(s + v);
  }
}


// ranges[f] = {0, lo_f, hi_f, count - hi_f} (binary searches in the ascending list), ranges[n_frames] = {0,count,0,0}
__global__ void sample_ranges_kernel(const int32_t* __restrict__ s_idx, const int32_t* __restrict__ s_count,
                                     int block_n, int n_frames, int32_t* __restrict__ ranges) {
  const int f = threadIdx.x;
  if (f > n_frames) return;
  const int count = __ldg(s_count);
  auto lower_bound = [&](int key) {
    int lo = 0, hi = count;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(s_idx + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  int4 r;
  if (f == n_frames) {
    r = make_int4(0, count, 0, 0);
  } else {
    const int lo = lower_bound(f * block_n);
    const int hi = lower_bound((f + 1) * block_n);
    r = make_int4(0, lo, hi, count - hi);
  }
  reinterpret_cast<int4*>(ranges)[f] = r;
}

// grid.y = group * 2 + (0: K, 1: V); one warp per destination row, 16 bytes per lane per step; rows in
// [count, count + CSA_TILE) are zero-filled
__global__ void __launch_bounds__(256) gather_kv_kernel(const uint8_t* __restrict__ k, const uint8_t* __restrict__ v,
                                                        int64_t ld, int group_rows, const int32_t* __restrict__ idx,
                                                        const int32_t* __restrict__ count, int max_rows,
                                                        uint8_t* __restrict__ k_out, uint8_t* __restrict__ v_out,
                                                        int64_t out_ld, int out_group_rows, int row_bytes) {
  int n = __ldg(count);
  n = n < max_rows ? n : max_rows;
  const int g = blockIdx.y >> 1;
  const uint8_t* src = (blockIdx.y & 1) ? v : k;
  uint8_t* dst = ((blockIdx.y & 1) ? v_out : k_out) + static_cast<int64_t>(g) * out_group_rows * out_ld;
  src += static_cast<int64_t>(g) * group_rows * ld;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = row_bytes >> 4;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n + CSA_TILE;
       i += gridDim.x * warps_per_block) {
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(i) * out_ld);
    if (i < n) {
      const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(__ldg(idx + i)) * ld);
      for (int c = lane; c < vecs; c += 32) d[c] = __ldg(s + c);
    } else {
      for (int c = lane; c < vecs; c += 32) d[c] = make_uint4(0, 0, 0, 0);
    }
  }
}

}  // namespace csa

using namespace csa;

extern "C" int csa_compact_rows(const uint8_t* mask, int64_t row_stride, int32_t n_rows, int32_t n_cols,
                                int32_t block_n, int32_t limit_cols, int32_t* idx, int64_t idx_stride, int32_t* counts,
                                void* stream) {
  if (!mask || !idx || !counts) return set_error(CSA_E_BADARG, "csa_compact_rows: null pointer");
  if (n_rows <= 0 || n_cols <= 0 || row_stride < 0 || block_n < 0)
    return set_error(CSA_E_BADARG, "csa_compact_rows: bad sizes (rows %d cols %d stride %lld)", n_rows, n_cols,
                     (long long)row_stride);
  const int64_t need = (static_cast<int64_t>(n_cols) + CSA_TILE - 1) / CSA_TILE * CSA_TILE;
  if (idx_stride < need || (idx_stride & 3))
    return set_error(CSA_E_BADARG, "csa_compact_rows: idx_stride %lld must be a multiple of 4 and >= %lld",
                     (long long)idx_stride, (long long)need);
  compact_rows_kernel<<<n_rows, kCompactThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      mask, row_stride, n_cols, block_n, limit_cols, idx, idx_stride, counts);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "compact_rows_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_validate_mask(const uint8_t* mask, int64_t row_stride, int32_t n_rows, int32_t n_cols,
                                 int32_t block_n, int32_t* n_bad, void* stream) {
  if (!mask || !n_bad) return set_error(CSA_E_BADARG, "csa_validate_mask: null pointer");
  if (n_rows <= 0 || n_cols <= 0 || block_n <= 0 || row_stride < n_cols)
    return set_error(CSA_E_BADARG, "csa_validate_mask: bad sizes");
  const int blocks = (n_rows + block_n - 1) / block_n;
  int split = (148 * 8 + blocks - 1) / blocks;  // ~8 CTAs per SM in flight
  if (split > block_n - 1) split = block_n - 1;
  if (split < 1) split = 1;
  validate_mask_kernel<<<dim3(blocks, split), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, row_stride, n_rows,
                                                                                         n_cols, block_n, n_bad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "validate_mask_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_gather_rows(const void* src, int64_t src_ld_bytes, int32_t row_base, const int32_t* idx,
                               const int32_t* count, int32_t count_adjust, int32_t max_rows, void* dst,
                               int64_t dst_ld_bytes, int32_t row_bytes, void* stream) {
  if (!src || !idx || !dst) return set_error(CSA_E_BADARG, "csa_gather_rows: null pointer");
  if (max_rows <= 0 || row_bytes <= 0 || (row_bytes & 15) || (src_ld_bytes & 15) || (dst_ld_bytes & 15) ||
      (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return set_error(CSA_E_BADARG, "csa_gather_rows: sizes/strides/pointers must be positive and 16-byte aligned");
  const int warps_per_block = 8;
  int grid = (max_rows + warps_per_block - 1) / warps_per_block;
  const int cap = 148 * 8;
  if (grid > cap) grid = cap;
  gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), src_ld_bytes, row_base, idx, count, count_adjust, max_rows,
      static_cast<uint8_t*>(dst), dst_ld_bytes, row_bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "gather_rows_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_sample_ranges(const int32_t* s_idx, const int32_t* s_count, int32_t block_n, int32_t n_frames,
                                 int32_t* ranges, void* stream) {
  if (!s_idx || !s_count || !ranges) return set_error(CSA_E_BADARG, "csa_sample_ranges: null pointer");
  if (block_n <= 0 || n_frames <= 0 || n_frames >= 1024 || (reinterpret_cast<uintptr_t>(ranges) & 15))
    return set_error(CSA_E_BADARG, "csa_sample_ranges: bad sizes or unaligned ranges");
  const int threads = (n_frames + 1 + 31) / 32 * 32;
  sample_ranges_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(s_idx, s_count, block_n, n_frames, ranges);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "sample_ranges_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_gather_kv(const void* k, const void* v, int64_t ld_bytes, int32_t group_rows, int32_t n_groups,
                             const int32_t* s_idx, const int32_t* s_count, int32_t max_rows, void* k_out, void* v_out,
                             int64_t out_ld_bytes, int32_t out_group_rows, int32_t row_bytes, void* stream) {
  if (!k || !v || !s_idx || !s_count || !k_out || !v_out) return set_error(CSA_E_BADARG, "csa_gather_kv: null pointer");
  if (max_rows <= 0 || n_groups <= 0 || group_rows <= 0 || row_bytes <= 0 || (row_bytes & 15) || (ld_bytes & 15) ||
      (out_ld_bytes & 15) || out_group_rows < max_rows + CSA_TILE || ((reinterpret_cast<uintptr_t>(k) |
      reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(k_out) | reinterpret_cast<uintptr_t>(v_out)) & 15))
    return set_error(CSA_E_BADARG, "csa_gather_kv: sizes/strides/pointers must be positive and 16-byte aligned, "
                                   "out_group_rows >= max_rows + CSA_TILE");
  const int warps_per_block = 8;
  int grid = (max_rows + CSA_TILE + warps_per_block - 1) / warps_per_block;
  const int cap = 148 * 8 / (2 * n_groups) > 0 ? 148 * 8 / (2 * n_groups) : 1;
  if (grid > cap) grid = cap;
  gather_kv_kernel<<<dim3(grid, 2 * n_groups), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(k), static_cast<const uint8_t*>(v), ld_bytes, group_rows, s_idx, s_count, max_rows,
      static_cast<uint8_t*>(k_out), static_cast<uint8_t*>(v_out), out_ld_bytes, out_group_rows, row_bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "gather_kv_kernel: %s", cudaGetErrorString(e));
  return 0;
}
