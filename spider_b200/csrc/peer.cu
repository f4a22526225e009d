// Multi-GPU exchange of the sampled K/V rows over peer memory (NVLink / NVSwitch), fused with their gather.
//
// The reference is single-GPU.  In the sharded write pass (spider_b200/dist.py) the F frames of one CFG half — one
// key sequence in the reference, StoryDiffusion/Comic_Generation.py:148 — live on several GPUs, and every GPU needs
// K[S], V[S]: the sampled rows (StoryDiffusion/utils/gradio_utils.py:257-261) of ALL frames of its half.  Instead of
// gather -> NCCL all-gather -> compaction, each GPU stores its sampled rows straight into the S-ordered buffers of
// every GPU of the half (one read of the local rows, n_peers remote writes), then raises a flag there; the
// attention kernel of the receiver (attn_sm100.cu, `ready`) works on its local keys until the flag is up.
#include "ptx.cuh"
#include "csa_internal.h"

namespace csa {

struct PeerScatterParams {
  const uint8_t* k;
  const uint8_t* v;
  int64_t ld;
  const int32_t* idx;
  int32_t count, dst_row0, n_peers, self, row_bytes;
  uint8_t* k_dst[CSA_MAX_PEERS];
  uint8_t* v_dst[CSA_MAX_PEERS];
  int64_t dst_ld;
  uint32_t* ready[CSA_MAX_PEERS];
  uint32_t epoch, done_epoch;
  const uint32_t* done;
  uint32_t* counter;
  uint32_t* dbg;
  const int32_t* ranges;
  int32_t frames_per_peer, idx_adjust;
  const uint32_t* epoch_base;
};

// one warp per (row, K|V): the row is read once (16 B per lane per step) and stored to every peer
__global__ void __launch_bounds__(256) peer_scatter_kernel(const __grid_constant__ PeerScatterParams p) {
  // epochs: host values, or offsets to a device-resident base (a step replayed from a CUDA graph)
  const uint32_t base = p.epoch_base != nullptr ? *p.epoch_base : 0u;
  const int64_t done_epoch = p.epoch_base != nullptr
                                 ? static_cast<int64_t>(base) + static_cast<int32_t>(p.done_epoch)
                                 : static_cast<int64_t>(p.done_epoch);
  // the buffers being overwritten were last read by the peers' attention launches of `done_epoch`
  if (threadIdx.x < p.n_peers && threadIdx.x != p.self && done_epoch > 0)
    flag_wait_ge(p.done + threadIdx.x, static_cast<uint32_t>(done_epoch), 0x400 + threadIdx.x, p.dbg);
  __syncthreads();

  // geometry: from the host, or from the device-resident runs of the sampled list
  int count = p.count, dst_row0 = p.dst_row0, adjust = 0;
  const int32_t* idx = p.idx;
  if (p.ranges != nullptr) {
    const int4 first = __ldg(reinterpret_cast<const int4*>(p.ranges) + p.self * p.frames_per_peer);
    const int4 last = __ldg(reinterpret_cast<const int4*>(p.ranges) + (p.self + 1) * p.frames_per_peer - 1);
    dst_row0 = first.y;
    count = last.z - first.y;
    idx = p.idx + dst_row0;
    adjust = p.idx_adjust;
  }
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = p.row_bytes >> 4;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < 2 * count; i += gridDim.x * warps_per_block) {
    const int row = i >> 1;
    const bool is_v = i & 1;
    const uint4* s =
        reinterpret_cast<const uint4*>((is_v ? p.v : p.k) + static_cast<int64_t>(__ldg(idx + row) + adjust) * p.ld);
    const int64_t off = static_cast<int64_t>(dst_row0 + row) * p.dst_ld;
    for (int c = lane; c < vecs; c += 32) {
      const uint4 x = __ldg(s + c);
#pragma unroll
      for (int r = 0; r < CSA_MAX_PEERS; ++r) {
        if (r < p.n_peers) reinterpret_cast<uint4*>((is_v ? p.v_dst[r] : p.k_dst[r]) + off)[c] = x;
      }
    }
  }

  // publish (the grid-barrier idiom): the block's stores are ordered before thread 0's system-scope fence by the
  // CTA barrier (fences are cumulative), the fence before the counter, and the last block's release stores come
  // after every block's counter increment.  One fence per block instead of one per thread: a system-scope fence
  // waits for the acknowledgement of all of the thread's remote stores.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t prev = atomicAdd(p.counter, 1u);
    if (prev == gridDim.x - 1) {
      *p.counter = 0u;
      __threadfence_system();   // orders every block's stores (observed through the counter) before the flags
      for (int r = 0; r < p.n_peers; ++r) st_relaxed_sys(p.ready[r] + p.self, base + p.epoch);
    }
  }
}

struct PeerSignalParams {
  uint32_t* done[CSA_MAX_PEERS];
  int32_t n_peers, self;
  uint32_t epoch;
  const uint32_t* epoch_base;
};

__global__ void peer_signal_kernel(const __grid_constant__ PeerSignalParams p) {
  const int r = threadIdx.x;
  if (r < p.n_peers && r != p.self) {
    const uint32_t base = p.epoch_base != nullptr ? *p.epoch_base : 0u;
    __threadfence_system();
    st_release_sys(p.done[r] + p.self, base + p.epoch);
  }
}

__global__ void epoch_advance_kernel(uint32_t* base, uint32_t delta) { *base += delta; }

}  // namespace csa

using namespace csa;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int csa_peer_scatter_kv(const csa_peer_scatter_args_t* a, void* stream) {
  if (!a) return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: null args");
  if (a->struct_size != sizeof(csa_peer_scatter_args_t))
    return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: struct_size %u != %zu (ABI mismatch)", a->struct_size,
                     sizeof(csa_peer_scatter_args_t));
  if (a->n_peers < 1 || a->n_peers > CSA_MAX_PEERS || a->self < 0 || a->self >= a->n_peers)
    return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: n_peers %d / self %d", a->n_peers, a->self);
  if (a->count < 0 || a->dst_row0 < 0 || a->row_bytes <= 0 || (a->row_bytes & 15) || (a->ld_bytes & 15) ||
      (a->dst_ld_bytes & 15) || a->ld_bytes < a->row_bytes || a->dst_ld_bytes < a->row_bytes)
    return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: sizes and strides must be non-negative multiples of 16 bytes");
  if (!a->k || !a->v || !a->done || !a->counter || (a->count > 0 && !a->idx) || !al16(a->k) || !al16(a->v))
    return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: null or misaligned pointer");
  if (a->epoch == 0) return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: epochs start at 1");
  if (a->epoch_base != nullptr && (reinterpret_cast<uintptr_t>(a->epoch_base) & 3))
    return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: epoch_base must be 4-byte aligned");
  PeerScatterParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < a->n_peers; ++r) {
    if (!a->k_dst[r] || !a->v_dst[r] || !a->ready[r] || !al16(a->k_dst[r]) || !al16(a->v_dst[r]))
      return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: null or misaligned buffer of peer %d", r);
    p.k_dst[r] = static_cast<uint8_t*>(a->k_dst[r]);
    p.v_dst[r] = static_cast<uint8_t*>(a->v_dst[r]);
    p.ready[r] = a->ready[r];
  }
  p.k = static_cast<const uint8_t*>(a->k);
  p.v = static_cast<const uint8_t*>(a->v);
  p.ld = a->ld_bytes;
  p.idx = a->idx;
  p.count = a->count;
  p.dst_row0 = a->dst_row0;
  p.n_peers = a->n_peers;
  p.self = a->self;
  p.row_bytes = a->row_bytes;
  p.dst_ld = a->dst_ld_bytes;
  p.epoch = a->epoch;
  p.done_epoch = a->done_epoch;
  p.done = a->done;
  p.counter = a->counter;
  p.epoch_base = a->epoch_base;
  p.dbg = debug_record_devptr();
  if (a->ranges != nullptr) {
    if (a->frames_per_peer <= 0 || (reinterpret_cast<uintptr_t>(a->ranges) & 15) || !a->idx)
      return set_error(CSA_E_BADARG, "csa_peer_scatter_kv: ranges need frames_per_peer > 0, 16-byte alignment and idx");
    p.ranges = a->ranges;
    p.frames_per_peer = a->frames_per_peer;
    p.idx_adjust = a->idx_adjust;
  }
  const int warps_per_block = 8;
  int grid = (2 * a->count + warps_per_block - 1) / warps_per_block;
  const int cap = 148 * 4;  // all blocks resident (the first thing a block does is wait for the peers)
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  peer_scatter_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "peer_scatter_kernel: %s", cudaGetErrorString(e));
  return 0;
}

static int peer_signal_launch(uint32_t* const* done, int32_t n_peers, int32_t self, uint32_t epoch,
                              const uint32_t* epoch_base, void* stream) {
  if (!done || n_peers < 1 || n_peers > CSA_MAX_PEERS || self < 0 || self >= n_peers)
    return set_error(CSA_E_BADARG, "csa_peer_signal: bad arguments");
  PeerSignalParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < n_peers; ++r) {
    if (!done[r]) return set_error(CSA_E_BADARG, "csa_peer_signal: null flag array of peer %d", r);
    p.done[r] = done[r];
  }
  p.n_peers = n_peers;
  p.self = self;
  p.epoch = epoch;
  p.epoch_base = epoch_base;
  peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "peer_signal_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_peer_signal(uint32_t* const* done, int32_t n_peers, int32_t self, uint32_t epoch, void* stream) {
  return peer_signal_launch(done, n_peers, self, epoch, nullptr, stream);
}

extern "C" int csa_peer_signal_ex(const csa_peer_signal_args_t* a, void* stream) {
  if (!a) return set_error(CSA_E_BADARG, "csa_peer_signal_ex: null args");
  return peer_signal_launch(a->done, a->n_peers, a->self, a->epoch, a->epoch_base, stream);
}

extern "C" int csa_epoch_advance(uint32_t* epoch_base, uint32_t delta, void* stream) {
  if (!epoch_base || (reinterpret_cast<uintptr_t>(epoch_base) & 3))
    return set_error(CSA_E_BADARG, "csa_epoch_advance: null or misaligned epoch_base");
  epoch_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(epoch_base, delta);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "epoch_advance_kernel: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int csa_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaGetDevice: %s", cudaGetErrorString(e));
  if (peer_device == dev) return 0;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess || !can)
    return set_error(CSA_E_DEVICE, "device %d cannot access memory of device %d (%s)", dev, peer_device,
                     e != cudaSuccess ? cudaGetErrorString(e) : "no P2P path");
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------- CUDA IPC
// The exchange buffers are ordinary device allocations of the owning process; the other processes of the box map
// them with the legacy CUDA IPC calls, opened with THEIR device current so that the mapping (and the lazily enabled
// peer access) belongs to the device whose kernels will store through it.
extern "C" int csa_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out) {
  if (!ptr || !handle64_out || !offset_out) return set_error(CSA_E_BADARG, "csa_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range range_fn = []() -> PFN_range {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_range>(f);
  }();
  if (!range_fn) return set_error(CSA_E_DRIVER, "csa_ipc_export: cuMemGetAddressRange not available");
  CUresult cr = range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (cr != CUDA_SUCCESS) return set_error(static_cast<int>(cr), "cuMemGetAddressRange failed (CUresult %d)", (int)cr);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle64_out, &h, 64);
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}

extern "C" int csa_ipc_open(const void* handle64, void** base_out) {
  if (!handle64 || !base_out) return set_error(CSA_E_BADARG, "csa_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *base_out = p;
  return 0;
}

extern "C" int csa_ipc_close(void* base) {
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  return 0;
}
