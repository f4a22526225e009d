// Hand-written projection GEMM for B200 (sm_100a):  y[M,N] = alpha * x[M,K] w[N,K]^T (+ bias[N]), 16-bit in/out, fp32
// accumulation in TMEM — the projections either side of the attention (attn.to_q / to_k|to_v / to_out[0] of diffusers'
// Attention module, StoryDiffusion/Comic_Generation.py:155,164-165,185) with torch's nn.Linear layouts, so the weights
// are read where the module keeps them (both operands are K-major: exactly the Q K^T operand shape of the attention
// kernel).  Replaces the cuBLASLt calls of csa_linear for the shapes of the path (K % 64 == 0, N % 128 == 0).
//
// Persistent CTAs (one per SM, 192 threads) in CLUSTERS OF TWO = one CTA PAIR on the two SMs of a TPC, warp-specialised.
// With 128 x 128 tiles and nothing shared the kernel is bound by the L2 -> SM fabric (32 KB of operands per 256
// tensor-clocks and SM: measured 11.4 TB/s of operand traffic, 730 TFLOP/s).  Two forms of sharing the w tile:
//   * 256-wide tiles: ONE tcgen05.mma.cta_group::2 of M = 256, N = 256 per 16-wide k step, issued by the pair's leader:
//     each CTA holds its own 128 rows of x and HALF of the w tile's rows in shared memory (32 KB per stage instead
//     of 48: six stages instead of four, and a third less shared-memory read traffic per FLOP), its own 128 x 256
//     accumulator in TMEM.  Both producers credit the leader's `full` barrier (cta_group::2 TMA loads), the leader's
//     commits free the stage and publish the accumulator in both CTAs (multicast), and both CTAs' epilogue warps
//     release the accumulator on the leader's barrier.  q|k|v of the 32x32 layer: 58.4 us, 0.84 of the cuBLAS-measured
//     peak (cuBLASLt on the same shape: 63.1 us); before, with two M = 128 MMAs and the w tile multicast: 63.9 us.
//   * 128-wide tiles (N % 256 != 0 and too ragged for 256): two M = 128 MMAs, each CTA loads half of the w tile and
//     multicasts it to both.
//   warp 0      TMA producer
//   warp 1      TMEM allocator + MMA issue: four k steps per 64-wide k-block; two accumulators (2 x BN TMEM columns)
//               so that the epilogue of tile i overlaps the main loop of i+1.
//   warps 2-5   epilogue: tcgen05.ld (thread == output row), alpha / bias, 16-bit pack, staged through shared memory so
//               that every store instruction writes whole 128-byte lines; and the FUSED K/V GATHER of the
//               consistent-attention write pass: a row whose position in the sampled key list S is known
//               (`scatter_pos[row] >= 0`, csa_sample_positions) is also stored, in S order, into the K[S] / V[S]
//               buffer pair the attention kernel streams — the separate csa_gather_kv launch and its re-read of K and V
//               from HBM disappear.
// Tile pairs are dealt round-robin to the clusters, n fastest: clusters that run at the same time share their x
// tiles through L2.
#include <cuda_fp16.h>

#include <cstdlib>

#include "ptx.cuh"
#include "csa_internal.h"

namespace csa {

constexpr int kGM = 128, kGK = 64;
constexpr int kGTileA = kGM * kGK * 2;  // 16 KB
constexpr int kGThreads = 192;
constexpr int kGSmemBudget = 196608;    // operand ring
constexpr int kStageRow = 144;          // 128 bytes of a row's 64 columns + 16 bytes of padding (bank-conflict free)
constexpr int kStageBytes = 32 * kStageRow;

// k2Sm: the CTA pair runs one M = 256 MMA (cta_group::2) and each CTA keeps only ITS half of the w tile's rows;
// otherwise two M = 128 MMAs, each CTA holding the whole w tile (the halves are multicast).
template <int kBN, bool k2Sm>
struct GemmCfg {
  static constexpr int kTileB = (k2Sm ? kBN / 2 : kBN) * kGK * 2;
  static constexpr int kStages = kGSmemBudget / (kGTileA + kTileB);   // 4 / 6 (BN = 256 / 128), 6 / 8 with k2Sm
};

template <int kBN, bool k2Sm>
struct __align__(1024) GemmSmem {
  uint8_t a[GemmCfg<kBN, k2Sm>::kStages][kGTileA];
  uint8_t b[GemmCfg<kBN, k2Sm>::kStages][GemmCfg<kBN, k2Sm>::kTileB];
  uint64_t full[GemmCfg<kBN, k2Sm>::kStages], empty[GemmCfg<kBN, k2Sm>::kStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  alignas(16) uint8_t stage[4][kStageBytes];   // epilogue staging, one block of 32 rows x 128 B (+ pad) per epilogue warp
};

struct GemmParams {
  CUtensorMap tm_x, tm_w;
  void* y;
  int64_t ldy;
  void* y2;           // optional second output: columns >= y_split go to y2[:, col - y_split]
  int64_t ldy2;
  int32_t y_split;    // multiple of 32; == n when there is no second output
  const void* bias;
  float alpha;
  int32_t m, n, k;
  int32_t pairs_m, tiles_n;     // work items: (pair of vertically adjacent 128-row blocks, n tile)
  // Tail halving (256-wide pair-MMA kernel): the items of the last, partial wave are cut in two along n — work items
  // [n_whole, n_whole + 2 * n_halved) are halves (128 columns) of items [n_whole, n_whole + n_halved): the wave that
  // would keep only a part of the clusters busy for a whole tile time takes half of it
  int32_t n_whole, n_halved;
  CUtensorMap tm_w_half;        // w with a box of kBN / 4 rows (a CTA's half of a 128-wide tile)
  // fused gather of the sampled rows (optional)
  const int32_t* scatter_pos;   // [scatter_group_rows]: position of a row in S, or -1
  void* scatter_k;              // K[S]: rows g * scatter_dst_group_rows + pos, columns [0, split_col)
  void* scatter_v;              // V[S]: same rows, columns [split_col, n) shifted down by split_col
  int64_t scatter_ld;
  int32_t scatter_group_rows, scatter_dst_group_rows, split_col;
  int32_t scatter_col0;  // columns below it are not gathered (the q part of a stacked q|k|v weight)
  uint32_t* dbg;
  // the gather as the multi-GPU exchange (optional, n_dst > 0): sampled rows go to every GPU's K[S] / V[S]
  int32_t n_dst, self;
  uint16_t* k_dst[CSA_MAX_PEERS];
  uint16_t* v_dst[CSA_MAX_PEERS];
  uint32_t* ready[CSA_MAX_PEERS];
  uint32_t epoch, done_epoch;
  const uint32_t* done;
  uint32_t* counter;
  const uint32_t* epoch_base;
};

#define GSB(field) (sb + static_cast<uint32_t>(offsetof(Smem, field)))

// work item t -> item (pair of m blocks x n tile), first column and width of its share of the tile
template <int kBN>
__device__ __forceinline__ void gemm_work(const GemmParams& p, int t, int& item, int& n0, int& width) {
  if (t < p.n_whole) {
    item = t;
    n0 = (t % p.tiles_n) * kBN;
    width = kBN;
  } else {
    const int h = t - p.n_whole;
    item = p.n_whole + (h >> 1);
    n0 = (item % p.tiles_n) * kBN + (h & 1) * (kBN / 2);
    width = kBN / 2;
  }
}

template <bool kBF16, int kBN, bool k2Sm>
__global__ void __launch_bounds__(kGThreads, 1) csa_gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr int kCl = 2;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kCl) - 1u);
  using Smem = GemmSmem<kBN, k2Sm>;
  constexpr int kStages = GemmCfg<kBN, k2Sm>::kStages;
  constexpr int kTileB = GemmCfg<kBN, k2Sm>::kTileB;
  constexpr int kAcc = kBN <= 128 ? 128 : 256;   // TMEM columns between the two accumulators
  constexpr int kTmemCols = 2 * kAcc;            // a power of two
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared-memory window starts at the same offset in both CTAs of the cluster, so rounding it up gives
  // the same offsets too (multicast loads and commits address the peer by offset)
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(&sm);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // which 128-row block of the cluster's kCl
  const int cluster_id = blockIdx.x / kCl;
  const int n_clusters = gridDim.x / kCl;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_w);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(GSB(full) + 8u * i, 1);
      // every MMA warp of the cluster has consumed the stage (k2Sm: the leader's commit arrives once in each CTA)
      mbar_init(GSB(empty) + 8u * i, k2Sm ? 1 : kCl);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(GSB(acc_full) + 8u * i, 1);
      // one arrival per epilogue warp (k2Sm: of both CTAs, on the leader's barrier — its MMA writes both accumulators)
      mbar_init(GSB(acc_empty) + 8u * i, k2Sm ? 8 : 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (k2Sm) tmem_alloc_2sm<kTmemCols>(GSB(tmem_base)); else tmem_alloc<kTmemCols>(GSB(tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();    // the peer's barriers exist before anything of ours can reach them
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel of the stream
  pdl_launch_dependents();
  pdl_wait();
  const int n_items = p.n_whole + 2 * p.n_halved;   // work items (== pairs_m * tiles_n when nothing is halved)
  const int kblocks = p.k / kGK;

  if (warp == 0) {
    // ================================================================================================ producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = cluster_id; t < n_items; t += n_clusters) {
        int item, n0, width;
        gemm_work<kBN>(p, t, item, n0, width);
        const int m0 = ((item / p.tiles_n) * kCl + static_cast<int>(crank)) * kGM;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(GSB(empty) + 8u * st, ph ^ 1, 0x500, p.dbg);
          if constexpr (k2Sm) {
            // both CTAs' tiles are credited to the leader's barrier: its MMA consumes both
            const bool half = width < kBN;
            if (crank == 0) mbar_arrive_expect_tx(GSB(full) + 8u * st, 2 * (kGTileA + (half ? kTileB / 2 : kTileB)));
            tma_load_2d_2sm(&p.tm_x, GSB(a) + static_cast<uint32_t>(kGTileA) * st, GSB(full) + 8u * st, kb * kGK, m0);
            // my half of the w tile's rows stays here
            tma_load_2d_2sm(half ? &p.tm_w_half : &p.tm_w, GSB(b) + static_cast<uint32_t>(kTileB) * st,
                            GSB(full) + 8u * st, kb * kGK, n0 + static_cast<int>(crank) * (width / 2));
          } else {
            mbar_arrive_expect_tx(GSB(full) + 8u * st, kGTileA + kTileB);   // x tile + every slice of the w tile
            tma_load_2d(&p.tm_x, GSB(a) + static_cast<uint32_t>(kGTileA) * st, GSB(full) + 8u * st, kb * kGK, m0);
            // my slice of the w tile (rows [n0 + rank * BN/kCl, + BN/kCl)) goes to every CTA, at the slice's place
            tma_load_2d_mc(&p.tm_w, GSB(b) + static_cast<uint32_t>(kTileB) * st + crank * (kTileB / kCl),
                           GSB(full) + 8u * st, kb * kGK, n0 + static_cast<int>(crank) * (kBN / kCl), kMask);
          }
          if (++st == kStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================================================ MMA issue
    constexpr uint32_t idesc = make_idesc(k2Sm ? 2 * kGM : kGM, kBN, kBF16 ? 1 : 0, 0, 0);
    constexpr uint32_t idesc_half = make_idesc(k2Sm ? 2 * kGM : kGM, kBN / 2, kBF16 ? 1 : 0, 0, 0);
    int st = 0;
    uint32_t ph = 0, aph[2] = {0, 0};
    int it = 0;
    for (int t = cluster_id; t < n_items && !(k2Sm && crank != 0); t += n_clusters, ++it) {
      const int buf = it & 1;
      // the epilogue has drained this accumulator (first use of each buffer: nothing to wait for)
      if (it >= 2) {
        mbar_wait(GSB(acc_empty) + 8u * buf, aph[buf], 0x510 + buf, p.dbg);
        aph[buf] ^= 1;
      }
      tc_fence_after();
      const uint32_t tacc = tmem + buf * kAcc;
      const uint32_t idesc_t = (k2Sm && t >= p.n_whole) ? idesc_half : idesc;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(GSB(full) + 8u * st, ph, 0x520, p.dbg);
        tc_fence_after();
        const uint64_t da = make_sw128_desc(GSB(a) + static_cast<uint32_t>(kGTileA) * st);
        const uint64_t db = make_sw128_desc(GSB(b) + static_cast<uint32_t>(kTileB) * st);
        if (elect_one()) {
          if constexpr (k2Sm) {
#pragma unroll
            for (int kk = 0; kk < kGK / 16; ++kk) mma_ss_2sm(tacc, da + kk * 2, db + kk * 2, idesc_t, (kb | kk) ? 1u : 0u);
            tc_commit_2sm(GSB(empty) + 8u * st);                        // free the stage in both CTAs
            if (kb == kblocks - 1) tc_commit_2sm(GSB(acc_full) + 8u * buf);  // both CTAs' epilogues
          } else {
#pragma unroll
            for (int kk = 0; kk < kGK / 16; ++kk) mma_ss(tacc, da + kk * 2, db + kk * 2, idesc, (kb | kk) ? 1u : 0u);
            tc_commit_mc(GSB(empty) + 8u * st, kMask);  // free the stage in every CTA of the cluster
            if (kb == kblocks - 1) tc_commit(GSB(acc_full) + 8u * buf);
          }
        }
        __syncwarp();
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ================================================================================================ epilogue
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint32_t fph[2] = {0, 0};
    int it = 0;
    uint32_t epoch_base = 0;
    if (p.n_dst > 0) {
      // exchange: the buffers about to be overwritten were last read by the peers' attention launches of done_epoch
      epoch_base = p.epoch_base != nullptr ? *p.epoch_base : 0u;
      const int64_t done_epoch = p.epoch_base != nullptr
                                     ? static_cast<int64_t>(epoch_base) + static_cast<int32_t>(p.done_epoch)
                                     : static_cast<int64_t>(p.done_epoch);
      if (lane < p.n_dst && lane != p.self && done_epoch > 0)
        flag_wait_ge(p.done + lane, static_cast<uint32_t>(done_epoch), 0x540 + lane, p.dbg);
      __syncwarp();
    }
    for (int t = cluster_id; t < n_items; t += n_clusters, ++it) {
      const int buf = it & 1;
      int item, n0, width;
      gemm_work<kBN>(p, t, item, n0, width);
      const int m0 = ((item / p.tiles_n) * kCl + static_cast<int>(crank)) * kGM;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < p.m;
      mbar_wait(GSB(acc_full) + 8u * buf, fph[buf], 0x530 + buf, p.dbg);
      fph[buf] ^= 1;
      tc_fence_after();
      // fused gather: where this row goes in K[S] / V[S] (or nowhere)
      int pos = -1, g = 0;
      if (p.scatter_pos != nullptr && row_ok) {
        g = row / p.scatter_group_rows;
        pos = __ldg(p.scatter_pos + (row - g * p.scatter_group_rows));
      }
      // 64 columns (128 bytes per row) at a time: TMEM -> registers (thread == row) -> this warp's staging rows in
      // shared memory -> global, 8 lanes per row: every store instruction writes four complete 128-byte lines
      // (thread-per-row 16-byte stores cost 8x the L2 write transactions and ran the fused gather at 1.5 TB/s).
      const uint32_t stage_u32 = GSB(stage) + static_cast<uint32_t>(warp - 2) * kStageBytes;
#pragma unroll 1
      for (int cp = 0; cp < (kBN + 63) / 64; ++cp) {
        const int col0 = n0 + cp * 64;
        if (col0 >= p.n || cp * 64 >= width) break;      // ragged last n tile (N % kBN != 0) / a halved tail item
        const bool two = cp * 64 + 32 < kBN;             // false for the last 32 columns of a 160-wide tile
        uint32_t acc[2][32];
        tmem_ld32(tmem + lane_base + buf * kAcc + cp * 64, acc[0]);
        if (two) tmem_ld32(tmem + lane_base + buf * kAcc + cp * 64 + 32, acc[1]);
        tc_wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !two) break;
          float bv[32];
          if (p.bias != nullptr) {
            const uint4* bp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias) + col0 + h * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 w4 = __ldg(bp + i);
              const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if constexpr (kBF16) {
                  bv[8 * i + 2 * j] = __uint_as_float(ws[j] << 16);
                  bv[8 * i + 2 * j + 1] = __uint_as_float(ws[j] & 0xffff0000u);
                } else {
                  const __half2 h2 = *reinterpret_cast<const __half2*>(&ws[j]);
                  bv[8 * i + 2 * j] = __low2float(h2);
                  bv[8 * i + 2 * j + 1] = __high2float(h2);
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) bv[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v0 = fmaf(__uint_as_float(acc[h][8 * i + 2 * j]), p.alpha, bv[8 * i + 2 * j]);
              const float v1 = fmaf(__uint_as_float(acc[h][8 * i + 2 * j + 1]), p.alpha, bv[8 * i + 2 * j + 1]);
              w[j] = pack2<kBF16>(v0, v1);
            }
            // my row's 16-byte piece (h * 4 + i) of the 128-byte segment; rows are 144 bytes apart (conflict-free)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + lane * kStageRow + (h * 4 + i) * 16),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                         : "memory");
          }
        }
        __syncwarp();
        // where the 64 columns go: y or y2, and (for sampled rows) K[S] or V[S]
        const bool to_y2 = col0 >= p.y_split;
        uint16_t* ybase = to_y2 ? reinterpret_cast<uint16_t*>(p.y2) + (col0 - p.y_split)
                                : reinterpret_cast<uint16_t*>(p.y) + col0;
        const int64_t yld = to_y2 ? p.ldy2 : p.ldy;
        const bool gather_cols = p.scatter_pos != nullptr && col0 >= p.scatter_col0;
        const bool is_v = col0 >= p.split_col;
        const int scol = col0 - (is_v ? p.split_col : p.scatter_col0);
        uint16_t* sbase = reinterpret_cast<uint16_t*>(is_v ? p.scatter_v : p.scatter_k) + scol;
        const int seg = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3);               // row of this warp's 32 that the lane helps to store
          uint4 v4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w)
                       : "r"(stage_u32 + r * kStageRow + seg * 16)
                       : "memory");
          const int grow = m0 + q * 32 + r;
          const int rpos = __shfl_sync(0xffffffffu, pos, r);
          const int rg = __shfl_sync(0xffffffffu, g, r);
          if (grow < p.m && (two || seg < 4)) {
            reinterpret_cast<uint4*>(ybase + static_cast<int64_t>(grow) * yld)[seg] = v4;
            if (gather_cols && rpos >= 0) {
              if (p.n_dst == 0) {
                reinterpret_cast<uint4*>(sbase + (static_cast<int64_t>(rg) * p.scatter_dst_group_rows + rpos) * p.scatter_ld)[seg] = v4;
              } else {
                const int64_t off = static_cast<int64_t>(rpos) * p.scatter_ld + scol;
#pragma unroll
                for (int r = 0; r < CSA_MAX_PEERS; ++r)
                  if (r < p.n_dst) reinterpret_cast<uint4*>((is_v ? p.v_dst[r] : p.k_dst[r]) + off)[seg] = v4;
              }
            }
          }
        }
        __syncwarp();                                      // the staging rows are rewritten by the next 64 columns
      }
      // this accumulator may be overwritten by the main loop of the tile after next
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (k2Sm) mbar_arrive_cluster(GSB(acc_empty) + 8u * buf, 0); else mbar_arrive(GSB(acc_empty) + 8u * buf);
      }
    }
    if (p.n_dst > 0) {
      // publish (the grid-barrier idiom of peer.cu): the CTA's stores are ordered before one thread's system-scope
      // fence by the barrier of its four epilogue warps, the fence before the counter, and the last CTA's flag stores
      // come after every CTA's increment
      named_bar_sync(1, 128);
      if (warp == 2 && lane == 0) {
        __threadfence_system();
        const uint32_t prev = atomicAdd(p.counter, 1u);
        if (prev == gridDim.x - 1) {
          *p.counter = 0u;
          __threadfence_system();
          for (int r = 0; r < p.n_dst; ++r) st_relaxed_sys(p.ready[r] + p.self, epoch_base + p.epoch);
        }
      }
    }
  }

  __syncthreads();
  cluster_sync_all();    // nobody leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (k2Sm) tmem_dealloc_2sm<kTmemCols>(tmem); else tmem_dealloc<kTmemCols>(tmem);
  }
}

static int gemm_encode(CUtensorMap* tm, int dtype, const void* base, int64_t rows, int64_t cols, int64_t ld,
                       uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_tiled();
  if (!fn) return set_error(CSA_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {kGK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype == CSA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

__global__ void sample_positions_kernel(const int32_t* __restrict__ s_idx, const int32_t* __restrict__ s_count,
                                        int32_t n_cols, int32_t* __restrict__ pos) {
  // pos[c] = i if s_idx[i] == c (i < count), else -1.  The list is ascending: every column has one writer.
  const int count = min(*s_count, n_cols);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += gridDim.x * blockDim.x) {
    // binary search of column i in the list
    int lo = 0, hi = count;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(s_idx + mid) < i) lo = mid + 1; else hi = mid;
    }
    pos[i] = (lo < count && __ldg(s_idx + lo) == i) ? lo : -1;
  }
}

}  // namespace csa

using namespace csa;

extern "C" int csa_gemm_supported(int64_t m, int64_t n, int64_t k) {
  return (m > 0 && n > 0 && k > 0 && (n % 128) == 0 && (k % kGK) == 0 && m < (1ll << 30) && n < (1ll << 30)) ? 1 : 0;
}

template <int kBN, bool k2Sm>
static int gemm_launch(const csa_gemm_args_t* a, GemmParams& p, int dev, int sms, void* stream) {
  constexpr int kCl = 2;
  int rc;
  if ((rc = gemm_encode(&p.tm_x, a->dtype, a->x, a->m, a->k, a->ldx, kGM))) return rc;
  if ((rc = gemm_encode(&p.tm_w, a->dtype, a->w, a->n, a->k, a->ldw, kBN / kCl))) return rc;
  p.pairs_m = static_cast<int32_t>((a->m + kCl * kGM - 1) / (kCl * kGM));
  p.tiles_n = static_cast<int32_t>((a->n + kBN - 1) / kBN);   // a ragged last tile reads zero rows of w (TMA OOB fill)
  const int n_items = p.pairs_m * p.tiles_n;
  p.n_whole = n_items;
  p.n_halved = 0;
  const size_t smem = sizeof(GemmSmem<kBN, k2Sm>) + 1024;
  auto kern = a->dtype == CSA_DTYPE_BF16 ? csa_gemm_kernel<true, kBN, k2Sm> : csa_gemm_kernel<false, kBN, k2Sm>;
  static bool smem_set[64][2];
  static int max_clusters[64][2];
  const int ki = a->dtype == CSA_DTYPE_BF16 ? 1 : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(kGThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dev < 0 || dev >= 64 || !smem_set[dev][ki]) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess)
      return set_error(static_cast<int>(ce), "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(ce));
    // how many clusters of this size the device can hold at once (GPC boundaries may strand a few SMs)
    int nc = sms / kCl;
    cfg.gridDim = dim3(static_cast<unsigned>(nc * kCl), 1, 1);
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0 && q < nc) nc = q;
    cudaGetLastError();
    if (dev >= 0 && dev < 64) {
      smem_set[dev][ki] = true;
      max_clusters[dev][ki] = nc;
    }
  }
  const int cap = (dev >= 0 && dev < 64) ? max_clusters[dev][ki] : sms / kCl;
  if constexpr (k2Sm && kBN == 256) {
    // tail halving: a last wave that fills at most half of the clusters is dealt as twice as many half-width items
    static const bool halve = []() {
      const char* e = getenv("CSA_GEMM_TAIL_HALVING");
      return !(e && atoi(e) == 0);
    }();
    const int rem = n_items % cap;
    if (halve && n_items > cap && rem > 0 && 2 * rem <= cap) {
      if ((rc = gemm_encode(&p.tm_w_half, a->dtype, a->w, a->n, a->k, a->ldw, kBN / 4))) return rc;
      p.n_whole = n_items - rem;
      p.n_halved = rem;
    }
  }
  const int n_work = p.n_whole + 2 * p.n_halved;
  const int clusters = n_work < cap ? n_work : cap;
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * kCl), 1, 1);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t ce = cudaLaunchKernelEx(&cfg, kern, p);
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "csa_gemm_kernel launch: %s", cudaGetErrorString(ce));
  return 0;
}

extern "C" int csa_gemm(const csa_gemm_args_t* a, void* stream) {
  if (!a) return set_error(CSA_E_BADARG, "csa_gemm: null args");
  if (a->struct_size != sizeof(csa_gemm_args_t))
    return set_error(CSA_E_BADARG, "csa_gemm: struct_size %u != %zu (ABI mismatch)", a->struct_size,
                     sizeof(csa_gemm_args_t));
  if (a->dtype != CSA_DTYPE_F16 && a->dtype != CSA_DTYPE_BF16) return set_error(CSA_E_BADARG, "csa_gemm: dtype %d", a->dtype);
  if (!csa_gemm_supported(a->m, a->n, a->k))
    return set_error(CSA_E_SHAPE, "csa_gemm: needs N %% 128 == 0 and K %% 64 == 0, got M=%lld N=%lld K=%lld",
                     (long long)a->m, (long long)a->n, (long long)a->k);
  if (a->ldx < a->k || a->ldw < a->k || (a->y2 == nullptr && a->ldy < a->n) || ((a->ldx | a->ldw | a->ldy) & 7))
    return set_error(CSA_E_BADARG, "csa_gemm: leading dimensions must cover the rows and be multiples of 8 elements");
  if (!a->x || !a->w || !a->y) return set_error(CSA_E_BADARG, "csa_gemm: null pointer");
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) | reinterpret_cast<uintptr_t>(a->y) |
       reinterpret_cast<uintptr_t>(a->bias)) & 15)
    return set_error(CSA_E_BADARG, "csa_gemm: pointers must be 16-byte aligned");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.y = a->y;
  p.ldy = a->ldy;
  p.y2 = a->y2;
  p.ldy2 = a->ldy2;
  p.y_split = static_cast<int32_t>(a->n);
  if (a->y2 != nullptr) {
    if (a->y_split <= 0 || a->y_split >= a->n || (a->y_split % 32) || a->ldy < a->y_split ||
        a->ldy2 < a->n - a->y_split || (a->ldy2 & 7) || (reinterpret_cast<uintptr_t>(a->y2) & 15))
      return set_error(CSA_E_BADARG, "csa_gemm: bad second output (y_split must be a multiple of 32 inside (0, N), "
                                     "leading dimensions must cover their column ranges)");
    p.y_split = a->y_split;
  }
  p.bias = a->bias;
  p.alpha = a->alpha;
  p.m = static_cast<int32_t>(a->m);
  p.n = static_cast<int32_t>(a->n);
  p.k = static_cast<int32_t>(a->k);
  if (a->scatter_pos != nullptr && a->exchange == nullptr) {
    if (!a->scatter_k || !a->scatter_v || a->scatter_group_rows <= 0 || a->scatter_dst_group_rows <= 0 ||
        a->split_col <= 0 || (a->split_col % 32) || a->split_col > a->n || (a->scatter_ld & 7) ||
        a->scatter_col0 < 0 || (a->scatter_col0 % 32) || a->scatter_col0 > a->split_col ||
        ((reinterpret_cast<uintptr_t>(a->scatter_k) | reinterpret_cast<uintptr_t>(a->scatter_v)) & 15) ||
        (a->m % a->scatter_group_rows))
      return set_error(CSA_E_BADARG, "csa_gemm: bad scatter arguments (scatter_col0 <= split_col, both multiples of 32, "
                                     "M a multiple of scatter_group_rows, buffers 16-byte aligned)");
    p.scatter_pos = a->scatter_pos;
    p.scatter_k = a->scatter_k;
    p.scatter_v = a->scatter_v;
    p.scatter_ld = a->scatter_ld;
    p.scatter_group_rows = a->scatter_group_rows;
    p.scatter_dst_group_rows = a->scatter_dst_group_rows;
    p.split_col = a->split_col;
    p.scatter_col0 = a->scatter_col0;
  }
  if (a->exchange != nullptr) {
    const csa_peer_exchange_t* x = a->exchange;
    if (x->struct_size != sizeof(csa_peer_exchange_t))
      return set_error(CSA_E_BADARG, "csa_gemm: exchange struct_size %u != %zu (ABI mismatch)", x->struct_size,
                       sizeof(csa_peer_exchange_t));
    if (a->scatter_pos == nullptr || a->scatter_group_rows <= 0 || a->m != a->scatter_group_rows || a->split_col <= 0 ||
        (a->split_col % 32) || a->split_col > a->n || a->scatter_col0 < 0 || (a->scatter_col0 % 32) ||
        a->scatter_col0 > a->split_col)
      return set_error(CSA_E_BADARG, "csa_gemm: the exchange needs scatter_pos, one group (M == scatter_group_rows) and "
                                     "scatter_col0 <= split_col, both multiples of 32");
    if (x->n_peers < 1 || x->n_peers > CSA_MAX_PEERS || x->self < 0 || x->self >= x->n_peers || x->epoch == 0 ||
        !x->done || !x->counter || (x->dst_ld & 7) || x->dst_ld < a->split_col - a->scatter_col0 ||
        (x->epoch_base != nullptr && (reinterpret_cast<uintptr_t>(x->epoch_base) & 3)))
      return set_error(CSA_E_BADARG, "csa_gemm: bad exchange (n_peers %d, self %d, epochs start at 1, dst_ld multiple "
                                     "of 8 covering the columns)", x->n_peers, x->self);
    p.scatter_pos = a->scatter_pos;
    p.scatter_group_rows = a->scatter_group_rows;
    p.scatter_dst_group_rows = 0;
    p.split_col = a->split_col;
    p.scatter_col0 = a->scatter_col0;
    p.scatter_ld = x->dst_ld;
    p.n_dst = x->n_peers;
    p.self = x->self;
    for (int r = 0; r < x->n_peers; ++r) {
      if (!x->k_dst[r] || !x->v_dst[r] || !x->ready[r] ||
          ((reinterpret_cast<uintptr_t>(x->k_dst[r]) | reinterpret_cast<uintptr_t>(x->v_dst[r])) & 15))
        return set_error(CSA_E_BADARG, "csa_gemm: null or misaligned exchange buffer of peer %d", r);
      p.k_dst[r] = static_cast<uint16_t*>(x->k_dst[r]);
      p.v_dst[r] = static_cast<uint16_t*>(x->v_dst[r]);
      p.ready[r] = x->ready[r];
    }
    p.epoch = x->epoch;
    p.done_epoch = x->done_epoch;
    p.done = x->done;
    p.counter = x->counter;
    p.epoch_base = x->epoch_base;
  }
  p.dbg = debug_record_devptr();
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "cudaGetDevice: %s", cudaGetErrorString(ce));
  const int sms = sm_count(dev);
  if (sms <= 1) return set_error(CSA_E_DEVICE, "csa_gemm: device %d is not sm_100", dev);
  // 128 x 256 tiles halve the operand traffic per FLOP; 128 x 128 when N is not a multiple of 256 (N = 640) or when
  // the wider tile would leave most of the chip without work
  const int64_t blocks_m = (a->m + kGM - 1) / kGM;
  // wide tiles also for N % 256 == 128 (the last n tile is half empty) as long as that wastes < 1/8 of the MMAs
  // (N = 1920: yes; N = 640: no — measured 45.5 us wide vs 42.0 us narrow)
  const int64_t tiles_wide = (a->n + 255) / 256;
  const bool wide = tiles_wide * 256 * 8 <= a->n * 9 && blocks_m * tiles_wide >= sms / 2;
  // CSA_GEMM_2SM=0: two M = 128 MMAs with the w tile multicast instead of one M = 256 MMA of the CTA pair (A/B knob)
  static const bool two_sm = []() {
    const char* e = getenv("CSA_GEMM_2SM");
    return !(e && atoi(e) == 0);
  }();
  // 160-wide tiles (pair MMA, N = 160) for plain projections whose N they divide when they need fewer tile-waves than
  // the choice above: N = 1280 is 2.16 waves of 256-wide tiles at M = 8192 (three passes of 256 columns) but 3.46 of
  // 160-wide ones (four passes of 160); N = 640 is four exact tiles instead of five 128-wide ones.  The model: passes
  // x tile width, +5 % for the 160-wide tile's extra operand traffic per FLOP, +15 % for the 128-wide one's.
  // OPT-IN (CSA_GEMM_BN160=1): timed alone the output projection of the 32x32 layer gains (29.4 -> 26.8 us, cuBLASLt
  // 26.0), but inside the power-capped denoise step it LOSES (14.80 vs 14.72 ms per step, three alternating runs,
  // profiles/r03h_gemm_bn160.log): 30 % more operand traffic per launch costs more than the shorter tail saves.
  static const bool allow160 = []() {
    const char* e = getenv("CSA_GEMM_BN160");
    return e && atoi(e) == 1;
  }();
  if (allow160 && two_sm && a->y2 == nullptr && a->scatter_pos == nullptr && a->n % 160 == 0) {
    const int64_t clusters = sms / 2;
    const int64_t pairs = (blocks_m + 1) / 2;
    auto passes = [&](int64_t tiles) { return (pairs * tiles + clusters - 1) / clusters; };
    const double cost160 = static_cast<double>(passes(a->n / 160)) * 160 * 1.05;
    const double cost_now = wide ? static_cast<double>(passes(tiles_wide)) * 256
                                 : static_cast<double>(passes(a->n / 128)) * 128 * 1.15;
    if (cost160 < cost_now * 0.97) return gemm_launch<160, true>(a, p, dev, sms, stream);
  }
  if (wide) return two_sm ? gemm_launch<256, true>(a, p, dev, sms, stream) : gemm_launch<256, false>(a, p, dev, sms, stream);
  // 128-wide tiles stay on the multicast form: the pair MMA with N = 128 measured slower (41.6 vs 37.5 us on
  // 32768 x 640 x 640; CSA_GEMM_2SM=2 forces it)
  static const bool two_sm_narrow = []() {
    const char* e = getenv("CSA_GEMM_2SM");
    return e && atoi(e) == 2;
  }();
  return two_sm_narrow ? gemm_launch<128, true>(a, p, dev, sms, stream) : gemm_launch<128, false>(a, p, dev, sms, stream);
}

extern "C" int csa_sample_positions(const int32_t* s_idx, const int32_t* s_count, int32_t n_cols, int32_t* pos,
                                    void* stream) {
  if (!s_idx || !s_count || !pos || n_cols <= 0) return set_error(CSA_E_BADARG, "csa_sample_positions: bad arguments");
  const int block = 256;
  int grid = (n_cols + block - 1) / block;
  if (grid > 148 * 4) grid = 148 * 4;
  sample_positions_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(s_idx, s_count, n_cols, pos);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "sample_positions_kernel: %s", cudaGetErrorString(e));
  return 0;
}
