// Consistent Self-Attention forward for B200 (sm_100a): flash attention over compacted key lists.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=dense_bool) of the reference
// (StoryDiffusion/Comic_Generation.py:175-177 and :248-250).  Because the sampled mask row is shared by every query
// of a frame (StoryDiffusion/utils/gradio_utils.py:257-286), the dense mask is never built: each (CFG half, frame)
// attends a *key list* = TMA-gathered sampled rows + contiguous rows, see include/csa_b200.h.
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//   warp 0      producer   TMA: Q tiles (box 64x128), K/V tiles either as one tiled load (contiguous run of keys) or
//                          as 32 lane-parallel `tile::gather4` loads (4 indexed rows each) into the same 128B-swizzled
//                          16 KB stage
//   warps 1, 3  MMA issue  one lane per Q tile issues tcgen05.mma:  S = Q K^T  (SS, M128 N128 K64),  O += P V (TS: P read from
//                          TMEM as the A operand, V read MN-major from the very tile TMA wrote, M128 N64 K128)
//   warp 2      TMEM allocator (512 columns: S0 S1 | O0 O1 | P0 P1)
//   warps 4-7   softmax for Q tile 0 (thread == query row, no shuffles), warps 8-11 for Q tile 1.
//               online softmax in the exp2 domain with lazy (threshold 8) rescaling of O, P written to TMEM as
//               16-bit, final 1/l normalisation and store.
// Two 128-row Q tiles are ping-ponged so that the tensor core works on one tile while the other is in softmax.
#include <type_traits>

#include "ptx.cuh"
#include "csa_internal.h"

// How many of every 16 element pairs take the polynomial exp2 (FMA/ALU pipes, ptx.cuh poly_exp2_fast_x2: clamp-free,
// degree 2 for a bf16 P, degree 3 for an fp16 P) instead of MUFU.EX2.  ptxas paces a warp's exp loop at one MUFU
// every 8 cycles and fills the gaps (decoded stall fields: 16.1 clk per pair with or without the polynomial), so the
// gain is small: timed alone (a burst at full clocks) 4/16 -> 864 vs 853 TFLOP/s on the 64x64 layer, 6/16 and more
// lose (profiles/r02h_classic_poly_sweep.log).  Inside the denoise step the GPU runs at its POWER CAP (1.70-1.78 of
// 1.965 GHz) and the extra FMA-pipe instructions cost more clock than they save MUFU time: 14.59-14.60 ms per step
// with 0/16 against 14.67-14.78 with 4/16, 15.06 with 6/16 (profiles/r03l_poly_in_step.log) — hence the default 0;
// -DCSA_POLY_PAIRS=4 is the burst-optimal build.  Round 2 also built and measured two reorganisations that were meant
// to lift the MUFU wall and did not: 16 softmax warps with the keys of a tile split between two warps per row
// (issue-bound: 475 instructions per tile and warp, profiles/experiments/r02_16warp_keysplit.patch) and a lazily
// updated reference point that takes the row max off the latency chain (the exp phase grows by what the chain loses,
// profiles/experiments/r02_lazy_max.patch).
#ifndef CSA_POLY_PAIRS
#define CSA_POLY_PAIRS 0
#endif
// Exp-phase ping-pong: the two softmax warps that share an SM sub-partition (same TMEM lane quarter, Q tile 0 and
// Q tile 1) hand a token back and forth through a pair of named barriers, so that one exponentiates (MUFU-bound)
// while the other waits for its PV, pulls the next scores and takes the row max.  Without it the two Q-tile streams
// run in phase (both start on the same K tile) and the MUFU idles whenever both are outside the exp phase.
#ifndef CSA_PINGPONG
#define CSA_PINGPONG 1
#endif
// The token is handed over after this 32-key chunk (0..3) of the exp phase has been issued: 3 = strict alternation,
// smaller = the two exp phases overlap, which puts two warps on the sub-partition's MUFU at once (a single warp
// cannot keep it busy: it issues nothing else while its own MUFU instruction is dispatched).  Measured
// (profiles/r01f_token_chunk_sweep.log, step-equivalent TFLOP/s): 0 -> 715, 1 -> 769, 2 -> 747, 3 -> 736.
#ifndef CSA_TOKEN_CHUNK
#define CSA_TOKEN_CHUNK 1
#endif
// Organisations that were built, measured slower on B200 and removed ("what did not work"):
// two threads per score row (row max exchanged through shared memory), software-pipelined score loads inside the exp
// phase (needs S double-buffered, TMEM is full), deferred PV wait inside the exp phase (a token holder that stalls
// blocks both Q tiles), 64-key half tiles with double-buffered S/P (per-step barrier latencies dominate).

// Timeline trace (debug builds only, -DCSA_TRACE=1): lane 0 of a few warps of CTA 0 records (event, clock) pairs
// into a device buffer set with csa_debug_set_trace(); tools/trace_timeline.py turns them into per-tile latencies.
#ifndef CSA_TRACE
#define CSA_TRACE 0
#endif

namespace csa {

#if CSA_TRACE
__device__ unsigned long long* g_trace = nullptr;
constexpr int kTraceSlots = 4;       // softmax warp of Q tile 0 / 1 (lane quarter 0), MMA warp of Q tile 0 / 1
constexpr int kTraceEvents = 8192;   // per slot
struct Tracer {
  unsigned long long* base;
  uint32_t n;
  __device__ __forceinline__ void init(int slot, bool on) {
    base = (on && g_trace != nullptr && blockIdx.x == 0) ? g_trace + slot * kTraceEvents : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void ev(uint32_t id) {
    if (base != nullptr && n < kTraceEvents) {
      base[n++] = (static_cast<unsigned long long>(id) << 48) | (static_cast<unsigned long long>(clock64()) & 0xffffffffffffull);
    }
  }
};
#define TRACE_INIT(slot, on) Tracer tracer; tracer.init(slot, on)
#define TRACE(id) tracer.ev(id)
// per-CTA unit boundaries (globaltimer, ns) behind the event slots: [blockIdx.x][64] — tools/trace_ctas.py
constexpr int kCtaMarks = 64;
#define TRACE_CTA(on, k, tag)                                                                                  \
  do {                                                                                                         \
    if ((on) && g_trace != nullptr && (k) < kCtaMarks)                                                         \
      g_trace[kTraceSlots * kTraceEvents + blockIdx.x * kCtaMarks + (k)] =                                     \
          (static_cast<unsigned long long>(tag) << 56) | (globaltimer_ns() & 0xffffffffffffffull);             \
  } while (0)
#else
#define TRACE_INIT(slot, on)
#define TRACE(id)
#define TRACE_CTA(on, k, tag)
#endif

constexpr int kBM = 128;  // query rows per Q tile
constexpr int kBN = 128;  // keys per K/V tile
constexpr int kHD = 64;   // head dim
constexpr int kTileBytes = kBN * kHD * 2;
constexpr int kKStages = 4;
constexpr int kVStages = 4;
constexpr int kThreads = 384;
constexpr int kRegsCtl = 72;       // producer / MMA / allocator warps after setmaxnreg.dec
constexpr int kRegsSoftmax = 216;  // softmax warps after setmaxnreg.inc: 72*128 + 216*256 == 168*384 (the CTA's pool)
constexpr int kSoftmaxWarpsPerTile = (kThreads - 128) / 64;  // arrivals on s_free / p_ready per Q tile

// TMEM column map (fp32 columns)
constexpr uint32_t kColS = 0;    // S0 at 0, S1 at 128
constexpr uint32_t kColO = 256;  // O0 at 256, O1 at 320
constexpr uint32_t kColP = 384;  // P0 at 384, P1 at 448 (128 16-bit values = 64 columns)

// pair i of a 16-pair chunk goes to the polynomial iff it is one of `n` evenly spread slots
__host__ __device__ constexpr bool poly_pair(int i, int n) { return ((i + 1) * n) / 16 != (i * n) / 16; }
#ifndef CSA_POLY_PAIRS_F16
#define CSA_POLY_PAIRS_F16 CSA_POLY_PAIRS
#endif

struct __align__(1024) AttnSmem {
  uint8_t q[2][kTileBytes];
  uint8_t k[kKStages][kTileBytes];
  uint8_t v[kVStages][kTileBytes];
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[kKStages], k_empty[kKStages];
  uint64_t v_full[kVStages], v_empty[kVStages];
  uint64_t s_full[2], s_free[2], p_ready[2], o_done[2];
  uint32_t tmem_base;
  uint32_t merge_flag;  // split units: this CTA delivered the last piece and merges (written by one thread)
  // per softmax warp: where the current unit's result goes {q_row0, head, Q pair, piece, split unit} — written at the
  // unit's start, read in its epilogue, so that nothing of it lives in registers across the tile loop
  int32_t epi[8][8];
};

struct AttnKernelParams {
  CUtensorMap tm_q;    // box 64 x 128
  CUtensorMap tm_ka;   // box 64 x 128, source A
  CUtensorMap tm_va;
  CUtensorMap tm_kag;  // box 64 x 1 (gather4), source A
  CUtensorMap tm_vag;
  CUtensorMap tm_kb;   // box 64 x 128, source B
  CUtensorMap tm_vb;
  void* o;
  int64_t o_ld;
  const int32_t* idx;
  const int32_t* counts;
  int64_t idx_stride;
  int32_t heads, n_groups, n_frames, n_q, n_qpairs, n_units;
  int32_t a_group_rows, b_group_rows;
  int32_t list_base, list_step, g_adjust;
  int32_t ca_start, ca_step, ca_len;
  int32_t cb_start, cb_step, cb_len;
  const int32_t* ranges;  // optional: per list 4 ints {start1, len1, start2, len2}: two runs of A (device-resident)
  int32_t range_base, range_step;
  float scale_log2;
  uint32_t* dbg;  // host-mapped watchdog record (may be null)
  // Tail split: units [0, n_whole) are processed whole; each of the remaining units (fewer than one per CTA) is cut
  // into `split` pieces along its key tiles, scheduled as units [n_whole, n_sched); the pieces leave unnormalised
  // partials in `ws` and the CTA that delivers a unit's last piece merges them.
  int32_t n_whole, split, n_sched;
  float* ws;           // split pieces: [piece][256 rows][64] O, then [256] m, then [256] l
  uint32_t* ws_count;  // one arrival counter per split unit (zero between launches)
  // Multi-GPU: rows [ready_bounds[r], ready_bounds[r+1]) of every group of A are delivered by peer rank r straight
  // into this GPU's memory (csa_peer_scatter_kv); they may be read once ready[r] >= ready_epoch.
  const uint32_t* ready;
  uint32_t ready_epoch;
  const uint32_t* epoch_base;  // optional: device word added to ready_epoch (steps replayed from a CUDA graph)
  int32_t ready_n;
  int32_t ready_bounds[CSA_MAX_PEERS + 1];
  int32_t ready_fpp;  // > 0: bounds come from `ranges` (frames per peer), not from ready_bounds
  // optional: the launch's last CTA tells the peers that the exchange buffers of this epoch have been read
  uint32_t* done_dst[CSA_MAX_PEERS];
  uint32_t* done_counter;
  int32_t peer_self;
  int32_t b_first;  // key order of a unit: contiguous B segment first, then the runs of A
};

constexpr int kPieceFloats = 2 * kBM * (kHD + 2);  // partial of one piece of a Q-tile pair
constexpr int kMaxSplit = 8;
constexpr int kWsHeaderBytes = 4096;  // arrival counters (one uint32 per split unit), zero between launches

constexpr int kMaxSeg = 3;  // contiguous runs per unit: up to two runs of source A and one of source B

struct Unit {
  int g, f, h, qp;
  int q_row0;  // first query row (in the q / o matrices) of the pair of Q tiles
  int ng;      // gathered keys
  const int32_t* gidx;
  int a_base;  // row offset of this group in A
  int seg_row[kMaxSeg], seg_len[kMaxSeg], seg_tiles[kMaxSeg];
  int seg_in_b;  // bit i set: segment i lives in source B (else in A)
  int tg, total;
  int t0, nt;     // this CTA's share of the unit's key tiles: [t0, t0 + nt) (the whole unit unless split)
  int piece;      // index of the partial in the workspace, or -1 for a whole unit
  int split_unit; // index of the unit among the split ones (arrival counter), or -1
};

__device__ __forceinline__ Unit decode_unit(const AttnKernelParams& p, int sched) {
  Unit w;
  int u = sched, pc = 0;
  w.piece = -1;
  w.split_unit = -1;
  if (sched >= p.n_whole) {
    const int v = sched - p.n_whole;
    w.split_unit = v / p.split;
    pc = v - w.split_unit * p.split;
    w.piece = v;
    u = p.n_whole + w.split_unit;
  }
  w.qp = u % p.n_qpairs;
  int r = u / p.n_qpairs;
  w.h = r % p.heads;
  r /= p.heads;
  w.f = r % p.n_frames;
  w.g = r / p.n_frames;
  w.q_row0 = (w.g * p.n_frames + w.f) * p.n_q + w.qp * (2 * kBM);
  w.ng = 0;
  w.gidx = nullptr;
  if (p.list_base >= 0) {
    const int list = p.list_base + w.f * p.list_step;
    int c = __ldg(p.counts + list) + p.g_adjust;
    w.ng = c > 0 ? c : 0;
    w.gidx = p.idx + static_cast<int64_t>(list) * p.idx_stride;
  }
  w.a_base = w.g * p.a_group_rows;
  int a_row[2], a_len[2];
  if (p.ranges != nullptr) {
    const int4 rg = __ldg(reinterpret_cast<const int4*>(p.ranges) + (p.range_base + w.f * p.range_step));
    a_row[0] = w.a_base + rg.x;
    a_len[0] = rg.y > 0 ? rg.y : 0;
    a_row[1] = w.a_base + rg.z;
    a_len[1] = rg.w > 0 ? rg.w : 0;
  } else {
    a_row[0] = w.a_base + p.ca_start + w.f * p.ca_step;
    a_len[0] = p.ca_len;
    a_row[1] = 0;
    a_len[1] = 0;
  }
  const int b_row = w.g * p.b_group_rows + p.cb_start + w.f * p.cb_step;
  // segment order = key order of the unit: A run 1, A run 2, B — or B first (multi-GPU: the frame's own block is
  // local, so its tiles are worked on while the sampled rows of the other GPUs are still in flight)
  const bool bf = p.b_first != 0;
  w.seg_row[0] = bf ? b_row : a_row[0];
  w.seg_len[0] = bf ? p.cb_len : a_len[0];
  w.seg_row[1] = bf ? a_row[0] : a_row[1];
  w.seg_len[1] = bf ? a_len[0] : a_len[1];
  w.seg_row[2] = bf ? a_row[1] : b_row;
  w.seg_len[2] = bf ? a_len[1] : p.cb_len;
  w.seg_in_b = bf ? 1 : 4;
  w.tg = (w.ng + kBN - 1) / kBN;
  w.total = w.tg;
#pragma unroll
  for (int i = 0; i < kMaxSeg; ++i) {
    w.seg_tiles[i] = (w.seg_len[i] + kBN - 1) / kBN;
    w.total += w.seg_tiles[i];
  }
  w.t0 = 0;
  w.nt = w.total;
  if (w.piece >= 0) {
    w.t0 = (w.total * pc) / p.split;
    w.nt = (w.total * (pc + 1)) / p.split - w.t0;
  }
  return w;
}

// tile t of the unit -> segment (-1 = gathered) and tile index inside it
__device__ __forceinline__ void locate_tile(const Unit& w, int t, int& seg, int& local) {
  seg = -1;
  local = t;
  if (t < w.tg) return;
  local -= w.tg;
#pragma unroll
  for (int i = 0; i < kMaxSeg; ++i) {
    if (seg < 0) {
      if (local < w.seg_tiles[i]) {
        seg = i;
      } else {
        local -= w.seg_tiles[i];
      }
    }
  }
}

// Walks the key tiles of a unit in order (gathered list, A run 1, A run 2, B run) and yields the number of valid
// keys of each: a few integer ops per tile instead of re-deriving the segment from the tile index.
struct TileWalker {
  int l1, l2, l3;  // lengths of the segments after the current one
  int rem;         // keys left in the current segment
  __device__ __forceinline__ void skip_empty() {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (rem <= 0) {
        rem = l1;
        l1 = l2;
        l2 = l3;
        l3 = 0;
      }
    }
  }
  __device__ __forceinline__ void init(const Unit& w) {
    rem = w.ng;
    l1 = w.seg_len[0];
    l2 = w.seg_len[1];
    l3 = w.seg_len[2];
    skip_empty();
  }
  __device__ __forceinline__ int next() {
    const int v = rem < kBN ? rem : kBN;
    rem -= kBN;
    skip_empty();
    return v;
  }
};

// number of valid keys in tile t of the unit
__device__ __forceinline__ int tile_valid(const Unit& w, int t) {
  int seg, local;
  locate_tile(w, t, seg, local);
  const int len = seg < 0 ? w.ng : (seg == 0 ? w.seg_len[0] : (seg == 1 ? w.seg_len[1] : w.seg_len[2]));
  const int rem = len - local * kBN;
  return rem < kBN ? rem : kBN;
}

// 32-bit shared-window address of a field of AttnSmem, given the address `sb` of the (aligned) struct
#define SB(field) (sb + static_cast<uint32_t>(offsetof(AttnSmem, field)))

// =============================================================================================== producer warp
// TMA producer shared by both kernel organisations: Q tiles (box 64x128) and 128-key K/V tiles, either as one tiled
// load (contiguous run of keys) or as 32 lane-parallel `tile::gather4` loads into the same 128B-swizzled stage.
__device__ __forceinline__ void producer_warp(const AttnKernelParams& p, const uint32_t sb, const int lane) {
  int ks = 0, vs = 0;
  uint32_t kph = 0, vph = 0, qph = 0;
  uint32_t confirmed = 0;  // peers whose rows of A are known to have landed (multi-GPU)
  const uint32_t ready_epoch = p.ready_epoch + (p.epoch_base != nullptr ? *p.epoch_base : 0u);
  for (int u = blockIdx.x; u < p.n_sched; u += gridDim.x) {
    const Unit w = decode_unit(p, u);
    if (w.nt == 0) continue;
    const int col = w.h * kHD;
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_wait((SB(q_empty) + 8u * (s)), qph ^ 1, 0x100 + s, p.dbg);
        mbar_arrive_expect_tx((SB(q_full) + 8u * (s)), kTileBytes);
        tma_load_2d(&p.tm_q, (SB(q) + static_cast<uint32_t>(kTileBytes) * (s)), (SB(q_full) + 8u * (s)), col, w.q_row0 + s * kBM);
      }
    }
    qph ^= 1;
    int last_idx = 0;
    if (w.ng > 0) last_idx = __ldg(w.gidx + w.ng - 1);

    for (int t = w.t0; t < w.t0 + w.nt; ++t) {
      // decide how tile t is fetched
      bool gather = false;
      const CUtensorMap* mk;
      const CUtensorMap* mv;
      int row0 = 0;
      int4 iv = make_int4(0, 0, 0, 0);
      if (t < w.tg) {
        const int valid = min(kBN, w.ng - t * kBN);
        const int pos = t * kBN + lane * 4;
        iv = __ldg(reinterpret_cast<const int4*>(w.gidx + pos));
        if (pos + 0 >= w.ng) iv.x = last_idx;
        if (pos + 1 >= w.ng) iv.y = last_idx;
        if (pos + 2 >= w.ng) iv.z = last_idx;
        if (pos + 3 >= w.ng) iv.w = last_idx;
        // contiguous run?  (lists are strictly ascending, so first/last decide)
        const int first = __shfl_sync(0xffffffffu, iv.x, 0);
        const int lpos = valid - 1;
        const int src_lane = lpos >> 2;
        const int c0 = __shfl_sync(0xffffffffu, iv.x, src_lane);
        const int c1 = __shfl_sync(0xffffffffu, iv.y, src_lane);
        const int c2 = __shfl_sync(0xffffffffu, iv.z, src_lane);
        const int c3 = __shfl_sync(0xffffffffu, iv.w, src_lane);
        const int sel = lpos & 3;
        const int lastv = sel == 0 ? c0 : (sel == 1 ? c1 : (sel == 2 ? c2 : c3));
        gather = (lastv - first) != lpos;
        row0 = w.a_base + first;
        mk = gather ? &p.tm_kag : &p.tm_ka;
        mv = gather ? &p.tm_vag : &p.tm_va;
        iv.x += w.a_base;
        iv.y += w.a_base;
        iv.z += w.a_base;
        iv.w += w.a_base;
      } else {
        int seg, local;
        locate_tile(w, t, seg, local);
        const int srow = seg == 0 ? w.seg_row[0] : (seg == 1 ? w.seg_row[1] : w.seg_row[2]);
        row0 = srow + local * kBN;
        const bool in_b = (w.seg_in_b >> seg) & 1;
        if (!in_b && p.ready != nullptr) {
          // rows [lo, hi) of this group's A are about to be read: wait for the peers that deliver them
          const int slen = seg == 0 ? w.seg_len[0] : (seg == 1 ? w.seg_len[1] : w.seg_len[2]);
          const int lo = row0 - w.a_base;
          const int hi = min(lo + kBN, srow - w.a_base + slen);
          for (int r = 0; r < p.ready_n; ++r) {
            if ((confirmed >> r) & 1u) continue;
            int b0, b1;
            if (p.ready_fpp > 0) {
              b0 = __ldg(p.ranges + 4 * (r * p.ready_fpp) + 1);
              b1 = __ldg(p.ranges + 4 * ((r + 1) * p.ready_fpp - 1) + 2);
            } else {
              b0 = p.ready_bounds[r];
              b1 = p.ready_bounds[r + 1];
            }
            if (b0 < hi && b1 > lo) {
              flag_wait_ge(p.ready + r, ready_epoch, 0x130 + r, p.dbg);
              confirmed |= 1u << r;
            }
          }
        }
        mk = in_b ? &p.tm_kb : &p.tm_ka;
        mv = in_b ? &p.tm_vb : &p.tm_va;
      }
      // ---- K
      if (lane == 0) {
        mbar_wait((SB(k_empty) + 8u * (ks)), kph ^ 1, 0x110, p.dbg);
        mbar_arrive_expect_tx((SB(k_full) + 8u * (ks)), kTileBytes);
      }
      __syncwarp();
      if (gather) {
        tma_gather4(mk, (SB(k) + static_cast<uint32_t>(kTileBytes) * (ks)) + lane * 512, (SB(k_full) + 8u * (ks)), col, iv.x, iv.y, iv.z, iv.w);
      } else if (lane == 0) {
        tma_load_2d(mk, (SB(k) + static_cast<uint32_t>(kTileBytes) * (ks)), (SB(k_full) + 8u * (ks)), col, row0);
      }
      if (++ks == kKStages) { ks = 0; kph ^= 1; }
      // ---- V
      if (lane == 0) {
        mbar_wait((SB(v_empty) + 8u * (vs)), vph ^ 1, 0x120, p.dbg);
        mbar_arrive_expect_tx((SB(v_full) + 8u * (vs)), kTileBytes);
      }
      __syncwarp();
      if (gather) {
        tma_gather4(mv, (SB(v) + static_cast<uint32_t>(kTileBytes) * (vs)) + lane * 512, (SB(v_full) + 8u * (vs)), col, iv.x, iv.y, iv.z, iv.w);
      } else if (lane == 0) {
        tma_load_2d(mv, (SB(v) + static_cast<uint32_t>(kTileBytes) * (vs)), (SB(v_full) + 8u * (vs)), col, row0);
      }
      if (++vs == kVStages) { vs = 0; vph ^= 1; }
    }
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1) csa_attn_kernel(const __grid_constant__ AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // warp, lane and the 32-bit shared-window address of the (aligned) storage are made opaque to the compiler:
  // otherwise it re-derives them (S2R SR_TID / SR_CgaCtaId / SR_SWINHI chains, ~25 clk each) in front of every
  // barrier operation of the latency-bound softmax loop instead of keeping three registers
  int warp_ = threadIdx.x >> 5;
  int lane_ = threadIdx.x & 31;
  uint32_t sb_ = smem_u32(&sm);
  asm volatile("" : "+r"(warp_), "+r"(lane_), "+r"(sb_));
  const int warp = warp_;
  const int lane = lane_;
  const uint32_t sb = sb_;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_ka);
    tma_prefetch_desc(&p.tm_va);
    tma_prefetch_desc(&p.tm_kag);
    tma_prefetch_desc(&p.tm_vag);
    tma_prefetch_desc(&p.tm_kb);
    tma_prefetch_desc(&p.tm_vb);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init((SB(q_full) + 8u * (i)), 1);
      mbar_init((SB(q_empty) + 8u * (i)), 1);
      mbar_init((SB(s_full) + 8u * (i)), 1);
      mbar_init((SB(s_free) + 8u * (i)), kSoftmaxWarpsPerTile);
      mbar_init((SB(p_ready) + 8u * (i)), kSoftmaxWarpsPerTile);
      mbar_init((SB(o_done) + 8u * (i)), 1);
    }
    for (int i = 0; i < kKStages; ++i) {
      mbar_init((SB(k_full) + 8u * (i)), 1);
      mbar_init((SB(k_empty) + 8u * (i)), 2);  // one commit per MMA stream
    }
    for (int i = 0; i < kVStages; ++i) {
      mbar_init((SB(v_full) + 8u * (i)), 1);
      mbar_init((SB(v_empty) + 8u * (i)), 2);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(SB(tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel of the stream
  pdl_launch_dependents();
  pdl_wait();

  // the softmax warpgroups keep a whole 128-column score row per thread: hand them the registers of warps 0-3
  // (setmaxnreg sits at the top of each role branch so that ptxas budgets every branch separately)
  if (warp == 0) {
    setmaxnreg_dec<kRegsCtl>();
    producer_warp(p, sb, lane);
  } else if (warp == 1 || warp == 3) {
    // =========================================================================================== MMA issue
    // One issuing thread per Q tile (warp 1: tile 0, warp 3: tile 1).  The two instruction streams are independent,
    // so a Q tile whose softmax is late never blocks the MMAs of the other tile (in-order issue from ONE thread
    // would serialise the two softmax warpgroups).  Per stream: QK(0), QK(1), then {QK(j+2), PV(j)} — the scores
    // run two tiles ahead of the PV so that S(j+1) is in TMEM while the softmax warps exponentiate tile j.
    setmaxnreg_dec<kRegsCtl>();
    // The whole warp runs the control flow (so that descriptors, stage indices and phases are warp-uniform and live
    // in uniform registers); one elected lane issues the tcgen05 instructions and their commits.
    {
      const int s = __shfl_sync(0xffffffffu, warp >> 1, 0);  // 0 or 1
      constexpr uint32_t idesc_qk = make_idesc(kBM, kBN, kBF16 ? 1 : 0, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kBM, kHD, kBF16 ? 1 : 0, 0, 1);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, qph = 0, pph = 0, fph = 0;
      const uint32_t tS = tmem + kColS + s * kBN;
      const uint32_t tO = tmem + kColO + s * kHD;
      const uint32_t tP = tmem + kColP + s * (kBN / 2);
      const uint32_t bar_qf = (SB(q_full) + 8u * (s)), bar_qe = (SB(q_empty) + 8u * (s));
      const uint32_t bar_sf = (SB(s_full) + 8u * (s)), bar_fr = (SB(s_free) + 8u * (s));
      const uint32_t bar_pr = (SB(p_ready) + 8u * (s)), bar_od = (SB(o_done) + 8u * (s));
      const uint64_t dq = make_sw128_desc((SB(q) + static_cast<uint32_t>(kTileBytes) * (s)));
      TRACE_INIT(2 + s, lane == 0);

      // S = Q K^T for the K tile in stage `ks`; releases the stage and, for the unit's last tile, the Q tile
      auto qk_step = [&](bool last) {
        TRACE(20);
        mbar_wait((SB(k_full) + 8u * (ks)), kph, 0x201 + s, p.dbg);
        TRACE(21);
        mbar_wait(bar_fr, fph, 0x203 + s, p.dbg);  // softmax has pulled the previous S into registers
        fph ^= 1;
        tc_fence_after();
        TRACE(22);
        const uint64_t dk = make_sw128_desc((SB(k) + static_cast<uint32_t>(kTileBytes) * (ks)));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHD / 16; ++kk) {
            // advance 16 elements (32 B) along the contraction dim inside the swizzled 128 B row
            mma_ss(tS, dq + kk * 2, dk + kk * 2, idesc_qk, kk > 0 ? 1u : 0u);
          }
          tc_commit(bar_sf);
          tc_commit((SB(k_empty) + 8u * (ks)));
          if (last) tc_commit(bar_qe);
        }
        __syncwarp();
        if (++ks == kKStages) { ks = 0; kph ^= 1; }
      };

      for (int u = blockIdx.x; u < p.n_sched; u += gridDim.x) {
        const Unit w = decode_unit(p, u);
        const int nt = w.nt;
        if (nt == 0) continue;
        mbar_wait(bar_qf, qph, 0x200 + s, p.dbg);
        qph ^= 1;
        qk_step(nt == 1);
        if (nt > 1) qk_step(nt == 2);
        for (int j = 0; j < nt; ++j) {
          TRACE(23);
          mbar_wait((SB(v_full) + 8u * (vs)), vph, 0x210 + s, p.dbg);
          TRACE(24);
          mbar_wait(bar_pr, pph, 0x212 + s, p.dbg);
          pph ^= 1;
          tc_fence_after();
          TRACE(25);
          const uint64_t dv = make_sw128_desc((SB(v) + static_cast<uint32_t>(kTileBytes) * (vs)));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kBN / 16; ++kk) {
              // 16 keys = 16 rows of 128 B = 2048 B along the contraction dim (MN-major B operand);
              // 16 16-bit P values = 8 TMEM columns
              mma_ts(tO, tP + kk * 8, dv + kk * (2048 >> 4), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(bar_od);
            tc_commit((SB(v_empty) + 8u * (vs)));
          }
          __syncwarp();
          TRACE(26);
          if (++vs == kVStages) { vs = 0; vph ^= 1; }
          if (j + 2 < nt) qk_step(j + 3 == nt);  // F(j+1) arrives after P(j): PV first
        }
      }
    }
  } else if (warp == 2) {
    setmaxnreg_dec<kRegsCtl>();
  } else {
    // =========================================================================================== softmax
    setmaxnreg_inc<kRegsSoftmax>();
    const int s = (warp - 4) >> 2;              // Q tile of this warpgroup
    const int row = ((warp & 3) << 5) | lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const uint32_t tS = tmem + lane_base + kColS + s * kBN;
    const uint32_t tO = tmem + lane_base + kColO + s * kHD;
    const uint32_t tP = tmem + lane_base + kColP + s * (kBN / 2);
    const uint32_t bar_s = (SB(s_full) + 8u * (s));
    const uint32_t bar_f = (SB(s_free) + 8u * (s));
    const uint32_t bar_p = (SB(p_ready) + 8u * (s));
    const uint32_t bar_o = (SB(o_done) + 8u * (s));
    const float sc = p.scale_log2;
    const uint64_t sc2 = pack_f2(sc, sc);
    uint32_t sph = 0;
    uint32_t od = 0;  // PV completions on o_done[s] before the current unit
#if CSA_PINGPONG
    // named barriers 1..8: two per lane quarter, one per direction of the exp token (Q tile 0 holds it first)
    const int tok_in = 1 + 2 * (warp & 3) + (s ^ 1);  // completed by the other Q tile's warp after its exp phase
    const int tok_out = 1 + 2 * (warp & 3) + s;
    bool have_token = (s == 0);
#endif
    TRACE_INIT(s, (warp & 3) == 0 && lane == 0);
#if CSA_TRACE
    const bool cta_mark = warp == 4 && lane == 0;
    int n_mark = 0;
    TRACE_CTA(cta_mark, n_mark, 1);
    ++n_mark;
#endif

    if (lane == 0) mbar_arrive(bar_f);  // S is free before the first QK of the kernel

#if CSA_PINGPONG
#define EPI_SLOT (tok_out - 1)   // 0..7, from a value that is live across the tile loop anyway
#else
#define EPI_SLOT (warp - 4)
#endif
    for (int u = blockIdx.x; u < p.n_sched; u += gridDim.x) {
      const Unit w = decode_unit(p, u);
      const int nt = w.nt;
      TRACE(10);
      if (w.total == 0) {  // no key at all (never split: every piece of such a unit lands here)
        const bool row_ok = w.qp * (2 * kBM) + s * kBM + row < p.n_q;
        uint4* optr = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.o) +
                                               static_cast<int64_t>(w.q_row0 + s * kBM + row) * p.o_ld + w.h * kHD);
        if (row_ok && (w.piece < 0 || w.t0 == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) optr[i] = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      if (lane == 0) {
        sm.epi[EPI_SLOT][0] = w.q_row0;
        sm.epi[EPI_SLOT][1] = w.h;
        sm.epi[EPI_SLOT][2] = w.qp;
        sm.epi[EPI_SLOT][3] = w.piece;
        sm.epi[EPI_SLOT][4] = w.split_unit;
      }
      __syncwarp();
      float m = -INFINITY;  // running max, already multiplied by scale*log2(e)
      float l = 0.f;
      TileWalker walk;
      walk.init(w);
      for (int t = 0; t < w.t0; ++t) walk.next();
      uint32_t sv[4][32];  // the score row of the current tile (a thread owns a whole 128-key row)
      constexpr int kPoly = kBF16 ? CSA_POLY_PAIRS : CSA_POLY_PAIRS_F16;
      constexpr int kDeg = kBF16 ? 2 : 3;

      for (int j = 0; j < nt; ++j) {
        TRACE(1);
        mbar_wait(bar_s, sph, 0x300 + s, p.dbg);
        sph ^= 1;
        tc_fence_after();
        TRACE(2);
        tmem_ld32(tS + 0, sv[0]);
        tmem_ld32(tS + 32, sv[1]);
        tmem_ld32(tS + 64, sv[2]);
        tmem_ld32(tS + 96, sv[3]);
        tc_wait_ld();
        // the score row now lives in registers: let the tensor core overwrite S with the next tile's scores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_f);
        TRACE(3);

        const int valid = walk.next();
        if (valid < kBN) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) sv[c][i] = 0xff800000u;  // -inf
        }
        // row max: 8 independent chains of 3-input max (8 deep each), then a 3-input tree
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
        float mn = INFINITY;  // minimum over the pairs that would take the polynomial (guards its exponent arithmetic)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx[(i >> 1) & 7] = fmax3(mx[(i >> 1) & 7], __uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1]));
            if (kPoly > 0 && poly_pair(i >> 1, kPoly))
              mn = fmin3(mn, __uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1]));
          }
        const float m_new =
            fmaxf(m, fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])) * sc);
#if CSA_TRACE
        asm volatile("" ::"f"(m_new));
        TRACE(4);
#endif

        // P and O of this Q tile are still being read by the PV of tile j-1 until o_done completes.  The wait sits
        // before the exp token is taken: a token holder that stalls blocks both Q tiles.
        if (j == 0) {
          m = m_new;
        } else {
          const bool need = m_new > m + 8.0f;
          mbar_wait(bar_o, (od + j - 1) & 1, 0x310 + s, p.dbg);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? fast_exp2(m - m_new) : 1.0f;
            if (need) m = m_new;
            l *= alpha;
            uint32_t ov[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              tmem_ld32(tO + c * 32, ov);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st32(tO + c * 32, ov);
            }
          }
        }

        TRACE(5);

        // P = exp2(S*scale - m) in chunks of 32 keys: FFMA2 -> MUFU.EX2 | polynomial -> FADD2 row sum -> pack -> TMEM.
        // The polynomial has no clamp: its arguments must stay above -120 (x <= 8 holds by construction); the tile's
        // minimum over its pairs decides — a ragged (-inf padded) or extreme tile takes the MUFU for every pair.
        uint64_t nm2 = pack_f2(-m, -m);
        uint64_t ls[2] = {0ull, 0ull};
        const bool poly_ok = kPoly > 0 && __all_sync(0xffffffffu, fmaf(mn, sc, -m) >= -120.0f);
#if CSA_PINGPONG
        if (have_token) {
          have_token = false;
        } else {
          named_bar_sync(tok_in, 64);  // the other Q tile's warp on this sub-partition has issued its exponentials
        }
        asm volatile("" : "+l"(nm2));  // every exponential depends on nm2: none may be hoisted above the token
#endif
        TRACE(6);
        auto exp_pass = [&](auto use_poly) {
          constexpr bool kUsePoly = decltype(use_poly)::value;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint64_t xs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              xs[i] = ffma2(pack_f2(__uint_as_float(sv[c][2 * i]), __uint_as_float(sv[c][2 * i + 1])), sc2, nm2);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p0, p1;
              if (kUsePoly && poly_pair(i, kPoly)) {
                poly_exp2_fast_x2<kDeg>(xs[i], p0, p1);
              } else {
                float x0, x1;
                unpack_f2(xs[i], x0, x1);
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
              }
              ls[i & 1] = fadd2(ls[i & 1], pack_f2(p0, p1));
              pk[i] = pack2<kBF16>(p0, p1);
            }
            tmem_st16(tP + c * 16, pk);
#if CSA_PINGPONG
            if (c == CSA_TOKEN_CHUNK) named_bar_arrive(tok_out, 64);  // the other Q tile may start its exponentials
#endif
          }
        };
        if (poly_ok) {
          exp_pass(std::true_type{});
        } else {
          exp_pass(std::false_type{});
        }
        TRACE(7);
        {
          float a0, a1, b0, b1;
          unpack_f2(ls[0], a0, a1);
          unpack_f2(ls[1], b0, b1);
          l += (a0 + a1) + (b0 + b1);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        TRACE(8);
      }

      // epilogue: wait for the last PV
      TRACE(11);
      if (nt > 0) {
        mbar_wait(bar_o, (od + nt - 1) & 1, 0x320 + s, p.dbg);
        tc_fence_after();
      }
      TRACE(12);
      od += nt;
      // Where the result goes was left in shared memory at the unit's start instead of being carried through the tile
      // loop: the loop runs at the register limit (a 128-key score row per thread), and anything live across it is
      // spilled or rematerialised on the latency chain of every tile.
      struct { int q_row0, h, qp, piece, split_unit; } e;
      e.q_row0 = sm.epi[EPI_SLOT][0];
      e.h = sm.epi[EPI_SLOT][1];
      e.qp = sm.epi[EPI_SLOT][2];
      e.piece = sm.epi[EPI_SLOT][3];
      e.split_unit = sm.epi[EPI_SLOT][4];
      const bool row_ok = e.qp * (2 * kBM) + s * kBM + row < p.n_q;
      uint4* optr = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.o) +
                                             static_cast<int64_t>(e.q_row0 + s * kBM + row) * p.o_ld + e.h * kHD);
#if CSA_TRACE
      asm volatile("" ::"l"(optr));
      TRACE(13);
#endif
      if (e.piece < 0) {
        // whole unit: normalise, store
        const float inv = 1.0f / l;
        uint32_t ov[2][32];
        tmem_ld32(tO, ov[0]);
        tmem_ld32(tO + 32, ov[1]);   // both halves in flight, one round trip
        tc_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 v4;
              v4.x = pack2<kBF16>(__uint_as_float(ov[c][8 * i + 0]) * inv, __uint_as_float(ov[c][8 * i + 1]) * inv);
              v4.y = pack2<kBF16>(__uint_as_float(ov[c][8 * i + 2]) * inv, __uint_as_float(ov[c][8 * i + 3]) * inv);
              v4.z = pack2<kBF16>(__uint_as_float(ov[c][8 * i + 4]) * inv, __uint_as_float(ov[c][8 * i + 5]) * inv);
              v4.w = pack2<kBF16>(__uint_as_float(ov[c][8 * i + 6]) * inv, __uint_as_float(ov[c][8 * i + 7]) * inv);
              optr[c * 4 + i] = v4;
            }
          }
        }
        tc_fence_before();
        TRACE(14);
      } else {
        // piece of a split unit: leave the unnormalised partial (O, m, l) of this row in the workspace; the CTA
        // that delivers the last piece of the unit (arrival counter) merges all of them — nobody waits for anybody.
        // A partial = [16 column chunks][256 rows] float4 of O (a warp touches 512 contiguous bytes), [256] m, [256] l.
        const int prow = s * kBM + row;
        const int my = e.piece % p.split;
        float* part = p.ws + static_cast<int64_t>(e.piece) * kPieceFloats;
        const float* base = p.ws + static_cast<int64_t>(e.piece - my) * kPieceFloats;
        // If every other piece has already been delivered this CTA is the one that merges, and it does so straight
        // from TMEM and registers: no partial of its own is written and read back.
        if (threadIdx.x == 128)
          sm.merge_flag = ld_acquire_sys(p.ws_count + e.split_unit) == static_cast<uint32_t>(p.split) - 1u ? 2u : 0u;
        named_bar_sync(9, kThreads - 128);
        const bool direct = sm.merge_flag == 2u;
        bool merge = direct;
        if (!direct) {
          if (nt > 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t ov[32];
              tmem_ld32(tO + c * 32, ov);
              tc_wait_ld();
              float4* dst = reinterpret_cast<float4*>(part) + (c * 8) * (2 * kBM) + prow;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                __stcg(dst + i * (2 * kBM), make_float4(__uint_as_float(ov[4 * i]), __uint_as_float(ov[4 * i + 1]),
                                                        __uint_as_float(ov[4 * i + 2]), __uint_as_float(ov[4 * i + 3])));
            }
            tc_fence_before();
          }
          __stcg(part + 2 * kBM * kHD + prow, m);
          __stcg(part + 2 * kBM * kHD + 2 * kBM + prow, l);
          __threadfence();
          named_bar_sync(9, kThreads - 128);  // all softmax threads of the CTA have published their rows
          if (threadIdx.x == 128) {
            const uint32_t prev = atomicAdd(p.ws_count + e.split_unit, 1u);
            sm.merge_flag = prev == static_cast<uint32_t>(p.split) - 1u ? 1u : 0u;
          }
          named_bar_sync(9, kThreads - 128);
          merge = sm.merge_flag != 0u;
        }
        if (merge) {
          __threadfence();
          if (threadIdx.x == 128) p.ws_count[e.split_unit] = 0u;  // ready for the next launch
          const bool own = direct && nt > 0 && l > 0.f;  // this thread's row is still on chip
          float mm = own ? m : -INFINITY;
          for (int i = 0; i < p.split; ++i)
            if (!(direct && i == my))
              mm = fmaxf(mm, __ldcg(base + static_cast<int64_t>(i) * kPieceFloats + 2 * kBM * kHD + prow));
          // pieces are accumulated in index order whoever merges (the result does not depend on the arrival order)
          float acc[kHD];
#pragma unroll
          for (int c = 0; c < kHD; ++c) acc[c] = 0.f;
          float lsum = 0.f;
          for (int i = 0; i < p.split; ++i) {
            if (direct && i == my) {
              if (own) {
                const float wgt = fast_exp2(m - mm);
                lsum += wgt * l;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  uint32_t ov[32];
                  tmem_ld32(tO + c * 32, ov);
                  tc_wait_ld();
#pragma unroll
                  for (int j = 0; j < 32; ++j) acc[c * 32 + j] += wgt * __uint_as_float(ov[j]);
                }
                tc_fence_before();
              }
              continue;
            }
            const float* pi = base + static_cast<int64_t>(i) * kPieceFloats;
            const float li = __ldcg(pi + 2 * kBM * kHD + 2 * kBM + prow);
            if (li > 0.f) {  // a piece without key tiles contributes nothing (its O rows were never written)
              const float wgt = fast_exp2(__ldcg(pi + 2 * kBM * kHD + prow) - mm);
              lsum += wgt * li;
              const float4* src = reinterpret_cast<const float4*>(pi) + prow;
#pragma unroll
              for (int c = 0; c < kHD / 4; ++c) {
                const float4 x = __ldcg(src + c * (2 * kBM));
                acc[4 * c + 0] += wgt * x.x;
                acc[4 * c + 1] += wgt * x.y;
                acc[4 * c + 2] += wgt * x.z;
                acc[4 * c + 3] += wgt * x.w;
              }
            }
          }
          if (row_ok) {
            const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 v4;
              v4.x = pack2<kBF16>(acc[8 * i + 0] * inv, acc[8 * i + 1] * inv);
              v4.y = pack2<kBF16>(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
              v4.z = pack2<kBF16>(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
              v4.w = pack2<kBF16>(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
              optr[i] = v4;
            }
          }
        }
        named_bar_sync(9, kThreads - 128);  // merge_flag may be rewritten by the next piece only after everyone read it
      }
#if CSA_TRACE
      TRACE_CTA(cta_mark, n_mark, e.piece < 0 ? 2 : 3);
      ++n_mark;
#endif
    }
#if CSA_PINGPONG
    // Q tile 1's last hand-over has no taker: absorb it so that no barrier is left half-arrived at exit
    if (s == 0 && !have_token) named_bar_sync(tok_in, 64);
#endif
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if (p.done_counter != nullptr && threadIdx.x == 0) {
    // every key tile this CTA fetched from the exchange buffers has been consumed (the grid-barrier idiom of peer.cu:
    // CTA barrier above, one system-scope fence, the counter; the last CTA's release stores follow every increment)
    __threadfence_system();
    const uint32_t prev = atomicAdd(p.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *p.done_counter = 0u;
      __threadfence_system();
      const uint32_t epoch = p.ready_epoch + (p.epoch_base != nullptr ? *p.epoch_base : 0u);
      for (int r = 0; r < p.ready_n; ++r)
        if (r != p.peer_self) st_release_sys(p.done_dst[r] + p.peer_self, epoch);
    }
  }
}


// ------------------------------------------------------------------------------------------------- host side
static int encode_2d(CUtensorMap* tm, int dtype, const void* base, int64_t rows, int64_t cols, int64_t ld_elems,
                     uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_tiled();
  if (!fn) return set_error(CSA_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {kHD, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype == CSA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static thread_local int32_t g_last_launch[4] = {0, 0, 0, 0};  // {grid, n_whole, split, n_sched} of the last csa_attn_fwd

}  // namespace csa

using namespace csa;

extern "C" int csa_attn_fwd(const csa_attn_args_t* a, void* stream_) {
  if (!a) return set_error(CSA_E_BADARG, "csa_attn_fwd: null args");
  if (a->struct_size != sizeof(csa_attn_args_t))
    return set_error(CSA_E_BADARG, "csa_attn_fwd: struct_size %u != %zu (ABI mismatch)", a->struct_size,
                     sizeof(csa_attn_args_t));
  if (a->head_dim != CSA_HEAD_DIM) return set_error(CSA_E_SHAPE, "csa_attn_fwd: head_dim %d != 64", a->head_dim);
  if (a->dtype != CSA_DTYPE_F16 && a->dtype != CSA_DTYPE_BF16)
    return set_error(CSA_E_BADARG, "csa_attn_fwd: dtype %d", a->dtype);
  if (a->heads <= 0 || a->n_groups <= 0 || a->n_frames <= 0 || a->n_q <= 0)
    return set_error(CSA_E_BADARG, "csa_attn_fwd: non-positive geometry");
  if (!a->q || !a->o) return set_error(CSA_E_BADARG, "csa_attn_fwd: null q/o");
  const bool use_g = a->list_base >= 0;
  const bool use_r = a->ranges != nullptr;
  const bool use_a = use_g || use_r || a->ca_len > 0;
  const bool use_b = a->cb_len > 0;
  if (!use_a && !use_b) return set_error(CSA_E_BADARG, "csa_attn_fwd: no key segment enabled");
  if (use_a && (!a->k_a || !a->v_a)) return set_error(CSA_E_BADARG, "csa_attn_fwd: null k_a/v_a");
  if (use_b && (!a->k_b || !a->v_b)) return set_error(CSA_E_BADARG, "csa_attn_fwd: null k_b/v_b");
  if (use_g && (!a->idx || !a->counts || (a->idx_stride & 3) || !aligned16(a->idx)))
    return set_error(CSA_E_BADARG, "csa_attn_fwd: idx/counts null or idx not 16-byte aligned / stride %% 4");
  const int64_t cols = static_cast<int64_t>(a->heads) * CSA_HEAD_DIM;
  auto ld_ok = [&](int64_t ld) { return ld >= cols && (ld % 8) == 0; };
  if (!ld_ok(a->q_ld) || !ld_ok(a->o_ld) || (use_a && !ld_ok(a->a_ld)) || (use_b && !ld_ok(a->b_ld)))
    return set_error(CSA_E_SHAPE, "csa_attn_fwd: row strides must be >= heads*64 and multiples of 8 elements");
  if (!aligned16(a->q) || !aligned16(a->o) || (use_a && (!aligned16(a->k_a) || !aligned16(a->v_a))) ||
      (use_b && (!aligned16(a->k_b) || !aligned16(a->v_b))))
    return set_error(CSA_E_BADARG, "csa_attn_fwd: pointers must be 16-byte aligned");
  if (a->ca_len < 0 || a->cb_len < 0) return set_error(CSA_E_BADARG, "csa_attn_fwd: negative segment length");
  if (use_r && (!aligned16(a->ranges) || a->ca_len != 0))
    return set_error(CSA_E_BADARG, "csa_attn_fwd: ranges must be 16-byte aligned and exclude ca_len");

  AttnKernelParams p;
  memset(&p, 0, sizeof(p));
  const int64_t q_rows = static_cast<int64_t>(a->n_groups) * a->n_frames * a->n_q;
  int rc;
  if ((rc = encode_2d(&p.tm_q, a->dtype, a->q, q_rows, cols, a->q_ld, kBM))) return rc;
  // Unused sources still need valid descriptors (they are prefetched); alias them to whatever is present.
  const void* ka = use_a ? a->k_a : a->k_b;
  const void* va = use_a ? a->v_a : a->v_b;
  const int64_t a_ld = use_a ? a->a_ld : a->b_ld;
  const int64_t a_rows = use_a ? a->a_rows : a->b_rows;
  const void* kb = use_b ? a->k_b : ka;
  const void* vb = use_b ? a->v_b : va;
  const int64_t b_ld = use_b ? a->b_ld : a_ld;
  const int64_t b_rows = use_b ? a->b_rows : a_rows;
  if (a_rows <= 0 || b_rows <= 0) return set_error(CSA_E_BADARG, "csa_attn_fwd: a_rows/b_rows must be positive");
  if ((rc = encode_2d(&p.tm_ka, a->dtype, ka, a_rows, cols, a_ld, kBN))) return rc;
  if ((rc = encode_2d(&p.tm_va, a->dtype, va, a_rows, cols, a_ld, kBN))) return rc;
  if ((rc = encode_2d(&p.tm_kag, a->dtype, ka, a_rows, cols, a_ld, 1))) return rc;
  if ((rc = encode_2d(&p.tm_vag, a->dtype, va, a_rows, cols, a_ld, 1))) return rc;
  if ((rc = encode_2d(&p.tm_kb, a->dtype, kb, b_rows, cols, b_ld, kBN))) return rc;
  if ((rc = encode_2d(&p.tm_vb, a->dtype, vb, b_rows, cols, b_ld, kBN))) return rc;

  p.o = a->o;
  p.o_ld = a->o_ld;
  p.idx = a->idx;
  p.counts = a->counts;
  p.idx_stride = a->idx_stride;
  p.heads = a->heads;
  p.n_groups = a->n_groups;
  p.n_frames = a->n_frames;
  p.n_q = a->n_q;
  p.n_qpairs = (a->n_q + 2 * kBM - 1) / (2 * kBM);
  p.n_units = a->n_groups * a->n_frames * a->heads * p.n_qpairs;
  p.a_group_rows = a->a_group_rows;
  p.b_group_rows = a->b_group_rows;
  p.list_base = use_g ? a->list_base : -1;
  p.list_step = a->list_step;
  p.g_adjust = a->g_adjust;
  p.ca_start = a->ca_start;
  p.ca_step = a->ca_step;
  p.ca_len = a->ca_len;
  p.cb_start = a->cb_start;
  p.cb_step = a->cb_step;
  p.cb_len = a->cb_len;
  p.ranges = a->ranges;
  p.range_base = a->range_base;
  p.range_step = a->range_step;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.b_first = (a->flags & CSA_ATTN_B_FIRST) ? 1 : 0;
  p.ready = nullptr;
  if (a->ready != nullptr) {
    if (a->ready_n <= 0 || a->ready_n > CSA_MAX_PEERS || (reinterpret_cast<uintptr_t>(a->ready) & 3))
      return set_error(CSA_E_BADARG, "csa_attn_fwd: ready_n must be in [1, %d] and ready 4-byte aligned", CSA_MAX_PEERS);
    if (use_g) return set_error(CSA_E_BADARG, "csa_attn_fwd: arrival flags cover contiguous runs of A only, not index lists");
    if (a->ready_frames_per_peer > 0) {
      if (!use_r) return set_error(CSA_E_BADARG, "csa_attn_fwd: ready_frames_per_peer needs ranges");
      p.ready_fpp = a->ready_frames_per_peer;
    } else {
      for (int r = 0; r < a->ready_n; ++r)
        if (a->ready_bounds[r] > a->ready_bounds[r + 1])
          return set_error(CSA_E_BADARG, "csa_attn_fwd: ready_bounds must be non-decreasing");
    }
    if (a->epoch_base != nullptr && (reinterpret_cast<uintptr_t>(a->epoch_base) & 3))
      return set_error(CSA_E_BADARG, "csa_attn_fwd: epoch_base must be 4-byte aligned");
    p.ready = a->ready;
    p.ready_epoch = a->ready_epoch;
    p.epoch_base = a->epoch_base;
    p.ready_n = a->ready_n;
    for (int r = 0; r <= a->ready_n; ++r) p.ready_bounds[r] = a->ready_bounds[r];
    if (a->done_counter != nullptr) {
      if (a->peer_self < 0 || a->peer_self >= a->ready_n || (reinterpret_cast<uintptr_t>(a->done_counter) & 3))
        return set_error(CSA_E_BADARG, "csa_attn_fwd: peer_self %d outside [0, ready_n) or misaligned done_counter", a->peer_self);
      for (int r = 0; r < a->ready_n; ++r) {
        if (r != a->peer_self && !a->done_dst[r])
          return set_error(CSA_E_BADARG, "csa_attn_fwd: null done_dst[%d]", r);
        p.done_dst[r] = a->done_dst[r];
      }
      p.done_counter = a->done_counter;
      p.peer_self = a->peer_self;
    }
  } else if (a->done_counter != nullptr) {
    return set_error(CSA_E_BADARG, "csa_attn_fwd: done_counter needs ready flags");
  }
  p.dbg = debug_record_devptr();

  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "cudaGetDevice: %s", cudaGetErrorString(ce));
  const int sms = sm_count(dev);
  if (sms <= 0) return set_error(CSA_E_DEVICE, "csa_attn_fwd: device %d is not sm_100", dev);
  int ctas = sms;
  if (a->max_ctas > 0 && ctas > a->max_ctas) ctas = a->max_ctas;

  // Tail split.  Units are dealt round-robin to `ctas` persistent CTAs; the last, partial round would keep only
  // `rem` of them busy for a whole unit time.  Cutting each of those `rem` units into k pieces along the keys turns
  // that round into ceil(rem * k / ctas) rounds of 1/k unit time.  k is the smallest value within 5 % of the best.
  p.n_whole = p.n_units;
  p.split = 1;
  p.n_sched = p.n_units;
  p.ws = nullptr;
  p.ws_count = nullptr;
  const int rem = p.n_units % ctas;
  const int64_t piece_bytes = static_cast<int64_t>(kPieceFloats) * 4;
  if (rem > 0 && a->workspace != nullptr && !(a->flags & CSA_ATTN_NO_SPLIT)) {
    if (!aligned16(a->workspace)) return set_error(CSA_E_BADARG, "csa_attn_fwd: workspace must be 16-byte aligned");
    const int64_t cap_pieces = (a->workspace_bytes - kWsHeaderBytes) / piece_bytes;
    // an upper estimate of the key tiles of a unit: pieces of fewer than ~3 tiles are not worth their prologue
    const int64_t est_keys = static_cast<int64_t>(a->ca_len) + a->cb_len + ((use_g || use_r) ? a->a_group_rows / 2 : 0);
    const int k_cap = static_cast<int>(est_keys / (3 * kBN));
    int best_k = 1;
    double best_cost = 1.0;
    for (int k = 2; k <= kMaxSplit && k <= k_cap; ++k) {
      const int64_t pieces = static_cast<int64_t>(rem) * k;
      if (pieces > cap_pieces || pieces > 2 * static_cast<int64_t>(ctas)) break;
      const double cost = static_cast<double>((pieces + ctas - 1) / ctas) / k;
      if (cost < best_cost * 0.95) {
        best_cost = cost;
        best_k = k;
      }
    }
    const int forced = (a->flags >> 8) & 0xff;  // test aid: CSA_ATTN_FORCE_SPLIT(k)
    if (forced > 0 && forced <= kMaxSplit && static_cast<int64_t>(rem) * forced <= cap_pieces) best_k = forced;
    if (best_k > 1 && rem * 4 <= kWsHeaderBytes) {
      p.split = best_k;
      p.n_whole = p.n_units - rem;
      p.n_sched = p.n_whole + rem * best_k;
      p.ws_count = static_cast<uint32_t*>(a->workspace);
      p.ws = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + kWsHeaderBytes);
    }
  }
  const int grid = p.n_sched < ctas ? p.n_sched : ctas;
  g_last_launch[0] = grid;
  g_last_launch[1] = p.n_whole;
  g_last_launch[2] = p.split;
  g_last_launch[3] = p.n_sched;

  const size_t smem = sizeof(AttnSmem) + 1024;
  auto kern = a->dtype == CSA_DTYPE_BF16 ? csa_attn_kernel<true> : csa_attn_kernel<false>;
  // per device and kernel instantiation, once (the attribute is sticky; the call costs a few microseconds)
  static bool smem_set[64][2];
  const int ki = a->dtype == CSA_DTYPE_BF16 ? 1 : 0;
  if (dev < 0 || dev >= 64 || !smem_set[dev][ki]) {
    ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess)
      return set_error(static_cast<int>(ce), "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(ce));
    if (dev >= 0 && dev < 64) smem_set[dev][ki] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ce = cudaLaunchKernelEx(&cfg, kern, p);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "csa_attn_kernel launch: %s", cudaGetErrorString(ce));
  return 0;
}

extern "C" int csa_debug_last_launch(int32_t* out4_host) {
  if (!out4_host) return set_error(CSA_E_BADARG, "csa_debug_last_launch: null out");
  for (int i = 0; i < 4; ++i) out4_host[i] = g_last_launch[i];
  return 0;
}

extern "C" int64_t csa_attn_workspace_bytes(int32_t ctas) {
  if (ctas <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    ctas = sm_count(dev);
    if (ctas <= 0) return 0;
  }
  return kWsHeaderBytes + 2 * static_cast<int64_t>(ctas) * kPieceFloats * 4;
}

// Debug builds only (-DCSA_TRACE=1): device buffer of kTraceSlots * kTraceEvents 64-bit words for the timeline trace.
extern "C" int csa_debug_set_trace(void* dev_buffer) {
#if CSA_TRACE
  cudaError_t ce = cudaMemcpyToSymbol(csa::g_trace, &dev_buffer, sizeof(dev_buffer));
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "csa_debug_set_trace: %s", cudaGetErrorString(ce));
  return 0;
#else
  (void)dev_buffer;
  return set_error(CSA_E_BADARG, "csa_debug_set_trace: library built without CSA_TRACE");
#endif
}
