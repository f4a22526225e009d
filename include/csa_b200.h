/*
 * csa_b200 — C ABI of the B200-native Consistent Self-Attention hot path.
 *
 * This is the drop-in boundary for the one data-parallel path of Layjins/Spider's story generation
 * (StoryDiffusion's `SpatialAttnProcessor2_0` + `cal_attn_mask_xl` + `id_bank`).  The reference has no FFI on
 * this path — it is pure PyTorch — so each entry point below replaces a *library call site* of the reference
 * (file:line relative to the reference tree) and is bound from Python with ctypes:
 *
 *   csa_compact_rows      <- StoryDiffusion/utils/gradio_utils.py:260-286  (dense (T*N)^2 bool mask build) and the
 *                            mask slicing at StoryDiffusion/Comic_Generation.py:105-114: instead of materialising
 *                            the mask, the per-frame rows are compacted into ascending key index lists
 *                            (== torch.nonzero(mask[f*N]), bit-exact).
 *   csa_validate_mask     <- the premise of the compaction (all N rows of a frame block are identical), checked on
 *                            a dense mask supplied by an unmodified driver (Comic_Generation.py:376).
 *   csa_attn_fwd          <- F.scaled_dot_product_attention(q, k, v, attn_mask=mask) at
 *                            Comic_Generation.py:175-177 (consistent) and :248-250 (standard / read-early), plus the
 *                            torch.cat of bank and current frames at :92 (two K/V sources are read in place).
 *   csa_gather_rows       <- the row selection implied by the mask when K/V rows have to be materialised
 *                            contiguously (multi-GPU exchange of the sampled rows; bank export).
 *   csa_linear            <- attn.to_q / to_k / to_v / to_out[0] (Comic_Generation.py:155,164-165,185): the projections
 *                            either side of the attention, as library GEMMs issued from inside the library.
 *   csa_run_batch         <- a whole processor call (:129-196) issued with one call into the library.
 *   csa_peer_scatter_kv,  <- (no reference counterpart: the reference is single-GPU) the per-layer exchange of the
 *   csa_peer_signal          sampled K/V rows between the GPUs that share one CFG half, fused with their gather:
 *                            rows are stored straight into every peer's K[S], V[S] buffer over NVLink and the
 *                            attention kernel of the receiving GPU starts on its local keys meanwhile.
 *   csa_sample_ranges,    <- the observation that the sampled vector of gradio_utils.py:257-261 is ONE list shared by
 *   csa_gather_kv            every frame, CFG half and head: the sampled K/V rows are made contiguous once per layer
 *                            (HBM-bound) and frame f then attends two runs of that buffer plus its own block, so the
 *                            attention kernel streams plain TMA tiles instead of re-gathering rows in every CTA.
 *
 * Conventions: all pointers are DEVICE pointers unless stated; no function allocates, frees or synchronises;
 * `stream` is a cudaStream_t passed as void*; every function returns 0 on success, a negative CSA_E_* on a bad
 * argument, or a positive cudaError_t / CUresult from the runtime.  csa_last_error() describes the last failure
 * on the calling thread.  There is no CPU fallback anywhere in this library.
 */
#ifndef CSA_B200_H_
#define CSA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSA_ABI_VERSION 9

#define CSA_E_BADARG (-1)   /* null pointer, non-positive size, misaligned pointer/stride */
#define CSA_E_SHAPE (-2)    /* unsupported geometry (head_dim != 64, row stride not 16-byte aligned, ...) */
#define CSA_E_DEVICE (-3)   /* not an sm_100 device */
#define CSA_E_DRIVER (-4)   /* could not resolve cuTensorMapEncodeTiled */

#define CSA_DTYPE_F16 0
#define CSA_DTYPE_BF16 1

#define CSA_HEAD_DIM 64
#define CSA_TILE 128 /* keys per tile; index-list rows must be padded to a multiple of this */
#define CSA_MAX_PEERS 8 /* GPUs that can share one K[S], V[S] buffer over peer memory (one CFG half of a 16-GPU job) */

int csa_abi_version(void);
const char* csa_last_error(void);

/* 0 if `device` is an sm_100 part this library can run on, CSA_E_DEVICE otherwise. */
int csa_device_supported(int device);

/* If a kernel trapped on its barrier watchdog, copies {tag, block, thread, parity} to out[4] (HOST pointer)
 * and returns 1; returns 0 if nothing is recorded.  Synchronises the device.  Debug aid only. */
int csa_debug_stuck(uint32_t* out4_host);

/* Timeline trace of the attention kernel (libraries built with -DCSA_TRACE=1 only; the shipped build returns
 * CSA_E_BADARG): `dev_buffer` receives 4 x 8192 64-bit words {event id << 48 | SM clock} written by CTA 0.
 * Pass NULL to switch tracing off.  Debug aid only (tools/trace_timeline.py). */
int csa_debug_set_trace(void* dev_buffer);

/*
 * Compact `n_rows` boolean rows of `n_cols` bytes each into ascending column index lists.
 *   mask        row r starts at mask + r*row_stride bytes; byte != 0 means "attend".  row_stride may be 0
 *               (every row is the same sampled vector, cf. gradio_utils.py:257-261).
 *   block_n     if > 0, row r additionally gets its own block [r*block_n, (r+1)*block_n) forced True and every
 *               column >= limit_cols outside that block forced False (gradio_utils.py:267-278); if 0 the bytes
 *               are used as they are (a row of an already-built dense mask).
 *   idx         out, row r at idx + r*idx_stride int32s; idx_stride must be a multiple of 4 and >= n_cols rounded
 *               up to CSA_TILE; entries at and beyond counts[r] are left untouched.
 *   counts      out, n_rows int32.
 */
int csa_compact_rows(const uint8_t* mask, int64_t row_stride, int32_t n_rows, int32_t n_cols, int32_t block_n,
                     int32_t limit_cols, int32_t* idx, int64_t idx_stride, int32_t* counts, void* stream);

/*
 * Check that a dense (n_rows x n_cols) bool mask, row stride `row_stride` bytes, consists of blocks of `block_n`
 * identical consecutive rows.  *n_bad (device int32, must be zeroed by the caller) receives the number of 16-byte
 * words (bytes, when base or stride is not 16-byte aligned) that differ from the first row of their block.
 * HBM-bound: reads the mask once.
 */
int csa_validate_mask(const uint8_t* mask, int64_t row_stride, int32_t n_rows, int32_t n_cols, int32_t block_n,
                      int32_t* n_bad, void* stream);

/*
 * dst[i, :] = src[row_base + idx[i], :] for i < min(*count + count_adjust, max_rows); rows are `row_bytes` wide
 * (multiple of 16), leading dimensions in bytes.  `count` is a device pointer so no host sync is needed.
 */
int csa_gather_rows(const void* src, int64_t src_ld_bytes, int32_t row_base, const int32_t* idx,
                    const int32_t* count, int32_t count_adjust, int32_t max_rows, void* dst, int64_t dst_ld_bytes,
                    int32_t row_bytes, void* stream);

/*
 * Per-frame runs of the sampled list.  `s_idx[0 .. *s_count)` is the ascending list S of sampled key positions
 * (columns < n_frames*block_n of the sample vector).  For f < n_frames, ranges[f] = {0, lo_f, hi_f, *s_count - hi_f}
 * with lo_f / hi_f the number of entries below f*block_n / (f+1)*block_n: the two runs of S that frame f attends
 * besides its own block (mask row f = S u block_f, gradio_utils.py:267-278).  ranges[n_frames] = {0, *s_count, 0, 0}
 * is the read-mode row (mask[F*N:], Comic_Generation.py:106-108).  ranges: device int32[(n_frames+1)*4], 16-byte
 * aligned.
 */
int csa_sample_ranges(const int32_t* s_idx, const int32_t* s_count, int32_t block_n, int32_t n_frames,
                      int32_t* ranges, void* stream);

/*
 * Make the sampled K and V rows contiguous: for g < n_groups and i < *s_count
 *   k_out[g*out_group_rows + i, :] = k[g*group_rows + s_idx[i], :]      (same for v)
 * and rows [*s_count, *s_count + CSA_TILE) of every group of k_out / v_out are zero-filled (a ragged last key tile
 * must read finite values).  Rows are `row_bytes` wide (multiple of 16); leading dimensions in bytes;
 * out_group_rows >= max_rows + CSA_TILE where max_rows bounds *s_count.  HBM-bound: 2 * (read + write) of the rows.
 */
int csa_gather_kv(const void* k, const void* v, int64_t ld_bytes, int32_t group_rows, int32_t n_groups,
                  const int32_t* s_idx, const int32_t* s_count, int32_t max_rows, void* k_out, void* v_out,
                  int64_t out_ld_bytes, int32_t out_group_rows, int32_t row_bytes, void* stream);

/*
 * Flash attention over compacted keys, head_dim 64, fp16 or bf16 in/out, fp32 softmax and accumulation.
 *
 * Work is organised as n_groups (the CFG halves, which never mix: Comic_Generation.py:148) x n_frames query
 * frames x heads.  Queries of (group g, frame f) are rows [(g*n_frames + f)*n_q, +n_q) of `q`; head h is columns
 * [64h, 64h+64).  `o` is indexed like `q`.  The keys of (g, f) are the concatenation of up to three segments:
 *
 *   gathered    rows  g*a_group_rows + idx[list][0 .. counts[list] + g_adjust)  of k_a / v_a,
 *               list = list_base + f*list_step  (disabled when list_base < 0)
 *   contig A    rows  g*a_group_rows + ca_start + f*ca_step + [0, ca_len)        of k_a / v_a
 *               or, when `ranges` is given, the two device-resident runs r = ranges[range_base + f*range_step]:
 *               g*a_group_rows + r[0] + [0, r[1])  and  g*a_group_rows + r[2] + [0, r[3])   (see csa_sample_ranges)
 *   contig B    rows  g*b_group_rows + cb_start + f*cb_step + [0, cb_len)        of k_b / v_b
 *
 * which covers the four branches of the reference processor:
 *   write, consistent (:129-196)   A = sampled K/V rows of this call (csa_gather_kv), ranges[f]; contig B = own
 *                                  frame.  Generic alternative: gathered with list f over A = this call's K/V
 *                                  (in-kernel TMA gather4)                                     (mask[:F*N,:F*N])
 *   read, consistent               A = sampled id_bank K/V rows, ranges[F]; contig B = this call's K/V.  Generic
 *                                  alternative: gathered list F minus the own block (g_adjust = -N) over A =
 *                                  id_bank K/V                                                 (mask[F*N:])
 *   read, early steps (:94-96)     contig A = all F*N bank rows, contig B = own frame
 *   standard (:198-268, enc=None)  contig B = own frame
 * Softmax is permutation invariant, so the segment order need not equal the reference's key order.
 */
typedef struct csa_attn_args {
  uint32_t struct_size; /* sizeof(csa_attn_args_t), checked */
  int32_t dtype;        /* CSA_DTYPE_* */
  int32_t head_dim;     /* must be 64 */
  int32_t heads;
  int32_t n_groups;
  int32_t n_frames;
  int32_t n_q;
  float scale; /* softmax scale, 1/sqrt(head_dim) in the reference (SDPA default) */

  const void* q;
  void* o;
  int64_t q_ld; /* row strides in elements */
  int64_t o_ld;

  const void* k_a;
  const void* v_a;
  int64_t a_ld;
  int64_t a_rows; /* total rows addressable in A (TMA bound) */
  int32_t a_group_rows;
  int32_t _pad0;

  const void* k_b;
  const void* v_b;
  int64_t b_ld;
  int64_t b_rows;
  int32_t b_group_rows;
  int32_t _pad1;

  const int32_t* idx;
  const int32_t* counts;
  int64_t idx_stride;
  int32_t list_base;
  int32_t list_step;
  int32_t g_adjust;

  int32_t ca_start, ca_step, ca_len;
  int32_t cb_start, cb_step, cb_len;

  int32_t max_ctas; /* 0 = one per SM */
  int32_t flags;    /* reserved, 0 */

  const int32_t* ranges; /* optional device int32[...][4], replaces ca_* (see above); 16-byte aligned */
  int32_t range_base;
  int32_t range_step;

  /* Optional scratch for the tail split (see below): device memory, 16-byte aligned, at least
   * csa_attn_workspace_bytes(max_ctas) bytes, ZERO when first handed to the library (the kernel leaves its header
   * zero again); one workspace must not be shared by launches that may run concurrently.  NULL = never split. */
  void* workspace;
  int64_t workspace_bytes;

  /* Multi-GPU (optional, NULL = everything is local): rows [ready_bounds[r], ready_bounds[r+1]) of every group of
   * k_a / v_a are written into this GPU's memory by peer rank r (csa_peer_scatter_kv on that GPU); the kernel
   * reads them only once ready[r] >= ready_epoch (device uint32[ready_n], polled with ld.acquire.sys).  Contiguous
   * runs of A only (ranges / ca_*), not index lists.  Combine with CSA_ATTN_B_FIRST so that the local B segment
   * is worked on while the peers' rows are in flight. */
  const uint32_t* ready;
  uint32_t ready_epoch;
  int32_t ready_n;
  int32_t ready_bounds[CSA_MAX_PEERS + 1];
  /* > 0: the bounds are NOT known on the host — peer r delivers the sampled rows of frames
   * [r*ready_frames_per_peer, (r+1)*ready_frames_per_peer), and the kernel derives the bounds from `ranges`
   * (ranges[f] = {0, lo_f, hi_f, ..}: bound r = lo of frame r*fpp, last bound = hi of the last frame), so that a step
   * needs no host read-back of the sampled counts.  ready_bounds is ignored. */
  int32_t ready_frames_per_peer;
  int32_t _pad2;
  /* Optional: device uint32 added to ready_epoch on the device (see "Epochs in device memory" below); NULL = 0. */
  const uint32_t* epoch_base;
  /* Optional (with `ready`): release the exchange buffers from INSIDE this launch — when its last CTA has finished,
   * done_dst[r][peer_self] = ready_epoch (+ *epoch_base) is published on every GPU r != peer_self (st.release.sys),
   * which is what csa_peer_signal does in a launch of its own.  done_dst[r]: GPU r's release flags (uint32[ready_n]),
   * done_counter: LOCAL zero-initialised uint32 (left zero again).  done_counter == NULL: no release. */
  uint32_t* done_dst[CSA_MAX_PEERS];
  uint32_t* done_counter;
  int32_t peer_self;
  int32_t _pad3;
} csa_attn_args_t;

#define CSA_ATTN_NO_SPLIT 1 /* flags: process every unit whole even if a workspace is given */
#define CSA_ATTN_B_FIRST 2  /* flags: key order of a unit = contiguous B segment first, then A (default: A, then B) */
#define CSA_ATTN_FORCE_SPLIT(k) (((k) & 0xff) << 8) /* flags, test aid: cut the tail units into exactly k <= 8 pieces */

/*
 * Work decomposition.  A unit = (group, frame, head, pair of 128-query tiles); units are dealt round-robin to one
 * persistent CTA per SM.  When the unit count is not a multiple of the CTA count, the units of the last, partial
 * round are cut into k <= 8 pieces along their keys so that every SM stays busy; pieces leave unnormalised partial
 * results (O, row max, row sum) in the workspace and the CTA that delivers the last piece of a unit merges them.
 * Results are independent of the split up to fp32 rounding of the merge.
 */
int csa_attn_fwd(const csa_attn_args_t* args, void* stream);

/* Work decomposition of this thread's last csa_attn_fwd: out[4] = {CTAs, whole units, pieces per split unit,
 * scheduled work items}.  Debug / test aid. */
int csa_debug_last_launch(int32_t* out4_host);

/* Bytes of workspace that let csa_attn_fwd split on `ctas` CTAs (0 = one per SM of the current device). */
int64_t csa_attn_workspace_bytes(int32_t ctas);

/*
 * Multi-GPU exchange of the sampled K/V rows, fused with their gather (replaces gather -> NCCL all-gather ->
 * compaction).  The reference is single-GPU; this is the scale-out of the write pass (Comic_Generation.py:148: the
 * F frames of one CFG half form one key sequence, so with the frames sharded over `n_peers` GPUs every GPU needs
 * the sampled rows of all of them).  Every GPU of the group owns a buffer pair K[S], V[S] laid out in the order of
 * the sampled list S; rank `self` holds the rows S[dst_row0 .. dst_row0 + count) (those of its own frames) and
 * this call stores them DIRECTLY into the buffers of all `n_peers` GPUs (peer pointers over NVLink; [self] is the
 * local buffer):
 *     k_dst[r][(dst_row0 + i) * dst_ld_bytes ...] = k[idx[i] * ld_bytes ...]   for every r, i < count    (same for v)
 * then publishes `epoch` to ready[r][self] on every GPU r (st.release.sys after a system-scope fence), which is what
 * csa_attn_fwd(ready = ...) on GPU r waits for.  Before writing, the kernel waits until done[r] >= done_epoch for
 * every r != self: GPU r has finished the attention launch that last read the buffers being overwritten
 * (csa_peer_signal).  `counter` is a zero-initialised local device uint32 (left zero again).
 * Epochs are monotonically increasing uint32 (start at 1, flags zero-initialised).  HBM/NVLink-bound.
 *
 * Epochs in device memory.  A denoise step captured in a CUDA graph replays the same kernel arguments, so the epoch a
 * launch uses cannot be a host value: with `epoch_base` (a device uint32, the same word for every call of a rank)
 * the kernels use  *epoch_base + epoch  (and  *epoch_base + (int32_t)done_epoch, waiting only if that is > 0), and
 * csa_epoch_advance — one single-thread kernel at the end of a step, captured with it — adds the number of exchange
 * calls of the step, so that every replay publishes and awaits fresh, still monotonic epochs.  NULL = host epochs.
 */
typedef struct csa_peer_scatter_args {
  uint32_t struct_size; /* sizeof(csa_peer_scatter_args_t), checked */
  int32_t n_peers;      /* 1 .. CSA_MAX_PEERS */
  int32_t self;         /* index of this GPU among them */
  int32_t row_bytes;    /* multiple of 16 */
  const void* k;
  const void* v;
  int64_t ld_bytes;
  const int32_t* idx; /* device int32[count]: rows of k / v that are sampled, ascending */
  int32_t count;
  int32_t dst_row0;
  void* k_dst[CSA_MAX_PEERS];
  void* v_dst[CSA_MAX_PEERS];
  int64_t dst_ld_bytes;
  uint32_t* ready[CSA_MAX_PEERS]; /* ready[r]: GPU r's arrival flags, uint32[n_peers] */
  uint32_t epoch;
  uint32_t done_epoch;
  const uint32_t* done; /* LOCAL uint32[n_peers], written by the peers */
  uint32_t* counter;    /* LOCAL, zero */
  /* Optional, device-side geometry (no host read-back of the sampled counts): when `ranges` (csa_sample_ranges
   * output) is given, `idx` is the WHOLE sampled list S, this GPU owns frames [self*frames_per_peer, +frames_per_peer)
   * and the kernel takes  dst_row0 = lo of its first frame,  count = hi of its last frame - dst_row0,  source row of
   * entry i = idx[dst_row0 + i] + idx_adjust  (idx_adjust = -first_frame*block_n turns a position in the half's
   * key sequence into a row of the local k / v); `count` is then only an upper bound used to size the grid. */
  const int32_t* ranges;
  int32_t frames_per_peer;
  int32_t idx_adjust;
  const uint32_t* epoch_base; /* optional, see above */
} csa_peer_scatter_args_t;

int csa_peer_scatter_kv(const csa_peer_scatter_args_t* args, void* stream);

/* done[r][self] = epoch on every GPU r != self (st.release.sys): everything this GPU enqueued on `stream` before
 * this call — in particular the attention launch that read the exchange buffers of `epoch` — has completed. */
int csa_peer_signal(uint32_t* const* done, int32_t n_peers, int32_t self, uint32_t epoch, void* stream);

/* *epoch_base += delta on `stream` (single-thread kernel): ends a step whose exchange calls used epochs relative to
 * *epoch_base (epochs in device memory, above). */
int csa_epoch_advance(uint32_t* epoch_base, uint32_t delta, void* stream);

/* Enable loads/stores from the current device to memory of `peer_device` (cudaDeviceEnablePeerAccess; already
 * enabled is not an error).  Host-side setup helper. */
int csa_enable_peer_access(int32_t peer_device);

/*
 * The projections either side of the attention: y[M,N] = x[M,K] * w[N,K]^T (+ bias[N]) — attn.to_q / to_k / to_v
 * (Comic_Generation.py:155,164-165; bias-free in SDXL) and attn.to_out[0] (:185, with bias), with the row-major layouts
 * of torch's nn.Linear (x, y: leading dimension in elements; w: the module's (out_features, in_features) weight).  A
 * plain library GEMM (cuBLASLt, fp32 accumulation, bias in the epilogue); a weight made of several modules' rows
 * (K and V stacked: N = 2C) projects them in one launch.  `workspace` (device, 16-byte aligned; may be NULL) is
 * cuBLASLt's scratch; nothing is allocated on the device.  Returns 1000 + cublasStatus_t on a cuBLAS failure.
 */
typedef struct csa_linear_args {
  uint32_t struct_size; /* sizeof(csa_linear_args_t), checked */
  int32_t dtype;        /* CSA_DTYPE_* of x, w, bias and y */
  int64_t m, n, k;
  const void* x;
  int64_t ldx;
  const void* w;
  int64_t ldw;
  const void* bias; /* NULL = none */
  void* y;
  int64_t ldy;
  void* workspace;
  int64_t workspace_bytes;
} csa_linear_args_t;

int csa_linear(const csa_linear_args_t* args, void* stream);

/*
 * The same projections as a HAND-WRITTEN sm_100a GEMM (csrc/gemm_sm100.cu: persistent CTA pairs, TMA-fed ring of
 * 4-8 stages, one tcgen05.mma.cta_group::2 of M256 N256 K16 per k step issued by the pair's leader — each CTA keeps its
 * own 128 rows of x and half of the w tile — two TMEM accumulators per CTA, epilogue staged through shared memory so
 * that stores cover whole 128-byte lines; N % 256 == 128 shapes that would waste more than 1/8 of the MMAs use
 * 128-wide tiles with the w tile multicast instead):  y = alpha * x w^T (+ bias), layouts as csa_linear.  Shapes of
 * the path only: N % 128 == 0 and K % 64 == 0 (csa_gemm_supported); anything else stays with csa_linear.  With a
 * stacked weight [w_q; w_k; w_v] (n = 3C) ONE launch projects q, K and V of a layer: y = q, y2 = K|V (y_split = C).
 * Launched with programmatic stream serialization (the prologue overlaps the previous kernel's tail; CSA_PDL=0 in
 * the environment turns that off), like csa_attn_fwd.
 *
 * Fused K/V gather (optional, the write pass of the consistent branch, Comic_Generation.py:164-165 feeding :175-177):
 * when x holds `m / scatter_group_rows` groups (CFG halves) of key rows and w = [w_k; w_v] (n = 2C, split_col = C),
 * every output row r whose position in the sampled key list S is known —  pos = scatter_pos[r % scatter_group_rows]
 * >= 0, from csa_sample_positions — is ALSO stored into the S-ordered buffers the attention kernel streams:
 *     scatter_k[(g * scatter_dst_group_rows + pos) * scatter_ld + c]      = y[r][c]              c <  split_col
 *     scatter_v[(g * scatter_dst_group_rows + pos) * scatter_ld + c - C]  = y[r][c]              c >= split_col
 * which replaces the separate csa_gather_kv launch and its second pass over K and V.
 */
/*
 * The fused gather AS THE MULTI-GPU EXCHANGE (csa_gemm_args_t.exchange): the epilogue stores every sampled row of this
 * GPU's frames straight into the S-ordered K[S] / V[S] buffers of ALL GPUs of the CFG half over peer memory (NVLink;
 * k_dst[self] is the local buffer) — projection, gather and exchange are one kernel, the transfer overlaps the main
 * loop tile by tile — and the launch's last CTA publishes `epoch` to ready[r][self] on every GPU r, which is what
 * csa_attn_fwd(ready = ...) there waits for.  Before its first remote store a CTA waits until done[r] >= done_epoch
 * for every r != self (GPU r has finished the attention launch that last read the buffers being overwritten).  Same
 * protocol, flags and epoch arithmetic as csa_peer_scatter_kv (above), which this replaces together with its re-read
 * of K and V.  One group only: m == scatter_group_rows, destination row = scatter_pos[row].
 */
typedef struct csa_peer_exchange {
  uint32_t struct_size; /* sizeof(csa_peer_exchange_t), checked */
  int32_t n_peers;      /* 1 .. CSA_MAX_PEERS */
  int32_t self;
  int32_t _pad0;
  void* k_dst[CSA_MAX_PEERS];
  void* v_dst[CSA_MAX_PEERS];
  int64_t dst_ld;                 /* elements */
  uint32_t* ready[CSA_MAX_PEERS]; /* ready[r]: GPU r's arrival flags, uint32[n_peers] */
  uint32_t epoch;
  uint32_t done_epoch;
  const uint32_t* done;       /* LOCAL uint32[n_peers], written by the peers */
  uint32_t* counter;          /* LOCAL, zero (left zero again) */
  const uint32_t* epoch_base; /* optional: epochs in device memory, see csa_peer_scatter_kv */
} csa_peer_exchange_t;

typedef struct csa_gemm_args {
  uint32_t struct_size; /* sizeof(csa_gemm_args_t), checked */
  int32_t dtype;        /* CSA_DTYPE_* of x, w, bias and y */
  int64_t m, n, k;
  const void* x;
  int64_t ldx;
  const void* w;
  int64_t ldw;
  const void* bias; /* NULL = none */
  void* y;
  int64_t ldy;
  float alpha;
  int32_t _pad0;
  const int32_t* scatter_pos; /* NULL = no fused gather; device int32[scatter_group_rows] */
  void* scatter_k;
  void* scatter_v;
  int64_t scatter_ld; /* elements */
  int32_t scatter_group_rows;
  int32_t scatter_dst_group_rows;
  int32_t split_col;
  int32_t scatter_col0; /* columns below it are not gathered: w = [w_q; w_k; w_v], scatter_col0 = C, split_col = 2C */
  /* Optional second output matrix: columns [y_split, n) of the product go to y2[:, col - y_split] (leading dimension
   * ldy2) instead of y — q and K|V of one stacked-weight GEMM land in two buffers (the id_bank keeps only K|V).
   * NULL = everything goes to y.  y_split: multiple of 32. */
  void* y2;
  int64_t ldy2;
  int32_t y_split;
  int32_t _pad1;
  /* Optional: the fused gather stores into the buffers of every GPU of the half (see csa_peer_exchange_t above);
   * scatter_k / scatter_v / scatter_ld / scatter_dst_group_rows are then ignored.  NULL = local gather. */
  const struct csa_peer_exchange* exchange;
} csa_gemm_args_t;

int csa_gemm(const csa_gemm_args_t* args, void* stream);
int csa_gemm_supported(int64_t m, int64_t n, int64_t k); /* 1 / 0 */

/* pos[c] = i if s_idx[i] == c for some i < *s_count, else -1, for c in [0, n_cols): the inverse of an ascending index
 * list (csa_compact_rows output), what the fused gather of csa_gemm looks rows up in. */
int csa_sample_positions(const int32_t* s_idx, const int32_t* s_count, int32_t n_cols, int32_t* pos, void* stream);

/*
 * One processor call = one call into the library: the entries are executed in order on `stream` (projections, K/V
 * gather or peer exchange, attention, output projection), stopping at the first failure (*failed_index = its
 * position, -1 if none; may be NULL).  Purely a host-overhead device: the launches are the ones the single entry
 * points make.  CSA_CALL_EVENT_RECORD records the cudaEvent_t passed as `args` (kernel timing inside a batch; on a
 * capturing stream it becomes an event-record node — cudaEventRecordExternal — whose time can be read after a replay).
 */
#define CSA_CALL_LINEAR 1
#define CSA_CALL_ATTN 2
#define CSA_CALL_GATHER_KV 3
#define CSA_CALL_PEER_SCATTER 4
#define CSA_CALL_PEER_SIGNAL 5
#define CSA_CALL_EVENT_RECORD 6
#define CSA_CALL_EPOCH_ADVANCE 7
#define CSA_CALL_GEMM 8

typedef struct csa_gather_kv_args { /* the arguments of csa_gather_kv, in its order */
  const void* k;
  const void* v;
  int64_t ld_bytes;
  int32_t group_rows;
  int32_t n_groups;
  const int32_t* s_idx;
  const int32_t* s_count;
  int32_t max_rows;
  int32_t _pad0;
  void* k_out;
  void* v_out;
  int64_t out_ld_bytes;
  int32_t out_group_rows;
  int32_t row_bytes;
} csa_gather_kv_args_t;

typedef struct csa_peer_signal_args { /* the arguments of csa_peer_signal (+ epochs in device memory) */
  uint32_t* done[CSA_MAX_PEERS];
  int32_t n_peers;
  int32_t self;
  uint32_t epoch;
  uint32_t _pad0;
  const uint32_t* epoch_base; /* optional: the flag value is *epoch_base + epoch */
} csa_peer_signal_args_t;

/* csa_peer_signal with the argument block (what csa_run_batch executes for CSA_CALL_PEER_SIGNAL). */
int csa_peer_signal_ex(const csa_peer_signal_args_t* args, void* stream);

typedef struct csa_epoch_advance_args { /* the arguments of csa_epoch_advance, for csa_run_batch */
  uint32_t* epoch_base;
  uint32_t delta;
  uint32_t _pad0;
} csa_epoch_advance_args_t;

typedef struct csa_call {
  int32_t kind; /* CSA_CALL_* */
  int32_t _pad0;
  const void* args; /* the matching *_args_t (or the cudaEvent_t itself for CSA_CALL_EVENT_RECORD) */
} csa_call_t;

int csa_run_batch(const csa_call_t* calls, int32_t n_calls, void* stream, int32_t* failed_index);

/*
 * Host-side setup helpers for the peer exchange (one process per GPU): export a device allocation of this process
 * (handle64_out: 64 bytes = cudaIpcMemHandle_t of the allocation that contains `ptr`, *offset_out = ptr - its base) and
 * map another process's allocation into the CURRENT device's address space with peer access enabled lazily
 * (*base_out = base of the mapped allocation; close it with csa_ipc_close before the owner frees the memory).
 */
int csa_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out);
int csa_ipc_open(const void* handle64, void** base_out);
int csa_ipc_close(void* base);

#ifdef __cplusplus
}
#endif
#endif /* CSA_B200_H_ */
