// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (tiled + gather4), tcgen05 (alloc / mma / ld / st /
// commit / fence).  Nothing here is portable and nothing is meant to be: the library targets B200 only.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace csa {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ----------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Watchdog: a kernel that would otherwise spin forever on a barrier (a protocol bug) records where it was
// stuck and traps, so a bad build costs one failed launch instead of a hung GPU box.
// `dbg` points at 4 words of host-mapped pinned memory {tag, blockIdx.x, threadIdx.x, parity} (or is null), so
// the record survives the trap that kills the context.  The limit is a poll count (each try_wait suspends the
// thread for a hardware time slice, so 2^26 polls is many seconds): no timer reads, a handful of instructions.
#ifndef CSA_WATCHDOG_POLLS
#define CSA_WATCHDOG_POLLS (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0,
                                          volatile uint32_t* dbg = nullptr) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls == CSA_WATCHDOG_POLLS) {
      if (dbg != nullptr && dbg[0] == 0u) {
        dbg[1] = blockIdx.x;
        dbg[2] = threadIdx.x;
        dbg[3] = parity;
        dbg[0] = tag | 0x80000000u;
      }
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}

// One lane of a converged warp (the same lane every time: the lowest active one).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates are {col (innermost), row}.
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t dst, uint32_t bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(col), "r"(row)
      : "memory");
}
// 2-D tiled load, multicast to every CTA of the cluster whose bit is set in `cta_mask`: the tile lands at the same
// shared-memory offset in each of them and each one's mbarrier (same offset) receives the byte count.
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint32_t dst, uint32_t bar, int col, int row,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(col), "r"(row), "h"(cta_mask)
      : "memory");
}
// gather4: four rows {r0..r3} of the 2-D tensor, each `box[0]` columns wide starting at `col`, land as four
// consecutive rows at dst.
__device__ __forceinline__ void tma_gather4(const CUtensorMap* m, uint32_t dst, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
// 2-D tiled store smem -> global (bulk group).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------- tcgen05
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Same, arriving on the mbarrier at this offset in EVERY CTA of `cta_mask` (a pipeline stage that a multicast load
// fills in several CTAs may be refilled only when all of them have consumed it).
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
// ---- programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start (prologue: barrier init, TMEM allocation, descriptor prefetch) while its predecessor in the stream drains;
// it must not touch global memory before pdl_wait(), which returns once the predecessor has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the successor be scheduled as SMs become free (it still waits for this grid's completion in its own pdl_wait)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA pair (cta_group::2): two CTAs of a cluster, on the two SMs of a TPC, run ONE tcgen05.mma of M = 256: each
// holds its own 128 rows of A and HALF of B's rows in shared memory, its own 128 rows of D in tensor memory; the
// leader (cluster rank 0) issues.  All of these are executed as described per function.
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {    // same warps, after a cluster barrier
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// 2-D tiled load into THIS CTA's shared memory whose byte count is credited to the mbarrier at the same offset in the
// pair's leader CTA (bit 24 of a shared-window address is the CTA's rank within the pair).
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t dst, uint32_t bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(col), "r"(row)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem halves of both]; issued by one thread of the leader CTA
__device__ __forceinline__ void mma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in both CTAs of the pair once the pair's MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor for a tile whose rows are 128 bytes wide and 128B-swizzled (what a TMA box of
// 64 16-bit columns with CU_TENSOR_MAP_SWIZZLE_128B produces).  Groups of 8 rows are 1024 B apart (SBO).
// The same tile can be read K-major (the 64 columns are the contraction dim) or MN-major (the rows are the
// contraction dim) — the major-ness is a bit of the *instruction* descriptor, the smem descriptor is shared.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused for SW128, 16 B)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // layout type: SWIZZLE_128B
  return d;
}

// Instruction descriptor, kind::f16, fp32 accumulate.  fmt: 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns; thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// 16-column load (rare paths that must stay small in registers)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 8 columns variant (16 keys of 16-bit P values)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

// 16 columns variant (one 32-key chunk of 16-bit P values)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------------------------- peer flags
// Flags in global memory written by OTHER GPUs over NVLink (st.release.sys after their data stores) and polled
// here; the data is then read through TMA (async proxy), hence the proxy fence after the acquire.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// relaxed system-scope store: the release half of a flag hand-over whose ordering comes from ONE preceding
// __threadfence_system() shared by several flag stores (a st.release each would repeat the fence)
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#ifndef CSA_FLAG_WATCHDOG_POLLS
#define CSA_FLAG_WATCHDOG_POLLS (1u << 25)
#endif
__device__ __forceinline__ void flag_wait_ge(const uint32_t* flag, uint32_t epoch, uint32_t tag = 0,
                                             volatile uint32_t* dbg = nullptr) {
  uint32_t polls = 0;
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(100);
    if (++polls == CSA_FLAG_WATCHDOG_POLLS) {
      if (dbg != nullptr && dbg[0] == 0u) {
        dbg[1] = blockIdx.x;
        dbg[2] = threadIdx.x;
        dbg[3] = epoch;
        dbg[0] = tag | 0x80000000u;
      }
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------- packed fp32 math
// Blackwell issues 2-wide fp32 FMA/ADD (FFMA2 / FADD2) and a 3-input max (FMNMX3): they halve the FMA-/ALU-pipe
// instruction count of the softmax inner loop, which competes with the MUFU for issue slots.
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// exp2 on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial, max relative error 8.6e-5 —
// far below the 2^-9 rounding of the 16-bit P it feeds).  Two elements at a time with packed fp32 ops.  Offloading a
// fraction of the exponentials from the MUFU (16 ex2/clk/SM) is what lets a head_dim-64 softmax keep up with the
// tensor core (which needs 32 exp/clk/SM).
__device__ __forceinline__ uint64_t fadd2_rm(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void poly_exp2_x2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack_f2(x2, x0, x1);
  x0 = fmaxf(x0, -125.0f);  // below this the result is flushed anyway; keeps the exponent arithmetic in range
  x1 = fmaxf(x1, -125.0f);
  x2 = pack_f2(x0, x1);
  const uint64_t magic = pack_f2(12582912.0f, 12582912.0f);  // 1.5 * 2^23: low mantissa bits hold floor(x)
  const uint64_t t2 = fadd2_rm(x2, magic);
  const uint64_t fl2 = fsub2(t2, magic);
  const uint64_t f2 = fsub2(x2, fl2);  // in [0, 1)
  const uint64_t c3 = pack_f2(0.07706618f, 0.07706618f);
  const uint64_t c2 = pack_f2(0.22764593f, 0.22764593f);
  const uint64_t c1 = pack_f2(0.69511657f, 0.69511657f);
  const uint64_t c0 = pack_f2(1.0f, 1.0f);
  uint64_t q2 = ffma2(c3, f2, c2);
  q2 = ffma2(q2, f2, c1);
  q2 = ffma2(q2, f2, c0);
  float q0, q1, t0, t1;
  unpack_f2(q2, q0, q1);
  unpack_f2(t2, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// Clamp-free variant for the softmax inner loop: round-to-nearest range reduction with the 1.5*2^23 trick
// (3 FADD2), minimax polynomial on [-0.5, 0.5] (degree 2: max relative error 1.73e-3, below the 2^-8 half-ulp of a
// bf16 P; degree 3: 7.5e-5, below the 2^-11 half-ulp of an fp16 P), exponent inserted with one LEA per element.
// The CALLER guarantees -120 <= x <= 120 (the softmax checks the tile's minimum and takes the MUFU path otherwise).
// Measured (profiles/r02a_pipe_probe2.log): the softmax mix costs 16.4 SMSP-clk per element pair with MUFU only and
// 13.2 with 3 of every 8 pairs through the degree-2 polynomial (4 warps per sub-partition).
template <int DEG>
__device__ __forceinline__ void poly_exp2_fast_x2(uint64_t x2, float& p0, float& p1) {
  const uint64_t magic = pack_f2(12582912.0f, 12582912.0f);
  const uint64_t t2 = fadd2(x2, magic);
  const uint64_t n2 = fsub2(t2, magic);
  const uint64_t f2 = fsub2(x2, n2);
  uint64_t q2;
  if constexpr (DEG == 2) {
    q2 = ffma2(pack_f2(0.23842891f, 0.23842891f), f2, pack_f2(0.70344800f, 0.70344800f));
    q2 = ffma2(q2, f2, pack_f2(1.0004431f, 1.0004431f));
  } else {
    q2 = ffma2(pack_f2(0.05517166f, 0.05517166f), f2, pack_f2(0.24261113f, 0.24261113f));
    q2 = ffma2(q2, f2, pack_f2(0.69326097f, 0.69326097f));
    q2 = ffma2(q2, f2, pack_f2(0.99992806f, 0.99992806f));
  }
  float q0, q1, t0, t1;
  unpack_f2(q2, q0, q1);
  unpack_f2(t2, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// Register re-distribution between warpgroups (the softmax warpgroups hold a whole 128-column score row per thread).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// pack two fp32 into one 32-bit register of 16-bit floats; `lo` goes to the low half.
template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if constexpr (kBF16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// non-blocking arrival on a named barrier (producer side of a bar.arrive / bar.sync pair)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace csa
