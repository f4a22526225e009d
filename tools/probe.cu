// GPU probe (debug tool, not part of the product): validates, one primitive at a time, the hardware conventions
// the attention kernel relies on.  Single CTA, 128 threads, strictly sequential.
//   A. TMA tiled load (box 64x128, SWIZZLE_128B) vs 32 x tile::gather4 (box 64x1) of the same rows: byte-identical
//      shared-memory images?  + the swizzle pattern itself.
//   B. S = Q K^T with tcgen05.mma SS (both K-major, SW128 descriptors, +32 B per K step) read back with
//      tcgen05.ld 32x32b.
//   C. O = P V with tcgen05.mma TS (P bf16 in TMEM written with tcgen05.st 32x32b, V MN-major from the TMA tile).
//   D. exp2 throughput: MUFU f32 vs packed f16x2/bf16x2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe tools/probe.cu -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../spider_b200/csrc/ptx.cuh"

using namespace csa;

#define CK(x)                                                                                    \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));    \
      exit(2);                                                                                   \
    }                                                                                            \
  } while (0)

struct ProbeParams {
  CUtensorMap tm_q, tm_k, tm_kg, tm_v;
  const int32_t* idx;   // 128 row indices for the gather test
  uint8_t* dump_tile;   // 16 KB: smem image of the tiled load
  uint8_t* dump_gath;   // 16 KB: smem image of the gather4 load
  float* dump_s;        // 128 x 128
  float* dump_o;        // 128 x 64
  uint32_t qk_kstep16;  // descriptor advance per K step (16-byte units) for QK
  uint32_t pv_kstep16;  // for PV
  uint32_t pv_bmajor;   // 1 = MN-major
  uint32_t pv_sbo16, pv_lbo16;
  uint32_t* dbg;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)lbo16 << 16;
  d |= (uint64_t)sbo16 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sq = base;
  uint8_t* sk = base + 16384;
  uint8_t* sg = base + 32768;
  uint8_t* sv = base + 49152;
  uint64_t* bars = (uint64_t*)(base + 65536);
  uint32_t* tmem_slot = (uint32_t*)(base + 65536 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- A: loads
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bars[0]), 3 * 16384);
    tma_load_2d(&p.tm_q, smem_u32(sq), smem_u32(&bars[0]), 0, 0);
    tma_load_2d(&p.tm_k, smem_u32(sk), smem_u32(&bars[0]), 0, 0);
    tma_load_2d(&p.tm_v, smem_u32(sv), smem_u32(&bars[0]), 0, 0);
    mbar_arrive_expect_tx(smem_u32(&bars[1]), 16384);
  }
  __syncthreads();
  if (warp == 0) {
    const int4 iv = *reinterpret_cast<const int4*>(p.idx + lane * 4);
    tma_gather4(&p.tm_kg, smem_u32(sg) + lane * 512, smem_u32(&bars[1]), 0, iv.x, iv.y, iv.z, iv.w);
  }
  mbar_wait(smem_u32(&bars[0]), 0, 1, p.dbg);
  mbar_wait(smem_u32(&bars[1]), 0, 2, p.dbg);
  for (int i = threadIdx.x; i < 16384 / 16; i += 128) {
    reinterpret_cast<uint4*>(p.dump_tile)[i] = reinterpret_cast<uint4*>(sk)[i];
    reinterpret_cast<uint4*>(p.dump_gath)[i] = reinterpret_cast<uint4*>(sg)[i];
  }
  __syncthreads();

  // ---- B: S = Q K^T
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 128, 1, 0, 0);
    const uint64_t dq = make_desc(smem_u32(sq), 1, 64);
    const uint64_t dk = make_desc(smem_u32(sk), 1, 64);
    for (int kk = 0; kk < 4; ++kk) mma_ss(tmem + 0, dq + kk * p.qk_kstep16, dk + kk * p.qk_kstep16, idesc, kk > 0);
    tc_commit(smem_u32(&bars[2]));
  }
  mbar_wait(smem_u32(&bars[2]), 0, 3, p.dbg);
  tc_fence_after();
  {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(t + c * 32, r);
      tc_wait_ld();
      for (int i = 0; i < 32; ++i) p.dump_s[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
    }
  }
  // ---- C: P (deterministic bf16 pattern) -> TMEM cols 384.., O = P V at cols 256..
  {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + 384;
    const int row = warp * 32 + lane;
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[32];
      for (int i = 0; i < 32; ++i) {
        const int col = c * 64 + 2 * i;
        const float p0 = (float)((row * 7 + col * 3) % 17) / 16.0f;
        const float p1 = (float)((row * 7 + (col + 1) * 3) % 17) / 16.0f;
        pk[i] = pack2<true>(p0, p1);
      }
      tmem_st32(t + c * 32, pk);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 64, 1, 0, p.pv_bmajor);
    const uint64_t dv = make_desc(smem_u32(sv), p.pv_lbo16, p.pv_sbo16);
    for (int kk = 0; kk < 8; ++kk) mma_ts(tmem + 256, tmem + 384 + kk * 8, dv + kk * p.pv_kstep16, idesc, kk > 0);
    tc_commit(smem_u32(&bars[3]));
  }
  mbar_wait(smem_u32(&bars[3]), 0, 4, p.dbg);
  tc_fence_after();
  {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    const int row = warp * 32 + lane;
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(t + c * 32, r);
      tc_wait_ld();
      for (int i = 0; i < 32; ++i) p.dump_o[row * 64 + c * 32 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ---- D: exp2 throughput
template <int MODE>
__global__ void exp_kernel(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b), uc = __float_as_uint(c), ud = __float_as_uint(d);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      a = fast_exp2(a); b = fast_exp2(b); c = fast_exp2(c); d = fast_exp2(d);
    } else if (MODE == 1) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ua));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(uc));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ud));
    } else {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ua));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ub));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(uc));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ud));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      a + b + c + d + __uint_as_float(ua) + __uint_as_float(ub) + __uint_as_float(uc) + __uint_as_float(ud);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled enc() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  return (PFN_encodeTiled)fp;
}
static void make_map(CUtensorMap* m, void* base, int rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d (box_rows %d)\n", (int)r, box_rows);
    exit(3);
  }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  const int R = 512, C = 64;  // K/V/Q source matrices: 512 rows x 64 cols bf16
  std::vector<__nv_bfloat16> hq(R * C), hk(R * C), hv(R * C);
  srand(1);
  for (int i = 0; i < R * C; ++i) {
    hq[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    hk[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    hv[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  }
  std::vector<int32_t> hidx(128);
  for (int i = 0; i < 128; ++i) hidx[i] = i;  // identity first: gather image must equal tile image
  __nv_bfloat16 *dq, *dk, *dv;
  int32_t* didx;
  uint8_t *dt, *dg;
  float *ds, *dof;
  CK(cudaMalloc(&dq, R * C * 2));
  CK(cudaMalloc(&dk, R * C * 2));
  CK(cudaMalloc(&dv, R * C * 2));
  CK(cudaMalloc(&didx, 128 * 4));
  CK(cudaMalloc(&dt, 16384));
  CK(cudaMalloc(&dg, 16384));
  CK(cudaMalloc(&ds, 128 * 128 * 4));
  CK(cudaMalloc(&dof, 128 * 64 * 4));
  CK(cudaMemcpy(dq, hq.data(), R * C * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), R * C * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), R * C * 2, cudaMemcpyHostToDevice));

  uint32_t* dbg_h;
  CK(cudaHostAlloc(&dbg_h, 16, cudaHostAllocMapped));
  memset(dbg_h, 0, 16);
  uint32_t* dbg_d;
  CK(cudaHostGetDevicePointer((void**)&dbg_d, dbg_h, 0));

  ProbeParams p;
  memset(&p, 0, sizeof(p));
  make_map(&p.tm_q, dq, R, C, 128);
  make_map(&p.tm_k, dk, R, C, 128);
  make_map(&p.tm_kg, dk, R, C, 1);
  make_map(&p.tm_v, dv, R, C, 128);
  p.idx = didx;
  p.dump_tile = dt;
  p.dump_gath = dg;
  p.dump_s = ds;
  p.dump_o = dof;
  p.dbg = dbg_d;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));

  // CPU references
  std::vector<float> rs(128 * 128), ro(128 * 64);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      float a = 0;
      for (int d = 0; d < 64; ++d) a += __bfloat162float(hq[i * C + d]) * __bfloat162float(hk[j * C + d]);
      rs[i * 128 + j] = a;
    }
  for (int i = 0; i < 128; ++i)
    for (int d = 0; d < 64; ++d) {
      float a = 0;
      for (int j = 0; j < 128; ++j) a += bf((float)((i * 7 + j * 3) % 17) / 16.0f) * __bfloat162float(hv[j * C + d]);
      ro[i * 64 + d] = a;
    }

  struct Variant { uint32_t qk_k, pv_k, bmaj, sbo, lbo; const char* name; };
  Variant vars[] = {
      {2, 128, 1, 64, 1, "expected: qk+32B, pv MN-major +2048B, SBO 1024"},
      {2, 128, 1, 64, 64, "pv LBO=1024 too"},
      {2, 128, 1, 1, 64, "pv SBO/LBO swapped"},
      {2, 256, 1, 64, 1, "pv kstep 4096B"},
      {2, 64, 1, 64, 1, "pv kstep 1024B"},
      {2, 2, 0, 64, 1, "pv K-major (wrong on purpose)"},
  };
  int rc = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1)
      for (int i = 0; i < 128; ++i) hidx[i] = (i * 37 + 11) % R;  // scattered rows
    CK(cudaMemcpy(didx, hidx.data(), 128 * 4, cudaMemcpyHostToDevice));
    for (size_t vi = 0; vi < (pass == 0 ? sizeof(vars) / sizeof(vars[0]) : 1); ++vi) {
      const Variant& v = vars[vi];
      p.qk_kstep16 = v.qk_k;
      p.pv_kstep16 = v.pv_k;
      p.pv_bmajor = v.bmaj;
      p.pv_sbo16 = v.sbo;
      p.pv_lbo16 = v.lbo;
      CK(cudaMemset(ds, 0, 128 * 128 * 4));
      CK(cudaMemset(dof, 0, 128 * 64 * 4));
      probe_kernel<<<1, 128, 70 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("probe kernel failed: %s; stuck record tag=%x blk=%u thr=%u par=%u\n", cudaGetErrorString(e), dbg_h[0],
               dbg_h[1], dbg_h[2], dbg_h[3]);
        return 4;
      }
      std::vector<uint8_t> ht(16384), hg(16384);
      std::vector<float> gs(128 * 128), go(128 * 64);
      CK(cudaMemcpy(ht.data(), dt, 16384, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hg.data(), dg, 16384, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(gs.data(), ds, 128 * 128 * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(go.data(), dof, 128 * 64 * 4, cudaMemcpyDeviceToHost));
      if (vi == 0) {
        // expected image of row r, 16-byte chunk c: at byte r*128 + ((c ^ (r & 7)) * 16)
        int bad_tile = 0, bad_gath = 0;
        for (int r = 0; r < 128; ++r)
          for (int c = 0; c < 8; ++c) {
            const uint8_t* src_t = (const uint8_t*)&hk[r * C + c * 8];
            const uint8_t* src_g = (const uint8_t*)&hk[hidx[r] * C + c * 8];
            const int off = r * 128 + ((c ^ (r & 7)) * 16);
            bad_tile += memcmp(&ht[off], src_t, 16) != 0;
            bad_gath += memcmp(&hg[off], src_g, 16) != 0;
          }
        printf("[A pass %d] tiled-load swizzle mismatches: %d/1024; gather4 image mismatches: %d/1024\n", pass,
               bad_tile, bad_gath);
        if (bad_tile || bad_gath) rc = 1;
      }
      double es = 0, eo = 0;
      for (int i = 0; i < 128 * 128; ++i) es = fmax(es, fabs(gs[i] - rs[i]));
      for (int i = 0; i < 128 * 64; ++i) eo = fmax(eo, fabs(go[i] - ro[i]));
      printf("[B/C pass %d] %-50s  S max-abs-err %.4g   O max-abs-err %.4g\n", pass, v.name, es, eo);
      if (vi == 0 && (es > 1e-2 || eo > 5e-2)) rc = 1;
    }
  }
  // D
  float* dout;
  CK(cudaMalloc(&dout, 148 * 8 * 256 * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) exp_kernel<0><<<148 * 8, 256>>>(dout, iters);
      if (mode == 1) exp_kernel<1><<<148 * 8, 256>>>(dout, iters);
      if (mode == 2) exp_kernel<2><<<148 * 8, 256>>>(dout, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = 148.0 * 8 * 256 * 4.0 * iters;
    printf("[D] exp2 mode %d (%s): %.3f ms, %.1f G warp-lane-instr/s (x2 elements for packed)\n", mode,
           mode == 0 ? "f32" : (mode == 1 ? "bf16x2" : "f16x2"), ms, instr / ms * 1e-6);
  }
  printf("probe rc=%d\n", rc);
  return rc;
}
