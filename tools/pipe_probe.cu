// Pipe-throughput probe (debug tool, not part of the product): how many SM-sub-partition cycles do the instructions of
// the softmax inner loop cost, alone and mixed, with 1 / 2 / 4 warps per sub-partition?  Answers the questions the
// attention kernel's softmax organisation depends on: MUFU.EX2 rate, whether F2FP (fp32 -> bf16x2 pack) shares the
// XU pipe with MUFU, FFMA2 / FADD2 / FMNMX3 rates, and what the full per-pair mix costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#include "../spider_b200/csrc/ptx.cuh"
using namespace csa;

constexpr int kIters = 2048;

__device__ __forceinline__ uint32_t prmt_hi(float lo, float hi) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(float* out, long long* cycles, float seed) {
  // 8 independent element pairs per iteration
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (i + 1) - 0.001f * threadIdx.x;
  uint32_t acc = 0;
  float mx = -1e30f;
  uint64_t ls = 0;
  const uint64_t sc2 = pack_f2(0.999f, 0.999f), nm2 = pack_f2(-0.01f, -0.01f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = x[2 * i], b = x[2 * i + 1];
      if constexpr (MODE == 0) {  // MUFU only
        a = fast_exp2(a);
        b = fast_exp2(b);
      } else if constexpr (MODE == 1) {  // F2FP only (2 per pair to keep the count equal to mode 0)
        acc += pack2<true>(a, b);
        acc += pack2<true>(b, a);
        a += 1.0f;
      } else if constexpr (MODE == 2) {  // 2 MUFU + 1 F2FP
        a = fast_exp2(a);
        b = fast_exp2(b);
        acc += pack2<true>(a, b);
      } else if constexpr (MODE == 3) {  // FFMA2 only (2 per pair)
        uint64_t v = pack_f2(a, b);
        v = ffma2(v, sc2, nm2);
        v = ffma2(v, sc2, nm2);
        unpack_f2(v, a, b);
      } else if constexpr (MODE == 4) {  // FMNMX3 only (2 per pair)
        mx = fmax3(mx, a, b);
        a = fmax3(a, b, mx);
      } else if constexpr (MODE == 5) {  // the softmax mix: FFMA2, 2 MUFU, FADD2, F2FP, FMNMX3
        mx = fmax3(mx, a, b);
        uint64_t v = ffma2(pack_f2(a, b), sc2, nm2);
        unpack_f2(v, a, b);
        a = fast_exp2(a);
        b = fast_exp2(b);
        ls = fadd2(ls, pack_f2(a, b));
        acc += pack2<true>(a, b);
      } else if constexpr (MODE == 6) {  // the mix with a PRMT (truncating) pack instead of F2FP
        mx = fmax3(mx, a, b);
        uint64_t v = ffma2(pack_f2(a, b), sc2, nm2);
        unpack_f2(v, a, b);
        a = fast_exp2(a);
        b = fast_exp2(b);
        ls = fadd2(ls, pack_f2(a, b));
        acc += prmt_hi(a, b);
      } else if constexpr (MODE == 7) {  // the mix, 2 of 8 pairs through the polynomial
        mx = fmax3(mx, a, b);
        uint64_t v = ffma2(pack_f2(a, b), sc2, nm2);
        if (i % 4 == 3) {
          poly_exp2_x2(v, a, b);
        } else {
          unpack_f2(v, a, b);
          a = fast_exp2(a);
          b = fast_exp2(b);
        }
        ls = fadd2(ls, pack_f2(a, b));
        acc += pack2<true>(a, b);
      } else if constexpr (MODE == 8) {  // the mix, 4 of 8 pairs through the polynomial
        mx = fmax3(mx, a, b);
        uint64_t v = ffma2(pack_f2(a, b), sc2, nm2);
        if (i % 2 == 1) {
          poly_exp2_x2(v, a, b);
        } else {
          unpack_f2(v, a, b);
          a = fast_exp2(a);
          b = fast_exp2(b);
        }
        ls = fadd2(ls, pack_f2(a, b));
        acc += pack2<true>(a, b);
      } else if constexpr (MODE == 9) {  // polynomial only
        uint64_t v = pack_f2(a, b);
        poly_exp2_x2(v, a, b);
        a -= 1.5f;
        b -= 1.5f;
      } else if constexpr (MODE == 10) {  // packed half exp2: ex2.approx.ftz.bf16x2 (one MUFU per pair?) + nothing else
        uint32_t h = pack2<true>(a, b);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h));
        acc += h;
        a += 1.0f;
      } else if constexpr (MODE == 11) {  // MUFU + PRMT pack
        a = fast_exp2(a);
        b = fast_exp2(b);
        acc += prmt_hi(a, b);
      }
      x[2 * i] = a;
      x[2 * i + 1] = b;
    }
  }
  const long long t1 = clock64();
  float s = mx;
  float l0, l1;
  unpack_f2(ls, l0, l1);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + l0 + l1 + __uint_as_float(acc);
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  for (int warps : {4, 8, 16}) {
    pipe_kernel<MODE><<<148, warps * 32>>>(out, cyc, -0.37f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: %s\n", MODE, cudaGetErrorString(e));
      exit(2);
    }
    long long h[148 * 32];
    cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
    long long mxc = 0;
    for (int i = 0; i < 148 * warps; ++i) mxc = h[i] > mxc ? h[i] : mxc;
    // cycles per element pair per warp, and per pair per sub-partition (what the SMSP spends per pair of one warp)
    const double per_pair = double(mxc) / (double(kIters) * 8);
    printf("[pipe] %-34s warps/SMSP %d: %7.2f clk per pair per warp, %6.2f SMSP-clk per pair\n", name, warps / 4,
           per_pair, per_pair / (warps / 4));
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  run<0>("2 MUFU.EX2", out, cyc);
  run<1>("2 F2FP.BF16 pack", out, cyc);
  run<2>("2 MUFU + 1 F2FP", out, cyc);
  run<11>("2 MUFU + 1 PRMT", out, cyc);
  run<3>("2 FFMA2", out, cyc);
  run<4>("2 FMNMX3", out, cyc);
  run<5>("softmax mix (F2FP)", out, cyc);
  run<6>("softmax mix (PRMT pack)", out, cyc);
  run<7>("softmax mix, 25% poly", out, cyc);
  run<8>("softmax mix, 50% poly", out, cyc);
  run<9>("poly exp2 x2 only", out, cyc);
  run<10>("F2FP + ex2.bf16x2", out, cyc);
  return 0;
}
