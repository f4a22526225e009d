// Pipe-throughput probe, round 2 (debug tool, not part of the product).  Questions:
//  * does a packed 16-bit ex2 (ex2.approx.f16x2 / ex2.approx.ftz.bf16x2) deliver two results per MUFU issue slot?
//  * what does a clamp-free exp2 polynomial cost on the FMA/ALU pipes (degree 2 for bf16 P, degree 3 for fp16 P), and
//    what does the softmax mix cost per element pair when k of every 8 pairs take it, with 1..4 warps per sub-partition?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/pipe_probe2 tools/pipe_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#include "../spider_b200/csrc/ptx.cuh"
using namespace csa;

constexpr int kIters = 2048;

// exp2 of a pair on the FMA/ALU pipes, no clamp (the caller guarantees x >= -120): round-to-nearest range reduction
// with the 1.5*2^23 trick, minimax polynomial on [-0.5, 0.5], exponent inserted with one shift-add per element.
template <int DEG>
__device__ __forceinline__ void poly2_exp2_x2(uint64_t x2, float& p0, float& p1) {
  const uint64_t magic = pack_f2(12582912.0f, 12582912.0f);
  const uint64_t t2 = fadd2(x2, magic);
  const uint64_t n2 = fsub2(t2, magic);
  const uint64_t f2 = fsub2(x2, n2);  // [-0.5, 0.5]
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

// MODE: 0 = ex2.f16x2 only, 1 = ex2.bf16x2 only, 2 = poly only, 3 = softmax mix with NPOLY of 8 pairs through the
// polynomial (max + min tracking, FFMA2 scale, exp, FADD2 row sum, F2FP pack), 4 = mix with the round-1 polynomial
template <int MODE, int NPOLY, int DEG>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(float* out, long long* cycles, float seed) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (i + 1) - 0.001f * threadIdx.x;
  uint32_t acc = 0;
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = 0x3c003c00u + i + threadIdx.x;
  float mx = -1e30f, mn = 1e30f;
  uint64_t ls = 0;
  const uint64_t sc2 = pack_f2(0.999f, 0.999f), nm2 = pack_f2(-0.01f, -0.01f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = x[2 * i], b = x[2 * i + 1];
      if constexpr (MODE == 0) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if constexpr (MODE == 1) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if constexpr (MODE == 2) {
        poly2_exp2_x2<DEG>(pack_f2(a, b), a, b);
        a -= 1.25f;
        b -= 1.25f;
      } else if constexpr (MODE == 3 || MODE == 4) {
        mx = fmax3(mx, a, b);
        const uint64_t v = ffma2(pack_f2(a, b), sc2, nm2);
        // evenly spread poly slots
        const bool poly = ((i + 1) * NPOLY) / 8 != (i * NPOLY) / 8;
        if (poly) {
          if constexpr (MODE == 3) {
            float lo = fminf(a, b);
            mn = fminf(mn, lo);
            poly2_exp2_x2<DEG>(v, a, b);
          } else {
            poly_exp2_x2(v, a, b);
          }
        } else {
          unpack_f2(v, a, b);
          a = fast_exp2(a);
          b = fast_exp2(b);
        }
        ls = fadd2(ls, pack_f2(a, b));
        acc += pack2<true>(a, b);
      }
      x[2 * i] = a;
      x[2 * i + 1] = b;
    }
  }
  const long long t1 = clock64();
  float s = mx + mn;
  float l0, l1;
  unpack_f2(ls, l0, l1);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + l0 + l1 + __uint_as_float(acc);
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE, int NPOLY, int DEG>
void run(const char* name, float* out, long long* cyc) {
  for (int warps : {4, 8, 12, 16}) {
    pipe_kernel<MODE, NPOLY, DEG><<<148, warps * 32>>>(out, cyc, -0.37f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: %s\n", MODE, cudaGetErrorString(e));
      exit(2);
    }
    long long h[148 * 32];
    cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
    long long mxc = 0;
    for (int i = 0; i < 148 * warps; ++i) mxc = h[i] > mxc ? h[i] : mxc;
    const double per_pair = double(mxc) / (double(kIters) * 8);
    printf("[pipe2] %-34s warps/SMSP %d: %7.2f clk per pair per warp, %6.2f SMSP-clk per pair\n", name, warps / 4,
           per_pair, per_pair / (warps / 4));
  }
}

// accuracy of the polynomials against exp2f over the range the softmax feeds them
template <int DEG>
__global__ void acc_kernel(float* worst) {
  float w = 0.f;
  for (int i = threadIdx.x; i < (1 << 20); i += blockDim.x) {
    const float x = -118.0f + 126.0f * (static_cast<float>(i) / (1 << 20));
    float p0, p1;
    poly2_exp2_x2<DEG>(pack_f2(x, x + 0.37f), p0, p1);
    const float r0 = exp2f(x), r1 = exp2f(x + 0.37f);
    w = fmaxf(w, fabsf(p0 - r0) / r0);
    w = fmaxf(w, fabsf(p1 - r1) / r1);
  }
  atomicMax(reinterpret_cast<int*>(worst), __float_as_int(w));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  float* worst;
  cudaMalloc(&worst, 8);
  cudaMemset(worst, 0, 8);
  acc_kernel<2><<<1, 256>>>(worst);
  acc_kernel<3><<<1, 256>>>(worst + 1);
  float hw[2];
  cudaMemcpy(hw, worst, 8, cudaMemcpyDeviceToHost);
  printf("[pipe2] polynomial max relative error on [-118, 8]: degree 2 %.3e, degree 3 %.3e\n", hw[0], hw[1]);
  run<0, 0, 0>("ex2.approx.f16x2 only", out, cyc);
  run<1, 0, 0>("ex2.approx.ftz.bf16x2 only", out, cyc);
  run<2, 0, 2>("poly deg2 only", out, cyc);
  run<2, 0, 3>("poly deg3 only", out, cyc);
  run<3, 0, 2>("mix, 0/8 poly", out, cyc);
  run<3, 2, 2>("mix, 2/8 poly deg2", out, cyc);
  run<3, 3, 2>("mix, 3/8 poly deg2", out, cyc);
  run<3, 4, 2>("mix, 4/8 poly deg2", out, cyc);
  run<3, 5, 2>("mix, 5/8 poly deg2", out, cyc);
  run<3, 2, 3>("mix, 2/8 poly deg3", out, cyc);
  run<3, 3, 3>("mix, 3/8 poly deg3", out, cyc);
  run<3, 4, 3>("mix, 4/8 poly deg3", out, cyc);
  run<4, 2, 3>("mix, 2/8 poly round-1", out, cyc);
  return 0;
}
