// Micro-benchmark: the element-wise phase of the attention kernels in isolation (no MMA, no barriers).
// Each warp streams its share of a 128 x 112 fp32 "score" tile out of TMEM, computes p = exp2(s*sc + b[col%28] + o[col/28]),
// accumulates the row sum / max and writes P back to TMEM as bf16 pairs -- with 8 or 16 warps per SM
// (2 or 4 per scheduler).  Answers: is 2 warps/scheduler enough to reach the MUFU / TMEM-read rate of 16 elements/clk/SM?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_warps_bench softmax_warps_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
  uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
         "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.055170976f, 0.24260971f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992818f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#define LDWAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define STWAIT() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

// GROUPS column groups per row (1: a thread owns the whole 112-column row; 2: 64 + 48; 4: 32+32+32+16)
template <int GROUPS, int MODE>  // MODE 0: full softmax step, 1: no MUFU (fma only), 2: no TMEM store
__global__ void __launch_bounds__(128 * GROUPS, 1) bench(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3, g = warp >> 2;
  const uint32_t base = slot + ((uint32_t)(q * 32) << 16);
  float bw[28];
#pragma unroll
  for (int i = 0; i < 28; ++i) bw[i] = -0.01f * (i + lane);
  float og[4] = {-1.f, -2.f, -3.f, -4.f};
  const float sc = 0.18f;
  // columns of this group
  constexpr int C0 = (GROUPS == 1) ? 0 : (GROUPS == 2 ? 64 : 32);
  const int c_begin = g * C0;
  float lsum = 0.f, xmax = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t sbase = base + (it & 1) * 128;   // two S buffers
    const uint32_t pbase = base + 256 + (it & 1) * 64;
    auto chunk32 = [&](int c) {
      float s[32];
      ld32(sbase + c, s);
      LDWAIT();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int col0 = (C0 == 0 ? 0 : 0) + i;  // compile-time pattern within the chunk
        float x0 = fmaf(s[i], sc, bw[(col0) % 28]) + og[(col0 / 28) & 3];
        float x1 = fmaf(s[i + 1], sc, bw[(col0 + 1) % 28]) + og[((col0 + 1) / 28) & 3];
        xmax = fmaxf(xmax, fmaxf(x0, x1));
        float p0 = MODE == 1 ? x0 * 0.5f : ex2(x0);
        float p1 = MODE == 1 ? x1 * 0.5f : ((MODE == 3 && ((i >> 1) & 1)) || MODE == 4) ? ex2_poly(x1) : ex2(x1);
        lsum += p0 + p1;
        pk[i >> 1] = pack(p0, p1);
      }
      if (MODE != 2) st16(pbase + (c >> 1), pk);
      else lsum += __uint_as_float(pk[3]);
    };
    auto chunk16 = [&](int c) {
      float s[16];
      ld16(sbase + c, s);
      LDWAIT();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0 = fmaf(s[i], sc, bw[i % 28]) + og[0];
        float x1 = fmaf(s[i + 1], sc, bw[(i + 1) % 28]) + og[1];
        xmax = fmaxf(xmax, fmaxf(x0, x1));
        float p0 = MODE == 1 ? x0 * 0.5f : ex2(x0), p1 = MODE == 1 ? x1 * 0.5f : ex2(x1);
        lsum += p0 + p1;
        pk[i >> 1] = pack(p0, p1);
      }
      if (MODE != 2) st8(pbase + (c >> 1), pk);
      else lsum += __uint_as_float(pk[3]);
    };
    if (GROUPS == 1) { chunk32(0); chunk32(32); chunk32(64); chunk16(96); }
    else if (GROUPS == 2) { if (g == 0) { chunk32(0); chunk32(32); } else { chunk32(64); chunk16(96); } }
    else { if (g < 3) chunk32(c_begin); else chunk16(96); }
    if (MODE != 2) STWAIT();
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = lsum + xmax;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int GROUPS, int MODE>
void run(const char* name, long long* cyc, float* sink) {
  const int iters = 4000;
  long long h[148];
  bench<GROUPS, MODE><<<148, 128 * GROUPS>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  bench<GROUPS, MODE><<<148, 128 * GROUPS>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double elems = (double)iters * 128 * 112;
  printf("%-44s %2d warps: %9lld cycles, %6.2f elements/clk/SM (%s)\n", name, 4 * GROUPS, h[0], elems / h[0],
         cudaGetErrorString(e));
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 1024 * 1024 * sizeof(float));
  run<1, 0>("softmax step, 1 thread per row", cyc, sink);
  run<2, 0>("softmax step, 2 threads per row (64+48)", cyc, sink);
  run<4, 0>("softmax step, 4 threads per row (32x3+16)", cyc, sink);
  run<2, 1>("no MUFU, 2 threads per row", cyc, sink);
  run<4, 1>("no MUFU, 4 threads per row", cyc, sink);
  run<2, 3>("25% polynomial exp2, 2 threads per row", cyc, sink);
  run<2, 4>("50% polynomial exp2, 2 threads per row", cyc, sink);
  run<4, 3>("25% polynomial exp2, 4 threads per row", cyc, sink);
  run<2, 2>("no TMEM store, 2 threads per row", cyc, sink);
  run<4, 2>("no TMEM store, 4 threads per row", cyc, sink);
  return 0;
}
