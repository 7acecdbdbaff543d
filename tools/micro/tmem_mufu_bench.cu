// Micro-benchmarks for the two rates that bound the attention kernel on B200:
//   (1) tcgen05.ld throughput (TMEM -> registers) per SM for 4 / 8 / 16 warps
//   (2) MUFU.EX2 throughput per SM for 4 / 8 / 16 warps
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mufu_bench tmem_mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t (&r)[32]);
template <>
__device__ __forceinline__ void ldtm<32>(uint32_t taddr, uint32_t (&r)[32]) {
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

__global__ void tmem_ld_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // 4 x 32 columns in flight, then one wait
      ldtm<32>(base + ((it * 4 + c) * 32) % 512, r);
      acc ^= r[0];  // keeps the load alive without serialising on it (the xor is after the wait below)
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

__global__ void mufu_kernel(int iters, long long* cycles, float* sink) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 1024 * 1024 * sizeof(uint32_t));
  long long h[148];
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    tmem_ld_kernel<<<148, warps * 32>>>(iters, cyc, sink);
    cudaDeviceSynchronize();
    tmem_ld_kernel<<<148, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * 4 * 32 * 32 * 4 * warps;  // per SM
    printf("tcgen05.ld x32: %2d warps: %lld cycles, %.1f B/clk/SM (%s)\n", warps, h[0], bytes / h[0], cudaGetErrorString(e));
  }
  for (int warps : {4, 8, 16, 32}) {
    mufu_kernel<<<148, warps * 32>>>(iters, cyc, (float*)sink);
    cudaDeviceSynchronize();
    mufu_kernel<<<148, warps * 32>>>(iters, cyc, (float*)sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 16 * 32 * warps;
    printf("MUFU.EX2: %2d warps: %lld cycles, %.2f ex2/clk/SM (%s)\n", warps, h[0], ops / h[0], cudaGetErrorString(e));
  }
  return 0;
}
