// Micro-benchmark: tcgen05.mma (bf16, M=128, K=16) issue/execute rate per SM for several N, with the A operand in
// shared memory (SS) or in tensor memory (TS).  One warp, one elected lane issues `batch` MMAs, commits, waits.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../beach_seg_b200/csrc \
//      -o umma_rate_bench umma_rate_bench.cu
#include <cstdio>
#include "common.cuh"
using namespace bseg;

// MODE: 0 = SS (A, B K-major in smem), 1 = TS (A in TMEM, B K-major), 2 = TS with B MN-major, 3 = SS with B MN-major,
//       4 = TS alternating between two accumulators
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench(int iters, int batch, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = uniform_warp_idx();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = uniform_u32(slot);
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t idesc_mn = umma_idesc_16bit(128, N, 1, 1, 0, 1);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        for (int j = 0; j < batch; ++j) {
          const int k = j & 3;
          if (MODE == 1) umma_bf16_ts(tm, tm + 256 + k * 8, umma_desc_sw128_kmajor(b_addr + k * 32), idesc, 1u);
          else if (MODE == 2) umma_bf16_ts(tm, tm + 256 + k * 8, umma_desc_sw128_mnmajor(b_addr + k * 2048), idesc_mn, 1u);
          else if (MODE == 3)
            umma_bf16_ss(tm, umma_desc_sw128_kmajor(a_addr + k * 32), umma_desc_sw128_mnmajor(b_addr + k * 2048), idesc_mn, 1u);
          else if (MODE == 4)
            umma_bf16_ts(tm + (j & 1) * 64, tm + 256 + k * 8, umma_desc_sw128_kmajor(b_addr + k * 32), idesc, 1u);
          else umma_bf16_ss(tm, umma_desc_sw128_kmajor(a_addr + k * 32), umma_desc_sw128_kmajor(b_addr + k * 32), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int MODE>
void run(long long* cyc) {
  const int iters = 2000;
  auto k = bench<N, MODE>;
  const char* names[5] = {"SS", "TS", "TS B=MN-major", "SS B=MN-major", "TS two accumulators"};
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  long long h[148];
  for (int batch : {1, 4, 8, 16}) {
    k<<<148, 128, 60000>>>(iters, batch, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=128 N=%3d K=16 %-20s batch=%2d: %7.1f cycles per MMA (%7.1f per batch) %s\n", N, names[MODE], batch,
           (double)h[0] / iters / batch, (double)h[0] / iters, cudaGetErrorString(e));
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  run<64, 0>(cyc);
  run<64, 1>(cyc);
  run<64, 2>(cyc);
  run<64, 3>(cyc);
  run<64, 4>(cyc);
  run<112, 0>(cyc);
  run<128, 0>(cyc);
  run<128, 1>(cyc);
  run<256, 0>(cyc);
  run<176, 0>(cyc);
  return 0;
}
