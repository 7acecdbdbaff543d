// Micro-benchmark: does the cost of a small tcgen05.mma (M=128, N=64, K=16, A in TMEM, B in smem) depend on what the
// other warps of the CTA do?  Warp 0 issues batches of MMAs; warps 4..11 run, per mode,
//   0: nothing   1: tcgen05.ld x32 (+wait) loop   2: ld + tcgen05.st loop   3: MUFU / FMUL loop   4: ld + MUFU + st
//   5: mode 4 plus an 8-KB TMA-like smem write stream (st.shared by warps 2-3)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../beach_seg_b200/csrc \
//      -o umma_contention_bench umma_contention_bench.cu
#include <cstdio>
#include "common.cuh"
using namespace bseg;

template <int MODE, bool SS>
__global__ void __launch_bounds__(384, 1) bench(int iters, int batch, long long* cycles, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = uniform_u32(slot);
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        for (int j = 0; j < batch; ++j) {
          const int k = j & 3;
          if (SS) umma_bf16_ss(tm + (j & 1) * 64, umma_desc_sw128_kmajor(a_addr + k * 32), umma_desc_sw128_kmajor(b_addr + k * 32), idesc, 1u);
          else umma_bf16_ts(tm + (j & 1) * 64, tm + 128 + k * 8, umma_desc_sw128_kmajor(b_addr + k * 32), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp >= 4 && MODE != 0) {
    const int quarter = warp & 3;
    const uint32_t base = tm + (static_cast<uint32_t>(quarter * 32) << 16) + 256 + ((warp - 4) >> 2) * 128;
    float acc = 0.f;
    while (!stop) {
      float s[32], d[32];
      if (MODE == 1 || MODE == 2 || MODE >= 4) {
        tmem_ld32(base, s);
        tmem_ld32(base + 64, d);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = acc + i; d[i] = acc - i; }
      }
      uint32_t pk[16], pd[16];
      if (MODE >= 3) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(s[i]), p1 = ex2_approx(s[i + 1]);
          float d0, d1;
          mul_f32x2(d0, d1, p0, p1, d[i], d[i + 1]);
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pd[i >> 1] = pack_bf16x2(d0, d1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { pk[i] = __float_as_uint(s[i]); pd[i] = __float_as_uint(d[i]); }
      }
      if (MODE == 2 || MODE >= 4) {
        tmem_st16u(base, pk);
        tmem_st16u(base + 64, pd);
        tmem_st_wait();
      }
      acc += __uint_as_float(pk[3]) + __uint_as_float(pd[5]);
    }
    if (acc == 1234.5f) sink[threadIdx.x] = acc;
  } else if ((warp == 2 || warp == 3) && MODE == 5) {
    uint4* dst = reinterpret_cast<uint4*>(smem + 16384 + 32768);
    uint32_t x = lane;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[(warp - 2) * 256 + i * 32 + lane] = make_uint4(x, x, x, x);
      ++x;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int MODE, bool SS>
void run(long long* cyc, float* sink) {
  const int iters = 2000;
  auto k = bench<MODE, SS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80000);
  const char* names[6] = {"idle", "ld", "ld+st", "mufu", "ld+mufu+st", "ld+mufu+st + smem writes"};
  long long h[148];
  double per[2];
  int bi = 0;
  for (int batch : {8, 16}) {
    k<<<148, 384, 80000>>>(iters, batch, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    per[bi++] = (double)h[0] / iters;
  }
  printf("%s N=64 other warps: %-26s  batch 8: %7.1f  batch 16: %7.1f  -> %5.1f cycles per additional MMA\n",
         SS ? "SS" : "TS", names[MODE], per[0], per[1], (per[1] - per[0]) / 8);
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 4096);
  run<0, false>(cyc, sink); run<1, false>(cyc, sink); run<2, false>(cyc, sink); run<3, false>(cyc, sink);
  run<4, false>(cyc, sink); run<5, false>(cyc, sink);
  run<0, true>(cyc, sink); run<1, true>(cyc, sink); run<2, true>(cyc, sink); run<3, true>(cyc, sink);
  run<4, true>(cyc, sink); run<5, true>(cyc, sink);
  return 0;
}
