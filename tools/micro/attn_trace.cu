// Timeline of one CTA of the fused attention kernel (clock64 stamps at every hand-off of the softmax / MMA protocol).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DBSEG_ATTN_TRACE \
//      -I../../beach_seg_b200/csrc -o attn_trace attn_trace.cu ../../beach_seg_b200/csrc/host_utils.cu
#include "../../beach_seg_b200/csrc/attention.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  using namespace bseg;
  const int nseq = argc > 1 ? atoi(argv[1]) : 32, heads = 16, T = 1568;
  const size_t n = (size_t)nseq * heads * T * 64;
  std::vector<__nv_bfloat16> h(n);
  __nv_bfloat16 *q, *k, *vt, *rel, *out;
  cudaMalloc(&q, n * 2); cudaMalloc(&k, n * 2); cudaMalloc(&vt, n * 2); cudaMalloc(&out, n * 2);
  cudaMalloc(&rel, 176 * 64 * 2);
  srand(1);
  auto fill = [&](__nv_bfloat16* d, size_t cnt, float s) {
    for (size_t i = 0; i < cnt; ++i) h[i] = __float2bfloat16(s * ((rand() % 2001) / 1000.0f - 1.0f));
    cudaMemcpy(d, h.data(), cnt * 2, cudaMemcpyHostToDevice);
  };
  fill(q, n, 1.5f * 0.18f); fill(k, n, 1.5f); fill(vt, n, 1.0f); fill(rel, 176 * 64, 0.3f * 8.0f);  // q pre-scaled, relcat8
  for (int it = 0; it < 3; ++it) {
    int rc = launch_attention(q, k, vt, rel, out, nullptr, nseq, heads, 56, 28, 0);
    if (rc) { printf("launch failed: %s\n", last_error_buf()); return 1; }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) launch_attention(q, k, vt, rel, out, nullptr, nseq, heads, 56, 28, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("attention nseq=%d: %.3f ms per launch (skip exp %d)\n", nseq, ms / 5, (int)BSEG_ATTN_SKIP_EXP);
  }
  long long tr[3][16][16];
  cudaMemcpyFromSymbol(tr, g_attn_trace, sizeof(tr));
  const long long t0 = tr[0][0][0];
  const char* names[3] = {"softmax WG0", "softmax WG1", "MMA issuer 0"};
  const char* name[6] = {"start", "S full seen", "S handed back", "exp done", "prevPV seen", "P handed"};
  for (int a = 0; a < 2; ++a) {
    printf("== softmax thread group %d (score columns [%d,%d)): cycles per step and key block\n", a, 56 * a, 56 * a + 56);
    printf("%6s", "kb");
    for (int e = 0; e < 6; ++e) printf(" %14s", name[e]);
    printf("\n");
    for (int kb = 0; kb < 14; ++kb) {
      printf("%6d", kb);
      long long prev = kb == 0 ? tr[a][0][0] : tr[a][kb - 1][5];
      for (int e = 0; e < 6; ++e) {
        const long long t = tr[a][kb][e];
        if (kb == 0 && e == 4) { printf(" %14s", "-"); continue; }
        printf(" %14lld", t - prev);
        prev = t;
      }
      printf("   | block end at %lld\n", tr[a][kb][5] - t0);
    }
  }
  printf("== MMA issuer (cycles since block 0 start): K ready, S free seen, V ready, P full seen\n");
  for (int kb = 0; kb < 14; ++kb)
    printf("%6d %10lld %10lld %10lld %10lld\n", kb, tr[2][kb][0] - t0, tr[2][kb][1] - t0, tr[2][kb][3] - t0, tr[2][kb][4] - t0);
  return 0;
}
