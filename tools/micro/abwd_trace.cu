// Timeline of one CTA of the attention-backward kernels (clock64 stamps at the hand-offs of the TMA / MMA / elementwise
// protocol).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DBSEG_ABWD_TRACE \
//      -I../../beach_seg_b200/csrc -I../../include -o abwd_trace abwd_trace.cu ../../beach_seg_b200/csrc/host_utils.cu -lcuda
#include "../../beach_seg_b200/csrc/attention_bwd.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  using namespace bseg;
  const int nseq = argc > 1 ? atoi(argv[1]) : 32, heads = 16, T = 1568;
  const size_t n = (size_t)nseq * heads * T * 64, nsh = (size_t)nseq * heads;
  std::vector<__nv_bfloat16> h(n);
  __nv_bfloat16 *q, *k, *v, *dO, *rel, *dqkv;
  float *lse, *dvec, *bias;
  cudaMalloc(&q, n * 2); cudaMalloc(&k, n * 2); cudaMalloc(&v, n * 2); cudaMalloc(&dO, n * 2);
  cudaMalloc(&dqkv, n * 2 * 3); cudaMalloc(&rel, 176 * 64 * 2);
  cudaMalloc(&lse, nsh * T * 4); cudaMalloc(&dvec, nsh * T * 4); cudaMalloc(&bias, nsh * T * 84 * 4);
  srand(1);
  auto fill = [&](__nv_bfloat16* d, size_t cnt, float s) {
    for (size_t i = 0; i < cnt; ++i) h[i] = __float2bfloat16(s * ((rand() % 2001) / 1000.0f - 1.0f));
    cudaMemcpy(d, h.data(), cnt * 2, cudaMemcpyHostToDevice);
  };
  fill(q, n, 1.5f * 0.18f); fill(k, n, 1.5f); fill(v, n, 1.0f); fill(dO, n, 0.01f); fill(rel, 176 * 64, 0.3f * 8.0f);
  {
    std::vector<float> f(nsh * T, 14.0f);  // lse of a typical row (log2 domain): P stays in range
    cudaMemcpy(lse, f.data(), f.size() * 4, cudaMemcpyHostToDevice);
    std::fill(f.begin(), f.end(), 0.001f);
    cudaMemcpy(dvec, f.data(), f.size() * 4, cudaMemcpyHostToDevice);
  }
  for (int it = 0; it < 3; ++it) {
    int rc = launch_attention_bwd(q, k, v, dO, lse, dvec, rel, bias, dqkv, nseq, heads, 0);
    if (rc) { printf("launch failed: %s\n", last_error_buf()); return 1; }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) launch_attention_bwd(q, k, v, dO, lse, dvec, rel, bias, dqkv, nseq, heads, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("attention bwd nseq=%d: %.3f ms per dq+dkv pair\n", nseq, ms / 5);
  }
  static long long tr[2][4][32][8];
  cudaMemcpyFromSymbol(tr, g_abwd_trace, sizeof(tr));
  for (int kern = 0; kern < 2; ++kern) {
    const int nblk = kern == 0 ? 14 : 25;
    long long t0 = tr[kern][1][0][0];
    if (t0 == 0) continue;
    printf("==== %s kernel, CTA (1,0,0); cycles since the MMA warp first looked at block 0\n", kern == 0 ? "dq" : "dkv");
    printf("-- TMA producer: [wait-empty start, wait done]\n");
    for (int j = 0; j < nblk; ++j) printf("%4d %8lld %8lld\n", j, tr[kern][0][j][0] - t0, tr[kern][0][j][1] - t0);
    printf("-- MMA issuer: score MMAs [wait-full start, full seen, issued] | gradient MMAs [wait start, operands seen, issued]\n");
    for (int j = 0; j < nblk; ++j)
      printf("%4d %8lld %8lld %8lld | %8lld %8lld %8lld\n", j, tr[kern][1][j][0] - t0, tr[kern][1][j][1] - t0,
             tr[kern][1][j][2] - t0, tr[kern][1][j][3] - t0, tr[kern][1][j][4] - t0, tr[kern][1][j][5] - t0);
    for (int g = 0; g < 2; ++g) {
      printf("-- elementwise warpgroup %d (warp %d): [wait start, scores seen, part 1 stored, part 2 stored, st drained, arrived]\n",
             g, 4 + 4 * g);
      for (int j = 0; j < nblk; ++j) {
        if (tr[kern][2 + g][j][0] == 0) continue;
        printf("%4d", j);
        for (int ev = 0; ev < 6; ++ev) printf(" %8lld", tr[kern][2 + g][j][ev] - t0);
        printf("\n");
      }
    }
  }
  return 0;
}
