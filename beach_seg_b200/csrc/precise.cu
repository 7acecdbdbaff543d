// fp32 mode ("1e-4 in fp32 mode" of the north star): the same SegGPT forward as api.cu's forward_impl with every
// operand, accumulator and activation in IEEE fp32 on the CUDA cores -- no tensor cores, no bf16 anywhere.  It is the
// accuracy mode of the library (bseg_forward_f32), about 30x slower than the tcgen05 path, and shares nothing with it
// but the input-independent embedding table and the fp32 residual-stream helpers (merge, feature ensemble).
//
//   sgemm_nt_kernel        C = A W^T (+bias | +bias,GELU | +bias,+residual | +embedding table), 128x128x16 tiles,
//                          8x8 outputs per thread, register-prefetched double buffer
//   layernorm_f32_kernel   one warp per 1024-wide row, two-pass statistics
//   attention_f32_kernel   flash-style streaming softmax over 64-key blocks, decomposed rel-pos bias tables
//                          (modeling_seggpt.py:268-311) built per query tile from the UNSCALED q (:324-329)
//   decoder_f32_kernel     pixel-shuffle addressing + conv3x3 + channel LayerNorm + GELU + 1x1 head
//                          (modeling_seggpt.py:533-585), one thread per pixel, 64 channels in registers
#include <algorithm>
#include <vector>

#include "../../include/bseg.h"
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

namespace {
constexpr int kT = 1568, kD = 1024, kGridH = 56, kGridW = 28;

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ----------------------------------------------------------------------------------------------
// patchify (fp32 copy of elementwise.cu's patchify_kernel): A[(s*B+b)*1568 + ph*28 + pw][c*256 + py*16 + px]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_f32_kernel(const float* __restrict__ px, const float* __restrict__ prompt_px,
                    const float* __restrict__ prompt_mask, float* __restrict__ A, int B) {
  const long long total = 2LL * B * 3 * 896 * 112;  // (stream, b, c, y, x4)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x4 = static_cast<int>(idx % 112);
    long long r = idx / 112;
    const int y = static_cast<int>(r % 896);
    r /= 896;
    const int c = static_cast<int>(r % 3);
    r /= 3;
    const int b = static_cast<int>(r % B);
    const int s = static_cast<int>(r / B);
    const int x = x4 * 4;
    const bool top = y < 448;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s == 0 || top) {
      const float* src = (s == 0) ? (top ? prompt_px : px) : prompt_mask;
      v = *reinterpret_cast<const float4*>(src + (((long long)b * 3 + c) * 448 + (top ? y : y - 448)) * 448 + x);
    }
    const long long row = ((long long)s * B + b) * kT + (y >> 4) * kGridW + (x >> 4);
    *reinterpret_cast<float4*>(A + row * 768 + c * 256 + (y & 15) * 16 + (x & 15)) = v;
  }
}

// ----------------------------------------------------------------------------------------------
// SGEMM.  NN = false: both operands K-major, C[m][n] = sum_k A[m][k] * W[n][k] (nn.Linear forward);
//         NN = true : C[m][n] = sum_k A[m][k] * W[k][n] -- the dgrad of the same nn.Linear with its weight as stored
//         ([out, in] row-major: k runs over `out`), so the fp32 backward needs no transposed weight copies.
// ----------------------------------------------------------------------------------------------
enum : int {
  SG_BIAS = 0, SG_BIAS_GELU = 1, SG_BIAS_RESID = 2, SG_EMBED = 3,
  SG_PLAIN = 4,           // C = acc                                  (dgrad GEMMs of the fp32 backward)
  SG_DGELU = 5,           // C = acc * gelu'(aux[m][n])               (lin2 dgrad through the GELU; aux = saved pre-activation)
  SG_BIAS_GELU_SAVE = 6,  // aux[m][n] = acc + bias; C = gelu(aux)    (training forward keeps the pre-activation)
};
struct SgemmEpi {
  const float* bias = nullptr;
  const float* resid = nullptr;  // [M, ldr]
  long long ldr = 0;
  const float* tab = nullptr;    // [2][T][N] additive table, stream = row / rows_per_stream
  long long rows_per_stream = 0;
  float* aux = nullptr;          // [M, ldaux] (SG_DGELU: input, SG_BIAS_GELU_SAVE: output)
  long long ldaux = 0;
};
// d/dx of the exact-erf GELU: Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;

template <int EPI, bool NN>
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                float* __restrict__ C, long long ldc, long long M, int N, int K, SgemmEpi ep) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[2][SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.y) * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  // loader: 512 float4 per operand tile, two per thread; float4 = 4 consecutive k of one row
  const int lrow = tid >> 2;          // 0..63 (+64 for the second)
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw[2];
  auto gload = [&](int kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long m = m0 + lrow + 64 * h;
      ra[h] = (m < M) ? *reinterpret_cast<const float4*>(A + m * lda + kt * SG_BK + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (NN)  // 16 x 128 tile of W[k][n]: row k = tid / 32 (+ 8), four consecutive n per thread
        rw[h] = *reinterpret_cast<const float4*>(W + static_cast<long long>(kt * SG_BK + (tid >> 5) + 8 * h) * ldw + n0 + (tid & 31) * 4);
      else
        rw[h] = *reinterpret_cast<const float4*>(W + static_cast<long long>(n0 + lrow + 64 * h) * ldw + kt * SG_BK + lk);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      if constexpr (NN) {
        *reinterpret_cast<float4*>(&Ws[buf][(tid >> 5) + 8 * h][(tid & 31) * 4]) = rw[h];
      } else {
        Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
      }
    }
  };

  const int nk = K / SG_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows ty*4 + i (+64), cols tx*4 + j (+64)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if constexpr (EPI == SG_EMBED) {
        const long long s = m / ep.rows_per_stream, t = m % kT;
        const float4 tb = *reinterpret_cast<const float4*>(ep.tab + (s * kT + t) * N + n);
        v[0] += tb.x; v[1] += tb.y; v[2] += tb.z; v[3] += tb.w;
      } else if constexpr (EPI == SG_PLAIN) {
      } else if constexpr (EPI == SG_DGELU) {
        const float4 z = *reinterpret_cast<const float4*>(ep.aux + m * ep.ldaux + n);
        v[0] *= gelu_exact_grad(z.x); v[1] *= gelu_exact_grad(z.y); v[2] *= gelu_exact_grad(z.z); v[3] *= gelu_exact_grad(z.w);
      } else {
        const float4 bb = *reinterpret_cast<const float4*>(ep.bias + n);
        v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
      }
      if constexpr (EPI == SG_BIAS_GELU_SAVE) {
        *reinterpret_cast<float4*>(ep.aux + m * ep.ldaux + n) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = gelu_exact(v[q]);
      }
      if constexpr (EPI == SG_BIAS_GELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = gelu_exact(v[q]);
      }
      if constexpr (EPI == SG_BIAS_RESID) {
        const float4 rr = *reinterpret_cast<const float4*>(ep.resid + m * ep.ldr + n);
        v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
      }
      *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int launch_sgemm(int epi, const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                 long long M, int N, int K, const SgemmEpi& ep, cudaStream_t stream) {
  BSEG_REQUIRE(M > 0 && N % SG_BN == 0 && K % SG_BK == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0,
               "sgemm: unsupported shape M=%lld N=%d K=%d", M, N, K);
  dim3 grid(N / SG_BN, static_cast<unsigned>((M + SG_BM - 1) / SG_BM));
  ProfScope prof(CAT_GEMM, 2.0 * M * N * K, 4.0 * (M * K + static_cast<double>(N) * K + M * N), stream, 8);
  switch (epi) {
    case SG_BIAS: sgemm_nt_kernel<SG_BIAS, false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_BIAS_GELU: sgemm_nt_kernel<SG_BIAS_GELU, false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_BIAS_RESID: sgemm_nt_kernel<SG_BIAS_RESID, false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_EMBED: sgemm_nt_kernel<SG_EMBED, false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_BIAS_GELU_SAVE:
      BSEG_REQUIRE(ep.aux != nullptr && ep.ldaux % 4 == 0, "sgemm: pre-activation buffer missing");
      sgemm_nt_kernel<SG_BIAS_GELU_SAVE, false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep);
      break;
    default: BSEG_REQUIRE(false, "sgemm: unknown epilogue %d", epi);
  }
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
// dgrad: C[M, N] = A[M, K] * W[K, N] with W the nn.Linear weight as stored ([out = K, in = N], ldw = N)
int launch_sgemm_nn(int epi, const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                    long long M, int N, int K, const SgemmEpi& ep, cudaStream_t stream) {
  BSEG_REQUIRE(M > 0 && N % SG_BN == 0 && K % SG_BK == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0,
               "sgemm_nn: unsupported shape M=%lld N=%d K=%d", M, N, K);
  dim3 grid(N / SG_BN, static_cast<unsigned>((M + SG_BM - 1) / SG_BM));
  ProfScope prof(CAT_GEMM, 2.0 * M * N * K, 4.0 * (M * K + static_cast<double>(N) * K + M * N), stream, 8);
  switch (epi) {
    case SG_PLAIN: sgemm_nt_kernel<SG_PLAIN, true><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_DGELU:
      BSEG_REQUIRE(ep.aux != nullptr && ep.ldaux % 4 == 0, "sgemm_nn: pre-activation buffer missing");
      sgemm_nt_kernel<SG_DGELU, true><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep);
      break;
    default: BSEG_REQUIRE(false, "sgemm_nn: unknown epilogue %d", epi);
  }
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm over 1024 features, fp32 in / fp32 out, one warp per row
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                     const float* __restrict__ b, float* __restrict__ out, long long ldo, long long rows, float eps) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  s = warp_sum(s);
  const float mean = s * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  q = warp_sum(q);
  const float rstd = 1.0f / sqrtf(q * (1.0f / kD) + eps);
  float4* o = reinterpret_cast<float4*>(out + row * ldo);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 ww = reinterpret_cast<const float4*>(w)[lane + 32 * i];
    const float4 bb = reinterpret_cast<const float4*>(b)[lane + 32 * i];
    o[lane + 32 * i] = make_float4((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y,
                                   (v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
  }
}
int launch_layernorm_f32(const float* x, long long ldx, const float* w, const float* b, float* out, long long ldo,
                         long long rows, float eps, cudaStream_t stream) {
  ProfScope prof(CAT_LAYERNORM, 0, static_cast<double>(rows) * kD * 8, stream);
  layernorm_f32_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, ldx, w, b, out, ldo, rows, eps);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// attention, fp32.  qkv: [nseq*T, 3072] with feature = which*1024 + head*64 + d (modeling_seggpt.py:318-322);
// out: [nseq*T, 1024] with feature = head*64 + d (:341-343).  grid = (ceil(T/64), heads, nseq), 256 threads.
// ----------------------------------------------------------------------------------------------
constexpr int AQ = 64, AK = 64;
constexpr int kAttnSmemFloats = AQ * 65 + AK * 65 + AK * 64 + AQ * 65 + AQ * kGridH + AQ * kGridW + 3 * AQ;

__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ rel_h /*[111,64]*/,
                     const float* __restrict__ rel_w /*[55,64]*/, float* __restrict__ out) {
  extern __shared__ float smf[];
  float* Qs = smf;                    // [64][65]
  float* Ks = Qs + AQ * 65;           // [64][65]
  float* Vs = Ks + AK * 65;           // [64][64]
  float* Ps = Vs + AK * 64;           // [64][65]
  float* Gh = Ps + AQ * 65;           // [64][56]
  float* Gw = Gh + AQ * kGridH;       // [64][28]
  float* row_m = Gw + AQ * kGridW;    // running max
  float* row_l = row_m + AQ;          // running sum
  float* row_a = row_l + AQ;          // rescale factor of the current block
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * AQ, head = blockIdx.y, seq = blockIdx.z;
  const float* base = qkv + static_cast<long long>(seq) * kT * 3072 + head * 64;

  for (int i = tid; i < AQ * 16; i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < kT) v = *reinterpret_cast<const float4*>(base + static_cast<long long>(q0 + r) * 3072 + c4);
    Qs[r * 65 + c4] = v.x; Qs[r * 65 + c4 + 1] = v.y; Qs[r * 65 + c4 + 2] = v.z; Qs[r * 65 + c4 + 3] = v.w;
  }
  if (tid < AQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }
  __syncthreads();
  // decomposed rel-pos tables: Gh[r][kh] = q_r . rel_h[qh - kh + 55], Gw[r][kw] = q_r . rel_w[qw - kw + 27]
  for (int i = tid; i < AQ * (kGridH + kGridW); i += 256) {
    const int r = i / (kGridH + kGridW), j = i % (kGridH + kGridW);
    const int q = min(q0 + r, kT - 1);
    const float* tab = (j < kGridH) ? rel_h + (q / kGridW - j + kGridH - 1) * 64
                                    : rel_w + (q % kGridW - (j - kGridH) + kGridW - 1) * 64;
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) acc = fmaf(Qs[r * 65 + d], tab[d], acc);
    if (j < kGridH) Gh[r * kGridH + j] = acc; else Gw[r * kGridW + (j - kGridH)] = acc;
  }

  const int ty = tid >> 4, tx = tid & 15;
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

  for (int k0 = 0; k0 < kT; k0 += AK) {
    __syncthreads();  // previous block's P*V reads are done (and the G tables are written, first iteration)
    for (int i = tid; i < AK * 16; i += 256) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < kT) {
        const float* p = base + static_cast<long long>(k0 + r) * 3072 + c4;
        kv = *reinterpret_cast<const float4*>(p + 1024);
        vv = *reinterpret_cast<const float4*>(p + 2048);
      }
      Ks[r * 65 + c4] = kv.x; Ks[r * 65 + c4 + 1] = kv.y; Ks[r * 65 + c4 + 2] = kv.z; Ks[r * 65 + c4 + 3] = kv.w;
      *reinterpret_cast<float4*>(Vs + r * 64 + c4) = vv;
    }
    __syncthreads();
    // S tile: rows ty*4 + i, columns tx + 16*j
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(ty * 4 + i) * 65 + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ks[(tx + 16 * j) * 65 + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty * 4 + i, key = k0 + tx + 16 * j;
        float v = -INFINITY;
        if (key < kT) v = (s[i][j] * 0.125f + Gh[r * kGridH + key / kGridW]) + Gw[r * kGridW + key % kGridW];
        Ps[r * 65 + tx + 16 * j] = v;
      }
    __syncthreads();
    // online softmax, 4 threads per row (columns c = part + 4*i)
    {
      const int r = tid >> 2, part = tid & 3;
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, Ps[r * 65 + part + 4 * i]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_old = row_m[r];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p = expf(Ps[r * 65 + part + 4 * i] - m_new);
        Ps[r * 65 + part + 4 * i] = p;
        sum += p;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp();
      if (part == 0) {
        const float a = expf(m_old - m_new);  // 0 for the first block (m_old = -inf)
        row_a[r] = a;
        row_l[r] = row_l[r] * a + sum;
        row_m[r] = m_new;
      }
    }
    __syncthreads();
    // O[rows ty*4+i][d = tx*4 .. tx*4+3] = O * a + P V
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = row_a[ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= a;
    }
#pragma unroll 4
    for (int kk = 0; kk < AK; ++kk) {
      const float4 vv = *reinterpret_cast<const float4*>(Vs + kk * 64 + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Ps[(ty * 4 + i) * 65 + kk];
        o[i][0] = fmaf(p, vv.x, o[i][0]);
        o[i][1] = fmaf(p, vv.y, o[i][1]);
        o[i][2] = fmaf(p, vv.z, o[i][2]);
        o[i][3] = fmaf(p, vv.w, o[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (q0 + r >= kT) continue;
    const float inv = 1.0f / row_l[r];
    *reinterpret_cast<float4*>(out + (static_cast<long long>(seq) * kT + q0 + r) * kD + head * 64 + tx * 4) =
        make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
  }
}

int launch_attention_f32(const float* qkv, const float* rel_h, const float* rel_w, float* out, int nseq,
                         cudaStream_t stream) {
  const size_t smem = kAttnSmemFloats * sizeof(float);
  static PerDeviceFlag attr_once;
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
  }
  dim3 grid((kT + AQ - 1) / AQ, BSEG_HEADS, nseq);
  ProfScope prof(CAT_ATTENTION, 4.0 * nseq * BSEG_HEADS * kT * static_cast<double>(kT) * 64,
                 static_cast<double>(nseq) * kT * 4096 * 4, stream);
  attention_f32_kernel<<<grid, 256, smem, stream>>>(qkv, rel_h, rel_w, out);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// decoder head, fp32 (modeling_seggpt.py:533-585).  dec: output of decoder_embed, [B*T, 16384] with feature
// (p1*16 + p2)*64 + c of token (gh, gw) == channel c of pixel (gh*16 + p1, gw*16 + p2) (the reshape/permute of
// :575-578 is pure addressing).  One CTA = 16x16 pixels, one thread = one pixel with its 64 conv outputs in registers.
// ----------------------------------------------------------------------------------------------
constexpr int DC = 8;  // input channels per smem chunk (in_s 11.7 KB + w_s 18.4 KB of static shared memory)
__global__ void __launch_bounds__(256)
decoder_f32_kernel(const float* __restrict__ dec, const float* __restrict__ conv_w /*[64,64,3,3]*/,
                   const float* __restrict__ conv_b, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                   const float* __restrict__ head_w /*[3,64]*/, const float* __restrict__ head_b,
                   float* __restrict__ pred /*[B,3,896,448]*/, float eps) {
  __shared__ float in_s[18 * 18][DC + 1];
  __shared__ __align__(16) float w_s[9][DC][64];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
  const int py = tid >> 4, pxl = tid & 15;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = conv_b[c];

  for (int c0 = 0; c0 < 64; c0 += DC) {
    __syncthreads();
    for (int i = tid; i < 18 * 18 * (DC / 4); i += 256) {
      const int pix = i / (DC / 4), c4 = (i % (DC / 4)) * 4;
      const int y = y0 - 1 + pix / 18, x = x0 - 1 + pix % 18;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < 896 && x >= 0 && x < 448) {
        const long long tok = static_cast<long long>(b) * kT + (y >> 4) * kGridW + (x >> 4);
        v = *reinterpret_cast<const float4*>(dec + tok * 16384 + ((y & 15) * 16 + (x & 15)) * 64 + c0 + c4);
      }
      in_s[pix][c4] = v.x; in_s[pix][c4 + 1] = v.y; in_s[pix][c4 + 2] = v.z; in_s[pix][c4 + 3] = v.w;
    }
    for (int i = tid; i < 9 * DC * 64; i += 256) {
      const int co = i & 63, ci = (i >> 6) % DC, tap = i / (64 * DC);
      w_s[tap][ci][co] = conv_w[(co * 64 + c0 + ci) * 9 + tap];
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* ip = in_s[(py + tap / 3) * 18 + pxl + tap % 3];
#pragma unroll 4
      for (int ci = 0; ci < DC; ++ci) {
        const float xv = ip[ci];
        const float4* wp = reinterpret_cast<const float4*>(w_s[tap][ci]);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 w = wp[c4];
          acc[c4 * 4 + 0] = fmaf(xv, w.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(xv, w.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(xv, w.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(xv, w.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  // channel LayerNorm (SegGptLayerNorm channels_first, :508-530) + GELU + 1x1 head
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) mean += acc[c];
  mean *= (1.0f / 64);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) var += (acc[c] - mean) * (acc[c] - mean);
  const float rstd = 1.0f / sqrtf(var * (1.0f / 64) + eps);
  float r0 = head_b[0], r1 = head_b[1], r2 = head_b[2];
#pragma unroll
  for (int c = 0; c < 64; ++c) {
    const float g = gelu_exact((acc[c] - mean) * rstd * ln_w[c] + ln_b[c]);
    r0 = fmaf(g, head_w[c], r0);
    r1 = fmaf(g, head_w[64 + c], r1);
    r2 = fmaf(g, head_w[128 + c], r2);
  }
  const int y = y0 + py, x = x0 + pxl;
  const long long plane = 896ll * 448;
  float* o = pred + static_cast<long long>(b) * 3 * plane + static_cast<long long>(y) * 448 + x;
  o[0] = r0; o[plane] = r1; o[2 * plane] = r2;
}

int launch_decoder_f32(const float* dec, const F32Weights& w, float* pred, int B, float eps, cudaStream_t stream) {
  dim3 grid(448 / 16, 896 / 16, B);
  ProfScope prof(CAT_DECODER_HEAD, 2.0 * B * 896 * 448 * 64 * (9 * 64 + 3), static_cast<double>(B) * 896 * 448 * 67 * 4,
                 stream);
  decoder_f32_kernel<<<grid, 256, 0, stream>>>(dec, w.dec_conv_w, w.dec_conv_b, w.dec_ln_w, w.dec_ln_b, w.dec_head_w,
                                               w.dec_head_b, pred, eps);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ==============================================================================================
// fp32 BACKWARD to the prompt pixels (autograd of modeling_seggpt.py under src/model.py:233-269, frozen backbone):
// the accuracy-mode counterpart of bseg_backward_to_prompt.  Everything IEEE fp32 on the CUDA cores, deterministic
// (no atomics), simple rather than fast.
// ==============================================================================================

// LayerNorm backward, one warp per 1024-wide row: dh[row] += d LN(x)[row] / dx for upstream gradient dy (statistics
// recomputed from the saved input).  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
__global__ void __launch_bounds__(256)
layernorm_bwd_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy,
                         const float* __restrict__ gamma, float* __restrict__ dh, long long rows, float eps) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const float4* gr = reinterpret_cast<const float4*>(dy + row * lddy);
  float4 v[8], g[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + eps);
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 ww = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
    const float4 d = gr[lane + 32 * i];
    g[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
    v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
    c1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    c2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
  }
  c1 = warp_sum(c1) * (1.0f / kD);
  c2 = warp_sum(c2) * (1.0f / kD);
  float4* o = reinterpret_cast<float4*>(dh + row * kD);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 a = o[lane + 32 * i];
    a.x += rstd * (g[i].x - c1 - v[i].x * c2);
    a.y += rstd * (g[i].y - c1 - v[i].y * c2);
    a.z += rstd * (g[i].z - c1 - v[i].z * c2);
    a.w += rstd * (g[i].w - c1 - v[i].w * c2);
    o[lane + 32 * i] = a;
  }
}
int launch_layernorm_bwd_f32(const float* x, long long ldx, const float* dy, long long lddy, const float* gamma, float* dh,
                             long long rows, float eps, cudaStream_t stream) {
  ProfScope prof(CAT_LAYERNORM, 0, static_cast<double>(rows) * kD * 16, stream);
  layernorm_bwd_f32_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, ldx, dy, lddy, gamma, dh, rows, eps);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

__global__ void scale_f32_kernel(float4* __restrict__ p, float sc, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = p[i];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    p[i] = v;
  }
}

// ---- attention backward, fp32 -----------------------------------------------------------------
// score[q][k] = 0.125 q.k + q.Rh[qh - kh + 55] + q.Rw[qw - kw + 27]   (q UNscaled in the bias terms, :324-329)
// P = softmax(score), O = P V.  With D[q] = dO[q].O[q] and dS = P * (dO V^T - D):
//   dQ[q] = 0.125 sum_k dS[q][k] K[k] + sum_kh (sum_kw dS[q][kh,kw]) Rh[qh - kh + 55] + sum_kw (sum_kh dS) Rw[qw - kw + 27]
//   dK[k] = 0.125 sum_q dS[q][k] Q[q],   dV[k] = sum_q P[q][k] dO[q]          (rel-pos tables are frozen)
// Kernel 1 (one CTA per 64 queries of one head): log-sum-exp pass, D, then dQ; writes lse and D for kernel 2.
// Kernel 2 (one CTA per 64 keys of one head): dK and dV, looping over the query blocks.
constexpr int kAbwdQSmemFloats = 5 * AQ * 65 + 2 * AQ * (kGridH + kGridW) + 3 * AQ;
__global__ void __launch_bounds__(256)
attention_bwd_dq_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ O, const float* __restrict__ dO,
                            const float* __restrict__ rel_h, const float* __restrict__ rel_w, float* __restrict__ dqkv,
                            float* __restrict__ lse_out, float* __restrict__ d_out) {
  extern __shared__ float smf[];
  float* Qs = smf;                    // [64][65]
  float* Ks = Qs + AQ * 65;           // [64][65]
  float* Vs = Ks + AK * 65;           // [64][65]
  float* dOs = Vs + AK * 65;          // [64][65]
  float* Ps = dOs + AQ * 65;          // [64][65]  scores, then dS
  float* Gh = Ps + AQ * 65;           // [64][56]
  float* Gw = Gh + AQ * kGridH;       // [64][28]
  float* dGh = Gw + AQ * kGridW;      // [64][56]  sum over kw of dS
  float* dGw = dGh + AQ * kGridH;     // [64][28]  sum over kh of dS
  float* row_m = dGw + AQ * kGridW;   // running max, then lse
  float* row_l = row_m + AQ;
  float* row_d = row_l + AQ;          // D
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * AQ, head = blockIdx.y, seq = blockIdx.z;
  const float* base = qkv + static_cast<long long>(seq) * kT * 3072 + head * 64;

  for (int i = tid; i < AQ * 16; i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = v;
    if (q0 + r < kT) {
      v = *reinterpret_cast<const float4*>(base + static_cast<long long>(q0 + r) * 3072 + c4);
      g = *reinterpret_cast<const float4*>(dO + (static_cast<long long>(seq) * kT + q0 + r) * kD + head * 64 + c4);
    }
    Qs[r * 65 + c4] = v.x; Qs[r * 65 + c4 + 1] = v.y; Qs[r * 65 + c4 + 2] = v.z; Qs[r * 65 + c4 + 3] = v.w;
    dOs[r * 65 + c4] = g.x; dOs[r * 65 + c4 + 1] = g.y; dOs[r * 65 + c4 + 2] = g.z; dOs[r * 65 + c4 + 3] = g.w;
  }
  if (tid < AQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }
  for (int i = tid; i < AQ * (kGridH + kGridW); i += 256) dGh[i] = 0.f;  // (dGh and dGw are contiguous)
  __syncthreads();
  for (int i = tid; i < AQ * (kGridH + kGridW); i += 256) {
    const int r = i / (kGridH + kGridW), j = i % (kGridH + kGridW);
    const int q = min(q0 + r, kT - 1);
    const float* tab = (j < kGridH) ? rel_h + (q / kGridW - j + kGridH - 1) * 64
                                    : rel_w + (q % kGridW - (j - kGridH) + kGridW - 1) * 64;
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) acc = fmaf(Qs[r * 65 + d], tab[d], acc);
    if (j < kGridH) Gh[r * kGridH + j] = acc; else Gw[r * kGridW + (j - kGridH)] = acc;
  }
  // D[r] = dO[r] . O[r]  (4 threads per row)
  {
    const int r = tid >> 2, part = tid & 3;
    float acc = 0.f;
    if (q0 + r < kT) {
      const float* op = O + (static_cast<long long>(seq) * kT + q0 + r) * kD + head * 64 + part * 16;
#pragma unroll
      for (int d = 0; d < 16; ++d) acc = fmaf(dOs[r * 65 + part * 16 + d], op[d], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) row_d[r] = acc;
  }
  const int ty = tid >> 4, tx = tid & 15;

  auto load_kv = [&](int k0, bool want_v) {
    for (int i = tid; i < AK * 16; i += 256) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < kT) {
        const float* p = base + static_cast<long long>(k0 + r) * 3072 + c4;
        kv = *reinterpret_cast<const float4*>(p + 1024);
        if (want_v) vv = *reinterpret_cast<const float4*>(p + 2048);
      }
      Ks[r * 65 + c4] = kv.x; Ks[r * 65 + c4 + 1] = kv.y; Ks[r * 65 + c4 + 2] = kv.z; Ks[r * 65 + c4 + 3] = kv.w;
      if (want_v) { Vs[r * 65 + c4] = vv.x; Vs[r * 65 + c4 + 1] = vv.y; Vs[r * 65 + c4 + 2] = vv.z; Vs[r * 65 + c4 + 3] = vv.w; }
    }
  };
  // biased score of (row ty*4+i, key k0 + tx + 16 j), -inf past the sequence
  auto scores = [&](int k0, float (&s)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(ty * 4 + i) * 65 + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ks[(tx + 16 * j) * 65 + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty * 4 + i, key = k0 + tx + 16 * j;
        s[i][j] = key < kT ? (s[i][j] * 0.125f + Gh[r * kGridH + key / kGridW]) + Gw[r * kGridW + key % kGridW] : -INFINITY;
      }
  };

  // ---- pass 1: log-sum-exp of every row ----
  for (int k0 = 0; k0 < kT; k0 += AK) {
    __syncthreads();
    load_kv(k0, false);
    __syncthreads();
    float s[4][4];
    scores(k0, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ps[(ty * 4 + i) * 65 + tx + 16 * j] = s[i][j];
    __syncthreads();
    {
      const int r = tid >> 2, part = tid & 3;
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, Ps[r * 65 + part + 4 * i]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_old = row_m[r];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) sum += expf(Ps[r * 65 + part + 4 * i] - m_new);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp();
      if (part == 0) {
        row_l[r] = row_l[r] * expf(m_old - m_new) + sum;
        row_m[r] = m_new;
      }
    }
  }
  __syncthreads();
  if (tid < AQ) {
    row_m[tid] = row_m[tid] + logf(row_l[tid]);  // lse
    if (q0 + tid < kT) {
      const long long o = (static_cast<long long>(seq) * BSEG_HEADS + head) * kT + q0 + tid;
      lse_out[o] = row_m[tid];
      d_out[o] = row_d[tid];
    }
  }
  __syncthreads();

  // ---- pass 2: dS, dQ, bias-gradient row sums ----
  float dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  for (int k0 = 0; k0 < kT; k0 += AK) {
    __syncthreads();
    load_kv(k0, true);
    __syncthreads();
    float s[4][4], dp[4][4];
    scores(k0, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = dOs[(ty * 4 + i) * 65 + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Vs[(tx + 16 * j) * 65 + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dp[i][j] = fmaf(a[i], b[j], dp[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty * 4 + i;
        const float p = expf(s[i][j] - row_m[r]);  // exp(-inf) = 0 past the sequence
        Ps[r * 65 + tx + 16 * j] = p * (dp[i][j] - row_d[r]);
      }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < AK; ++kk) {
      float kv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = Ks[kk * 65 + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ds = Ps[(ty * 4 + i) * 65 + kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = fmaf(ds, kv[j], dq[i][j]);
      }
    }
    if (tid < AQ) {  // one thread per row, keys in order: deterministic
      for (int c = 0; c < AK; ++c) {
        const int key = k0 + c;
        if (key < kT) {
          const float ds = Ps[tid * 65 + c];
          dGh[tid * kGridH + key / kGridW] += ds;
          dGw[tid * kGridW + key % kGridW] += ds;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i, q = q0 + r;
    if (q >= kT) continue;
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = dq[i][j] * 0.125f;
    const int qh = q / kGridW, qw = q % kGridW;
    for (int kh = 0; kh < kGridH; ++kh) {
      const float g = dGh[r * kGridH + kh];
      const float4 t = *reinterpret_cast<const float4*>(rel_h + (qh - kh + kGridH - 1) * 64 + tx * 4);
      acc[0] = fmaf(g, t.x, acc[0]); acc[1] = fmaf(g, t.y, acc[1]); acc[2] = fmaf(g, t.z, acc[2]); acc[3] = fmaf(g, t.w, acc[3]);
    }
    for (int kw = 0; kw < kGridW; ++kw) {
      const float g = dGw[r * kGridW + kw];
      const float4 t = *reinterpret_cast<const float4*>(rel_w + (qw - kw + kGridW - 1) * 64 + tx * 4);
      acc[0] = fmaf(g, t.x, acc[0]); acc[1] = fmaf(g, t.y, acc[1]); acc[2] = fmaf(g, t.z, acc[2]); acc[3] = fmaf(g, t.w, acc[3]);
    }
    *reinterpret_cast<float4*>(dqkv + (static_cast<long long>(seq) * kT + q) * 3072 + head * 64 + tx * 4) =
        make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

constexpr int kAbwdKSmemFloats = 5 * AQ * 65 + AQ * 4 + AQ * kGridW + 2 * AQ;
__global__ void __launch_bounds__(256)
attention_bwd_dkv_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, const float* __restrict__ lse,
                             const float* __restrict__ dvec, const float* __restrict__ rel_h,
                             const float* __restrict__ rel_w, float* __restrict__ dqkv) {
  extern __shared__ float smf[];
  float* Ks = smf;                    // [64][65] keys of this CTA
  float* Vs = Ks + AK * 65;
  float* Qs = Vs + AK * 65;           // [64][65] current query block
  float* dOs = Qs + AQ * 65;
  float* Ps = dOs + AQ * 65;          // P, then dS
  float* Gh = Ps + AQ * 65;           // [64][4]   q . Rh for the (at most four) key rows of this CTA
  float* Gw = Gh + AQ * 4;            // [64][28]
  float* ls = Gw + AQ * kGridW;       // lse of the query block
  float* ds_ = ls + AQ;               // D of the query block
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * AK, head = blockIdx.y, seq = blockIdx.z;
  const float* base = qkv + static_cast<long long>(seq) * kT * 3072 + head * 64;
  const int kh0 = k0 / kGridW;
  for (int i = tid; i < AK * 16; i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (k0 + r < kT) {
      const float* p = base + static_cast<long long>(k0 + r) * 3072 + c4;
      kv = *reinterpret_cast<const float4*>(p + 1024);
      vv = *reinterpret_cast<const float4*>(p + 2048);
    }
    Ks[r * 65 + c4] = kv.x; Ks[r * 65 + c4 + 1] = kv.y; Ks[r * 65 + c4 + 2] = kv.z; Ks[r * 65 + c4 + 3] = kv.w;
    Vs[r * 65 + c4] = vv.x; Vs[r * 65 + c4 + 1] = vv.y; Vs[r * 65 + c4 + 2] = vv.z; Vs[r * 65 + c4 + 3] = vv.w;
  }
  const int ty = tid >> 4, tx = tid & 15;
  float dk[4][4], dv[4][4];  // key rows ty*4 + i, features tx*4 + j
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }

  for (int q0 = 0; q0 < kT; q0 += AQ) {
    __syncthreads();
    for (int i = tid; i < AQ * 16; i += 256) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = v;
      if (q0 + r < kT) {
        v = *reinterpret_cast<const float4*>(base + static_cast<long long>(q0 + r) * 3072 + c4);
        g = *reinterpret_cast<const float4*>(dO + (static_cast<long long>(seq) * kT + q0 + r) * kD + head * 64 + c4);
      }
      Qs[r * 65 + c4] = v.x; Qs[r * 65 + c4 + 1] = v.y; Qs[r * 65 + c4 + 2] = v.z; Qs[r * 65 + c4 + 3] = v.w;
      dOs[r * 65 + c4] = g.x; dOs[r * 65 + c4 + 1] = g.y; dOs[r * 65 + c4 + 2] = g.z; dOs[r * 65 + c4 + 3] = g.w;
    }
    if (tid < AQ) {
      const bool ok = q0 + tid < kT;
      const long long o = (static_cast<long long>(seq) * BSEG_HEADS + head) * kT + q0 + tid;
      ls[tid] = ok ? lse[o] : INFINITY;  // exp(x - inf) = 0: rows past the sequence contribute nothing
      ds_[tid] = ok ? dvec[o] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < AQ * 32; i += 256) {
      const int r = i >> 5, j = i & 31;
      const int q = min(q0 + r, kT - 1);
      const float* tab;
      if (j < 4) {
        const int kh = min(kh0 + j, kGridH - 1);
        tab = rel_h + (q / kGridW - kh + kGridH - 1) * 64;
      } else {
        tab = rel_w + (q % kGridW - (j - 4) + kGridW - 1) * 64;
      }
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < 64; ++d) acc = fmaf(Qs[r * 65 + d], tab[d], acc);
      if (j < 4) Gh[r * 4 + j] = acc; else Gw[r * kGridW + (j - 4)] = acc;
    }
    __syncthreads();
    // S and dP for (query row ty*4+i, key tx + 16 j)
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float a[4], b[4], g[4], vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Qs[(ty * 4 + i) * 65 + d]; g[i] = dOs[(ty * 4 + i) * 65 + d]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { b[j] = Ks[(tx + 16 * j) * 65 + d]; vv[j] = Vs[(tx + 16 * j) * 65 + d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[i][j] = fmaf(a[i], b[j], s[i][j]); dp[i][j] = fmaf(g[i], vv[j], dp[i][j]); }
    }
    float pds[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty * 4 + i, c = tx + 16 * j, key = k0 + c;
        float p = 0.f;
        if (key < kT) p = expf(((s[i][j] * 0.125f + Gh[r * 4 + key / kGridW - kh0]) + Gw[r * kGridW + key % kGridW]) - ls[r]);
        Ps[r * 65 + c] = p;
        pds[i][j] = p * (dp[i][j] - ds_[r]);
      }
    __syncthreads();
    // dV[key ty*4+i][tx*4+j] += sum_r P[r][key] dO[r][.]
#pragma unroll 4
    for (int r = 0; r < AQ; ++r) {
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = dOs[r * 65 + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Ps[r * 65 + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) dv[i][j] = fmaf(p, g[j], dv[i][j]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ps[(ty * 4 + i) * 65 + tx + 16 * j] = pds[i][j];
    __syncthreads();
    // dK[key ty*4+i][tx*4+j] += sum_r dS[r][key] Q[r][.]
#pragma unroll 4
    for (int r = 0; r < AQ; ++r) {
      float qv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) qv[j] = Qs[r * 65 + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = Ps[r * 65 + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) dk[i][j] = fmaf(d, qv[j], dk[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + ty * 4 + i;
    if (key >= kT) continue;
    float* o = dqkv + (static_cast<long long>(seq) * kT + key) * 3072 + head * 64 + tx * 4;
    *reinterpret_cast<float4*>(o + 1024) = make_float4(dk[i][0] * 0.125f, dk[i][1] * 0.125f, dk[i][2] * 0.125f, dk[i][3] * 0.125f);
    *reinterpret_cast<float4*>(o + 2048) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
  }
}

int launch_attention_bwd_f32(const float* qkv, const float* O, const float* dO, const float* rel_h, const float* rel_w,
                             float* lse, float* dvec, float* dqkv, int nseq, cudaStream_t stream) {
  const size_t smem_q = kAbwdQSmemFloats * sizeof(float), smem_k = kAbwdKSmemFloats * sizeof(float);
  static PerDeviceFlag attr_once;
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dq_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_q)));
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_k)));
  }
  dim3 grid((kT + AQ - 1) / AQ, BSEG_HEADS, nseq);
  ProfScope prof(CAT_ATTENTION, 14.0 * nseq * BSEG_HEADS * kT * static_cast<double>(kT) * 64,
                 static_cast<double>(nseq) * kT * 4096 * 4 * 3, stream);
  attention_bwd_dq_f32_kernel<<<grid, 256, smem_q, stream>>>(qkv, O, dO, rel_h, rel_w, dqkv, lse, dvec);
  BSEG_CHECK_CUDA(cudaGetLastError());
  attention_bwd_dkv_f32_kernel<<<grid, 256, smem_k, stream>>>(qkv, dO, lse, dvec, rel_h, rel_w, dqkv);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

// ---- decoder head backward, fp32 ----------------------------------------------------------------
// (1) per pixel of the query half (image rows >= 448; the loss never looks at the prompt half, src/model.py:48-57):
//     recompute conv3x3 -> LayerNorm(C=64) -> GELU, then d pred -> d(conv output), written NHWC.
// (2) conv3x3 dgrad (flipped taps, transposed weights) -> rows of the decoder_embed output gradient; the pixel
//     un-shuffle (modeling_seggpt.py:575-578) is the addressing.
__global__ void __launch_bounds__(256)
decoder_head_bwd_f32_kernel(const float* __restrict__ dec, const float* __restrict__ conv_w,
                            const float* __restrict__ conv_b, const float* __restrict__ ln_w,
                            const float* __restrict__ ln_b, const float* __restrict__ head_w,
                            const float* __restrict__ d_pred /*[B,3,896,448]*/, float* __restrict__ dconv /*[B,896,448,64]*/,
                            float eps) {
  __shared__ float in_s[18 * 18][DC + 1];
  __shared__ __align__(16) float w_s[9][DC][64];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = 448 + blockIdx.y * 16, x0 = blockIdx.x * 16;
  const int py = tid >> 4, pxl = tid & 15;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = conv_b[c];
  for (int c0 = 0; c0 < 64; c0 += DC) {
    __syncthreads();
    for (int i = tid; i < 18 * 18 * (DC / 4); i += 256) {
      const int pix = i / (DC / 4), c4 = (i % (DC / 4)) * 4;
      const int y = y0 - 1 + pix / 18, x = x0 - 1 + pix % 18;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < 896 && x >= 0 && x < 448) {
        const long long tok = static_cast<long long>(b) * kT + (y >> 4) * kGridW + (x >> 4);
        v = *reinterpret_cast<const float4*>(dec + tok * 16384 + ((y & 15) * 16 + (x & 15)) * 64 + c0 + c4);
      }
      in_s[pix][c4] = v.x; in_s[pix][c4 + 1] = v.y; in_s[pix][c4 + 2] = v.z; in_s[pix][c4 + 3] = v.w;
    }
    for (int i = tid; i < 9 * DC * 64; i += 256) {
      const int co = i & 63, ci = (i >> 6) % DC, tap = i / (64 * DC);
      w_s[tap][ci][co] = conv_w[(co * 64 + c0 + ci) * 9 + tap];
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* ip = in_s[(py + tap / 3) * 18 + pxl + tap % 3];
#pragma unroll 4
      for (int ci = 0; ci < DC; ++ci) {
        const float xv = ip[ci];
        const float4* wp = reinterpret_cast<const float4*>(w_s[tap][ci]);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 w = wp[c4];
          acc[c4 * 4 + 0] = fmaf(xv, w.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(xv, w.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(xv, w.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(xv, w.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) mean += acc[c];
  mean *= (1.0f / 64);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) var += (acc[c] - mean) * (acc[c] - mean);
  const float rstd = 1.0f / sqrtf(var * (1.0f / 64) + eps);
  const int y = y0 + py, x = x0 + pxl;
  const long long plane = 896ll * 448;
  const float* gp = d_pred + static_cast<long long>(b) * 3 * plane + static_cast<long long>(y) * 448 + x;
  const float g0 = gp[0], g1 = gp[plane], g2 = gp[2 * plane];
  // acc[c] <- gy[c] = d/d(LN output before the affine) ; xhat kept implicitly as (acc - mean) * rstd
  float c1 = 0.f, c2 = 0.f;
  float xh[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) {
    xh[c] = (acc[c] - mean) * rstd;
    const float u = xh[c] * ln_w[c] + ln_b[c];
    const float dg = (g0 * head_w[c] + g1 * head_w[64 + c]) + g2 * head_w[128 + c];
    const float gy = dg * gelu_exact_grad(u) * ln_w[c];
    acc[c] = gy;
    c1 += gy;
    c2 += gy * xh[c];
  }
  c1 *= (1.0f / 64);
  c2 *= (1.0f / 64);
  float* o = dconv + ((static_cast<long long>(b) * 896 + y) * 448 + x) * 64;
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4)
    *reinterpret_cast<float4*>(o + c4 * 4) =
        make_float4(rstd * (acc[c4 * 4] - c1 - xh[c4 * 4] * c2), rstd * (acc[c4 * 4 + 1] - c1 - xh[c4 * 4 + 1] * c2),
                    rstd * (acc[c4 * 4 + 2] - c1 - xh[c4 * 4 + 2] * c2), rstd * (acc[c4 * 4 + 3] - c1 - xh[c4 * 4 + 3] * c2));
}

__global__ void __launch_bounds__(256)
decoder_conv_dgrad_f32_kernel(const float* __restrict__ dconv /*[B,896,448,64]*/, const float* __restrict__ conv_w,
                              float* __restrict__ ddec /*[B*T,16384]*/, int y_first) {
  __shared__ float in_s[18 * 18][DC + 1];
  __shared__ __align__(16) float w_s[9][DC][64];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = y_first + blockIdx.y * 16, x0 = blockIdx.x * 16;
  const int py = tid >> 4, pxl = tid & 15;
  float acc[64];  // d(input channel ci) of this pixel
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = 0.f;
  for (int c0 = 0; c0 < 64; c0 += DC) {  // chunk of OUTPUT channels co
    __syncthreads();
    for (int i = tid; i < 18 * 18 * (DC / 4); i += 256) {
      const int pix = i / (DC / 4), c4 = (i % (DC / 4)) * 4;
      const int y = y0 - 1 + pix / 18, x = x0 - 1 + pix % 18;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < 896 && x >= 0 && x < 448)
        v = *reinterpret_cast<const float4*>(dconv + ((static_cast<long long>(b) * 896 + y) * 448 + x) * 64 + c0 + c4);
      in_s[pix][c4] = v.x; in_s[pix][c4 + 1] = v.y; in_s[pix][c4 + 2] = v.z; in_s[pix][c4 + 3] = v.w;
    }
    for (int i = tid; i < 9 * DC * 64; i += 256) {
      const int ci = i & 63, co = (i >> 6) % DC, tap = i / (64 * DC);
      w_s[tap][co][ci] = conv_w[((c0 + co) * 64 + ci) * 9 + tap];
    }
    __syncthreads();
    // out[y][x][co] = sum in[y + ky - 1][x + kx - 1][ci] w[co][ci][ky][kx]
    //   => d in[y][x][ci] = sum_{ky,kx,co} dconv[y - ky + 1][x - kx + 1][co] w[co][ci][ky][kx]; halo index = py + 2 - ky
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* ip = in_s[(py + 2 - tap / 3) * 18 + pxl + 2 - tap % 3];
#pragma unroll 4
      for (int co = 0; co < DC; ++co) {
        const float gv = ip[co];
        const float4* wp = reinterpret_cast<const float4*>(w_s[tap][co]);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 w = wp[c4];
          acc[c4 * 4 + 0] = fmaf(gv, w.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(gv, w.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(gv, w.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(gv, w.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  const int y = y0 + py, x = x0 + pxl;
  const long long tok = static_cast<long long>(b) * kT + (y >> 4) * kGridW + (x >> 4);
  float* o = ddec + tok * 16384 + ((y & 15) * 16 + (x & 15)) * 64;
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4)
    *reinterpret_cast<float4*>(o + c4 * 4) = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// workspace layout and the forward itself (mirrors forward_impl in api.cu step by step)
// ----------------------------------------------------------------------------------------------
size_t f32_workspace_bytes(int B) {
  const size_t rows2 = 2ull * B * kT, rows1 = 1ull * B * kT;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  return al(rows2 * kD * 4) * 3 + al(rows2 * 3072 * 4) + al(rows2 * 4096 * 4) + al(rows1 * 4096 * 4) +
         al(rows1 * 16384 * 4);
}

namespace {
// Saved activations of a training forward in fp32 (what backward_f32_impl reads).  Rows: 2*B*T for the two-stream
// layers, B*T afterwards; the image stream (the only one with a path to the prompt pixels) is rows [0, B*T).
struct F32Saved {
  struct Layer { float *h_mid, *h_out, *qkv, *att, *z; };
  float* h_emb = nullptr;
  float* dec = nullptr;
  std::vector<Layer> layers;
};
int forward_f32_core(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int B, int embedding_type, int P, float* hbuf, float* xn, float* att,
                     float* qkv, float* mlp, float* inter, float* dec, const F32Saved* save, float* pred_masks,
                     cudaStream_t stream);
}  // namespace

int forward_f32_impl(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int B, int embedding_type, int P, void* workspace, float* pred_masks,
                     cudaStream_t stream) {
  const size_t rows2 = 2ull * B * kT, rows1 = 1ull * B * kT;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* hbuf = reinterpret_cast<float*>(ws); ws += al(rows2 * kD * 4);
  float* xn = reinterpret_cast<float*>(ws);   ws += al(rows2 * kD * 4);
  float* att = reinterpret_cast<float*>(ws);  ws += al(rows2 * kD * 4);
  float* qkv = reinterpret_cast<float*>(ws);  ws += al(rows2 * 3072 * 4);
  float* mlp = reinterpret_cast<float*>(ws);  ws += al(rows2 * 4096 * 4);
  float* inter = reinterpret_cast<float*>(ws); ws += al(rows1 * 4096 * 4);
  float* dec = reinterpret_cast<float*>(ws);
  return forward_f32_core(w, pixel_values, prompt_pixel_values, prompt_masks, B, embedding_type, P, hbuf, xn, att, qkv,
                          mlp, inter, dec, nullptr, pred_masks, stream);
}

namespace {
// One forward pass.  save == nullptr: inference (one in-place residual stream `hbuf`, per-layer-reused qkv / att);
// otherwise every layer writes into its own slots of `save`.
int forward_f32_core(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int B, int embedding_type, int P, float* hbuf, float* xn, float* att_shared,
                     float* qkv_shared, float* mlp, float* inter, float* dec_shared, const F32Saved* save,
                     float* pred_masks, cudaStream_t stream) {
  const size_t rows2 = 2ull * B * kT, rows1 = 1ull * B * kT;
  float* dec = save ? save->dec : dec_shared;
  float* h_in = save ? save->h_emb : hbuf;
  int rc;

  // ---- embeddings (modeling_seggpt.py:713-737, 163-206) ----
  {
    const long long total = 2LL * B * 3 * 896 * 112;
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * 32, stream);
    patchify_f32_kernel<<<static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 32)), 256, 0, stream>>>(
        pixel_values, prompt_pixel_values, prompt_masks, mlp, B);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
    SgemmEpi ep;
    ep.tab = w.embed_tab[embedding_type == 0 ? 0 : 1];
    ep.rows_per_stream = static_cast<long long>(B) * kT;
    if ((rc = launch_sgemm(SG_EMBED, mlp, 768, w.patch_w, 768, h_in, kD, rows2, kD, 768, ep, stream))) return rc;
  }

  // ---- encoder (modeling_seggpt.py:453-501) ----
  for (int i = 0; i < w.num_layers; ++i) {
    const F32Layer& lw = w.layers[i];
    const int nstreams = (i <= w.merge_index) ? 2 : 1;
    const int nseq = nstreams * B;
    const long long M = static_cast<long long>(nseq) * kT;
    float* h_mid = save ? save->layers[i].h_mid : hbuf;
    float* h_out = save ? save->layers[i].h_out : hbuf;
    float* qkv = save ? save->layers[i].qkv : qkv_shared;
    float* att = save ? save->layers[i].att : att_shared;
    if ((rc = launch_layernorm_f32(h_in, kD, lw.ln1_w, lw.ln1_b, xn, kD, M, w.eps, stream))) return rc;
    {
      SgemmEpi ep;
      ep.bias = lw.qkv_b;
      if ((rc = launch_sgemm(SG_BIAS, xn, kD, lw.qkv_w, kD, qkv, 3072, M, 3072, kD, ep, stream))) return rc;
    }
    if ((rc = launch_attention_f32(qkv, lw.rel_pos_h, lw.rel_pos_w, att, nseq, stream))) return rc;
    bool ens = false;
    if (P > 0) ens = (i == w.merge_index) ? true : (P >= 2);
    if (!ens) {
      SgemmEpi ep;
      ep.bias = lw.proj_b; ep.resid = h_in; ep.ldr = kD;
      if ((rc = launch_sgemm(SG_BIAS_RESID, att, kD, lw.proj_w, kD, h_mid, kD, M, kD, kD, ep, stream))) return rc;
    } else {
      BSEG_REQUIRE(save == nullptr, "feature ensemble is an inference-only path");
      SgemmEpi ep;
      ep.bias = lw.proj_b;
      if ((rc = launch_sgemm(SG_BIAS, att, kD, lw.proj_w, kD, mlp, kD, M, kD, kD, ep, stream))) return rc;
      if ((rc = launch_ensemble_residual(hbuf, mlp, nstreams, B / P, P, i == w.merge_index ? 1 : 0, kT, kD, stream)))
        return rc;
    }
    if ((rc = launch_layernorm_f32(h_mid, kD, lw.ln2_w, lw.ln2_b, xn, kD, M, w.eps, stream))) return rc;
    {
      SgemmEpi ep;
      ep.bias = lw.lin1_b;
      if (save) { ep.aux = save->layers[i].z; ep.ldaux = 4096; }
      if ((rc = launch_sgemm(save ? SG_BIAS_GELU_SAVE : SG_BIAS_GELU, xn, kD, lw.lin1_w, kD, mlp, 4096, M, 4096, kD, ep, stream)))
        return rc;
    }
    {
      SgemmEpi ep;
      ep.bias = lw.lin2_b; ep.resid = h_mid; ep.ldr = kD;
      if ((rc = launch_sgemm(SG_BIAS_RESID, mlp, 4096, lw.lin2_w, 4096, h_out, kD, M, kD, 4096, ep, stream))) return rc;
    }
    if (i == w.merge_index)
      if ((rc = launch_merge_streams(h_out, static_cast<long long>(B) * kT * kD, stream))) return rc;
    for (int j = 0; j < 4; ++j)
      if (w.inter[j] == i)
        if ((rc = launch_layernorm_f32(h_out, kD, w.enc_ln_w, w.enc_ln_b, inter + j * kD, 4 * kD,
                                       static_cast<long long>(B) * kT, w.eps, stream)))
          return rc;
    h_in = h_out;
  }

  // ---- decoder (modeling_seggpt.py:555-585) ----
  {
    SgemmEpi ep;
    ep.bias = w.dec_embed_b;
    if ((rc = launch_sgemm(SG_BIAS, inter, 4 * kD, w.dec_embed_w, 4 * kD, dec, 16384, rows1, 16384, 4 * kD, ep, stream)))
      return rc;
  }
  return launch_decoder_f32(dec, w, pred_masks, B, w.eps, stream);
}

// ---- training workspace of the fp32 mode: forward transients, saved activations, backward scratch ----
struct F32TrainLayout {
  size_t xn, mlp, inter;                                   // forward transients
  size_t h_emb, dec;                                       // saved
  std::vector<size_t> h_mid, h_out, qkv, att, z;           // saved, per layer
  size_t dh, dxn, dz, datt, dqkv, lse, dvec, dinter, ddec, dconv, dpatch;  // backward scratch (B*T rows)
  size_t total;
};
F32TrainLayout f32_train_layout(int num_layers, int merge_index, int B) {
  F32TrainLayout L;
  const size_t rows2 = 2ull * B * kT, rows1 = 1ull * B * kT;
  size_t off = 0;
  auto carve = [&](size_t floats) {
    const size_t o = off;
    off = (off + floats * 4 + 255) / 256 * 256;
    return o;
  };
  L.xn = carve(rows2 * kD);
  L.mlp = carve(rows2 * 4096);
  L.inter = carve(rows1 * 4096);
  L.h_emb = carve(rows2 * kD);
  L.dec = carve(rows1 * 16384);
  for (int i = 0; i < num_layers; ++i) {
    const size_t rows = (i <= merge_index) ? rows2 : rows1;
    L.h_mid.push_back(carve(rows * kD));
    L.h_out.push_back(carve(rows * kD));
    L.qkv.push_back(carve(rows * 3072));
    L.att.push_back(carve(rows * kD));
    L.z.push_back(carve(rows * 4096));
  }
  L.dh = carve(rows1 * kD);
  L.dxn = carve(rows1 * kD);
  L.dz = carve(rows1 * 4096);
  L.datt = carve(rows1 * kD);
  L.dqkv = carve(rows1 * 3072);
  L.lse = carve(rows1 * BSEG_HEADS);
  L.dvec = carve(rows1 * BSEG_HEADS);
  L.dinter = carve(rows1 * 4096);
  L.ddec = carve(rows1 * 16384);
  L.dconv = carve(static_cast<size_t>(B) * 896 * 448 * 64);
  L.dpatch = carve(rows1 * 768);
  L.total = off;
  return L;
}
F32Saved f32_saved(const F32TrainLayout& L, uint8_t* ws) {
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  F32Saved sv;
  sv.h_emb = fp(L.h_emb);
  sv.dec = fp(L.dec);
  for (size_t i = 0; i < L.h_mid.size(); ++i)
    sv.layers.push_back({fp(L.h_mid[i]), fp(L.h_out[i]), fp(L.qkv[i]), fp(L.att[i]), fp(L.z[i])});
  return sv;
}
}  // namespace

size_t f32_train_workspace_bytes(const F32Weights& w, int B) { return f32_train_layout(w.num_layers, w.merge_index, B).total; }

int forward_f32_train_impl(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                           const float* prompt_masks, int B, int embedding_type, void* workspace, float* pred_masks,
                           cudaStream_t stream) {
  const F32TrainLayout L = f32_train_layout(w.num_layers, w.merge_index, B);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const F32Saved sv = f32_saved(L, ws);
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  return forward_f32_core(w, pixel_values, prompt_pixel_values, prompt_masks, B, embedding_type, 0, nullptr, fp(L.xn), nullptr,
                          nullptr, fp(L.mlp), fp(L.inter), nullptr, &sv, pred_masks, stream);
}

// Gradient of sum(pred_masks * d_pred_masks) w.r.t. prompt_pixel_values, fp32 throughout.  Mirrors
// bseg_backward_to_prompt (api.cu) step by step; d_pred_masks must be zero for image rows < 448.
int backward_f32_impl(const F32Weights& w, const float* d_pred_masks, int B, void* workspace, float* d_prompt,
                      cudaStream_t stream) {
  const F32TrainLayout L = f32_train_layout(w.num_layers, w.merge_index, B);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const F32Saved sv = f32_saved(L, ws);
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  float *dh = fp(L.dh), *dxn = fp(L.dxn), *dz = fp(L.dz), *datt = fp(L.datt), *dqkv = fp(L.dqkv), *lse = fp(L.lse),
        *dvec = fp(L.dvec), *dinter = fp(L.dinter), *ddec = fp(L.ddec), *dconv = fp(L.dconv), *dpatch = fp(L.dpatch);
  const long long M = static_cast<long long>(B) * kT;
  int rc;
  // ---- decoder head and conv3x3 dgrad (modeling_seggpt.py:546-552); gradient exists for image rows >= 448 only, so
  // the conv's input gradient is non-zero from row 447: rows 432.. (token row 27) are computed, the rest is zero ----
  BSEG_CHECK_CUDA(cudaMemsetAsync(dconv, 0, static_cast<size_t>(B) * 896 * 448 * 64 * 4, stream));
  BSEG_CHECK_CUDA(cudaMemsetAsync(ddec, 0, static_cast<size_t>(M) * 16384 * 4, stream));
  {
    ProfScope prof(CAT_DECODER_HEAD, 0, 0, stream);
    decoder_head_bwd_f32_kernel<<<dim3(448 / 16, 448 / 16, B), 256, 0, stream>>>(
        sv.dec, w.dec_conv_w, w.dec_conv_b, w.dec_ln_w, w.dec_ln_b, w.dec_head_w, d_pred_masks, dconv, w.eps);
    BSEG_CHECK_CUDA(cudaGetLastError());
    decoder_conv_dgrad_f32_kernel<<<dim3(448 / 16, (896 - 432) / 16, B), 256, 0, stream>>>(dconv, w.dec_conv_w, ddec, 432);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
  }
  {
    SgemmEpi ep;  // decoder_embed dgrad: [M,16384] x W[16384,4096]
    if ((rc = launch_sgemm_nn(SG_PLAIN, ddec, 16384, w.dec_embed_w, 4 * kD, dinter, 4 * kD, M, 4 * kD, 16384, ep, stream)))
      return rc;
  }
  BSEG_CHECK_CUDA(cudaMemsetAsync(dh, 0, static_cast<size_t>(M) * kD * 4, stream));
  // ---- encoder, last layer first (image stream only: rows [0, B*T) of every saved buffer) ----
  for (int i = w.num_layers - 1; i >= 0; --i) {
    const F32Layer& lw = w.layers[i];
    const F32Saved::Layer& sl = sv.layers[i];
    for (int j = 0; j < 4; ++j)
      if (w.inter[j] == i)
        if ((rc = launch_layernorm_bwd_f32(sl.h_out, kD, dinter + j * kD, 4 * kD, w.enc_ln_w, dh, M, w.eps, stream)))
          return rc;
    if (i == w.merge_index) {  // merge (modeling_seggpt.py:476-479): the image stream receives half of the gradient
      scale_f32_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<float4*>(dh), 0.5f, M * kD / 4);
      BSEG_CHECK_CUDA(cudaGetLastError());
      count_launch();
    }
    {
      SgemmEpi ep;  // lin2 dgrad through the GELU: dz = (dh W2) * gelu'(z)
      ep.aux = sl.z; ep.ldaux = 4096;
      if ((rc = launch_sgemm_nn(SG_DGELU, dh, kD, lw.lin2_w, 4096, dz, 4096, M, 4096, kD, ep, stream))) return rc;
    }
    {
      SgemmEpi ep;  // lin1 dgrad
      if ((rc = launch_sgemm_nn(SG_PLAIN, dz, 4096, lw.lin1_w, kD, dxn, kD, M, kD, 4096, ep, stream))) return rc;
    }
    if ((rc = launch_layernorm_bwd_f32(sl.h_mid, kD, dxn, kD, lw.ln2_w, dh, M, w.eps, stream))) return rc;
    {
      SgemmEpi ep;  // proj dgrad
      if ((rc = launch_sgemm_nn(SG_PLAIN, dh, kD, lw.proj_w, kD, datt, kD, M, kD, kD, ep, stream))) return rc;
    }
    if ((rc = launch_attention_bwd_f32(sl.qkv, sl.att, datt, lw.rel_pos_h, lw.rel_pos_w, lse, dvec, dqkv, B, stream)))
      return rc;
    {
      SgemmEpi ep;  // qkv dgrad
      if ((rc = launch_sgemm_nn(SG_PLAIN, dqkv, 3072, lw.qkv_w, kD, dxn, kD, M, kD, 3072, ep, stream))) return rc;
    }
    const float* h_in = (i == 0) ? sv.h_emb : sv.layers[i - 1].h_out;
    if ((rc = launch_layernorm_bwd_f32(h_in, kD, dxn, kD, lw.ln1_w, dh, M, w.eps, stream))) return rc;
  }
  {
    SgemmEpi ep;  // patch-embedding dgrad (Conv2d k16 s16 as a GEMM over im2col rows), then un-patchify the prompt half
    if ((rc = launch_sgemm_nn(SG_PLAIN, dh, kD, w.patch_w, 768, dpatch, 768, M, 768, kD, ep, stream))) return rc;
  }
  return launch_unpatchify_prompt_grad(dpatch, d_prompt, B, stream);
}

}  // namespace bseg
