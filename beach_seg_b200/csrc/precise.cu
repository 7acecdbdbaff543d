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
// SGEMM, both operands K-major: C[m][n] = sum_k A[m][k] * W[n][k]
// ----------------------------------------------------------------------------------------------
enum : int { SG_BIAS = 0, SG_BIAS_GELU = 1, SG_BIAS_RESID = 2, SG_EMBED = 3 };
struct SgemmEpi {
  const float* bias = nullptr;
  const float* resid = nullptr;  // [M, ldr]
  long long ldr = 0;
  const float* tab = nullptr;    // [2][T][N] additive table, stream = row / rows_per_stream
  long long rows_per_stream = 0;
};

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;

template <int EPI>
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
      rw[h] = *reinterpret_cast<const float4*>(W + static_cast<long long>(n0 + lrow + 64 * h) * ldw + kt * SG_BK + lk);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
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
      } else {
        const float4 bb = *reinterpret_cast<const float4*>(ep.bias + n);
        v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
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
    case SG_BIAS: sgemm_nt_kernel<SG_BIAS><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_BIAS_GELU: sgemm_nt_kernel<SG_BIAS_GELU><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_BIAS_RESID: sgemm_nt_kernel<SG_BIAS_RESID><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    case SG_EMBED: sgemm_nt_kernel<SG_EMBED><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep); break;
    default: BSEG_REQUIRE(false, "sgemm: unknown epilogue %d", epi);
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
    if ((rc = launch_sgemm(SG_EMBED, mlp, 768, w.patch_w, 768, hbuf, kD, rows2, kD, 768, ep, stream))) return rc;
  }

  // ---- encoder (modeling_seggpt.py:453-501) ----
  for (int i = 0; i < w.num_layers; ++i) {
    const F32Layer& lw = w.layers[i];
    const int nstreams = (i <= w.merge_index) ? 2 : 1;
    const int nseq = nstreams * B;
    const long long M = static_cast<long long>(nseq) * kT;
    if ((rc = launch_layernorm_f32(hbuf, kD, lw.ln1_w, lw.ln1_b, xn, kD, M, w.eps, stream))) return rc;
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
      ep.bias = lw.proj_b; ep.resid = hbuf; ep.ldr = kD;
      if ((rc = launch_sgemm(SG_BIAS_RESID, att, kD, lw.proj_w, kD, hbuf, kD, M, kD, kD, ep, stream))) return rc;
    } else {
      SgemmEpi ep;
      ep.bias = lw.proj_b;
      if ((rc = launch_sgemm(SG_BIAS, att, kD, lw.proj_w, kD, mlp, kD, M, kD, kD, ep, stream))) return rc;
      if ((rc = launch_ensemble_residual(hbuf, mlp, nstreams, B / P, P, i == w.merge_index ? 1 : 0, kT, kD, stream)))
        return rc;
    }
    if ((rc = launch_layernorm_f32(hbuf, kD, lw.ln2_w, lw.ln2_b, xn, kD, M, w.eps, stream))) return rc;
    {
      SgemmEpi ep;
      ep.bias = lw.lin1_b;
      if ((rc = launch_sgemm(SG_BIAS_GELU, xn, kD, lw.lin1_w, kD, mlp, 4096, M, 4096, kD, ep, stream))) return rc;
    }
    {
      SgemmEpi ep;
      ep.bias = lw.lin2_b; ep.resid = hbuf; ep.ldr = kD;
      if ((rc = launch_sgemm(SG_BIAS_RESID, mlp, 4096, lw.lin2_w, 4096, hbuf, kD, M, kD, 4096, ep, stream))) return rc;
    }
    if (i == w.merge_index)
      if ((rc = launch_merge_streams(hbuf, static_cast<long long>(B) * kT * kD, stream))) return rc;
    for (int j = 0; j < 4; ++j)
      if (w.inter[j] == i)
        if ((rc = launch_layernorm_f32(hbuf, kD, w.enc_ln_w, w.enc_ln_b, inter + j * kD, 4 * kD,
                                       static_cast<long long>(B) * kT, w.eps, stream)))
          return rc;
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

}  // namespace bseg
