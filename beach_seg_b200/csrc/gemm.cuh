// Persistent, warp-specialised tcgen05 GEMM:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//   A : bf16 row-major activations (K contiguous)          -> TMA, 128B swizzle
//   W : bf16 row-major nn.Linear weight [out=N, in=K]      -> TMA, 128B swizzle (already K-major)
//   D : fp32 accumulators in TMEM (double buffered), drained by 4 epilogue warps.
// Replaces the reference's CPU addmm calls behind every nn.Linear of HF SegGPT
// (transformers/models/seggpt/modeling_seggpt.py:224,225,355,356,558) and the patch-embed conv (:108).
#pragma once
#include "common.cuh"

namespace bseg {

enum GemmEpiMode : int {
  EPI_BF16 = 0,       // out_bf16[m, n]  = acc + bias[n]
  EPI_BF16_GELU = 1,  // out_bf16[m, n]  = gelu_erf(acc + bias[n])
  EPI_F32 = 2,        // out_f32[m, n]   = acc + bias[n]
  EPI_RESID_F32 = 3,  // out_f32[m, n]   = resid[m, n] + acc + bias[n]      (out may alias resid)
  EPI_QKV = 4,        // scatter to q[seq,h,t,64], k[seq,h,t,64], vT[seq,h,64,T]  (+bias)
  EPI_EMBED = 5,      // out_f32[m, n]   = acc + tab[(m / rows_per_stream) * T + m % T, n]
  EPI_PIXSHUF = 6,    // decoder pixel shuffle: nhwc[b, ph*16+py, pw*16+px, c] = acc + bias[n], n = (py*16+px)*64+c
  EPI_DGELU = 7,      // backward of lin1's GELU: out_bf16[m, n] = acc * gelu'(aux_bf16[m, n])   (aux = saved pre-activation)
  // EPI_RESID_F32 + the LayerNorm that reads the updated residual stream next (N == 1024 == the whole row, BLOCK_N = 256):
  //   out_f32[m, :] = resid[m, :] + acc + bias ;  ln_out_bf16[m, :] = LN(out_f32[m, :]) * gamma + beta
  // The fp32 row is normalised while it is still in L2 instead of being read back from HBM by layernorm1024_kernel.
  EPI_RESID_LN = 8,
};

struct GemmEpiParams {
  void* out = nullptr;  // bf16* or float*
  long long ldc = 0;
  const float* bias = nullptr;
  const float* resid = nullptr;
  long long ldr = 0;
  // EPI_BF16_GELU: optional bf16 copy of the pre-activation (saved for the backward pass); EPI_DGELU: that copy (input)
  __nv_bfloat16* aux = nullptr;
  // EPI_EMBED
  const float* tab = nullptr;
  int rows_per_stream = 0;
  // EPI_QKV / EPI_EMBED / EPI_PIXSHUF : tokens per sequence and token grid width
  int T = 1568;
  int grid_w = 28;
  // EPI_QKV
  __nv_bfloat16* q = nullptr;
  __nv_bfloat16* k = nullptr;
  __nv_bfloat16* vt = nullptr;
  int heads = 16;
  float q_scale = 1.0f;  // q is stored as bf16((acc + bias) * q_scale): the attention kernels take q * scale * log2(e)
  // EPI_RESID_LN: LayerNorm parameters / output, and the cross-CTA exchange of the row statistics
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  __nv_bfloat16* ln_out = nullptr;
  long long ld_ln = 0;
  float ln_eps = 1e-6f;
  // [M][8] 64-bit entries {mean, M2} of the eight 128-column slices of a row; the low 4 mantissa bits of each value hold
  // a nibble of the launch tag.  Consecutive launches on the same buffer must use different tags (0..254); 0xFF bytes =
  // never written.
  unsigned long long* ln_stats = nullptr;
  unsigned int ln_tag = 0;
};

// Which rows of A (== rows of the output) a launch covers: `nbatch` entries of `rows_per_batch` rows each, of which
// rows [row_begin, row_begin + rows) are computed.  A plain GEMM is {M, 1, 0, M}; the decoder's query-half slice is
// {1568, B, 756, 812}.
struct GemmRows {
  long long rows_per_batch;
  int nbatch;
  int row_begin;
  int rows;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each draining half of the tile's columns
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4.. epilogue

// kCtas = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BLOCK_N tile with one M = 256 MMA; each CTA stages
// its 128 rows of A and BLOCK_N / 2 rows of W, so a stage is 32 KB per CTA instead of 48 KB (6 stages instead of 4) and
// every W tile is fetched from L2 once per 256 output rows instead of once per 128.
template <int BLOCK_N, int kCtas = 1>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N == 256 && kCtas == 1) ? 4 : 6;
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBRows = BLOCK_N / kCtas;  // rows of W this CTA stages
  static constexpr int kBBytes = kBRows * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiStageBytes = GEMM_EPI_WARPS * 32 * 32 * 4;  // one 32x32 fp32 transposing tile per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiStageBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // 512 or 256
};

// ---- fp32-output epilogues (residual stream / embedding table / plain) ---------------------------------------
// The accumulator arrives with thread <-> row (TMEM lane).  A row-per-thread global access touches 32 different
// 128-byte lines per warp instruction, which makes the K=1024 GEMMs epilogue-bound on L1 wavefronts; so a 32x32
// chunk is transposed through a per-warp swizzled smem tile and rows are then read / written by 8 adjacent lanes
// (128 contiguous bytes per row, 4 rows per instruction).
struct EpiRows {
  float4 r[8];  // addend of rows (i*4 + lane/8), columns 4*(lane%8)..+3
};
template <int MODE>
__device__ __forceinline__ void gemm_epi_f32_prefetch(const GemmEpiParams& ep, EpiRows& pr, long long row0, int nvalid,
                                                      int n, int lane) {
  if constexpr (MODE == EPI_RESID_F32 || MODE == EPI_EMBED) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = row0 + i * 4 + (lane >> 3);
      if (i * 4 + (lane >> 3) < nvalid) {
        const float* add;
        if constexpr (MODE == EPI_RESID_F32) {
          add = ep.resid + m * ep.ldr;
        } else {
          add = ep.tab + ((m / ep.rows_per_stream) * ep.T + m % ep.T) * ep.ldc;
        }
        pr.r[i] = *reinterpret_cast<const float4*>(add + n + 4 * (lane & 7));
      }
    }
  }
}
template <int MODE>
__device__ __forceinline__ void gemm_epi_f32_chunk(const GemmEpiParams& ep, const float (&v)[32], const EpiRows& pr,
                                                   float* stg, long long row0, int nvalid, int n, int lane) {
  // 1) accumulator row of this thread -> smem (16-byte chunk index XOR row&7: conflict free both ways)
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  // 2) coalesced read-back, add, store
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (MODE != EPI_EMBED) {
    if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4 * (lane & 7)));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    const long long m = row0 + rr;
    float4 o = *reinterpret_cast<const float4*>(stg + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
    o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
    if constexpr (MODE != EPI_F32) {
      o.x += pr.r[i].x; o.y += pr.r[i].y; o.z += pr.r[i].z; o.w += pr.r[i].w;
    }
    if (rr < nvalid) *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + m * ep.ldc + n + 4 * (lane & 7)) = o;
  }
  __syncwarp();
}

// ---- EPI_RESID_LN: residual epilogue + LayerNorm of the finished row ---------------------------------------------
// A 256-column tile sees a quarter of a row, so the row statistics are exchanged between the four CTAs (pairs) that own
// the N tiles of the same rows, through L2: every epilogue warp publishes (mean, M2) of its 32 rows x 128 columns as ONE
// 64-bit relaxed store per row which carries the launch's 8-bit tag in its lowest mantissa bits, and a reader polls the eight entries of a row
// until all carry the tag.  A 64-bit scalar access is single-copy atomic, so no fence, no counter and no ordering with
// any other access is needed (a first version with arrival counters + release/acquire ran 3x slower: each fence waits
// for the warp's ~100 outstanding residual loads and stores).  The peers are resident by construction (persistent kernel,
// at most one wave of CTAs), and a warp publishes its tile BEFORE it waits for anybody, so there is no cyclic wait.
// The warp then normalises ITS OWN 32 x 128 slice, which it wrote a tile ago and reads back from L2 with the same
// thread <-> element mapping.
// Statistics are carried as (mean, M2 = sum of squared deviations) and merged pairwise between equal counts (Chan et
// al.): no E[x^2] - mean^2 cancellation, and the merge is bitwise symmetric, so every warp that merges the same eight
// partials in the same tree gets the same bits (one row is normalised consistently by four different CTAs).
__device__ __forceinline__ void ln_merge_equal(float& mean, float& m2, float mean_b, float m2_b, float half_count) {
  const float d = mean_b - mean;
  m2 = __fmaf_rn(d * d, half_count, m2 + m2_b);
  mean = 0.5f * (mean + mean_b);
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// pass 1 for one 32 x 32 chunk: gemm_epi_f32_chunk<EPI_RESID_F32> + running statistics of rows (i*4 + lane/8) over
// this thread's four columns per chunk.  `ci` = number of chunks already merged (the running side holds 4*ci values).
// All three row-major operands of this mode (residual in, stream out, LayerNorm out) have leading dimension 1024 (checked
// by the launcher), so every access is `thread base + compile-time offset`.
constexpr int kLnLd = 1024;
// -DBSEG_LN_TRACE: per-phase clock64 sums of the epilogue loop of a few warps, printed at kernel end (tools/ln_trace.py)
#ifdef BSEG_LN_TRACE
#define LN_T(k) do { const long long t_ = clock64(); ln_t[k] += t_ - ln_last; ln_last = t_; } while (0)
#else
#define LN_T(k) do { } while (0)
#endif
__device__ __forceinline__ void gemm_epi_resid_ln_prefetch(const float* resid_thread /*row0 + lane/8, n + 4*(lane%8)*/,
                                                           EpiRows& pr, int nvalid, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 4 + (lane >> 3) < nvalid) pr.r[i] = *reinterpret_cast<const float4*>(resid_thread + i * 4 * kLnLd);
}
__device__ __forceinline__ void sts128_f32(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128_f32(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
// stg_w / stg_r: shared-window byte addresses of this thread's row in the per-warp 32 x 32 transposing tile
// (stg + lane * 128) and of its read-back position (stg + (lane / 8) * 128); the 16-byte chunk index is XORed with row % 8.
__device__ __forceinline__ void gemm_epi_resid_ln_chunk(const float* bias_thread /*bias + n + 4*(lane%8), or null*/,
                                                        float* out_thread /*row0 + lane/8, n + 4*(lane%8)*/,
                                                        const float (&v)[32], const EpiRows& pr, uint32_t stg_w,
                                                        uint32_t stg_r, int nvalid, int lane, int ci, float (&rmean)[8],
                                                        float (&rm2)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts128_f32(stg_w + (((j ^ lane) & 7) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias_thread != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias_thread));
  // merging 4 new values into 4*ci: mean += delta / (ci + 1), M2 += M2_new + delta^2 * 4 ci / (ci + 1)
  const float w_mean = ci == 0 ? 1.0f : (ci == 1 ? 0.5f : (ci == 2 ? (1.0f / 3.0f) : 0.25f));
  const float w_m2 = ci == 0 ? 0.0f : (ci == 1 ? 2.0f : (ci == 2 ? (8.0f / 3.0f) : 3.0f));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // row rr = i * 4 + lane / 8: rr % 8 = (i % 2) * 4 + lane / 8 (lane / 8 < 4)
    const int rr = i * 4 + (lane >> 3);
    float4 o = lds128_f32(stg_r + i * 4 * 128 + ((((lane & 7) ^ rr) & 7) << 4));
    o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
    o.x += pr.r[i].x; o.y += pr.r[i].y; o.z += pr.r[i].z; o.w += pr.r[i].w;
    if (rr < nvalid) *reinterpret_cast<float4*>(out_thread + i * 4 * kLnLd) = o;
    const float m4 = 0.25f * ((o.x + o.y) + (o.z + o.w));
    const float da = o.x - m4, db = o.y - m4, dc = o.z - m4, dd = o.w - m4;
    const float q4 = (da * da + db * db) + (dc * dc + dd * dd);
    const float delta = m4 - rmean[i];
    rmean[i] = __fmaf_rn(delta, w_mean, rmean[i]);
    rm2[i] += __fmaf_rn(delta * delta, w_m2, q4);
  }
  __syncwarp();
}
// After the last chunk of pass 1: merge the eight lanes that share a row (16 values each -> 128) and publish the warp's
// 32 x (mean, M2 | tag).  `slot` = 2 * (N tile) + (column half of the warp).
__device__ __forceinline__ void gemm_epi_resid_ln_publish(const GemmEpiParams& ep, float (&rmean)[8], float (&rm2)[8],
                                                          long long row0, int nvalid, int slot, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float mb = __shfl_xor_sync(0xffffffffu, rmean[i], o);
      const float qb = __shfl_xor_sync(0xffffffffu, rm2[i], o);
      ln_merge_equal(rmean[i], rm2[i], mb, qb, 8.0f * o);
    }
  }
  if ((lane & 7) == 0) {  // every lane of a row's group holds the same bits; the first one stores them
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + (lane >> 3);
      // the 8-bit tag replaces the low 4 mantissa bits of each value (relative error 2^-20: ~1e-3 of a bf16 ulp)
      const unsigned long long e =
          (static_cast<unsigned long long>((__float_as_uint(rm2[i]) & 0xfffffff0u) | (ep.ln_tag & 15u)) << 32) |
          ((__float_as_uint(rmean[i]) & 0xfffffff0u) | (ep.ln_tag >> 4));
      if (rr < nvalid) st_relaxed_gpu_u64(ep.ln_stats + (row0 + rr) * 8 + slot, e);
    }
  }
}
// The LayerNorm of a tile happens one tile later (the peers published long ago, so the entries normally carry the tag
// on the first look).  _issue starts the loads of the first half (64 columns) of the warp's own 32 x 128 slice of the
// fp32 stream, which the same thread wrote a tile ago and which now sits in L2; _finish, called after the next tile's
// pass 1, fetches and merges the eight statistics entries per row (128 -> 1024 columns), normalises the first half,
// fetches and normalises the second half (register budget: the next tile's residual prefetch is in flight as well).  8 lanes x 8 B = 64 contiguous bytes of bf16 per row.
struct LnPending {
  float4 x[16];  // columns 32 j + 4 (lane % 8) .. + 3 of rows i * 4 + lane / 8, j = 0, 1
};
__device__ __forceinline__ void gemm_epi_resid_ln_load_half(const float* xin, float4 (&x)[16], int nvalid, int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[j * 8 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i * 4 + (lane >> 3) < nvalid) x[j * 8 + i] = *reinterpret_cast<const float4*>(xin + i * 4 * kLnLd + 32 * j);
    }
}
__device__ __forceinline__ void gemm_epi_resid_ln_load_stats(const unsigned long long* mine, unsigned long long (&e)[8],
                                                             unsigned int tag, int nvalid, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    e[i] = (static_cast<unsigned long long>(tag & 15u) << 32) | (tag >> 4);  // rows past the range: "ready", zeros
    if (i * 4 + (lane >> 3) < nvalid) e[i] = ld_relaxed_gpu_u64(mine + i * 4 * 8);
  }
}
__device__ __forceinline__ void gemm_epi_resid_ln_issue(const GemmEpiParams& ep, LnPending& p, long long row0, int nvalid,
                                                        int ncol0, int lane) {
  const size_t toff = static_cast<size_t>(row0 + (lane >> 3)) * kLnLd + ncol0 + 4 * (lane & 7);
  gemm_epi_resid_ln_load_half(reinterpret_cast<const float*>(ep.out) + toff, p.x, nvalid, lane);
}
__device__ __forceinline__ void gemm_epi_resid_ln_store_half(const float4 (&x)[16], const float (&mean)[8],
                                                             const float (&rstd)[8], const float* gam, const float* bet,
                                                             __nv_bfloat16* yout, int nvalid, int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gam + 32 * j));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bet + 32 * j));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 xv = x[j * 8 + i];
      const uint2 o = make_uint2(pack_bf16x2((xv.x - mean[i]) * rstd[i] * g.x + b.x, (xv.y - mean[i]) * rstd[i] * g.y + b.y),
                                 pack_bf16x2((xv.z - mean[i]) * rstd[i] * g.z + b.z, (xv.w - mean[i]) * rstd[i] * g.w + b.w));
      if (i * 4 + (lane >> 3) < nvalid) *reinterpret_cast<uint2*>(yout + i * 4 * kLnLd + 32 * j) = o;
    }
  }
}
__device__ __forceinline__ void gemm_epi_resid_ln_finish(const GemmEpiParams& ep, LnPending& p, long long row0, int nvalid,
                                                         int ncol0, int lane) {
  float mean[8], rstd[8];
  {
    const unsigned long long* mine = ep.ln_stats + (row0 + (lane >> 3)) * 8 + (lane & 7);
    unsigned long long e[8];  // statistics entry (slot lane % 8) of rows i * 4 + lane / 8
    long long t0 = 0;
    uint32_t spins = 0;
    for (;;) {
      gemm_epi_resid_ln_load_stats(mine, e, ep.ln_tag, nvalid, lane);
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        ok = ok && ((((static_cast<uint32_t>(e[i]) & 15u) << 4) | (static_cast<uint32_t>(e[i] >> 32) & 15u)) == ep.ln_tag);
      if (__all_sync(0xffffffffu, ok)) break;
      __nanosleep(64);
      if ((++spins & 63u) == 0u) {  // same kind of watchdog as mbar_wait: a protocol bug traps instead of hanging the box
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        if (now - t0 > 4000000000LL) {
          if (lane == 0)
            printf("bseg: LayerNorm statistics watchdog block=%d thread=%d row0=%lld tag=%u\n", blockIdx.x, threadIdx.x,
                   row0, ep.ln_tag);
          __trap();
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = __uint_as_float(static_cast<uint32_t>(e[i]) & 0xfffffff0u);
      rstd[i] = __uint_as_float(static_cast<uint32_t>(e[i] >> 32) & 0xfffffff0u);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float mb = __shfl_xor_sync(0xffffffffu, mean[i], o);
      const float qb = __shfl_xor_sync(0xffffffffu, rstd[i], o);
      ln_merge_equal(mean[i], rstd[i], mb, qb, 64.0f * o);
    }
    rstd[i] = rsqrtf(rstd[i] * (1.0f / 1024.0f) + ep.ln_eps);
  }
  const size_t toff = static_cast<size_t>(row0 + (lane >> 3)) * kLnLd + ncol0 + 4 * (lane & 7);
  const float* xin = reinterpret_cast<const float*>(ep.out) + toff;
  __nv_bfloat16* yout = ep.ln_out + toff;
  const float* gam = ep.ln_gamma + ncol0 + 4 * (lane & 7);
  const float* bet = ep.ln_beta + ncol0 + 4 * (lane & 7);
  gemm_epi_resid_ln_store_half(p.x, mean, rstd, gam, bet, yout, nvalid, lane);
  asm volatile("" ::: "memory");  // (the second half reuses the first half's registers: do not hoist its loads)
  gemm_epi_resid_ln_load_half(xin + 64, p.x, nvalid, lane);
  gemm_epi_resid_ln_store_half(p.x, mean, rstd, gam + 64, bet + 64, yout + 64, nvalid, lane);
}

// ---- bf16-output epilogues with a row-major destination (bias / GELU / GELU') ---------------------------------------
// Same problem as above: thread <-> row means one 16-byte store per lane hits 32 different lines per instruction, and
// for the K = 1024 GEMMs (8192 MMA cycles per tile) those L1 wavefronts do not hide behind the MMA.  The 32x32 chunk is
// packed to bf16, transposed through the per-warp swizzled smem tile and written 8 rows x 64 contiguous bytes per
// instruction (4 lanes per row).
__device__ __forceinline__ void gemm_epi_bf16_store(__nv_bfloat16* out, long long ldc, const float (&v)[32],
                                                    uint32_t* stg, long long row0, int nvalid, int n, int lane) {
  // 1) this thread's row (32 bf16 = 4 x 16 B) -> smem, 16-byte chunk index XOR ((row >> 1) & 3)
  const int swz = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(stg + lane * 16 + ((j ^ swz) << 2)) =
        make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                   pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  __syncwarp();
  // 2) read back 8 rows per instruction, 4 lanes per row, and store
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = i * 8 + (lane >> 2);
    const uint4 o = *reinterpret_cast<const uint4*>(stg + rr * 16 + ((c ^ ((rr >> 1) & 3)) << 2));
    if (rr < nvalid) *reinterpret_cast<uint4*>(out + (row0 + rr) * ldc + n + c * 8) = o;
  }
  __syncwarp();
}
template <int MODE>
__device__ __forceinline__ void gemm_epi_bf16_chunk(const GemmEpiParams& ep, float (&v)[32], uint32_t* stg,
                                                    long long row0, int nvalid, int n, int lane) {
  const long long m = row0 + lane;
  if (ep.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + i));
      add_f32x2(v[i], v[i + 1], b.x, b.y);
      add_f32x2(v[i + 2], v[i + 3], b.z, b.w);
    }
  }
  if constexpr (MODE == EPI_BF16_GELU) {
    if (ep.aux != nullptr) gemm_epi_bf16_store(ep.aux, ep.ldc, v, stg, row0, nvalid, n, lane);  // saved pre-activation
#pragma unroll
    for (int i = 0; i < 32; i += 2) gelu_erf_x2(v[i], v[i + 1]);
  }
  if constexpr (MODE == EPI_DGELU) {
    if (lane < nvalid) {
      const __nv_bfloat16* z = ep.aux + m * ep.ldc + n;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 pk = *reinterpret_cast<const uint4*>(z + i);
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g0, g1;
          gelu_erf_grad_x2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u), g0, g1);
          mul_f32x2(v[i + 2 * j], v[i + 2 * j + 1], v[i + 2 * j], v[i + 2 * j + 1], g0, g1);
        }
      }
    }
  }
  gemm_epi_bf16_store(reinterpret_cast<__nv_bfloat16*>(ep.out), ep.ldc, v, stg, row0, nvalid, n, lane);
}

template <int MODE>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmEpiParams& ep, float (&v)[32], long long m, int n) {
  // v: 32 consecutive accumulator columns n..n+31 of row m
  if constexpr (MODE != EPI_EMBED) {
    if (ep.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    }
  }
  if constexpr (MODE == EPI_BF16 || MODE == EPI_BF16_GELU || MODE == EPI_DGELU) {
    if constexpr (MODE == EPI_BF16_GELU) {
      if (ep.aux != nullptr) {
        __nv_bfloat16* z = ep.aux + m * ep.ldc + n;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
          *reinterpret_cast<uint4*>(z + i) = pk;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf_x2(v[i], v[i + 1]);
    }
    if constexpr (MODE == EPI_DGELU) {
      const __nv_bfloat16* z = ep.aux + m * ep.ldc + n;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 pk = *reinterpret_cast<const uint4*>(z + i);
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float g0, g1;
          gelu_erf_grad_x2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u), g0, g1);
          mul_f32x2(v[i + 2 * j], v[i + 2 * j + 1], v[i + 2 * j], v[i + 2 * j + 1], g0, g1);
        }
      }
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.out) + m * ep.ldc + n;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 pk = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                            pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
      *reinterpret_cast<uint4*>(dst + i) = pk;
    }
  } else if constexpr (MODE == EPI_QKV) {
    const int D = ep.heads * 64;
    const int which = n / D;
    const int head = (n % D) >> 6;
    const int d0 = n & 63;
    const long long seq = m / ep.T;
    const long long t = m % ep.T;
    if (which < 2) {
      __nv_bfloat16* base = (which == 0) ? ep.q : ep.k;
      __nv_bfloat16* dst = base + ((seq * ep.heads + head) * ep.T + t) * 64 + d0;
      if (which == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= ep.q_scale;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                              pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
    } else {
      // V is stored transposed ([d, t]) so that P*V consumes it as a K-major B operand
      __nv_bfloat16* dst = ep.vt + ((seq * ep.heads + head) * 64 + d0) * (long long)ep.T + t;
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[(long long)i * ep.T] = __float2bfloat16_rn(v[i]);
    }
  } else if constexpr (MODE == EPI_PIXSHUF) {
    const long long b = m / ep.T;
    const int t = static_cast<int>(m % ep.T);
    const int ph = t / ep.grid_w, pw = t % ep.grid_w;
    const int py = n >> 10, px = (n >> 6) & 15, c0 = n & 63;
    const int H = (ep.T / ep.grid_w) * 16, W = ep.grid_w * 16;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.out) +
                         ((b * H + (ph * 16 + py)) * W + (pw * 16 + px)) * 64 + c0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 pk = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                            pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
      *reinterpret_cast<uint4*>(dst + i) = pk;
    }
  }
}

template <int BLOCK_N, int MODE, int kCtas = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         GemmRows gr, int N, int K, GemmEpiParams ep) {
  using Cfg = GemmCfg<BLOCK_N, kCtas>;
  static_assert(kCtas == 1 || kCtas == 2, "one CTA or a CTA pair");
  constexpr int kTileM = GEMM_BLOCK_M * kCtas;  // output rows per (pair) tile
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // launched as clusters of kCtas CTAs: rank 0 is the leader (MMA issuer, owner of the full / tmem_empty barriers)
  const uint32_t cta_rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const long long tile_first = blockIdx.x / kCtas;
  const long long tile_step = gridDim.x / kCtas;
  const int num_n_tiles = N / BLOCK_N;
  const int tiles_per_batch = (gr.rows + kTileM - 1) / kTileM;
  const long long num_m_tiles = static_cast<long long>(gr.nbatch) * tiles_per_batch;
  const long long num_tiles = num_m_tiles * num_n_tiles;
  const int num_k_blocks = K / GEMM_BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], GEMM_EPI_WARPS * kCtas);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCtas == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  // everything above (barriers, TMEM, tensor-map prefetch) may run under the tail of the previous kernel in the stream
  pdl_wait();

  // EPI_RESID_LN: the epilogue warps carry the residual prefetch, the LayerNorm loads and the row statistics, so the control
  // warpgroup gives its registers away (128 x 56 + 256 x 224 = 384 x 168, the CTA's pool).  ptxas sizes each side of
  // this branch by the setmaxnreg it starts with, hence the nesting.
  if (warp < 4) {
  if constexpr (MODE == EPI_RESID_LN) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
      const long long mt = tile / num_n_tiles;
      const int bidx = static_cast<int>(mt / tiles_per_batch);
      const int m0 = gr.row_begin + static_cast<int>(mt % tiles_per_batch) * kTileM + cta_rank * GEMM_BLOCK_M;
      const int n0 = static_cast<int>(tile % num_n_tiles) * BLOCK_N + cta_rank * Cfg::kBRows;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if constexpr (kCtas == 2) {
            // both CTAs' bytes are credited to the leader's barrier, which expects the pair's total
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_3d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, m0, bidx);
            tma_load_2d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_3d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, m0, bidx);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom
            const uint64_t da = umma_desc_sw128_kmajor(a_addr + k * 32);
            const uint64_t db = umma_desc_sw128_kmajor(b_addr + k * 32);
            if constexpr (kCtas == 2) umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (kCtas == 2) {
            umma_commit_2sm(&empty_bar[stage]);  // frees the smem slot in both CTAs when these MMAs retire
            if (kb == num_k_blocks - 1) umma_commit_2sm(&tmem_full[as]);
          } else {
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (kb == num_k_blocks - 1) umma_commit(&tmem_full[as]);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    if constexpr (MODE == EPI_RESID_LN) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int kColsPerWarp = BLOCK_N / (GEMM_EPI_WARPS / 4);
    const int col0 = ((warp - 4) >> 2) * kColsPerWarp;
    int as = 0;
    uint32_t aphase = 0;
    if constexpr (MODE == EPI_RESID_LN) {
      // Software-pipelined over the warp's tiles, so that every L2 / HBM round trip of the epilogue overlaps other work:
      //   top of tile i : issue the loads for the LayerNorm of tile i-1 (its eight statistics entries per row and the
      //                   first half of the warp's own slice of the fp32 stream, written a tile ago, now in L2)
      //   pass 1 of i   : accumulator + bias + residual -> fp32 stream, running row statistics; the residual chunks are
      //                   fetched two chunks ahead (chunks 0, 1 of tile i were issued at the end of tile i-1)
      //   then          : TMEM buffer back to the MMA warp, publish the statistics of tile i, issue the residual loads of
      //                   tile i+1's first two chunks, LayerNorm of tile i-1 (second half of the slice fetched here).
      // No warp waits for another CTA before it has published everything it owes, so there is no cyclic wait.
      static_assert(MODE != EPI_RESID_LN || (BLOCK_N == 256 && kColsPerWarp == 128),
                    "the LayerNorm exchange assumes eight 128-column slices per row");
      const uint32_t stg_s = smem_u32(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * 4096;
      const uint32_t stg_w = stg_s + lane * 128, stg_r = stg_s + (lane >> 3) * 128;
      // N == 1024 and one contiguous row range (checked by the launcher): tile -> (row block, N tile) by shifts
      const int warp_row = cta_rank * GEMM_BLOCK_M + q * 32;
      const size_t lane_off = static_cast<size_t>(lane >> 3) * kLnLd + col0 + 4 * (lane & 7);
      EpiRows pr[2];
      long long tile = tile_first;
      long long row0 = 0, row0_p = 0;
      int nvalid = 0, n0 = 0, nvalid_p = 0, n0_p = 0;
      if (tile < num_tiles) {
        row0 = (tile >> 2) * kTileM + warp_row;
        nvalid = gr.rows - static_cast<int>(row0);
        n0 = static_cast<int>(tile & 3) * BLOCK_N;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
          gemm_epi_resid_ln_prefetch(ep.resid + static_cast<size_t>(row0) * kLnLd + n0 + lane_off + 32 * ci, pr[ci], nvalid, lane);
      }
#ifdef BSEG_LN_TRACE
      long long ln_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ln_last = clock64();
      int ln_tiles = 0;
#endif
      while (tile < num_tiles) {
        // ---- loads for the LayerNorm of the previous tile (consumed after this tile's pass 1) ----
        LnPending pend;
        if (nvalid_p > 0) gemm_epi_resid_ln_issue(ep, pend, row0_p, nvalid_p, n0_p + col0, lane);
        LN_T(7);
        float rmean[8], rm2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { rmean[i] = 0.f; rm2[i] = 0.f; }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + col0;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        LN_T(0);
        const float* resid_thread = ep.resid + static_cast<size_t>(row0) * kLnLd + n0 + lane_off;
        float* out_thread = reinterpret_cast<float*>(ep.out) + static_cast<size_t>(row0) * kLnLd + n0 + lane_off;
        const float* bias_thread = ep.bias != nullptr ? ep.bias + n0 + col0 + 4 * (lane & 7) : nullptr;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          float v[32];
          tmem_ld32(taddr + 32 * ci, v);
          tmem_ld_wait();
          gemm_epi_resid_ln_chunk(bias_thread != nullptr ? bias_thread + 32 * ci : nullptr, out_thread + 32 * ci, v, pr[ci & 1],
                                  stg_w, stg_r, nvalid, lane, ci, rmean, rm2);
          if (ci < 2) gemm_epi_resid_ln_prefetch(resid_thread + 32 * (ci + 2), pr[ci & 1], nvalid, lane);
        }
        LN_T(1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtas == 2) mbar_arrive_cluster(&tmem_empty[as], 0);
          else mbar_arrive(&tmem_empty[as]);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        // (nvalid is warp-uniform; a 32-row group past the end of the row range publishes and normalises nothing)
        if (nvalid > 0) gemm_epi_resid_ln_publish(ep, rmean, rm2, row0, nvalid, (n0 >> 8) * 2 + ((warp - 4) >> 2), lane);
        LN_T(2);
        const long long row0_c = row0;
        const int nvalid_c = nvalid, n0_c = n0;
        tile += tile_step;
        if (tile < num_tiles) {
          row0 = (tile >> 2) * kTileM + warp_row;
          nvalid = gr.rows - static_cast<int>(row0);
          n0 = static_cast<int>(tile & 3) * BLOCK_N;
#pragma unroll
          for (int ci = 0; ci < 2; ++ci)
            gemm_epi_resid_ln_prefetch(ep.resid + static_cast<size_t>(row0) * kLnLd + n0 + lane_off + 32 * ci, pr[ci], nvalid, lane);
        }
        LN_T(3);
        if (nvalid_p > 0) {
          gemm_epi_resid_ln_finish(ep, pend, row0_p, nvalid_p, n0_p + col0, lane);
#ifdef BSEG_LN_TRACE
          ++ln_tiles;
#endif
        }
        LN_T(5);
        row0_p = row0_c; nvalid_p = nvalid_c; n0_p = n0_c;
      }
      if (nvalid_p > 0) {
        LnPending pend;
        gemm_epi_resid_ln_issue(ep, pend, row0_p, nvalid_p, n0_p + col0, lane);
        gemm_epi_resid_ln_finish(ep, pend, row0_p, nvalid_p, n0_p + col0, lane);
      }
#ifdef BSEG_LN_TRACE
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && (warp == 4 || warp == 9) && ln_tiles > 0)
        printf("LNTRACE block %d warp %d tiles %d | per tile: wait_acc %lld pass1 %lld release+publish %lld prefetch %lld "
               "layernorm %lld issue %lld\n", blockIdx.x, warp, ln_tiles, ln_t[0] / ln_tiles, ln_t[1] / ln_tiles,
               ln_t[2] / ln_tiles, ln_t[3] / ln_tiles, ln_t[5] / ln_tiles, ln_t[7] / ln_tiles);
#endif
    } else
    for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
      const long long mt = tile / num_n_tiles;
      const int local0 = gr.row_begin + static_cast<int>(mt % tiles_per_batch) * kTileM + cta_rank * GEMM_BLOCK_M + q * 32;
      const long long row0 = (mt / tiles_per_batch) * gr.rows_per_batch + local0;  // first row of this warp
      const int nvalid = gr.row_begin + gr.rows - local0;                          // rows of this warp inside the range
      const int n0 = static_cast<int>(tile % num_n_tiles) * BLOCK_N;
      const long long m = row0 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      if constexpr (MODE == EPI_F32 || MODE == EPI_RESID_F32 || MODE == EPI_EMBED) {
        float* stg = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * 1024;
        EpiRows pr, pn;
        gemm_epi_f32_prefetch<MODE>(ep, pr, row0, nvalid, n0 + col0, lane);  // does not depend on the accumulator
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = col0; c < col0 + kColsPerWarp; c += 32) {
          float v[32];
          tmem_ld32(taddr + c, v);
          if (c + 32 < col0 + kColsPerWarp) gemm_epi_f32_prefetch<MODE>(ep, pn, row0, nvalid, n0 + c + 32, lane);
          tmem_ld_wait();
          gemm_epi_f32_chunk<MODE>(ep, v, pr, stg, row0, nvalid, n0 + c, lane);
          pr = pn;
        }
      } else if constexpr (MODE == EPI_BF16 || MODE == EPI_BF16_GELU || MODE == EPI_DGELU) {
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * 1024;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = col0; c < col0 + kColsPerWarp; c += 32) {
          float v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          gemm_epi_bf16_chunk<MODE>(ep, v, stg, row0, nvalid, n0 + c, lane);
        }
      } else if constexpr (MODE == EPI_QKV) {
        // q / k: the 32 rows of a warp are 32 consecutive tokens of one sequence (T is a multiple of 32), i.e. 32
        // consecutive 128-byte rows of [seq, head, t, 64]: same coalesced store as the row-major epilogues.
        // v^T keeps the per-thread transposed scatter (one 64-byte run along t per instruction).
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * 1024;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const int D = ep.heads * 64;
        const long long seq = row0 / ep.T;
        const long long t0 = row0 % ep.T;
        const bool rows_in_seq = (ep.T % 32) == 0;
#pragma unroll 1
        for (int c = col0; c < col0 + kColsPerWarp; c += 32) {
          float v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          const int n = n0 + c;
          const int which = n / D;
          if (which < 2 && rows_in_seq) {
            if (ep.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + i));
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
              }
            }
            const int head = (n % D) >> 6;
            if (which == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= ep.q_scale;
            }
            __nv_bfloat16* base = (which == 0 ? ep.q : ep.k) + ((seq * ep.heads + head) * ep.T + t0) * 64;
            gemm_epi_bf16_store(base, 64, v, stg, 0, nvalid, n & 63, lane);
          } else if (lane < nvalid) {
            gemm_epilogue_chunk<MODE>(ep, v, m, n);
          }
        }
      } else {
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = col0; c < col0 + kColsPerWarp; c += 32) {
          float v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (lane < nvalid) gemm_epilogue_chunk<MODE>(ep, v, m, n0 + c);
        }
      }
      if constexpr (MODE != EPI_RESID_LN) {  // (that mode runs its own loop above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtas == 2) mbar_arrive_cluster(&tmem_empty[as], 0);  // the leader's barrier counts both CTAs
          else mbar_arrive(&tmem_empty[as]);
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();  // the peer's shared memory / TMEM stay valid until both are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCtas == 2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace bseg
