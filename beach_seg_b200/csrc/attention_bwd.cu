// Backward of the fused SegGPT attention (forward: attention.cu; reference: torch autograd through
// modeling_seggpt.py:268-348).  With S = scale*q.k + q.Rh[qh-kh] + q.Rw[qw-kw], P = softmax(S), O = P v:
//     dP = dO v^T,   dS = P * (dP - D),  D = rowsum(dO * O),
//     dv = P^T dO,   dk = scale * dS^T q,
//     dq = scale * dS k  +  sum_kh dSh[q,kh] Rh[qh-kh]  +  sum_kw dSw[q,kw] Rw[qw-kw]
// (dSh / dSw = dS summed over the key columns / key rows of the 56x28 token grid; the rel-pos tables are frozen).
// Operand conventions (shared with the forward, attention.cu): q arrives as qs = bf16(q * scale * log2 e) and the rel-pos
// tables as relcat8 = 8 * rel (8 = 1 / scale), so qs.k and qs.relcat8 are the score and the bias in the log2 domain;
// dk = dS^T qs / log2(e) and dq = scale * (dS k + dG relcat8) are the gradients w.r.t. the UNSCALED projections.
// Nothing of size T x T touches HBM.  Two kernels, both with S, dP and the accumulators in TMEM:
//
//   attention_bwd_dq_kernel   one CTA per (seq, head, 128 queries); thread <-> query row (as in the forward), loops
//                             over 14 key blocks of 112 keys.  Also writes the per-query bias tables for the second
//                             kernel.
//   attention_bwd_dkv_kernel  one CTA per (seq, head, 128 keys); thread <-> key row (S^T = k q^T, so that P^T and dS^T
//                             are TMEM A operands of dv += P^T dO and dk += dS^T q), loops over 25 query blocks of 64.
//                             Bias, -lse and -D come out of the tensor core with the scores: S^T = k qs^T + A_h TabH +
//                             A_w TabW and dP^T - D = v dO^T + A_1 TabD, where A_* are constant one-hot / all-one
//                             columns per key row (smem) and Tab* the per-query 16-bit tables of the block, TMA-loaded
//                             and read as MN-major B operands.  Per element the key thread is left with one ex2, half
//                             a packed multiply and two halves of a bf16 pack.
//
// Each CTA: warp 0 TMA producer, warp 1 tcgen05 issuer, warps 2-3 idle, warps 4-11
// elementwise (two warps per TMEM lane quarter, splitting the columns).  S / dP are double buffered in TMEM so that the
// tensor core works on block j+1 while the elementwise warps are on block j.
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

namespace abwd {
constexpr int kGridW = 28, kGridH = 56;
constexpr int kT = kGridW * kGridH;  // 1568
constexpr int kThreads = 384;
constexpr int kRegsControl = 64, kRegsWork = 216;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kBiasRows = kGridH + kGridW;  // 84 rows of the per-(seq,head) bias table [84][T]: bh (56) then bw (28)

// ---------------- dq kernel ----------------
constexpr int kQTile = 128;
constexpr int kKB = 112, kNumKB = kT / kKB;  // 14
constexpr int kStagesQ = 4;
constexpr int kRelRows = 176;
constexpr int kQBytes = kQTile * 128;      // 16384
constexpr int kKBytes = kKB * 128;         // 14336
constexpr int kStageQBytes = 2 * kKBytes;  // 28672: K block, V block
constexpr int kRelRegion = 3 * 64 * 128;   // 24576: relcat8 [176 x 64]: K-major B of G = Q rel^T, MN-major B of dQ += dG rel
// Per-query tables, 16-bit, [row][T] per (seq, head), written here and read by the dkv kernel as MN-major B operands
// (K = table row, N = query): fp16 rows 0..55 bh[kh], 56..83 bw[kw], 84 / 85 = -lse split hi / lo, 86 / 87 = 0;
// bf16 rows 0..2 of a second table = -D split three ways.
constexpr int kTabRows = 88, kDRows = 3;
constexpr int kBiasTileBytes = kTabRows * kQTile * 2;  // 22528: the CTA's [88][128 queries] slice of the fp16 table
constexpr int kDshStride = 57;
constexpr int kDshBytes = kQTile * kDshStride * 4;
constexpr int kQOffQ = 0;
constexpr int kQOffdO = kQOffQ + kQBytes;
constexpr int kQOffRel = kQOffdO + kQBytes;
constexpr int kQOffRing = kQOffRel + kRelRegion;
constexpr int kQOffBias = kQOffRing + kStagesQ * kStageQBytes;
constexpr int kQOffDsh = kQOffBias + kBiasTileBytes;
constexpr int kQOffBar = kQOffDsh + kDshBytes;
constexpr int kQSmemBytes = kQOffBar + 256 + 1024;
static_assert(kQOffRing % 1024 == 0 && kStageQBytes % 1024 == 0 && kKBytes % 1024 == 0, "swizzle alignment");
static_assert(kQSmemBytes <= 227 * 1024, "dq kernel shared memory");
constexpr int kDswStride = 29;  // staging of the dSw partials in the (dead) ring after the loop
// TMEM columns: S0 [0,112) dP0 [112,224) S1 [224,336) dP1 [336,448) dQ [448,512); G (176) overlays S1/dP1 in the
// prologue, dG (88) overlays S0 at the end
constexpr uint32_t kQColBuf = 224, kQColdP = 112, kQColdQ = 448, kQColG = 224;

// ---------------- dkv kernel ----------------
constexpr int kKTile = 128;
constexpr int kQB = 64, kNumQB = (kT + kQB - 1) / kQB;  // 25 (the last block has 32 live queries)
constexpr int kStagesK = 5;
constexpr int kBufsK = 3;                  // S^T / dP^T buffers in TMEM = query blocks in flight
constexpr int kTileBytes = 64 * 128;       // 8192: [64 x 64] bf16
constexpr int kTabHBytes = 16 * 128, kTabWBytes = 32 * 128, kTabDBytes = 16 * 128;  // TMA boxes of 16 / 32 / 16 rows
// a stage: Q block, dO block (each both the K-major B of the score MMAs and the MN-major B of the gradient MMAs), tables
constexpr int kTabHOff = 2 * kTileBytes, kTabWOff = kTabHOff + kTabHBytes, kTabDOff = kTabWOff + kTabWBytes;
constexpr int kStageKBytes = kTabDOff + kTabDBytes;  // 24576
constexpr int kKOffK = 0;
constexpr int kKOffV = kKOffK + kKTile * 128;
constexpr int kKOffA = kKOffV + kKTile * 128;        // constant A operand of the fold MMAs: [128 keys][64 x 16 bit]
constexpr int kKOffRing = kKOffA + kKTile * 128;
constexpr int kKOffBar = kKOffRing + kStagesK * kStageKBytes;
constexpr int kKSmemBytes = kKOffBar + 256 + 1024;
static_assert(kStageKBytes % 1024 == 0 && kTabHOff % 1024 == 0 && kTabWOff % 1024 == 0 && kTabDOff % 1024 == 0,
              "swizzle alignment");
static_assert(kKSmemBytes <= 227 * 1024, "dkv kernel shared memory");
// TMEM columns: three S^T / dP^T buffers [b * 128, +64) / [b * 128 + 64, +64), dV [384,448), dK [448,512)
constexpr uint32_t kKColBuf = 128, kKColdP = 64, kKColdV = 384, kKColdK = 448;
}  // namespace abwd

// Optional timeline instrumentation (tools/micro/abwd_trace.cu defines BSEG_ABWD_TRACE): clock64 stamps of CTA (0,0,0).
#ifdef BSEG_ABWD_TRACE
__device__ long long g_abwd_trace[2][4][32][8];  // [kernel: dq, dkv][actor: tma, mma, wg0, wg1][block][event]
#define ABWD_TRACE(kern, actor, blk, ev)                                                              \
  do {                                                                                                \
    if (trace_cta && lane == 0) g_abwd_trace[kern][actor][blk][ev] = clock64();                       \
  } while (0)
#else
#define ABWD_TRACE(kern, actor, blk, ev) do {} while (0)
#endif

namespace {
__device__ __forceinline__ uint32_t tmem_lane_base(uint32_t tmem_base, int quarter) {
  return tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
}
}  // namespace

// =====================================================================================================
// dq
// =====================================================================================================
__global__ void __launch_bounds__(abwd::kThreads, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                        const __grid_constant__ CUtensorMap tmap_rel, const float* __restrict__ lse,
                        const float* __restrict__ Dvec, __half* tab16, __nv_bfloat16* dtab,
                        __nv_bfloat16* __restrict__ dqkv, int heads) {
  using namespace abwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kQOffQ;
  uint8_t* sdO = smem + kQOffdO;
  uint8_t* sRel = smem + kQOffRel;
  uint8_t* sRing = smem + kQOffRing;
  __half* sBias = reinterpret_cast<__half*>(smem + kQOffBias);
  float* sDsh = reinterpret_cast<float*>(smem + kQOffDsh);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kQOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;     // MMA -> all: G = Q rel^T is in TMEM
  uint64_t* g_free = bars + 2;     // elementwise -> MMA: G consumed, S/dP buffers may be written
  uint64_t* k_full = bars + 3;     // [4]
  uint64_t* kv_empty = bars + 7;   // [4]
  uint64_t* sdp_full = bars + 11;  // [2]
  uint64_t* ds_full = bars + 13;   // [2]
  uint64_t* dq_done = bars + 15;   // all scale*dS*K MMAs retired
  uint64_t* dg_full = bars + 16;   // elementwise -> MMA: dG is in TMEM
  uint64_t* dq_final = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y, seq = blockIdx.z;
  const int sh = seq * heads + head;
#ifdef BSEG_ABWD_TRACE
  const bool trace_cta = blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    mbar_init(g_full, 1);
    mbar_init(g_free, 8);
    for (int i = 0; i < kStagesQ; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], 4);   // the four warps of the warpgroup that owns the buffer
    }
    mbar_init(dq_done, 1);
    mbar_init(dg_full, 8);
    mbar_init(dq_final, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
    if (warp == 0) {
      // ============================ TMA producer (warp-uniform loop, one elected lane issues) ============================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, 2 * kQBytes + kRelRows * 128);
        tma_load_3d(sQ, &tmap_q, q_full, 0, q0, sh);
        tma_load_4d(sdO, &tmap_do, q_full, 0, q0, head, seq);
        tma_load_2d(sRel, &tmap_rel, q_full, 0, 0);
      }
      __syncwarp();
      for (int kb = 0; kb < kNumKB; ++kb) {
        const int st = kb % kStagesQ;
        ABWD_TRACE(0, 0, kb, 0);
        if (kb >= kStagesQ) mbar_wait(&kv_empty[st], ((kb / kStagesQ) & 1) ^ 1);
        ABWD_TRACE(0, 0, kb, 1);
        uint8_t* base = sRing + st * kStageQBytes;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&k_full[st], kStageQBytes);
          tma_load_3d(base, &tmap_k, &k_full[st], 0, kb * kKB, sh);
          tma_load_3d(base + kKBytes, &tmap_v, &k_full[st], 0, kb * kKB, sh);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ============================ MMA issuer (warp-uniform loop, one elected lane issues) ============================
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
      // B operands whose K dimension is the ROW index of a [rows][64 x bf16] tile (keys of the K block, rows of relcat8)
      // are read straight from the tile the K-major MMAs use, as MN-major operands: no transposed copies
      constexpr uint32_t idesc_o = umma_idesc_16bit(128, 64, 1, 1, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), rel_addr = smem_u32(sRel);
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + kQColG, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(rel_addr + k * 32),
                       idesc_g, k != 0);
        umma_commit(g_full);
      }
      __syncwarp();

      auto issue_sdp = [&](int kb) {
        const int st = kb % kStagesQ;
        ABWD_TRACE(0, 1, kb, 0);
        mbar_wait(&k_full[st], (kb / kStagesQ) & 1);
        ABWD_TRACE(0, 1, kb, 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t k_addr = smem_u32(sRing + st * kStageQBytes), v_addr = k_addr + kKBytes;
          const uint32_t d = tmem_base + (kb & 1) * kQColBuf;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(d, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                         k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(d + kQColdP, umma_desc_sw128_kmajor(do_addr + k * 32),
                         umma_desc_sw128_kmajor(v_addr + k * 32), idesc_s, k != 0);
          umma_commit(&sdp_full[kb & 1]);
        }
        __syncwarp();
        ABWD_TRACE(0, 1, kb, 2);
      };

      issue_sdp(0);  // buffer 0 does not overlap G: block 0's scores are computed under the bias-table prologue
      mbar_wait(g_free, 0);
      tc_fence_after();
      issue_sdp(1);
      for (int kb = 0; kb < kNumKB; ++kb) {
        const int st = kb % kStagesQ, buf = kb & 1;
        ABWD_TRACE(0, 1, kb, 3);
        mbar_wait(&ds_full[buf], (kb >> 1) & 1);
        ABWD_TRACE(0, 1, kb, 4);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t k_addr = smem_u32(sRing + st * kStageQBytes);
          const uint32_t a_base = tmem_base + buf * kQColBuf;
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k) {
            // dS (bf16 pairs) of columns [0,64) sits at S columns [0,32), of columns [64,112) at S columns [64,88)
            const uint32_t a = a_base + (k < 4 ? k * 8 : 64 + (k - 4) * 8);
            umma_bf16_ts(tmem_base + kQColdQ, a, umma_desc_sw128_mnmajor(k_addr + k * 2048), idesc_o, (kb | k) != 0);
          }
          umma_commit(&kv_empty[st]);
          if (kb == kNumKB - 1) umma_commit(dq_done);
        }
        __syncwarp();
        ABWD_TRACE(0, 1, kb, 5);
        if (kb + 2 < kNumKB) issue_sdp(kb + 2);
      }
      // bias gradient: dQ_acc += dG relcat8   (relcat8 = 8 rel = rel / scale)
      mbar_wait(dg_full, 0);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < kRelRows / 16; ++k)
          umma_bf16_ts(tmem_base + kQColdQ, tmem_base + k * 8, umma_desc_sw128_mnmajor(rel_addr + k * 2048), idesc_o,
                       1u);
        umma_commit(dq_final);
      }
      __syncwarp();
    }
  } else {
    // ============================ elementwise warps ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsWork));
    const int quarter = warp & 3;
    const int g = (warp - 4) >> 2;  // warpgroup: owns the key blocks kb with (kb & 1) == g (and S / dP buffer g)
    const int r = quarter * 32 + lane;
    const int qi_raw = q0 + r;
    const bool valid = qi_raw < kT;
    const int qi = valid ? qi_raw : kT - 1;
    const int qh = qi / kGridW, qw = qi % kGridW;
    const uint32_t lane_base = tmem_lane_base(tmem_base, quarter);
    __half* tab = tab16 + static_cast<long long>(sh) * kTabRows * kT;
    float* dsh_row = sDsh + r * kDshStride;

    // ---- prologue: the CTA's [88 rows][128 queries] slice of the fp16 bias table goes to shared memory (warpgroup 0:
    //      the 56 token-row biases bh, warpgroup 1: the 28 token-column biases bw and the -lse rows), is copied from there
    //      to the global table with 16-byte stores, and stays resident for this kernel's own bias lookups ----
    mbar_wait(g_full, 0);
    tc_fence_after();
    const float lse_q = valid ? lse[static_cast<long long>(sh) * kT + qi] : 0.f;
    const float d_q = valid ? Dvec[static_cast<long long>(sh) * kT + qi] : 0.f;
    if (g == 0) {
      const int off_h = 55 - qh;  // bh[kh] = G[off_h + kh]  (already in the log2 domain)
#pragma unroll
      for (int c = 0; c < 112; c += 16) {
        float v[16];
        tmem_ld16(lane_base + kQColG + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int kh = c + i - off_h;
          if (kh >= 0 && kh < kGridH) sBias[kh * kQTile + r] = __float2half_rn(v[i]);
        }
      }
    } else {
      const int off_w = 27 - qw;  // bw[kw] = G[112 + off_w + kw]
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        float v[16];
        tmem_ld16(lane_base + kQColG + 112 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int kw = c + i - off_w;
          if (kw >= 0 && kw < kGridW) sBias[(kGridH + kw) * kQTile + r] = __float2half_rn(v[i]);
        }
      }
      // rows that depend on the query only: -lse as fp16 hi + lo (rows 84, 85; 86, 87 = 0), -D as three bf16 terms
      const __half l_hi = __float2half_rn(-lse_q);
      sBias[(kBiasRows + 0) * kQTile + r] = l_hi;
      sBias[(kBiasRows + 1) * kQTile + r] = __float2half_rn(-lse_q - __half2float(l_hi));
      sBias[(kBiasRows + 2) * kQTile + r] = __float2half_rn(0.f);
      sBias[(kBiasRows + 3) * kQTile + r] = __float2half_rn(0.f);
      if (valid) {
        __nv_bfloat16* drow = dtab + static_cast<long long>(sh) * kDRows * kT + qi;
        float rest = -d_q;
#pragma unroll
        for (int i = 0; i < kDRows; ++i) {
          const __nv_bfloat16 part = __float2bfloat16_rn(rest);
          drow[i * kT] = part;
          rest -= __bfloat162float(part);
        }
      }
    }
    tc_fence_before();
    named_bar_sync(1, 256);  // the table slice is complete; every G column has been read
    if (lane == 0) mbar_arrive(g_free);
    {
      // 88 rows x 16 chunks of 8 queries; T - q0 is a multiple of 8, so a chunk is all valid or all out of range
      const int t = threadIdx.x - 128;
      for (int idx = t; idx < kTabRows * (kQTile / 8); idx += 256) {
        const int row = idx >> 4, c8 = (idx & 15) * 8;
        if (q0 + c8 < kT)
          *reinterpret_cast<uint4*>(tab + row * kT + q0 + c8) = *reinterpret_cast<const uint4*>(sBias + row * kQTile + c8);
      }
    }
    float bw[kGridW];
#pragma unroll
    for (int kw = 0; kw < kGridW; ++kw) bw[kw] = __half2float(sBias[(kGridH + kw) * kQTile + r]);

    float dsw[kGridW];
#pragma unroll
    for (int i = 0; i < kGridW; ++i) dsw[i] = 0.f;

    // One chunk of N columns starting at compile-time column C0: P = exp2(S + bias - lse), dS = P * (dP - D); pairs of
    // columns go through the packed fp32 pipe (a pair never straddles a token row: 28 is even).
    auto chunk = [&](auto c0_tag, auto n_tag, uint32_t sbase, const float (&boff)[4], float (&acc)[8]) {
      constexpr int C0 = decltype(c0_tag)::value, N = decltype(n_tag)::value;
      float s[32], dp[32];
      if constexpr (N == 32) {
        tmem_ld32(sbase + C0, s);
        tmem_ld32(sbase + kQColdP + C0, dp);
      } else {
        tmem_ld16(sbase + C0, *reinterpret_cast<float(*)[16]>(&s[0]));
        tmem_ld16(sbase + kQColdP + C0, *reinterpret_cast<float(*)[16]>(&dp[0]));
      }
      tmem_ld_wait();
      uint32_t pk[N / 2];
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const int c0 = C0 + i;
        const int wi = c0 % kGridW, hi = c0 / kGridW;
        float x0 = s[i], x1 = s[i + 1];
        add_f32x2(x0, x1, bw[wi], bw[wi + 1]);
        add_f32x2(x0, x1, boff[hi], boff[hi]);
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        float t0 = dp[i], t1 = dp[i + 1];
        add_f32x2(t0, t1, -d_q, -d_q);
        float ds0, ds1;
        mul_f32x2(ds0, ds1, p0, p1, t0, t1);
        add_f32x2(dsw[wi], dsw[wi + 1], ds0, ds1);
        add_f32x2(acc[2 * hi], acc[2 * hi + 1], ds0, ds1);
        pk[i >> 1] = pack_bf16x2(ds0, ds1);
      }
      // dS (bf16 pairs) in place over the S columns this thread has just consumed
      if constexpr (N == 32) {
        tmem_st16u(sbase + (C0 < 64 ? C0 / 2 : 64 + (C0 - 64) / 2), pk);
      } else {
        tmem_st8u(sbase + 64 + (C0 - 64) / 2, pk);
      }
    };
    using I0 = std::integral_constant<int, 0>;
    using I16 = std::integral_constant<int, 16>;
    using I32 = std::integral_constant<int, 32>;
    using I64 = std::integral_constant<int, 64>;
    using I96 = std::integral_constant<int, 96>;

    // The two warpgroups take alternate key blocks (warpgroup g <-> S / dP buffer g), each thread a whole row of 112
    // columns, so that one warpgroup's hand-off latencies (barrier wait, tcgen05.ld / st round trips) sit under the
    // other's exponentials.
    for (int kb = g; kb < kNumKB; kb += 2) {
      float boff[4], acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) boff[i] = __half2float(sBias[(kb * 4 + i) * kQTile + r]) - lse_q;
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 0);
      mbar_wait(&sdp_full[g], (kb >> 1) & 1);
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 1);
      tc_fence_after();
      const uint32_t sbase = lane_base + g * kQColBuf;
      chunk(I0{}, I32{}, sbase, boff, acc);
      chunk(I32{}, I32{}, sbase, boff, acc);
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 2);
      chunk(I64{}, I32{}, sbase, boff, acc);
      chunk(I96{}, I16{}, sbase, boff, acc);
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 3);
      tmem_st_wait();
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[g]);
      if (quarter == 0) ABWD_TRACE(0, 2 + g, kb, 5);
#pragma unroll
      for (int i = 0; i < 4; ++i) dsh_row[kb * 4 + i] = acc[2 * i] + acc[2 * i + 1];
    }

    // ---- bias gradient: dG[q, :] (176 columns) = dSh scattered at off_h + kh, dSw scattered at 112 + off_w + kw ----
    mbar_wait(dq_done, 0);  // every MMA that read the S buffers / the ring has retired
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(sRing) + (g * kQTile + r) * kDswStride;
#pragma unroll
    for (int i = 0; i < kGridW; ++i) stage[i] = dsw[i];
    named_bar_sync(1, 256);
    if (g == 0) {
      // columns 0..111 -> packed words 0..55
      const int off_h = 55 - qh;
#pragma unroll 1
      for (int w0 = 0; w0 < 56; w0 += 8) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ka = 2 * (w0 + j) - off_h, kb2 = ka + 1;
          const float a = (ka >= 0 && ka < kGridH) ? dsh_row[ka] : 0.f;
          const float b = (kb2 >= 0 && kb2 < kGridH) ? dsh_row[kb2] : 0.f;
          pk[j] = pack_bf16x2(a, b);
        }
        tmem_st8u(lane_base + w0, pk);
      }
    } else {
      // columns 112..175 -> packed words 56..87
      const int off_w = 27 - qw;
      const float* other = reinterpret_cast<const float*>(sRing) + r * kDswStride;  // group 0's partials
#pragma unroll 1
      for (int w0 = 0; w0 < 32; w0 += 8) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ka = 2 * (w0 + j) - off_w, kb2 = ka + 1;
          const float a = (ka >= 0 && ka < kGridW) ? (stage[ka] + other[ka]) : 0.f;
          const float b = (kb2 >= 0 && kb2 < kGridW) ? (stage[kb2] + other[kb2]) : 0.f;
          pk[j] = pack_bf16x2(a, b);
        }
        tmem_st8u(lane_base + 56 + w0, pk);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(dg_full);

    // ---- epilogue: dq = scale * accumulator -> bf16, token-major [seq*T + t][0*D + head*64 + d] ----
    mbar_wait(dq_final, 0);
    tc_fence_after();
    {
      float v[32];
      tmem_ld32(lane_base + kQColdQ + g * 32, v);
      tmem_ld_wait();
      if (valid) {
        __nv_bfloat16* dst = dqkv + (static_cast<long long>(seq) * kT + qi) * (3 * heads * 64) + head * 64 + g * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(
              pack_bf16x2(v[i] * 0.125f, v[i + 1] * 0.125f), pack_bf16x2(v[i + 2] * 0.125f, v[i + 3] * 0.125f),
              pack_bf16x2(v[i + 4] * 0.125f, v[i + 5] * 0.125f), pack_bf16x2(v[i + 6] * 0.125f, v[i + 7] * 0.125f));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================================
// dk, dv
// =====================================================================================================
__global__ void __launch_bounds__(abwd::kThreads, 1)
attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                         const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_tabh, const __grid_constant__ CUtensorMap tmap_tabw,
                         const __grid_constant__ CUtensorMap tmap_tabd, __nv_bfloat16* __restrict__ dqkv, int heads) {
  using namespace abwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + kKOffK;
  uint8_t* sV = smem + kKOffV;
  uint8_t* sA = smem + kKOffA;
  uint8_t* sRing = smem + kKOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kKOffBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* full = bars + 1;        // [5]
  uint64_t* empty = bars + 6;       // [5]
  uint64_t* sdp_full = bars + 11;   // [3]
  uint64_t* pds_full = bars + 14;   // [3]
  uint64_t* dkv_done = bars + 17;
  uint64_t* buf_free = bars + 18;   // [3] gradient MMAs of the block that used the buffer have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kKTile;
  const int head = blockIdx.y, seq = blockIdx.z;
  const int sh = seq * heads + head;
  const int kh_lo = k0 / kGridW;
#ifdef BSEG_ABWD_TRACE
  const bool trace_cta = blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_tabh);
    tma_prefetch_desc(&tmap_tabw);
    tma_prefetch_desc(&tmap_tabd);
    mbar_init(kv_full, 1);
    for (int i = 0; i < kStagesK; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kBufsK; ++i) {
      mbar_init(&sdp_full[i], 2);   // one commit from the S^T issuer, one from the dP^T issuer
      mbar_init(&pds_full[i], 4);   // the four warps of the warpgroup that works on the block
      mbar_init(&buf_free[i], 1);
    }
    mbar_init(dkv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // Constant A operand of the fold MMAs, one K-major row of 64 16-bit values per key (128-byte swizzle):
  //   fp16 [0,16)   one-hot at the key's token row relative to kh_lo          x TabH rows kh_lo ..
  //   fp16 [16,48)  one-hot at 16 + the key's token column; 1 at 44 and 45     x TabW rows 56 .. 87 (bw, -lse hi, lo)
  //   bf16 [48,64)  1 at 48, 49, 50                                           x TabD rows 0 .. 2 (-D split)
  for (int idx = threadIdx.x; idx < kKTile * 8; idx += kThreads) {
    const int n = idx >> 3, c = idx & 7;
    const int ki = min(k0 + n, kT - 1);
    const int khi = ki / kGridW - kh_lo, kw = ki % kGridW;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pair = 0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int col = 8 * c + 2 * e + hf;
        uint32_t bits = 0;
        if (col < 16) bits = (col == khi) ? 0x3C00u : 0u;
        else if (col < 48) bits = (col - 16 == kw || col == 44 || col == 45) ? 0x3C00u : 0u;
        else bits = (col < 48 + kDRows) ? 0x3F80u : 0u;
        pair |= bits << (16 * hf);
      }
      w[e] = pair;
    }
    *reinterpret_cast<uint4*>(sA + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_proxy_async_smem();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp < 4) {
    if (warp == 0) {
      // ============================ TMA producer (warp-uniform loop, one elected lane issues) ============================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(kv_full, 2 * kKTile * 128);
        tma_load_3d(sK, &tmap_k, kv_full, 0, k0, sh);
        tma_load_3d(sV, &tmap_v, kv_full, 0, k0, sh);
      }
      __syncwarp();
      for (int j = 0; j < kNumQB; ++j) {
        const int st = j % kStagesK;
        ABWD_TRACE(1, 0, j, 0);
        if (j >= kStagesK) mbar_wait(&empty[st], ((j / kStagesK) & 1) ^ 1);
        ABWD_TRACE(1, 0, j, 1);
        if (elect_one_sync()) {
          uint8_t* base = sRing + st * kStageKBytes;
          mbar_arrive_expect_tx(&full[st], kStageKBytes);  // full boxes: out-of-range rows / queries arrive as zeros
          tma_load_3d(base, &tmap_q, &full[st], 0, j * kQB, sh);
          tma_load_4d(base + kTileBytes, &tmap_do, &full[st], 0, j * kQB, head, seq);
          // 16 table rows from kh_lo (the tile touches at most 6; the others meet zero columns of A), the 32 rows
          // bw / -lse, the 3 (+13 out-of-range = zero) rows of -D
          tma_load_3d(base + kTabHOff, &tmap_tabh, &full[st], j * kQB, kh_lo, sh);
          tma_load_3d(base + kTabWOff, &tmap_tabw, &full[st], j * kQB, kGridH, sh);
          tma_load_3d(base + kTabDOff, &tmap_tabd, &full[st], j * kQB, 0, sh);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 3) {
      // ============================ score-MMA issuers (warp-uniform loops, one elected lane issues) ============================
      // Three issuing warps (S^T in warp 1, dP^T in warp 3, the gradients in warp 2): a tcgen05.mma of N = 64 costs its
      // issuer ~57 cycles and an mbarrier wait ~170 even when the phase is long complete, so one warp doing all of it
      // was the kernel's critical path.
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      constexpr uint32_t idesc_tab = umma_idesc_16bit(128, 64, 0, 0, 0, 1);   // fp16 x fp16, B MN-major
      constexpr uint32_t idesc_tabd = umma_idesc_16bit(128, 64, 1, 1, 0, 1);  // bf16 x bf16, B MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), a_addr = smem_u32(sA);
      mbar_wait(kv_full, 0);
      for (int j = 0; j < kNumQB; ++j) {
        const int st = j % kStagesK, buf = j % kBufsK;
        if (warp == 1) ABWD_TRACE(1, 1, j, 0);
        mbar_wait(&full[st], (j / kStagesK) & 1);
        if (j >= kBufsK) mbar_wait(&buf_free[buf], (j / kBufsK - 1) & 1);  // P^T / dS^T of block j - 3 consumed
        if (warp == 1) ABWD_TRACE(1, 1, j, 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t q_addr = smem_u32(sRing + st * kStageKBytes), do_addr = q_addr + kTileBytes;
          const uint32_t d = tmem_base + buf * kKColBuf;
          if (warp == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // S^T = K Q_blk^T
              umma_bf16_ss(d, umma_desc_sw128_kmajor(k_addr + k * 32), umma_desc_sw128_kmajor(q_addr + k * 32), idesc,
                           k != 0);
            // ... + bias(key, query) - lse(query)
            umma_bf16_ss(d, umma_desc_sw128_kmajor(a_addr), umma_desc_sw128_mnmajor(q_addr + kTabHOff), idesc_tab, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16_ss(d, umma_desc_sw128_kmajor(a_addr + 32 + k * 32),
                           umma_desc_sw128_mnmajor(q_addr + kTabWOff + k * 2048), idesc_tab, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // dP^T = V dO_blk^T
              umma_bf16_ss(d + kKColdP, umma_desc_sw128_kmajor(v_addr + k * 32),
                           umma_desc_sw128_kmajor(do_addr + k * 32), idesc, k != 0);
            // ... - D(query)
            umma_bf16_ss(d + kKColdP, umma_desc_sw128_kmajor(a_addr + 96), umma_desc_sw128_mnmajor(q_addr + kTabDOff),
                         idesc_tabd, 1u);
          }
          umma_commit(&sdp_full[buf]);
        }
        __syncwarp();
        if (warp == 1) ABWD_TRACE(1, 1, j, 2);
      }
    } else if (warp == 2) {
      // ============================ gradient-MMA issuer ============================
      constexpr uint32_t idesc_grad = umma_idesc_16bit(128, 64, 1, 1, 0, 1);  // K = query = ROW of the dO / Q tile
      for (int j = 0; j < kNumQB; ++j) {
        const int st = j % kStagesK, buf = j % kBufsK;
        ABWD_TRACE(1, 1, j, 3);
        mbar_wait(&full[st], (j / kStagesK) & 1);   // complete long ago; observed here for the operand tiles' visibility
        mbar_wait(&pds_full[buf], (j / kBufsK) & 1);
        ABWD_TRACE(1, 1, j, 4);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t q_addr = smem_u32(sRing + st * kStageKBytes), do_addr = q_addr + kTileBytes;
          const uint32_t a_base = tmem_base + buf * kKColBuf;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // P^T (bf16 pairs) of queries [0,32) at S columns [0,16), of queries [32,64) at S columns [32,48);
            // the B operand is the block's dO / Q tile read MN-major: 16 query rows = 2048 bytes per K step
            const uint32_t a_off = (k < 2 ? k * 8 : 32 + (k - 2) * 8);
            umma_bf16_ts(tmem_base + kKColdV, a_base + a_off, umma_desc_sw128_mnmajor(do_addr + k * 2048), idesc_grad,
                         (j | k) != 0);
            umma_bf16_ts(tmem_base + kKColdK, a_base + kKColdP + a_off, umma_desc_sw128_mnmajor(q_addr + k * 2048),
                         idesc_grad, (j | k) != 0);
          }
          umma_commit(&empty[st]);
          umma_commit(&buf_free[buf]);
          if (j == kNumQB - 1) umma_commit(dkv_done);
        }
        __syncwarp();
        ABWD_TRACE(1, 1, j, 5);
      }
    }
  } else {
    // ============================ elementwise warps ============================
    const int quarter = warp & 3;
    const int g = (warp - 4) >> 2;  // warpgroup: takes the query blocks j with (j & 1) == g
    const int r = quarter * 32 + lane;
    const int ki_raw = k0 + r;
    const bool valid = ki_raw < kT;
    const int ki = valid ? ki_raw : kT - 1;
    const uint32_t lane_base = tmem_lane_base(tmem_base, quarter);

    // The two warpgroups take alternate query blocks, each thread the whole 64-column row of its key, and three
    // blocks are in flight (buffer j % 3): while a warpgroup exponentiates block j the tensor core finishes the
    // gradient MMAs of j - 1 / j - 2 and the score MMAs of j + 1 / j + 2, so neither side waits for a round trip.
    for (int j = g; j < kNumQB; j += 2) {
      const int buf = j % kBufsK;
      if (quarter == 0) ABWD_TRACE(1, 2 + g, j, 0);
      mbar_wait(&sdp_full[buf], (j / kBufsK) & 1);
      if (quarter == 0) ABWD_TRACE(1, 2 + g, j, 1);
      tc_fence_after();
      const uint32_t sbase = lane_base + buf * kKColBuf;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s[32], dp[32];
        tmem_ld32(sbase + h * 32, s);              // S^T + bias - lse  (log2 domain)
        tmem_ld32(sbase + kKColdP + h * 32, dp);   // dP^T - D
        tmem_ld_wait();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(s[i]), p1 = ex2_approx(s[i + 1]);
          float d0, d1;
          mul_f32x2(d0, d1, p0, p1, dp[i], dp[i + 1]);
          pp[i >> 1] = pack_bf16x2(p0, p1);
          pd[i >> 1] = pack_bf16x2(d0, d1);
        }
        tmem_st16u(sbase + h * 32, pp);
        tmem_st16u(sbase + kKColdP + h * 32, pd);
        if (quarter == 0) ABWD_TRACE(1, 2 + g, j, 2 + h);
      }
      tmem_st_wait();
      if (quarter == 0) ABWD_TRACE(1, 2 + g, j, 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[buf]);
      if (quarter == 0) ABWD_TRACE(1, 2 + g, j, 5);
    }

    // ---- epilogue: dv -> columns [2D, 3D), dk * scale -> columns [D, 2D) of the token-major dqkv rows ----
    mbar_wait(dkv_done, 0);
    tc_fence_after();
    const int D = heads * 64;
    __nv_bfloat16* row = dqkv + (static_cast<long long>(seq) * kT + ki) * (3 * D) + head * 64 + g * 32;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float v[32];
      tmem_ld32(lane_base + (which == 0 ? kKColdV : kKColdK) + g * 32, v);
      tmem_ld_wait();
      const float a = which == 0 ? 1.0f : 0.6931471805599453f;  // dk = dS^T qs / log2(e)
      if (valid) {
        __nv_bfloat16* dst = row + (which == 0 ? 2 * D : D);
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) =
              make_uint4(pack_bf16x2(v[i] * a, v[i + 1] * a), pack_bf16x2(v[i + 2] * a, v[i + 3] * a),
                         pack_bf16x2(v[i + 4] * a, v[i + 5] * a), pack_bf16x2(v[i + 6] * a, v[i + 7] * a));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================================
// launcher
// =====================================================================================================
int launch_attention_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* dO,
                         const float* lse, const float* Dvec, const __nv_bfloat16* relcat, float* bias_tab,
                         __nv_bfloat16* dqkv, int nseq, int heads, cudaStream_t stream) {
  using namespace abwd;
  BSEG_REQUIRE(nseq > 0 && heads > 0, "attention_bwd: empty problem");
  const uint64_t nsh = static_cast<uint64_t>(nseq) * heads;
  const uint64_t D = static_cast<uint64_t>(heads) * 64;
  CUtensorMap tq128, tq64, tk128, tk112, tv128, tv112, tdo128, tdo64, trel, ttabh, ttabw, ttabd;
  // the scratch the caller sizes as [nsh, 84, T] floats holds both 16-bit tables
  __half* tab16 = reinterpret_cast<__half*>(bias_tab);
  __nv_bfloat16* dtab = reinterpret_cast<__nv_bfloat16*>(tab16 + nsh * kTabRows * kT);
  static_assert((kTabRows + kDRows) * 2 <= kBiasRows * 4, "16-bit tables must fit in the bias scratch");
  int rc;
  {
    uint64_t dims[3] = {64, static_cast<uint64_t>(kT), nsh};
    uint64_t strides[2] = {128, static_cast<uint64_t>(kT) * 128};
    uint32_t b128[3] = {64, 128, 1}, b112[3] = {64, kKB, 1}, b64[3] = {64, 64, 1};
    if ((rc = make_tmap_bf16(&tq128, q, 3, dims, strides, b128))) return rc;
    if ((rc = make_tmap_bf16(&tq64, q, 3, dims, strides, b64))) return rc;
    if ((rc = make_tmap_bf16(&tk128, k, 3, dims, strides, b128))) return rc;
    if ((rc = make_tmap_bf16(&tk112, k, 3, dims, strides, b112))) return rc;
    if ((rc = make_tmap_bf16(&tv128, v, 3, dims, strides, b128))) return rc;
    if ((rc = make_tmap_bf16(&tv112, v, 3, dims, strides, b112))) return rc;
  }
  {
    // dO is token-major [nseq*T, heads*64]: dims (d, t, head, seq)
    uint64_t dims[4] = {64, static_cast<uint64_t>(kT), static_cast<uint64_t>(heads), static_cast<uint64_t>(nseq)};
    uint64_t strides[3] = {D * 2, 128, static_cast<uint64_t>(kT) * D * 2};
    uint32_t b128[4] = {64, 128, 1, 1}, b64[4] = {64, 64, 1, 1};
    if ((rc = make_tmap_bf16(&tdo128, dO, 4, dims, strides, b128))) return rc;
    if ((rc = make_tmap_bf16(&tdo64, dO, 4, dims, strides, b64))) return rc;
  }
  {
    // per-query tables [nsh, 88, T] fp16 and [nsh, 3, T] bf16 (written by the dq kernel); 16-bit data either way
    uint64_t dims[3] = {static_cast<uint64_t>(kT), kTabRows, nsh};
    uint64_t strides[2] = {static_cast<uint64_t>(kT) * 2, static_cast<uint64_t>(kT) * kTabRows * 2};
    uint32_t b16[3] = {kQB, 16, 1}, b32[3] = {kQB, 32, 1};
    if ((rc = make_tmap_bf16(&ttabh, tab16, 3, dims, strides, b16))) return rc;
    if ((rc = make_tmap_bf16(&ttabw, tab16, 3, dims, strides, b32))) return rc;
    uint64_t ddims[3] = {static_cast<uint64_t>(kT), kDRows, nsh};
    uint64_t dstrides[2] = {static_cast<uint64_t>(kT) * 2, static_cast<uint64_t>(kT) * kDRows * 2};
    if ((rc = make_tmap_bf16(&ttabd, dtab, 3, ddims, dstrides, b16))) return rc;
  }
  if ((rc = make_tmap_bf16_2d(&trel, relcat, 64, kRelRows, 64, 64, kRelRows))) return rc;
  static PerDeviceFlag attr_once;
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(
        cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmemBytes));
    BSEG_CHECK_CUDA(
        cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKSmemBytes));
  }
  const double pair = static_cast<double>(nseq) * heads * kT * kT;
  {
    dim3 grid((kT + kQTile - 1) / kQTile, heads, nseq);
    ProfScope prof(CAT_ATTENTION, pair * 64 * 2 * 3 + static_cast<double>(nseq) * heads * kT * 2.0 * 176 * 64 * 2,
                   static_cast<double>(nseq) * heads * kT * (64 * 2 * 5 + (kTabRows + kDRows) * 2), stream);
    attention_bwd_dq_kernel<<<grid, kThreads, kQSmemBytes, stream>>>(tq128, tdo128, tk112, tv112, trel, lse, Dvec, tab16,
                                                                     dtab, dqkv, heads);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    dim3 grid((kT + kKTile - 1) / kKTile, heads, nseq);
    ProfScope prof(CAT_ATTENTION, pair * 64 * 2 * 4, static_cast<double>(nseq) * heads * kT * (64 * 2 * 6), stream);
    attention_bwd_dkv_kernel<<<grid, kThreads, kKSmemBytes, stream>>>(tk128, tv128, tq64, tdo64, ttabh, ttabw, ttabd, dqkv,
                                                                      heads);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

}  // namespace bseg
