// Fused SegGPT attention, one CTA per (sequence, head, 128-query tile), two CTAs per SM (BSEG_ATTN_WG=1, the default;
// the warp map below shows the older BSEG_ATTN_WG=2 shape: one CTA per SM with two softmax warpgroups / 256 queries):
//     out = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v
// with the decomposed relative-position bias of modeling_seggpt.py:268-311 computed from the UNSCALED q
// (modeling_seggpt.py:324-329) and an fp32 softmax (:331).  The reference materialises a
// (16n,1568,1568) fp32 score tensor; here S, P and O live in TMEM and never touch HBM.
//
//   warp 0      TMA producer   (two 128-query Q tiles + rel tables once; 112-key K blocks and V^T blocks through
//                               two independent 3-stage rings)
//   warps 1-2   tcgen05 issuers, one per softmax warpgroup (G = Q*Rel^T once; per key block S = Q*K^T in two
//                               column halves (N=64, N=48) and O += P*V (N=64, P read from TMEM))
//   warp 3      idle (completes the control warpgroup, which gives its registers away with setmaxnreg)
//   warps 4-7   softmax warpgroup 0 (query rows   0..127 of the tile; thread <-> row == TMEM lane)
//   warps 8-11  softmax warpgroup 1 (query rows 128..255)
//
// Streaming softmax: a thread walks its S row in 16-column chunks straight out of TMEM and exponentiates them on the
// fly against the running reference m (the row max seen in EARLIER blocks), so TMEM loads, FMAs and MUFU.EX2 of one
// warp interleave instead of running in phases.  That is exact: a stale reference only changes the common scale of
// P, l and O.  The reference is raised lazily (when a block exceeds it by 2^16) and O in TMEM is rescaled then; if a
// half block exceeds it by 2^100 the half is redone against the new reference so that P cannot overflow.
// Each S half is handed back to the tensor core as soon as it has been consumed, so the next block's Q*K^T
// overlaps the current block's exponentials.  Key blocks are 112 keys = 4 rows of the 28-wide token grid: a score
// column maps to (kh, kw) at compile time and 1568 = 14 * 112 needs no key masking.
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

// experiment hooks of tools/micro/attn_trace.cu (how sensitive is the kernel to tensor-core work?); defaults = full work
#ifndef BSEG_ATTN_S_KSTEPS
#define BSEG_ATTN_S_KSTEPS 4
#endif
#ifndef BSEG_ATTN_PV_KSTEPS
#define BSEG_ATTN_PV_KSTEPS (kKB / 16)
#endif
#ifndef BSEG_ATTN_SKIP_EXP
#define BSEG_ATTN_SKIP_EXP 0
#endif

// BSEG_ATTN_WG = 2: one CTA per SM, two softmax warpgroups (256 queries) sharing the K/V stages.
// BSEG_ATTN_WG = 1: two CTAs per SM, one softmax warpgroup (128 queries) each, 2-stage rings, the rel tables overlay the
//                   V stages: 13 instead of 14 warpgroup tiles per (sequence, head), and one CTA's prologue / epilogue
//                   overlaps the other's main loop.
// Measured on B200 (tools/ab_attention_wg.sh, nseq 128): WG=2 2.716 ms (487 TFLOP/s), WG=1 2.522 ms (525 TFLOP/s).
#ifndef BSEG_ATTN_WG
#define BSEG_ATTN_WG 1
#endif

namespace attn {
constexpr int kWG = BSEG_ATTN_WG;
constexpr int kQTile = 128;            // queries per softmax warpgroup
constexpr int kCtaQ = kWG * kQTile;    // 256
constexpr int kKB = 112;               // keys per block
constexpr int kHalfLo = 64;            // S columns [0,64) and [64,112) are produced / released separately
constexpr int kGridW = 28;
constexpr int kGridH = 56;
constexpr int kT = kGridW * kGridH;    // 1568
constexpr int kNumKB = kT / kKB;       // 14
constexpr int kStages = kWG == 2 ? 3 : 2;
constexpr int kThreads = 128 + kWG * 128;  // 384 | 256
constexpr int kCtasPerSm = kWG == 2 ? 1 : 2;
constexpr int kRegsControl = kWG == 2 ? 64 : 40;
constexpr int kRegsSoftmax = 216;  // 128*64 + 256*216 = 63488 <= 65536 | 2 * (128*40 + 128*216) = 65536
constexpr int kRelRows = 176;  // 112 (reversed rel_pos_h, 111 used) + 64 (reversed rel_pos_w, 55 used)

constexpr int kQBytes = kQTile * 128;        // 16384 per warpgroup
constexpr int kKBytes = kKB * 128;           // 14336
constexpr int kVBytes = 2 * 64 * 128;        // 16384 (two 64-key halves)
constexpr int kRelBytes = kRelRows * 128;    // 22528
constexpr int kBhStride = 57;                // fp32 words per row (odd -> conflict free)
constexpr int kBhBytes = kQTile * kBhStride * 4;
constexpr int kBwStride = 29;
constexpr int kBwBytes = kQTile * kBwStride * 4;   // staging of the per-query width bias (then kept in registers)

constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kWG * kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffBh = kOffV + kStages * kVBytes;
// rel tables, then reused as bw staging; with one warpgroup per CTA the region overlays the (not yet used) V stages
constexpr int kRelRegion = (kWG * kBwBytes > kRelBytes) ? kWG * kBwBytes : kRelBytes;
constexpr bool kRelOverlaysV = (kWG == 1);
static_assert(!kRelOverlaysV || kRelRegion <= kStages * kVBytes, "rel overlay does not fit in the V stages");
constexpr int kOffRel = kRelOverlaysV ? kOffV : (kOffBh + kWG * kBhBytes + 1023) / 1024 * 1024;
constexpr int kOffBar = kRelOverlaysV ? (kOffBh + kWG * kBhBytes + 1023) / 1024 * 1024 : kOffRel + kRelRegion;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
static_assert(kSmemBytes * kCtasPerSm + 1024 * kCtasPerSm <= 228 * 1024, "shared memory budget");

// TMEM columns: warpgroup w owns [w*256, w*256+256): S at +0 (112 of 128), P at +128 (56 of 64, packed bf16 pairs),
// O at +192 (64); G (176) overlays S and P in the prologue
constexpr uint32_t kTmemCols = 256 * kWG;
constexpr uint32_t kColsPerWG = 256;
constexpr uint32_t kColP = 128;
constexpr uint32_t kColO = 192;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLazyThreshold = 16.0f;     // raise the reference when a block exceeds it by 2^16
constexpr long long kStaggerCycles = 1300;  // about half of a warpgroup's key-block period
constexpr float kOverflowGuard = 100.0f;    // redo a half block whose scores exceed the reference by 2^100
}  // namespace attn

// Optional timeline instrumentation (tools/micro/attn_trace.cu defines BSEG_ATTN_TRACE): clock64 stamps of one CTA.
#ifdef BSEG_ATTN_TRACE
__device__ long long g_attn_trace[3][16][16];  // [actor: wg0, wg1, mma0][block][event]
#define ATTN_TRACE(actor, kb, ev)                                                              \
  do {                                                                                         \
    if (trace_cta && lane == 0) g_attn_trace[actor][kb][ev] = clock64();                       \
  } while (0)
#else
#define ATTN_TRACE(actor, kb, ev) do {} while (0)
#endif

namespace {
__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t v, float a) {
  const float lo = __uint_as_float(v << 16) * a, hi = __uint_as_float(v & 0xffff0000u) * a;
  return pack_bf16x2(lo, hi);
}
}  // namespace

__global__ void __launch_bounds__(attn::kThreads, attn::kCtasPerSm)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_rel,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, int heads) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sRel = smem + kOffRel;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* k_full = bars + 2;     // [3]
  uint64_t* k_empty = bars + 5;    // [3]
  uint64_t* v_full = bars + 8;     // [3]
  uint64_t* v_empty = bars + 11;   // [3]
  uint64_t* s_full = bars + 14;    // [kWG][2]  MMA -> softmax: half h of S_j is in TMEM
  uint64_t* s_free = bars + 18;    // [kWG][2]  softmax -> MMA: half h of the S region may be overwritten
  uint64_t* p_full = bars + 22;    // [kWG]     softmax -> MMA: P_j is in TMEM (and O rescaled if it had to be)
  uint64_t* pv_done = bars + 24;   // [kWG]     MMA -> softmax: O += P_j V_j retired (P region free, O stable)
  uint64_t* rel_free = bars + 26;  // softmax -> TMA: the rel / bw staging region is dead (it overlays the V stages)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kCtaQ;
  const int head = blockIdx.y;
  const int seq = blockIdx.z;
  const int sh = seq * heads + head;
  const int n_active = (kWG == 2 && q0 + kQTile < kT) ? 2 : 1;  // the last tile of a sequence has one live warpgroup
#ifdef BSEG_ATTN_TRACE
  const bool trace_cta = blockIdx.x == 2 && blockIdx.y == 5 && blockIdx.z == gridDim.z / 2;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    mbar_init(g_full, n_active);         // one commit per MMA issuer
    mbar_init(rel_free, 4 * n_active);   // one arrive per live softmax warp
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], n_active);  // a stage is free once every live warpgroup's MMAs on it have retired
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], n_active);
    }
    for (int i = 0; i < kWG; ++i) {
      mbar_init(&s_full[2 * i], 1);
      mbar_init(&s_full[2 * i + 1], 1);
      mbar_init(&s_free[2 * i], 4);   // one arrive per softmax warp
      mbar_init(&s_free[2 * i + 1], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
    if (warp == 0) {
      // ============================ TMA producer (warp-uniform loop, one elected lane issues) ============================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, n_active * kQBytes + kRelBytes);
        for (int w = 0; w < n_active; ++w) tma_load_3d(sQ + w * kQBytes, &tmap_q, q_full, 0, q0 + w * kQTile, sh);
        tma_load_2d(sRel, &tmap_rel, q_full, 0, 0);
      }
      __syncwarp();
      for (int kb = 0; kb < kNumKB; ++kb) {
        const int st = kb % kStages;
        if (kb >= kStages) mbar_wait(&k_empty[st], ((kb / kStages) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&k_full[st], kKBytes);
          tma_load_3d(sK + st * kKBytes, &tmap_k, &k_full[st], 0, kb * kKB, sh);
        }
        __syncwarp();
        if (kb >= kStages) mbar_wait(&v_empty[st], ((kb / kStages) & 1) ^ 1);
        if (kRelOverlaysV && kb == 0) mbar_wait(rel_free, 0);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&v_full[st], kVBytes);
          tma_load_3d(sV + st * kVBytes, &tmap_vt, &v_full[st], kb * kKB, 0, sh);
          tma_load_3d(sV + st * kVBytes + 8192, &tmap_vt, &v_full[st], kb * kKB + 64, 0, sh);
        }
        __syncwarp();
      }
    } else if (warp - 1 < n_active) {
      // ============================ MMA issuers: warp 1 -> warpgroup 0, warp 2 -> warpgroup 1 ============================
      // One issuing warp per softmax warpgroup, blocking on that warpgroup's barriers in the order in which the
      // warpgroup arrives on them (S_lo free, S_hi free, P full), so neither warpgroup ever waits for the other's turn.
      // The whole warp runs the (warp-uniform) loop and one elected lane issues, which keeps every tcgen05.mma operand
      // in uniform registers.
      {
        const int w = warp - 1;
        constexpr uint32_t idesc_lo = umma_idesc_bf16(128, kHalfLo);
        constexpr uint32_t idesc_hi = umma_idesc_bf16(128, kKB - kHalfLo);
        constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
        const uint32_t q_addr = smem_u32(sQ) + w * kQBytes;
        const uint32_t rel_addr = smem_u32(sRel);
        const uint32_t tm = tmem_base + w * kColsPerWG;

        mbar_wait(q_full, 0);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(rel_addr + k * 32),
                         idesc_g, k != 0);
          umma_commit(g_full);
        }
        __syncwarp();

        auto issue_s = [&](int kb) {
          const int st = kb % kStages;
          mbar_wait(&k_full[st], (kb / kStages) & 1);
          if (w == 0) ATTN_TRACE(2, kb, 0);  // K block in smem
          const uint32_t k_addr = smem_u32(sK + st * kKBytes);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&s_free[2 * w + half], kb & 1);
            if (w == 0) ATTN_TRACE(2, kb, 1 + half);  // S half free -> issue
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < BSEG_ATTN_S_KSTEPS; ++k)
                umma_bf16_ss(tm + half * kHalfLo, umma_desc_sw128_kmajor(q_addr + k * 32),
                             umma_desc_sw128_kmajor(k_addr + half * (kHalfLo * 128) + k * 32),
                             half ? idesc_hi : idesc_lo, k != 0);
              umma_commit(&s_full[2 * w + half]);
              if (half == 1) umma_commit(&k_empty[st]);
            }
            __syncwarp();
          }
        };

        issue_s(0);
        for (int kb = 0; kb < kNumKB; ++kb) {
          if (kb + 1 < kNumKB) issue_s(kb + 1);
          const int st = kb % kStages;
          mbar_wait(&v_full[st], (kb / kStages) & 1);
          if (w == 0) ATTN_TRACE(2, kb, 3);  // V block in smem
          mbar_wait(&p_full[w], kb & 1);
          if (w == 0) ATTN_TRACE(2, kb, 4);  // P full -> issue PV
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t v_addr = smem_u32(sV + st * kVBytes);
#pragma unroll
            for (int k = 0; k < BSEG_ATTN_PV_KSTEPS; ++k) {
              const uint32_t va = v_addr + (k >> 2) * 8192 + (k & 3) * 32;
              umma_bf16_ts(tm + kColO, tm + kColP + k * 8, umma_desc_sw128_kmajor(va), idesc_o, (kb | k) != 0);
            }
            umma_commit(&pv_done[w]);
            umma_commit(&v_empty[st]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================ softmax warpgroups ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    const int w = (warp - 4) >> 2;
    if (w < n_active) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;  // query row in the warpgroup tile == TMEM lane
      const int qi_raw = q0 + w * kQTile + r;
      const bool valid = qi_raw < kT;
      const int qi = valid ? qi_raw : kT - 1;
      const int qh = qi / kGridW, qw = qi % kGridW;
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + w * kColsPerWG;
      float* bh_row = reinterpret_cast<float*>(smem + kOffBh + w * kBhBytes) + r * kBhStride;
      float* stage = reinterpret_cast<float*>(sRel + w * kBwBytes) + r * kBwStride;

      // ---- prologue: decomposed rel-pos bias of this query, pre-multiplied by log2(e) ----
      mbar_wait(g_full, 0);  // both G MMAs have retired: the rel tables in smem are dead, G is in TMEM
      tc_fence_after();
      {
        const int off_h = 55 - qh;  // bh[kh] = G[off_h + kh]
#pragma unroll
        for (int c = 0; c < 112; c += 16) {
          float v[16];
          tmem_ld16(lane_base + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int kh = c + i - off_h;
            if (kh >= 0 && kh < kGridH) bh_row[kh] = v[i] * kLog2e;
          }
        }
        const int off_w = 27 - qw;  // bw[kw] = G[112 + off_w + kw]
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          float v[16];
          tmem_ld16(lane_base + 112 + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int kw = c + i - off_w;
            if (kw >= 0 && kw < kGridW) stage[kw] = v[i] * kLog2e;
          }
        }
      }
      float bw[kGridW];
#pragma unroll
      for (int i = 0; i < kGridW; ++i) bw[i] = stage[i];
      // the staging area is about to be overwritten by TMA (it overlays the V stages): order this thread's generic-proxy
      // stores to it before the async-proxy writes that follow the rel_free hand-off
      if (kRelOverlaysV) fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // G consumed: both halves of the S region are free for S_0
        mbar_arrive(&s_free[2 * w]);
        mbar_arrive(&s_free[2 * w + 1]);
        if (kRelOverlaysV) mbar_arrive(rel_free);  // ... and the bw staging has been read
      }

      // De-phase the two warpgroups by about half a key block: both share the SM's MUFU (16 ex2/clk) and both have the
      // same compute / hand-off rhythm, so in lock-step they fight over MUFU and then idle together; staggered, one
      // exponentiates while the other waits for its barriers (the offset is neutrally stable, so it persists).
      if (w == 1 && n_active == 2) {
        const long long t_start = clock64();
        while (clock64() - t_start < kStaggerCycles) {
        }
      }

      const float sc = 0.125f * kLog2e;  // head_dim^-0.5 * log2(e)
      float m_run = 0.f, l_run = 0.f;
      float alpha_pending = 1.0f;  // factor still to be applied to O (after the P*V that is in flight retires)
      float bh4[4];
      float og[4];
      uint32_t pk[kKB / 2];  // P of the current block: bf16 pairs

      // exponentiate columns [c0, c1) of the S row against the current reference; returns sum and max exponent.
      // 32-column TMEM loads, the next one in flight while the current chunk is processed.
      auto process = [&](const float* cur, int c, int n, float& lsum, float& xmax) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (i < n) {
            const int col0 = c + i, col1 = c + i + 1;
            const float x0 = fmaf(cur[i], sc, bw[col0 % kGridW]) + og[col0 / kGridW];
            const float x1 = fmaf(cur[i + 1], sc, bw[col1 % kGridW]) + og[col1 / kGridW];
            xmax = fmaxf(xmax, fmaxf(x0, x1));
            const float p0 = BSEG_ATTN_SKIP_EXP ? x0 * 0.001f : ex2_approx(x0);
            const float p1 = BSEG_ATTN_SKIP_EXP ? x1 * 0.001f : ex2_approx(x1);
            lsum += p0 + p1;
            pk[col0 >> 1] = pack_bf16x2(p0, p1);
          }
        }
      };
      auto stream = [&](auto c0_tag, auto c1_tag, float& lsum, float& xmax) {
        constexpr int c0 = decltype(c0_tag)::value, c1 = decltype(c1_tag)::value;
        static_assert(c1 - c0 == 64 || c1 - c0 == 48, "half sizes");
        float bufa[32], bufb[32];
        tmem_ld32(lane_base + c0, bufa);
        tmem_ld_wait();
        if constexpr (c1 - c0 == 64) {
          tmem_ld32(lane_base + c0 + 32, bufb);
        } else {
          tmem_ld16(lane_base + c0 + 32, *reinterpret_cast<float(*)[16]>(&bufb[0]));
        }
        process(bufa, c0, 32, lsum, xmax);
        tmem_ld_wait();
        process(bufb, c0 + 32, c1 - c0 - 32, lsum, xmax);
      };
      using I0 = std::integral_constant<int, 0>;
      using I64 = std::integral_constant<int, kHalfLo>;
      using I112 = std::integral_constant<int, kKB>;

      for (int kb = 0; kb < kNumKB; ++kb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bh4[i] = lds32(bh_row + kb * 4 + i);

        // ---------------- lower half: columns 0..63 ----------------
        if (quarter == 0) ATTN_TRACE(w, kb, 0);  // block start
        mbar_wait(&s_full[2 * w], kb & 1);
        if (quarter == 0) ATTN_TRACE(w, kb, 1);  // S lower half ready
        tc_fence_after();
        if (kb == 0) {  // initial reference: row max over the first 64 keys
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < kHalfLo; c += 16) {
            float v[16];
            tmem_ld16(lane_base + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              mx = fmaxf(mx, fmaf(v[i], sc, bw[(c + i) % kGridW]) + bh4[(c + i) / kGridW]);
          }
          m_run = mx;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) og[i] = bh4[i] - m_run;
        float lsum = 0.f, xmax = -INFINITY;
        stream(I0{}, I64{}, lsum, xmax);
        if (__any_sync(0xffffffffu, xmax > kOverflowGuard)) {  // (practically never) redo against a safe reference
          const float up = fmaxf(xmax, 0.f);
          const float a = ex2_approx(-up);
          m_run += up;
          l_run *= a;
          alpha_pending *= a;
#pragma unroll
          for (int i = 0; i < 4; ++i) og[i] = bh4[i] - m_run;
          lsum = 0.f;
          xmax = -INFINITY;
          stream(I0{}, I64{}, lsum, xmax);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[2 * w]);  // the next block's lower S half may be issued

        // ---------------- upper half: columns 64..111 ----------------
        if (quarter == 0) ATTN_TRACE(w, kb, 2);  // lower half processed
        mbar_wait(&s_full[2 * w + 1], kb & 1);
        if (quarter == 0) ATTN_TRACE(w, kb, 3);  // S upper half ready
        tc_fence_after();
        float lsum_hi = 0.f, xmax_hi = -INFINITY;
        stream(I64{}, I112{}, lsum_hi, xmax_hi);
        if (__any_sync(0xffffffffu, xmax_hi > kOverflowGuard)) {
          const float up = fmaxf(xmax_hi, 0.f);
          const float a = ex2_approx(-up);
          m_run += up;
          l_run *= a;
          alpha_pending *= a;
          lsum *= a;
          xmax -= up;
#pragma unroll
          for (int i = 0; i < kHalfLo / 2; ++i) pk[i] = scale_bf16x2(pk[i], a);
#pragma unroll
          for (int i = 0; i < 4; ++i) og[i] = bh4[i] - m_run;
          lsum_hi = 0.f;
          xmax_hi = -INFINITY;
          stream(I64{}, I112{}, lsum_hi, xmax_hi);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[2 * w + 1]);
        lsum += lsum_hi;
        xmax = fmaxf(xmax, xmax_hi);

        // ---------------- hand P to the tensor core ----------------
        if (quarter == 0) ATTN_TRACE(w, kb, 4);  // upper half processed
        if (kb > 0) {
          // the P region and O are ours again once the previous P*V has retired
          mbar_wait(&pv_done[w], (kb - 1) & 1);
          if (quarter == 0) ATTN_TRACE(w, kb, 5);  // previous PV retired
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha_pending != 1.0f)) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
              float v[16];
              tmem_ld16(lane_base + kColO + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= alpha_pending;
              tmem_st16(lane_base + kColO + c, v);
            }
          }
        }
        alpha_pending = 1.0f;
#pragma unroll
        for (int c = 0; c < kKB / 2; c += 8)
          tmem_st8(lane_base + kColP + c, *reinterpret_cast<uint32_t(*)[8]>(&pk[c]));
        l_run += lsum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w]);
        if (quarter == 0) ATTN_TRACE(w, kb, 6);  // P handed over

        // lazily raise the reference for the following blocks
        if (xmax > kLazyThreshold) {
          const float a = ex2_approx(-xmax);
          m_run += xmax;
          l_run *= a;
          alpha_pending = a;  // applied to O once this block's P*V has retired
        }
      }

      // ---- epilogue: O / l -> bf16, token-major [seq, t, heads*64] ----
      mbar_wait(&pv_done[w], (kNumKB - 1) & 1);
      tc_fence_after();
      const float inv = alpha_pending / l_run;
      // log2-domain log-sum-exp of the row (saved for the backward pass): P = exp2(x - lse)
      if (lse_out != nullptr && valid) lse_out[static_cast<long long>(sh) * kT + qi] = m_run + log2f(l_run);
      __nv_bfloat16* dst = out + (static_cast<long long>(seq) * kT + qi) * (heads * 64) + head * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        float v[16];
        tmem_ld16(lane_base + kColO + c, v);
        tmem_ld_wait();
        if (valid) {
          *reinterpret_cast<uint4*>(dst + c) =
              make_uint4(pack_bf16x2(v[0] * inv, v[1] * inv), pack_bf16x2(v[2] * inv, v[3] * inv),
                         pack_bf16x2(v[4] * inv, v[5] * inv), pack_bf16x2(v[6] * inv, v[7] * inv));
          *reinterpret_cast<uint4*>(dst + c + 8) =
              make_uint4(pack_bf16x2(v[8] * inv, v[9] * inv), pack_bf16x2(v[10] * inv, v[11] * inv),
                         pack_bf16x2(v[12] * inv, v[13] * inv), pack_bf16x2(v[14] * inv, v[15] * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<attn::kTmemCols>(tmem_base);
  }
}

int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                     const __nv_bfloat16* relcat, __nv_bfloat16* out, float* lse_out, int nseq, int heads, int grid_h,
                     int grid_w, cudaStream_t stream) {
  using namespace attn;
  BSEG_REQUIRE(grid_h == kGridH && grid_w == kGridW, "attention: only the 56x28 token grid is supported (got %dx%d)",
               grid_h, grid_w);
  BSEG_REQUIRE(nseq > 0 && heads > 0, "attention: empty problem");
  CUtensorMap tq, tk, tv, tr;
  const uint64_t nsh = static_cast<uint64_t>(nseq) * heads;
  {
    uint64_t dims[3] = {64, static_cast<uint64_t>(kT), nsh};
    uint64_t strides[2] = {128, static_cast<uint64_t>(kT) * 128};
    uint32_t boxq[3] = {64, kQTile, 1};
    uint32_t boxk[3] = {64, kKB, 1};
    int rc = make_tmap_bf16(&tq, q, 3, dims, strides, boxq);
    if (rc) return rc;
    rc = make_tmap_bf16(&tk, k, 3, dims, strides, boxk);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(kT), 64, nsh};
    uint64_t strides[2] = {static_cast<uint64_t>(kT) * 2, static_cast<uint64_t>(kT) * 128};
    uint32_t box[3] = {64, 64, 1};
    int rc = make_tmap_bf16(&tv, vt, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    int rc = make_tmap_bf16_2d(&tr, relcat, 64, kRelRows, 64, 64, kRelRows);
    if (rc) return rc;
  }
  static PerDeviceFlag attr_once;
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(
        cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  dim3 grid((kT + kCtaQ - 1) / kCtaQ, heads, nseq);
  ProfScope prof(CAT_ATTENTION, static_cast<double>(nseq) * heads * (4.0 * kT * kT * 64 + 2.0 * kT * 84 * 64),
                 static_cast<double>(nseq) * heads * kT * 64 * 2 * 4, stream);
  attention_fwd_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, tr, out, lse_out, heads);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bseg
