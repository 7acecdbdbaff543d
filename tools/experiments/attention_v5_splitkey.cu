// EXPERIMENT, NOT BUILT INTO libbseg.so (kept for the record; compiles against beach_seg_b200/csrc headers).
// Result on B200 (r01): numerically correct (all tests of tests/test_gpu_attention.py pass), but 0.916 ms vs 0.699 ms
// for the two-warpgroup kernel at nseq = 32: with 640 threads the softmax code has to live in 104 registers (spills in
// the hot loop) and every per-block fixed cost (barrier waits, fences, reference bookkeeping) is paid by twice as many
// threads for half as many elements each.  Lesson recorded in profiles/r01_attn_fwd_sensitivity.txt.
// Fused SegGPT attention, forward, "split-key" variant (v5).  Same math as attention.cu (modeling_seggpt.py:268-348):
//     out = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v
// One CTA = two 128-query tiles of one (sequence, head).  Each tile is served by TWO softmax warpgroups that split the
// KEYS of every 112-key block between them (columns [0,64) and [64,112)) and run the streaming softmax independently on
// their key subset -- own running reference m, own denominator l, own accumulator O in TMEM -- exactly like split-K
// "flash decoding"; the two partial results are merged once at the end:
//     out = (O_lo 2^(m_lo-M) + O_hi 2^(m_hi-M)) / (l_lo 2^(m_lo-M) + l_hi 2^(m_hi-M)),  M = max(m_lo, m_hi).
// Why: the forward was limited by the issue efficiency of 8 softmax warps (2 per scheduler) and by ~5 barrier hand-offs
// per key block (profiles/r01_attn_fwd_sensitivity.txt).  Here 16 softmax warps (4 per scheduler) hide each other's
// waits, no state is shared between the two threads of a row during the loop, P is written in place over the consumed
// S columns, and a warpgroup has two hand-offs per block (S ready, P ready).
//
//   warp 0        TMA producer (Q tiles + rel tables once; K / V^T blocks through 3-stage rings)
//   warps 1-2     tcgen05 issuers, one per tile; each serves its two warpgroups in arrival order
//   warp 3        idle
//   warps 4-19    softmax: tile = (warp-4)/8, key half = ((warp-4)/4)&1, TMEM lane quarter = warp&3
//
// TMEM per tile (256 columns): S_lo [0,64) (P_lo in place, 32 cols), S_hi [64,112) (P_hi in place at 64, 24 cols),
// O_lo [112,176), O_hi [176,240); the prologue's G = Q*relcat^T (176 columns) overlays [0,176).
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

namespace attn5 {
constexpr int kTiles = 2;
constexpr int kQTile = 128;
constexpr int kCtaQ = kTiles * kQTile;  // 256
constexpr int kKB = 112;
constexpr int kHalfLo = 64;
constexpr int kGridW = 28, kGridH = 56;
constexpr int kT = kGridW * kGridH;     // 1568
constexpr int kNumKB = kT / kKB;        // 14
constexpr int kStages = 3;
constexpr int kThreads = 128 + 16 * 32;  // 640
// setmaxnreg moves registers inside the CTA's own pool: the kernel launches with 96 registers per thread (the cap for
// 640 threads), the four control warps give (96-56)*128 = 5120 back and the sixteen softmax warps may take
// (104-96)*512 = 4096 of them.  (Asking for more than was released makes setmaxnreg.inc spin forever.)
constexpr int kRegsControl = 56;
constexpr int kRegsSoftmax = 104;
constexpr int kRelRows = 176;

constexpr int kQBytes = kQTile * 128;
constexpr int kKBytes = kKB * 128;
constexpr int kVBytes = 2 * 64 * 128;
constexpr int kRelBytes = kRelRows * 128;
constexpr int kBhStride = 57;
constexpr int kBhBytes = kQTile * kBhStride * 4;
constexpr int kBwStride = 29;

constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kTiles * kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffBh = kOffV + kStages * kVBytes;
constexpr int kOffRel = (kOffBh + kTiles * kBhBytes + 1023) / 1024 * 1024;
// the rel region holds relcat [176 x 128 B] for the prologue MMA, afterwards the bw staging [2 tiles][128][29] fp32
constexpr int kStageBytes = kTiles * kQTile * kBwStride * 4;            // 29696
constexpr int kRelRegion = (kStageBytes + 1023) / 1024 * 1024;          // 29696 (29 KB)
constexpr int kExchStride = 4;                                           // floats per (row, half): m, l, alpha, pad
constexpr int kOffExch = kOffRel + kRelRegion;                           // merge exchange [2 tiles][128][2][4] fp32
constexpr int kExchBytes = kTiles * kQTile * 2 * kExchStride * 4;        // 8192
constexpr int kOffBar = kOffExch + kExchBytes;
constexpr int kSmemBytes = kOffBar + 512 + 1024;
static_assert(kRelRegion >= kRelBytes, "rel tables must fit");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColsPerTile = 256;
constexpr uint32_t kColSHi = 64;
constexpr uint32_t kColO = 112;  // O_lo at +112, O_hi at +176

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLazyThreshold = 16.0f;
constexpr float kOverflowGuard = 100.0f;
}  // namespace attn5

namespace {
__device__ __forceinline__ uint32_t scale_bf16x2_v5(uint32_t v, float a) {
  const float lo = __uint_as_float(v << 16) * a, hi = __uint_as_float(v & 0xffff0000u) * a;
  return pack_bf16x2(lo, hi);
}
}  // namespace

__global__ void __launch_bounds__(attn5::kThreads, 1)
attention_fwd_v5_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_rel,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, int heads) {
  using namespace attn5;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sRel = smem + kOffRel;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* g_free = bars + 2;     // [kTiles]  softmax -> MMA: G consumed, S regions may be written
  uint64_t* k_full = bars + 4;     // [3]
  uint64_t* k_empty = bars + 7;    // [3]
  uint64_t* v_full = bars + 10;    // [3]
  uint64_t* v_empty = bars + 13;   // [3]
  uint64_t* s_full = bars + 16;    // [kTiles][2]  MMA -> softmax: S half of block kb is in TMEM (and P*V of kb-1 retired)
  uint64_t* p_full = bars + 20;    // [kTiles][2]  softmax -> MMA: P half of block kb is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kCtaQ;
  const int head = blockIdx.y;
  const int seq = blockIdx.z;
  const int sh = seq * heads + head;
  const int n_active = (q0 + kQTile < kT) ? 2 : 1;  // the last CTA of a sequence has one live tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    mbar_init(g_full, n_active);
    for (int t = 0; t < kTiles; ++t) mbar_init(&g_free[t], 8);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2 * n_active);  // S_lo and S_hi of every live tile
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2 * n_active);  // P*V of both halves of every live tile
    }
    for (int i = 0; i < kTiles * 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
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
      // ============================ TMA producer ============================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, n_active * kQBytes + kRelBytes);
        for (int t = 0; t < n_active; ++t) tma_load_3d(sQ + t * kQBytes, &tmap_q, q_full, 0, q0 + t * kQTile, sh);
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
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&v_full[st], kVBytes);
          tma_load_3d(sV + st * kVBytes, &tmap_vt, &v_full[st], kb * kKB, 0, sh);
          tma_load_3d(sV + st * kVBytes + 8192, &tmap_vt, &v_full[st], kb * kKB + 64, 0, sh);
        }
        __syncwarp();
      }
    } else if (warp - 1 < n_active) {
      // ============================ MMA issuer of tile t ============================
      const int t = warp - 1;
      constexpr uint32_t idesc_lo = umma_idesc_bf16(128, kHalfLo);
      constexpr uint32_t idesc_hi = umma_idesc_bf16(128, kKB - kHalfLo);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint32_t q_addr = smem_u32(sQ) + t * kQBytes;
      const uint32_t rel_addr = smem_u32(sRel);
      const uint32_t tm = tmem_base + t * kColsPerTile;

      mbar_wait(q_full, 0);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tm, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(rel_addr + k * 32), idesc_g,
                       k != 0);
        umma_commit(g_full);
      }
      __syncwarp();

      // S half h of block kb: Q * K[kb, keys of the half]^T into the half's S region (K block already in smem)
      auto issue_s = [&](int h, int kb) {
        const int st = kb % kStages;
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t k_addr = smem_u32(sK + st * kKBytes) + h * (kHalfLo * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm + h * kColSHi, umma_desc_sw128_kmajor(q_addr + k * 32),
                         umma_desc_sw128_kmajor(k_addr + k * 32), h ? idesc_hi : idesc_lo, k != 0);
          umma_commit(&s_full[2 * t + h]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      // O_h += P_h * V[kb, keys of the half]; P_h sits in place at the start of the half's S region
      auto issue_pv = [&](int h, int kb) {
        const int st = kb % kStages;
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t v_addr = smem_u32(sV + st * kVBytes) + h * 8192;
          const int nk = h ? 3 : 4;  // 64 or 48 keys
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nk)
              umma_bf16_ts(tm + kColO + h * 64, tm + h * kColSHi + k * 8, umma_desc_sw128_kmajor(v_addr + k * 32),
                           idesc_o, (kb | k) != 0);
          umma_commit(&v_empty[st]);
          if (kb == kNumKB - 1) umma_commit(&s_full[2 * t + h]);  // "last P*V of this half retired"
        }
        __syncwarp();
      };

      mbar_wait(&g_free[t], 0);
      mbar_wait(&k_full[0], 0);
      issue_s(0, 0);
      issue_s(1, 0);
      // Serve the two halves in arrival order, never blocking on one of them (the K / V rings are shared by all four
      // warpgroups of the CTA, so a blocking wait for one half's K block could depend on the other half's progress).
#ifdef BSEG_V5_DEBUG
      unsigned idle_polls = 0;
#endif
      int n_pv[2] = {0, 0};  // blocks whose P*V has been issued
      int n_s[2] = {1, 1};   // blocks whose S has been issued
      while (n_pv[0] < kNumKB || n_pv[1] < kNumKB) {
        bool progressed = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kb = n_pv[h];
          if (kb < kNumKB && n_s[h] > kb) {  // S(kb) issued: next is P*V(kb) once P and V are there
            const bool ready = mbar_test(&p_full[2 * t + h], kb & 1) &&
                               mbar_test(&v_full[kb % kStages], (kb / kStages) & 1);
            if (__all_sync(0xffffffffu, ready)) {
              issue_pv(h, kb);
              n_pv[h] = kb + 1;
              progressed = true;
            }
          }
          const int ks = n_s[h];
          if (ks < kNumKB && n_pv[h] == ks) {  // P*V(ks-1) issued: S(ks) may overwrite the region (in order behind it)
            if (__all_sync(0xffffffffu, mbar_test(&k_full[ks % kStages], (ks / kStages) & 1))) {
              issue_s(h, ks);
              n_s[h] = ks + 1;
              progressed = true;
            }
          }
        }
        if (!progressed) {
          __nanosleep(20);
#ifdef BSEG_V5_DEBUG
          if (++idle_polls == (1u << 22) && lane == 0)
            printf("v5 issuer stuck: block=(%d,%d,%d) t=%d n_pv=(%d,%d) n_s=(%d,%d) p=(%d,%d) v=(%d,%d) k=(%d,%d)\n",
                   blockIdx.x, blockIdx.y, blockIdx.z, t, n_pv[0], n_pv[1], n_s[0], n_s[1],
                   (int)mbar_test(&p_full[2 * t], n_pv[0] & 1), (int)mbar_test(&p_full[2 * t + 1], n_pv[1] & 1),
                   (int)mbar_test(&v_full[n_pv[0] % kStages], (n_pv[0] / kStages) & 1),
                   (int)mbar_test(&v_full[n_pv[1] % kStages], (n_pv[1] / kStages) & 1),
                   (int)mbar_test(&k_full[n_s[0] % kStages], (n_s[0] / kStages) & 1),
                   (int)mbar_test(&k_full[n_s[1] % kStages], (n_s[1] / kStages) & 1));
#endif
        } else {
#ifdef BSEG_V5_DEBUG
          idle_polls = 0;
#endif
        }
      }
    }
  } else {
    // ============================ softmax warpgroups ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    const int ws = warp - 4;
    const int t = ws >> 3;
    const int h = (ws >> 2) & 1;
    if (t < n_active) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      const int qi_raw = q0 + t * kQTile + r;
      const bool valid = qi_raw < kT;
      const int qi = valid ? qi_raw : kT - 1;
      const int qh = qi / kGridW, qw = qi % kGridW;
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * kColsPerTile;
      float* bh_row = reinterpret_cast<float*>(smem + kOffBh + t * kBhBytes) + r * kBhStride;
      // after the prologue MMA the rel region is the bw staging of both tiles
      float* stage = reinterpret_cast<float*>(sRel) + (t * kQTile + r) * kBwStride;

      // ---- prologue: decomposed rel-pos bias of this query (x log2 e): the lo warpgroup extracts bh (-> smem), the hi
      // warpgroup bw (-> staging); both then keep bw in registers ----
      mbar_wait(g_full, 0);
      tc_fence_after();
      if (h == 0) {
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
      } else {
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
      tc_fence_before();
      named_bar_sync(1 + t, 256);  // both warpgroups of the tile: rel tables in smem are dead, bh / bw are staged
      float bw[kGridW];
#pragma unroll
      for (int i = 0; i < kGridW; ++i) bw[i] = stage[i];
      named_bar_sync(1 + t, 256);  // staging fully read: the rows may be reused for the merge exchange
      if (lane == 0) mbar_arrive(&g_free[t]);

      const float sc = 0.125f * kLog2e;
      float m_run = 0.f, l_run = 0.f;
      float alpha_pending = 1.0f;  // factor still to be applied to O_h (after the P*V in flight retires)
      const uint32_t s_col = h * kColSHi;            // first S column of this half (P goes in place from here)
      const uint32_t o_col = kColO + h * 64;

      auto rescale_o = [&](float a) {
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          float v[16];
          tmem_ld16(lane_base + o_col + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= a;
          tmem_st16(lane_base + o_col + c, v);
        }
        tmem_st_wait();
      };

      for (int kb = 0; kb < kNumKB; ++kb) {
        float og[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) og[i] = lds32(bh_row + kb * 4 + i);
        mbar_wait(&s_full[2 * t + h], kb & 1);  // S half ready; the previous block's P*V (same issuer, in order) retired
        tc_fence_after();
        if (kb > 0 && __any_sync(0xffffffffu, alpha_pending != 1.0f)) rescale_o(alpha_pending);
        alpha_pending = 1.0f;
        if (kb == 0) {  // initial reference: row max over this half's keys of the first block
          float mx = -INFINITY;
          if (h == 0) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
              float v[16];
              tmem_ld16(lane_base + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaf(v[i], sc, bw[(c + i) % kGridW]) + og[(c + i) / kGridW]);
            }
          } else {
#pragma unroll
            for (int c = 64; c < 112; c += 16) {
              float v[16];
              tmem_ld16(lane_base + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fmaf(v[i], sc, bw[(c + i) % kGridW]) + og[(c + i) / kGridW]);
            }
          }
          m_run = mx;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) og[i] -= m_run;
        float lsum = 0.f, xmax = -INFINITY;

        // one chunk of N columns starting at compile-time column C0; `done` = packed P columns already stored for
        // this block (they are rescaled if this chunk forces the reference up)
        auto chunk = [&](auto c0_tag, auto n_tag, int done) {
          constexpr int C0 = decltype(c0_tag)::value, N = decltype(n_tag)::value;
          float x[32];
          if constexpr (N == 32) {
            tmem_ld32(lane_base + C0, x);
          } else {
            tmem_ld16(lane_base + C0, *reinterpret_cast<float(*)[16]>(&x[0]));
          }
          tmem_ld_wait();
          float cmax = -INFINITY;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            x[i] = fmaf(x[i], sc, bw[(C0 + i) % kGridW]) + og[(C0 + i) / kGridW];
            cmax = fmaxf(cmax, x[i]);
          }
          if (__any_sync(0xffffffffu, cmax > kOverflowGuard)) {  // (practically never) raise the reference right now
            const float up = fmaxf(cmax, 0.f);
            const float a = ex2_approx(-up);
            m_run += up;
            l_run *= a;
            lsum *= a;
            xmax -= up;
            cmax -= up;
#pragma unroll
            for (int i = 0; i < 4; ++i) og[i] -= up;
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] -= up;
            rescale_o(a);  // the previous P*V has retired (s_full), this block's has not been issued yet
            if (done > 0) {  // P columns of this block that are already in TMEM
              float v[16];
              tmem_ld16(lane_base + s_col, v);
              tmem_ld_wait();
              uint32_t* pv = reinterpret_cast<uint32_t*>(v);
#pragma unroll
              for (int i = 0; i < 16; ++i) pv[i] = scale_bf16x2_v5(pv[i], a);
              tmem_st16(lane_base + s_col, v);
              tmem_st_wait();
            }
          }
          xmax = fmaxf(xmax, cmax);
          uint32_t pk[N / 2];
#pragma unroll
          for (int i = 0; i < N; i += 2) {
            const float p0 = ex2_approx(x[i]), p1 = ex2_approx(x[i + 1]);
            lsum += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          if constexpr (N == 32) {
            tmem_st16u(lane_base + s_col + done, pk);
          } else {
            tmem_st8u(lane_base + s_col + done, pk);
          }
        };
        using I0 = std::integral_constant<int, 0>;
        using I16 = std::integral_constant<int, 16>;
        using I32 = std::integral_constant<int, 32>;
        using I64 = std::integral_constant<int, 64>;
        using I96 = std::integral_constant<int, 96>;
        if (h == 0) {
          chunk(I0{}, I32{}, 0);
          chunk(I32{}, I32{}, 16);
        } else {
          chunk(I64{}, I32{}, 0);
          chunk(I96{}, I16{}, 16);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + h]);
        l_run += lsum;
        // lazily raise the reference for the following blocks
        if (xmax > kLazyThreshold) {
          const float a = ex2_approx(-xmax);
          m_run += xmax;
          l_run *= a;
          alpha_pending = a;  // applied to O_h once this block's P*V has retired
        }
      }

      // ---- merge the two key halves of the row and write O / l (bf16, token-major [seq, t, heads*64]) ----
      mbar_wait(&s_full[2 * t + h], kNumKB & 1);  // the last P*V of this half retired
      tc_fence_after();
      float* exch = reinterpret_cast<float*>(smem + kOffExch) + ((t * kQTile + r) * 2) * kExchStride;  // [row][half]
      exch[h * kExchStride + 0] = m_run;
      exch[h * kExchStride + 1] = l_run;
      exch[h * kExchStride + 2] = alpha_pending;
      tc_fence_before();
      named_bar_sync(1 + t, 256);
      tc_fence_after();
      const float m_o = exch[(h ^ 1) * kExchStride + 0], l_o = exch[(h ^ 1) * kExchStride + 1],
                  a_o = exch[(h ^ 1) * kExchStride + 2];
      const float M = fmaxf(m_run, m_o);
      const float w_me = ex2_approx(m_run - M), w_ot = ex2_approx(m_o - M);
      const float denom = l_run * w_me + l_o * w_ot;
      const float inv = 1.0f / denom;
      const float f_me = alpha_pending * w_me * inv, f_ot = a_o * w_ot * inv;
      if (h == 0 && lse_out != nullptr && valid) lse_out[static_cast<long long>(sh) * kT + qi] = M + log2f(denom);
      // this thread writes head-dim columns [32h, 32h+32): needs both partial accumulators of those columns
      {
        float mine[32], other[32];
        tmem_ld32(lane_base + o_col + 32 * h, mine);
        tmem_ld32(lane_base + kColO + (h ^ 1) * 64 + 32 * h, other);
        tmem_ld_wait();
        if (valid) {
          __nv_bfloat16* dst = out + (static_cast<long long>(seq) * kT + qi) * (heads * 64) + head * 64 + 32 * h;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = mine[c + i] * f_me + other[c + i] * f_ot;
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                            pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<attn5::kTmemCols>(tmem_base);
  }
}

int launch_attention_v5(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                        const __nv_bfloat16* relcat, __nv_bfloat16* out, float* lse_out, int nseq, int heads,
                        cudaStream_t stream) {
  using namespace attn5;
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
  static bool attr_set = false;
  if (!attr_set) {
    BSEG_CHECK_CUDA(
        cudaFuncSetAttribute(attention_fwd_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  dim3 grid((kT + kCtaQ - 1) / kCtaQ, heads, nseq);
  ProfScope prof(CAT_ATTENTION, static_cast<double>(nseq) * heads * (4.0 * kT * kT * 64 + 2.0 * kT * 84 * 64),
                 static_cast<double>(nseq) * heads * kT * 64 * 2 * 4, stream);
  attention_fwd_v5_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, tr, out, lse_out, heads);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bseg
