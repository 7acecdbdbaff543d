// Fused SegGPT attention: one CTA per (sequence, head, 128-query tile), two CTAs per SM.
//     out = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v
// with the decomposed relative-position bias of modeling_seggpt.py:268-311 computed from the UNSCALED q
// (modeling_seggpt.py:324-329) and an fp32 softmax (:331).  The reference materialises a (16n,1568,1568) fp32 score
// tensor; here S, P and O live in TMEM and never touch HBM.
//
// Inputs: qs = bf16(q * head_dim^-0.5 * log2 e) (written so by the QKV GEMM epilogue), k, v^T, and relcat8 = 8 x the
// reversed rel-pos tables (8 = 1 / head_dim^-0.5, exact in bf16), so that qs.k and qs.relcat8 are the score and the
// bias in the log2 domain without any per-element scaling.
//
// The score tile arrives from the tensor core COMPLETE -- scaled, biased and relative to the running softmax reference
// -- so that a softmax thread does one MUFU.EX2, half a pack and half a packed add per element (round 1: 6.7
// instructions per element):
//     S = Qs K^T                       4 K-steps, bf16 operands from shared memory
//       + Ew Bw                        2 K-steps, fp16: Ew[q, kw] = width bias of query q (TMEM, A operand),
//                                                       Bw[kw, key] = [key % 28 == kw]   (constant one-hot, smem)
//       + Eh Bh                        1 K-step,  fp16: Eh[q, 0..3] = height bias of q for the block's 4 token rows,
//                                                       Eh[q, 4]    = -m (the row's softmax reference, a multiple of
//                                                       16, exact in fp16), Bh[j, key] = [key / 28 == j], Bh[4, :] = 1
// Key blocks are 112 keys = 4 rows of the 28-wide token grid (1568 = 14 * 112: no key masking).  The bias operands
// are fp16 (11 significant bits; they are products q.rel of bf16 factors) and the one-hot factors are exact.
//
//   warp 0      TMA producer (Q tile + rel tables once; K blocks and V^T blocks through two 2-stage rings)
//   warp 1      tcgen05 issuer (G = Qs relcat8^T once; per key block S, then O += P V)
//   warps 2-3   idle (complete the control warpgroup, which gives its registers away with setmaxnreg)
//   warps 4-7   softmax warpgroup (thread <-> query row == TMEM lane)
//
// Hand-offs: a thread copies its whole S row (112 fp32) to registers in one go and hands the S region straight back
// -- with the next block's bias row and reference already in Eh -- so that the next S MMA runs under ALL of this
// block's exponentials; a key block has two mbarrier waits (S full, previous P V retired; the second is probed while
// the exponentials run).  Measured alternatives that lost (profiles/r02_attn_fwd_timeline_exp*.txt, sources under
// tools/experiments/): strict MUFU turn taking between two warpgroups, a packed-FMA polynomial exp2 for part of the
// elements, handing S back later to pipeline the stores, two threads per query row (8 softmax warps per CTA).
//
// Streaming softmax against a lazily raised reference m: exact, because a stale reference only changes the common
// scale of P, l and O.  m is folded into the MMA (column 4 of Eh); a block whose MMA was issued before m was raised
// is processed through a slow path that adds the difference per element.  O in TMEM is rescaled only when m moved.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

// experiment hook of tools/micro (how sensitive is the kernel to the exponentials?); default = full work
#ifndef BSEG_ATTN_SKIP_EXP
#define BSEG_ATTN_SKIP_EXP 0
#endif

// Token grid GH x GW of the stacked (prompt over query) image and ROWS token rows per key block:
//   56 x 28, 4 rows  (896 x 448 px: the resized path, T = 1568, key blocks of 112)
//   64 x 32, 2 rows  (1024 x 512 px: native 512-px tiles, SURVEY 8(f) rank 4, T = 2048, key blocks of 64)
//  128 x 64, 1 row   (2048 x 1024 px: native 1024-px tiles, T = 8192, key blocks of 64; the rel tables (255 + 127 rows) do
//                     not fit TMEM's 256 columns at once, so G is computed in two passes)
template <int GH, int GW, int ROWS>
struct AttnCfg {
  static constexpr int kQTile = 128;            // queries per CTA
  static constexpr int kGridW = GW;
  static constexpr int kGridH = GH;
  static constexpr int kRowsPerKB = ROWS;       // token rows per key block
  static constexpr int kKB = ROWS * GW;         // keys per block: 112 | 64
  static constexpr int kT = GW * GH;            // 1568 | 2048
  static constexpr int kNumKB = kT / kKB;       // 14 | 32
  static constexpr int kStages = 2;
  static constexpr int kThreads = 256;
  static constexpr int kCtasPerSm = 2;
  static constexpr int kRegsControl = 40;
  static constexpr int kRegsSoftmax = 216;      // 2 * (128*40 + 128*216) = 65536
  static constexpr int kRelH = (2 * GH - 1 + 15) / 16 * 16;  // rows of the reversed rel_pos_h table: 111 -> 112 | 127 -> 128
  static constexpr int kRelW = (2 * GW - 1 + 15) / 16 * 16;  // rows of the reversed rel_pos_w table: 55 -> 64 | 63 -> 64
  static constexpr int kRelRows = kRelH + kRelW;             // 176 | 192
  static constexpr bool kTwoPassG = kRelRows > 256;          // G = Qs relcat8^T in two passes (height, then width)
  static constexpr int kEwSteps = (GW + 15) / 16;            // K-steps of the width-bias MMA: 2 | 2 | 4
  static constexpr int kEhBase = kEwSteps * 16;              // first K-column of the Eh step in the one-hot operand
  static constexpr int kOneHotAtoms = (kEhBase + 16 + 63) / 64;  // 128-byte-row tiles of the one-hot operand: 1 | 1 | 2
  static_assert(GW % 2 == 0 && ROWS <= 4 && kKB % 16 == 0 && kT % kKB == 0 && kRelH <= 256 && kRelW <= 256, "token grid");

  static constexpr int kQBytes = kQTile * 128;                  // 16384
  static constexpr int kKBytes = kKB * 128;                     // 14336 | 8192
  static constexpr int kVHalves = (kKB + 63) / 64;              // V^T arrives in 64-key halves: 2 | 1
  static constexpr int kVBytes = kVHalves * 64 * 128;           // 16384 | 8192
  static constexpr int kRelBytes = (kTwoPassG ? kRelH : kRelRows) * 128;  // 22528 | 24576 | 32768 (height table, then width)
  static constexpr int kOneHotBytes = kOneHotAtoms * kKB * 128; // one-hot B operand [keys][64 fp16] per atom
  static constexpr int kBhStride = GH / 2 + 2;                  // 32-bit words per row of the packed height-bias table (30 | 34)
  static constexpr int kBhBytes = kQTile * kBhStride * 4;
  static constexpr int kBwStride = GW % 32 == 0 && GW > 32 ? GW : GW + 1;
  static constexpr int kBwBytes = kQTile * kBwStride * 4;       // staging of the per-query width bias in the prologue

  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKBytes;
  static constexpr int kOffOneHot = kOffV + kStages * kVBytes;
  static constexpr int kOffBh = kOffOneHot + kOneHotBytes;
  // the rel tables, then the bw staging, overlay the (not yet used) K and V stages
  static constexpr int kOffRel = kOffK;
  static_assert(kRelBytes <= kStages * (kKBytes + kVBytes) && kBwBytes <= kStages * (kKBytes + kVBytes),
                "rel overlay does not fit in the K / V stages");
  static constexpr int kOffBar = (kOffBh + kBhBytes + 1023) / 1024 * 1024;
  static constexpr int kSmemBytes = kOffBar + 256 + 1024;
  static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kOffOneHot % 1024 == 0 && kKBytes % 1024 == 0, "swizzle alignment");
  static_assert(kSmemBytes * kCtasPerSm + 1024 * kCtasPerSm <= 228 * 1024, "shared memory budget");

  // TMEM columns (256 per CTA): S [0,KB)  O [KB,KB+64)  P (bf16 pairs, KB/2)  Ew 16  Eh 8 (fp16 pairs);
  // G = Qs relcat8^T (kRelRows columns) overlays them in the prologue
  static constexpr uint32_t kTmemCols = 256;
  static constexpr uint32_t kColO = kKB;
  static constexpr uint32_t kColP = kColO + 64;
  static constexpr uint32_t kColEw = kColP + kKB / 2;
  static constexpr uint32_t kColEh = kColEw + kEhBase / 2;
  static_assert(kColEh + 8 <= kTmemCols, "TMEM budget");
};

namespace attn {
constexpr float kRaiseThreshold = 65536.0f;     // raise the reference when a block's row sum exceeds 2^16
constexpr float kOverflowGuard = 1.0e30f;       // redo a half block whose row sum exceeds this (or is inf / nan)
constexpr float kMaxEncodedRef = 32768.0f;      // |m| that fp16 holds exactly in steps of 16
}  // namespace attn

// Optional timeline instrumentation (tools/micro/attn_trace.cu defines BSEG_ATTN_TRACE): clock64 stamps of one CTA.
#ifdef BSEG_ATTN_TRACE
__device__ long long g_attn_trace[3][16][16];  // [actor: softmax wg0, wg1, mma issuer 0][block][event]
#define ATTN_TRACE(actor, kb, ev)                                                              \
  do {                                                                                         \
    if (trace_cta && lane == 0) g_attn_trace[actor][kb][ev] = clock64();                       \
  } while (0)
#else
#define ATTN_TRACE(actor, kb, ev) do {} while (0)
#endif

namespace {
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}
// smallest multiple of 16 that is >= x (as a float)
__device__ __forceinline__ float ceil16(float x) { return 16.0f * ceilf(x * 0.0625f); }
// exponentiate columns [C0, C1) of a score row (kAdjust adds the -- rare -- reference correction per element): P as
// bf16 pairs, partial row sums in ls[4]
template <bool kAdjust, int C0, int C1, int KB>
__device__ __forceinline__ void exp_cols(const float (&x)[KB], float delta, float (&ls)[4], uint32_t (&pk)[KB / 2]) {
#pragma unroll
  for (int i = C0; i < C1; i += 4) {
    float x0 = x[i], x1 = x[i + 1], x2 = x[i + 2], x3 = x[i + 3];
    if (kAdjust) {
      x0 += delta; x1 += delta; x2 += delta; x3 += delta;
    }
    const float p0 = BSEG_ATTN_SKIP_EXP ? x0 * 0.001f : ex2_approx(x0);
    const float p1 = BSEG_ATTN_SKIP_EXP ? x1 * 0.001f : ex2_approx(x1);
    const float p2 = BSEG_ATTN_SKIP_EXP ? x2 * 0.001f : ex2_approx(x2);
    const float p3 = BSEG_ATTN_SKIP_EXP ? x3 * 0.001f : ex2_approx(x3);
    add_f32x2(ls[0], ls[1], p0, p1);
    add_f32x2(ls[2], ls[3], p2, p3);
    pk[i >> 1] = pack_bf16x2(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
  }
}
template <int KB>
__device__ __forceinline__ float row_max(const float (&x)[KB]) {
  float mx = x[0];
#pragma unroll
  for (int i = 1; i < KB; ++i) mx = fmaxf(mx, x[i]);
  return mx;
}
// packed fp16 height bias of key block kb from a row of the table: ROWS values (two words for 4 rows, one for 2)
template <int ROWS>
__device__ __forceinline__ void bh_block(const uint32_t* bh_row, int kb, uint32_t& w0, uint32_t& w1) {
  static_assert(ROWS == 4 || ROWS == 2 || ROWS == 1, "token rows per key block");
  if (ROWS == 4) {
    const uint2 g = *reinterpret_cast<const uint2*>(bh_row + 2 * kb);
    w0 = g.x;
    w1 = g.y;
  } else if (ROWS == 2) {
    w0 = bh_row[kb];
    w1 = 0u;
  } else {
    w0 = reinterpret_cast<const unsigned short*>(bh_row)[kb];  // (upper half zero)
    w1 = 0u;
  }
}
}  // namespace

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kCtasPerSm)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_rel,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, int heads) {
  using namespace attn;
  pdl_launch_dependents();
  constexpr int kQTile = Cfg::kQTile, kGridW = Cfg::kGridW, kGridH = Cfg::kGridH, kRows = Cfg::kRowsPerKB;
  constexpr int kKB = Cfg::kKB, kT = Cfg::kT, kNumKB = Cfg::kNumKB, kStages = Cfg::kStages;
  constexpr int kRelH = Cfg::kRelH, kRelW = Cfg::kRelW, kRelRows = Cfg::kRelRows;
  constexpr int kKBytes = Cfg::kKBytes, kVBytes = Cfg::kVBytes;
  constexpr uint32_t kColO = Cfg::kColO, kColP = Cfg::kColP, kColEw = Cfg::kColEw, kColEh = Cfg::kColEh;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Cfg::kOffQ;
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sV = smem + Cfg::kOffV;
  uint8_t* sRel = smem + Cfg::kOffRel;
  uint8_t* sOneHot = smem + Cfg::kOffOneHot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* k_full = bars + 2;     // [2]
  uint64_t* k_empty = bars + 4;    // [2]
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_empty = bars + 8;    // [2]
  uint64_t* s_full = bars + 10;    // MMA -> softmax: S_j is in TMEM
  uint64_t* s_free = bars + 11;    // softmax -> MMA: the S region may be overwritten (and Eh is set)
  uint64_t* p_full = bars + 12;    // softmax -> MMA: P_j is in TMEM (and O rescaled if it had to be)
  uint64_t* pv_done = bars + 13;   // MMA -> softmax: O += P_j V_j retired (P region free, O stable)
  uint64_t* rel_free = bars + 14;  // softmax -> TMA: the rel / bw staging region is dead (it overlays the K / V stages)
  uint64_t* gh_done = bars + 15;   // (two-pass G) softmax -> MMA: the height part of G has been read out of TMEM
  uint64_t* q2_full = bars + 16;   // (two-pass G) TMA -> MMA: the width rel table is in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int seq = blockIdx.z;
  const int sh = seq * heads + head;
#ifdef BSEG_ATTN_TRACE
  const bool trace_cta = blockIdx.x == 2 && blockIdx.y == 5 && blockIdx.z == gridDim.z / 2;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    mbar_init(g_full, 1);
    mbar_init(rel_free, 4);   // one arrive per softmax warp
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);     // one arrive per softmax warp
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    mbar_init(gh_done, 4);
    mbar_init(q2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // One-hot B operand of the bias MMAs, K-major rows of 128 B with the 128-byte swizzle (one tile of kKB rows per 64
  // K-columns): row n = key n of a block, fp16 columns 0..GW-1 = [n % GW == column], kEhBase..kEhBase+ROWS-1 =
  // [n / GW == column - kEhBase], kEhBase + 4 = 1 (the -m column), rest 0.
  constexpr int kEhBase = Cfg::kEhBase;
  for (int idx = threadIdx.x; idx < Cfg::kOneHotAtoms * kKB * 8; idx += Cfg::kThreads) {
    const int atom = idx / (kKB * 8), n = (idx >> 3) % kKB, c = idx & 7;
    const int kw = n % kGridW, j = n / kGridW;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pair = 0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int col = 64 * atom + 8 * c + 2 * e + hf;
        const bool one = (col < kEhBase) ? (col < kGridW && col == kw)
                                         : (col < kEhBase + 4) ? (col - kEhBase == j) : (col == kEhBase + 4);
        if (one) pair |= 0x3C00u << (16 * hf);
      }
      w[e] = pair;
    }
    *reinterpret_cast<uint4*>(sOneHot + atom * (kKB * 128) + n * 128 + ((c ^ (n & 7)) << 4)) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_proxy_async_smem();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  pdl_wait();  // (barriers, TMEM and the one-hot operands above were set up under the tail of the QKV GEMM)

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegsControl));
    if (warp == 0) {
      // ============================ TMA producer (warp-uniform loop, one elected lane issues) ============================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, Cfg::kQBytes + Cfg::kRelBytes);
        tma_load_3d(sQ, &tmap_q, q_full, 0, q0, sh);
        tma_load_2d(sRel, &tmap_rel, q_full, 0, 0);  // all of relcat8, or its height part (two-pass G)
      }
      __syncwarp();
      if constexpr (Cfg::kTwoPassG) {
        mbar_wait(g_full, 0);  // the height MMAs have read their table: the width table may take its place
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(q2_full, Cfg::kRelBytes);  // (a full box: the rows past the table arrive as zeros)
          tma_load_2d(sRel, &tmap_rel, q2_full, 0, kRelH);
        }
        __syncwarp();
      }
      mbar_wait(rel_free, 0);  // the rel tables / bw staging overlay the K and V stages
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
#pragma unroll
          for (int hv = 0; hv < Cfg::kVHalves; ++hv)
            tma_load_3d(sV + st * kVBytes + hv * 8192, &tmap_vt, &v_full[st], kb * kKB + hv * 64, 0, sh);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ============================ MMA issuer ============================
      // The whole warp runs the (warp-uniform) loop and one elected lane issues, which keeps every tcgen05.mma operand
      // in uniform registers.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB);
      constexpr uint32_t idesc_e = umma_idesc_f16(128, kKB);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, Cfg::kTwoPassG ? kRelH : kRelRows);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t rel_addr = smem_u32(sRel);
      const uint32_t onehot_addr = smem_u32(sOneHot);
      const uint32_t tm = tmem_base;

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
      if constexpr (Cfg::kTwoPassG) {  // second pass: the width part of G into TMEM columns [0, kRelW)
        constexpr uint32_t idesc_gw = umma_idesc_bf16(128, kRelW);
        mbar_wait(q2_full, 0);
        mbar_wait(gh_done, 0);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(rel_addr + k * 32),
                         idesc_gw, k != 0);
          umma_commit(g_full);
        }
        __syncwarp();
      }

      auto issue_s = [&](int kb) {
        const int st = kb % kStages;
        mbar_wait(&k_full[st], (kb / kStages) & 1);
        ATTN_TRACE(2, kb, 0);  // K block in smem
        const uint32_t k_addr = smem_u32(sK + st * kKBytes);
        mbar_wait(s_free, kb & 1);
        ATTN_TRACE(2, kb, 1);  // S free -> issue
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                         k != 0);
          // + width bias (2 K-steps), + height bias and -m (1 K-step): fp16 A operands from TMEM, one-hot B from smem
#pragma unroll
          for (int k = 0; k <= Cfg::kEwSteps; ++k)  // (the last step is Eh)
            umma_bf16_ts(tm, tm + kColEw + k * 8,
                         umma_desc_sw128_kmajor(onehot_addr + (k >> 2) * (kKB * 128) + (k & 3) * 32), idesc_e, 1u);
          umma_commit(s_full);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };

      issue_s(0);
      for (int kb = 0; kb < kNumKB; ++kb) {
        if (kb + 1 < kNumKB) issue_s(kb + 1);
        const int st = kb % kStages;
        mbar_wait(&v_full[st], (kb / kStages) & 1);
        ATTN_TRACE(2, kb, 3);  // V block in smem
        mbar_wait(p_full, kb & 1);
        ATTN_TRACE(2, kb, 4);  // P full -> issue PV
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t v_addr = smem_u32(sV + st * kVBytes);
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k) {
            const uint32_t va = v_addr + (k >> 2) * 8192 + (k & 3) * 32;
            umma_bf16_ts(tm + kColO, tm + kColP + k * 8, umma_desc_sw128_kmajor(va), idesc_o, (kb | k) != 0);
          }
          umma_commit(pv_done);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ softmax warpgroup ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kRegsSoftmax));
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row in the tile == TMEM lane
    const int qi_raw = q0 + r;
    const bool valid = qi_raw < kT;
    const int qi = valid ? qi_raw : kT - 1;
    const int qh = qi / kGridW, qw = qi % kGridW;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t* bh_row = reinterpret_cast<uint32_t*>(smem + Cfg::kOffBh) + r * Cfg::kBhStride;
    float* stage = reinterpret_cast<float*>(sRel) + r * Cfg::kBwStride;

    // ---- prologue: decomposed rel-pos bias of this query (log2 domain), as fp16 MMA operands ----
    mbar_wait(g_full, 0);  // the G MMAs have retired: the rel tables in smem are dead, G is in TMEM
    tc_fence_after();
    {
      const int off_h = (kGridH - 1) - qh;  // bh[kh] = G[off_h + kh]
      __half* bh_half = reinterpret_cast<__half*>(bh_row);
#pragma unroll
      for (int c = 0; c < kRelH; c += 16) {
        float v[16];
        tmem_ld16(lane_base + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int kh = c + i - off_h;
          if (kh >= 0 && kh < kGridH) bh_half[kh] = __float2half_rn(v[i]);
        }
      }
      if constexpr (Cfg::kTwoPassG) {  // hand the G columns back: the width pass overwrites them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(gh_done);
        mbar_wait(g_full, 1);
        tc_fence_after();
      }
      constexpr int kGwCol0 = Cfg::kTwoPassG ? 0 : kRelH;  // first TMEM column of the width part of G
      const int off_w = (kGridW - 1) - qw;  // bw[kw] = G[kGwCol0 + off_w + kw]
#pragma unroll
      for (int c = 0; c < kRelW; c += 16) {
        float v[16];
        tmem_ld16(lane_base + kGwCol0 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int kw = c + i - off_w;
          if (kw >= 0 && kw < kGridW) stage[kw] = v[i];
        }
      }
    }
    {
      constexpr int kEwWords = Cfg::kEhBase / 2;  // 16 | 32 packed fp16 pairs
#pragma unroll
      for (int w0 = 0; w0 < kEwWords; w0 += 16) {
        uint32_t ew[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          ew[i] = (2 * (w0 + i) < kGridW) ? pack_f16x2(stage[2 * (w0 + i)], stage[2 * (w0 + i) + 1]) : 0u;
        tmem_st16u(lane_base + kColEw + w0, ew);
      }
      uint32_t g0, g1;
      bh_block<kRows>(bh_row, 0, g0, g1);
      tmem_st4u(lane_base + kColEh, g0, g1, 0u, 0u);       // height bias of key block 0, -m = 0
      tmem_st4u(lane_base + kColEh + 4, 0u, 0u, 0u, 0u);
    }
    tmem_st_wait();
    // the staging area is about to be overwritten by TMA (it overlays the K / V stages): order this thread's generic-proxy
    // accesses to it before the async-proxy writes that follow the rel_free hand-off
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {  // G consumed and E written: the S region is free for S_0
      mbar_arrive(s_free);
      mbar_arrive(rel_free);
    }

    float m_run = 0.f;           // the row's softmax reference (a multiple of 16)
    float m_in_next = 0.f;       // the reference that is in Eh for the NEXT S block to be issued
    float l_run = 0.f;
    float alpha_pending = 1.0f;  // factor still to be applied to O (after the P*V that is in flight retires)
    uint32_t pk[kKB / 2];        // P of the current block: bf16 pairs

    // raise the reference by `up` (>= 0, a multiple of 16): everything accumulated so far shrinks by 2^-up
    auto raise = [&](float up) {
      const float a = ex2_approx(-up);
      m_run += up;
      l_run *= a;
      alpha_pending *= a;
    };
    constexpr int kProbeAt = (kKB / 2 + 8) / 16 * 16;  // the P*V barrier is probed after this many exponentials

    float x[kKB];
    uint32_t gn0 = 0u, gn1 = 0u;
    bh_block<kRows>(bh_row, 1, gn0, gn1);  // packed height bias of key block 1
    for (int kb = 0; kb < kNumKB; ++kb) {
      const float m_in_s = m_in_next;  // the reference that was in Eh when THIS block's S was issued

      // ---------------- S row -> registers, S region (with the next block's Eh) straight back to the tensor core ----------------
      if (quarter == 0) ATTN_TRACE(0, kb, 0);  // block start
      mbar_wait(s_full, kb & 1);
      if (quarter == 0) ATTN_TRACE(0, kb, 1);  // S ready
      tc_fence_after();
#pragma unroll
      for (int c = 0; c + 32 <= kKB; c += 32) tmem_ld32(lane_base + c, *reinterpret_cast<float(*)[32]>(&x[c]));
      if constexpr (kKB % 32 == 16) tmem_ld16(lane_base + kKB - 16, *reinterpret_cast<float(*)[16]>(&x[kKB - 16]));
      tmem_ld_wait();
      if (kb == 0) m_run = ceil16(row_max(x));  // initial reference: row max over the first key block
      if (kb + 1 < kNumKB) {
        const float m_enc = fminf(fmaxf(m_run, -kMaxEncodedRef), kMaxEncodedRef);
        tmem_st4u(lane_base + kColEh, gn0, gn1, pack_f16x2(-m_enc, 0.f), 0u);
        m_in_next = m_enc;
        if (kb + 2 < kNumKB) bh_block<kRows>(bh_row, kb + 2, gn0, gn1);  // (used one block later)
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);  // the next block's S may be issued: it runs under this block's exponentials
      if (quarter == 0) ATTN_TRACE(0, kb, 2);  // S handed back

      // ---------------- exponentials ----------------
      float delta = m_in_s - m_run;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      const bool adjust = __any_sync(0xffffffffu, delta != 0.f);
      if (adjust) exp_cols<true, 0, kProbeAt>(x, delta, ls, pk);
      else exp_cols<false, 0, kProbeAt>(x, 0.f, ls, pk);
      // probe the barrier the end of the block needs now: the probe's latency runs under the remaining exponentials
      const bool pv_ready = kb > 0 ? mbar_test(pv_done, (kb - 1) & 1) : true;
      if (adjust) exp_cols<true, kProbeAt, kKB>(x, delta, ls, pk);
      else exp_cols<false, kProbeAt, kKB>(x, 0.f, ls, pk);
      float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (__any_sync(0xffffffffu, !(lsum < kOverflowGuard))) {  // (practically never) redo against a safe reference
        raise(fmaxf(ceil16(row_max(x) + delta), 0.f));
        delta = m_in_s - m_run;
        ls[0] = ls[1] = ls[2] = ls[3] = 0.f;
        exp_cols<true, 0, kKB>(x, delta, ls, pk);
        lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      }
      if (quarter == 0) ATTN_TRACE(0, kb, 3);  // exponentials done

      // ---------------- hand P to the tensor core ----------------
      if (kb > 0) {
        // the P region and O are ours again once the previous P*V has retired
        if (!__all_sync(0xffffffffu, pv_ready)) mbar_wait(pv_done, (kb - 1) & 1);
        if (quarter == 0) ATTN_TRACE(0, kb, 4);  // previous PV retired
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
      for (int c = 0; c + 16 <= kKB / 2; c += 16) tmem_st16u(lane_base + kColP + c, &pk[c]);
      if constexpr ((kKB / 2) % 16 == 8) tmem_st8u(lane_base + kColP + kKB / 2 - 8, &pk[kKB / 2 - 8]);
      l_run += lsum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (quarter == 0) ATTN_TRACE(0, kb, 5);  // P handed over

      // lazily raise the reference for the following blocks: row sum < 2^e and max P >= row sum / KB
      if (lsum > kRaiseThreshold) {
        const int e = ((__float_as_int(lsum) >> 23) & 0xff) - 126;
        raise(static_cast<float>((e + 15) & ~15));  // applied to O once this block's P*V has retired
      }
    }

    // ---- epilogue: O / l -> bf16, token-major [seq, t, heads*64] ----
    mbar_wait(pv_done, (kNumKB - 1) & 1);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <class Cfg>
static int launch_attention_t(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                              const __nv_bfloat16* relcat, __nv_bfloat16* out, float* lse_out, int nseq, int heads,
                              cudaStream_t stream) {
  constexpr int kT = Cfg::kT;
  CUtensorMap tq, tk, tv, tr;
  const uint64_t nsh = static_cast<uint64_t>(nseq) * heads;
  {
    uint64_t dims[3] = {64, static_cast<uint64_t>(kT), nsh};
    uint64_t strides[2] = {128, static_cast<uint64_t>(kT) * 128};
    uint32_t boxq[3] = {64, Cfg::kQTile, 1};
    uint32_t boxk[3] = {64, Cfg::kKB, 1};
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
    // one box = the whole table, or (two-pass G) its height part; the width part is a second box of kRelW rows
    int rc = make_tmap_bf16_2d(&tr, relcat, 64, Cfg::kRelRows, 64, 64, Cfg::kTwoPassG ? Cfg::kRelH : Cfg::kRelRows);
    if (rc) return rc;
  }
  auto kern = attention_fwd_kernel<Cfg>;
  static PerDeviceFlag attr_once;  // per template instantiation
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  dim3 grid((kT + Cfg::kQTile - 1) / Cfg::kQTile, heads, nseq);
  ProfScope prof(CAT_ATTENTION,
                 static_cast<double>(nseq) * heads * (4.0 * kT * kT * 64 + 2.0 * kT * (Cfg::kGridH + Cfg::kGridW) * 64),
                 static_cast<double>(nseq) * heads * kT * 64 * 2 * 4, stream);
  BSEG_CHECK_CUDA(launch_pdl(kern, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, tq, tk, tv, tr, out, lse_out, heads));
  count_launch();
  return 0;
}

int attention_relcat_rows(int grid_h, int grid_w) {
  return (2 * grid_h - 1 + 15) / 16 * 16 + (2 * grid_w - 1 + 15) / 16 * 16;
}

int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                     const __nv_bfloat16* relcat, __nv_bfloat16* out, float* lse_out, int nseq, int heads, int grid_h,
                     int grid_w, cudaStream_t stream) {
  BSEG_REQUIRE(nseq > 0 && heads > 0, "attention: empty problem");
  if (grid_h == 56 && grid_w == 28)
    return launch_attention_t<AttnCfg<56, 28, 4>>(q, k, vt, relcat, out, lse_out, nseq, heads, stream);
  if (grid_h == 64 && grid_w == 32)
    return launch_attention_t<AttnCfg<64, 32, 2>>(q, k, vt, relcat, out, lse_out, nseq, heads, stream);
  if (grid_h == 128 && grid_w == 64)
    return launch_attention_t<AttnCfg<128, 64, 1>>(q, k, vt, relcat, out, lse_out, nseq, heads, stream);
  BSEG_REQUIRE(false, "attention: token grid %dx%d is not built (56x28 = the 448-px path, 64x32 / 128x64 = native 512- / "
               "1024-px tiles)", grid_h, grid_w);
  return 0;
}

}  // namespace bseg
