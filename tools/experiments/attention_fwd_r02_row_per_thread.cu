// Fused SegGPT attention.  BSEG_ATTN_WG = 1 (default): one CTA per (sequence, head, 128 queries), two CTAs per SM;
// BSEG_ATTN_WG = 2: one CTA per 256 queries and SM with two softmax warpgroups sharing the K / V stages.
//     out = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v
// with the decomposed relative-position bias of modeling_seggpt.py:268-311 computed from the UNSCALED q
// (modeling_seggpt.py:324-329) and an fp32 softmax (:331).  The reference materialises a (16n,1568,1568) fp32 score
// tensor; here S, P and O live in TMEM and never touch HBM.
//
// Inputs: qs = bf16(q * head_dim^-0.5 * log2 e) (written so by the QKV GEMM epilogue), k, v^T, and relcat8 = 8 x the
// reversed rel-pos tables (8 = 1 / head_dim^-0.5, exact in bf16), so that qs.k and qs.relcat8 are the score and the
// bias in the log2 domain without any per-element scaling.
//
// Round-2 structure: the score tile arrives from the tensor core COMPLETE -- scaled, biased and already relative to
// the running softmax reference -- so that the softmax warps do one MUFU.EX2, half a pack and half a packed add per
// element (round 1: 6.7 instructions per element, issue-bound at 0.42 of the MUFU rate):
//     S = Qs K^T                       4 K-steps, bf16 operands from shared memory
//       + Ew Bw                        2 K-steps, fp16: Ew[q, kw] = width bias of query q (TMEM, A operand),
//                                                       Bw[kw, key] = [key % 28 == kw]   (constant one-hot, smem)
//       + Eh Bh                        1 K-step,  fp16: Eh[q, 0..3] = height bias of q for the block's 4 token rows,
//                                                       Eh[q, 4]    = -m (the row's softmax reference, a multiple of
//                                                       16, exact in fp16), Bh[j, key] = [key / 28 == j], Bh[4, :] = 1
// Key blocks are 112 keys = 4 rows of the 28-wide token grid (1568 = 14 * 112: no key masking).  The bias operands
// are fp16 (11 significant bits; they are products q.rel of bf16 factors) and the one-hot factors are exact.
//
//   warp 0      TMA producer (Q tiles + rel tables once; K blocks and V^T blocks through two rings shared by the tiles)
//   warps 1-2   tcgen05 issuers, one per softmax warpgroup (G = Qs relcat8^T once; per key block S, then O += P V),
//               each blocking on its own warpgroup's barriers only
//   warp 3      idle (completes the control warpgroup, which gives its registers away with setmaxnreg)
//   warps 4-7   softmax warpgroup 0 (thread <-> query row == TMEM lane), warps 8-11 softmax warpgroup 1 (WG = 2)
//
// Hand-offs (profiles/r02_attn_fwd_timeline_*): every element costs one MUFU.EX2 (16 /clk/SM), i.e. >= 896 MUFU cycles
// per warpgroup and key block; round 1 additionally exposed ~750 cycles of hand-off latency per block (five mbarrier
// waits at >= 100 cycles each, tcgen05.commit -> wake-up, the S MMA behind a half-block of exponentials), during which
// both co-resident warpgroups tend to wait at the same time (lock-step: 63 % MUFU utilisation).  Now a thread copies its
// whole S row (112 fp32) to registers in one go and hands the S region straight back -- with the next block's bias row
// and reference already in Eh -- so the next S MMA runs under ALL of this block's exponentials; a block has two waits
// (S full, previous P V retired) instead of five.
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

#ifndef BSEG_ATTN_WG
#define BSEG_ATTN_WG 1
#endif
// of every 4 element pairs of a row, this many take the polynomial exp2 (FMA pipe) instead of MUFU.EX2
#ifndef BSEG_ATTN_POLY
#define BSEG_ATTN_POLY 0
#endif
// 1: software-pipelined hand-offs (next S row fetched under the P stores, barriers probed early); 0: plain order
#ifndef BSEG_ATTN_PIPELINED
#define BSEG_ATTN_PIPELINED 0
#endif

namespace attn {
constexpr int kWG = BSEG_ATTN_WG;      // softmax warpgroups (query tiles) per CTA
constexpr int kQTile = 128;            // queries per softmax warpgroup
constexpr int kCtaQ = kWG * kQTile;
constexpr int kGridW = 28;
constexpr int kGridH = 56;
constexpr int kRowsPerKB = 4;          // token rows per key block
constexpr int kKB = kRowsPerKB * kGridW;  // 112 keys per block
constexpr int kT = kGridW * kGridH;    // 1568
constexpr int kNumKB = kT / kKB;       // 14
constexpr int kStages = kWG == 2 ? 3 : 2;
constexpr int kThreads = 128 + kWG * 128;  // 384 | 256
constexpr int kCtasPerSm = kWG == 2 ? 1 : 2;
constexpr int kRegsControl = kWG == 2 ? 56 : 40;
constexpr int kRegsSoftmax = kWG == 2 ? 224 : 216;  // 128*56 + 256*224 = 64512 | 2 * (128*40 + 128*216) = 65536
constexpr int kRelRows = 176;  // 112 (reversed rel_pos_h, 111 used) + 64 (reversed rel_pos_w, 55 used)

constexpr int kQBytes = kQTile * 128;        // 16384
constexpr int kKBytes = kKB * 128;           // 14336
constexpr int kVBytes = 2 * 64 * 128;        // 16384 (two 64-key halves)
constexpr int kRelBytes = kRelRows * 128;    // 22528
constexpr int kOneHotBytes = kKB * 128;      // 14336: one-hot B operand [112 keys][64 fp16], 48 columns used
constexpr int kBhStride = 30;                // 32-bit words per row of the packed height-bias table (15 x uint2: LDS.64
                                             // of 16 consecutive rows hits 32 distinct banks)
constexpr int kBhBytes = kQTile * kBhStride * 4;   // 15360
constexpr int kBwStride = 29;
constexpr int kBwBytes = kQTile * kBwStride * 4;   // staging of the per-query width bias in the prologue

constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kWG * kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffOneHot = kOffV + kStages * kVBytes;
constexpr int kOffBh = kOffOneHot + kOneHotBytes;
// the rel tables, then the bw staging, overlay the (not yet used) V stages
constexpr int kOffRel = kOffV;
static_assert(kRelBytes <= kStages * kVBytes && kWG * kBwBytes <= kStages * kVBytes,
              "rel overlay does not fit in the V stages");
constexpr int kOffBar = (kOffBh + kWG * kBhBytes + 1023) / 1024 * 1024;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kOffOneHot % 1024 == 0 && kKBytes % 1024 == 0, "swizzle alignment");
static_assert(kSmemBytes * kCtasPerSm + 1024 * kCtasPerSm <= 228 * 1024, "shared memory budget");

// TMEM columns (256 per warpgroup): S [0,112)  O [112,176)  P [176,232) (bf16 pairs)  Ew [232,248)  Eh [248,256) (fp16
// pairs); G = Qs relcat8^T (176 columns) overlays S and O in the prologue
constexpr uint32_t kColsPerWG = 256;
constexpr uint32_t kTmemCols = kColsPerWG * kWG;
constexpr uint32_t kColO = 112;
constexpr uint32_t kColP = 176;
constexpr uint32_t kColEw = 232;
constexpr uint32_t kColEh = 248;

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
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 adds in one instruction (sm_100 packed fp32)
__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(d));
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// 2^x for two elements on the FMA / ALU pipes instead of the MUFU (the MUFU.EX2 rate, 16 /clk/SM, is what bounds this
// kernel): Cody-Waite split x = j + f, j = round(x) via the 1.5*2^23 trick, f in [-0.5, 0.5], 2^f by a degree-3
// minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P: 3.9e-3), 2^j by adding j to the
// exponent field.  x is clamped to >= -126 (results below 2^-126 are flushed by the MUFU path as well).
__device__ __forceinline__ void exp2_poly_x2(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);
  const uint64_t nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
  const uint64_t x = pack_f32x2(x0, x1);
  const uint64_t t = fadd2(x, magic);                                   // low mantissa bits = round(x)
  const uint64_t jf = fadd2(t, nmagic);                                 // round(x) as a float
  const uint64_t f = fadd2(x, jf ^ 0x8000000080000000ull);              // x - round(x)
  uint64_t p = ffma2(pack_f32x2(0.0551716685f, 0.0551716685f), f, pack_f32x2(0.242611125f, 0.242611125f));
  p = ffma2(p, f, pack_f32x2(0.693260968f, 0.693260968f));
  p = ffma2(p, f, pack_f32x2(0.999928057f, 0.999928057f));
  float t0, t1, q0, q1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(p));
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// Instruction descriptor: A = B = fp16, D = fp32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}
// smallest multiple of 16 that is >= x (as a float)
__device__ __forceinline__ float ceil16(float x) { return 16.0f * ceilf(x * 0.0625f); }
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
  uint8_t* sOneHot = smem + kOffOneHot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* k_full = bars + 2;     // [3]
  uint64_t* k_empty = bars + 5;    // [3]
  uint64_t* v_full = bars + 8;     // [3]
  uint64_t* v_empty = bars + 11;   // [3]
  uint64_t* s_full = bars + 14;    // [kWG]     MMA -> softmax: S_j is in TMEM
  uint64_t* s_free = bars + 18;    // [kWG]     softmax -> MMA: the S region may be overwritten (and Eh is set)
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
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);   // one arrive per softmax warp
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  // One-hot B operand of the bias MMAs, K-major rows of 128 B with the 128-byte swizzle: row n = key n of a block,
  // fp16 columns 0..27 = [n % 28 == column], 32..35 = [n / 28 == column - 32], 36 = 1 (the -m column), rest 0.
  for (int idx = threadIdx.x; idx < kKB * 8; idx += kThreads) {
    const int n = idx >> 3, c = idx & 7;
    const int kw = n % kGridW, j = n / kGridW;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pair = 0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int col = 8 * c + 2 * e + hf;
        const bool one = (col < kGridW) ? (col == kw) : (col >= 32 && col < 36) ? (col - 32 == j) : (col == 36);
        if (one) pair |= 0x3C00u << (16 * hf);
      }
      w[e] = pair;
    }
    *reinterpret_cast<uint4*>(sOneHot + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_proxy_async_smem();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
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
        if (kb == 0) mbar_wait(rel_free, 0);
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
      // warpgroup arrives on them, so neither warpgroup ever waits for the other's turn.  The whole warp runs the
      // (warp-uniform) loop and one elected lane issues, which keeps every tcgen05.mma operand in uniform registers.
      const int w = warp - 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB);
      constexpr uint32_t idesc_e = umma_idesc_f16(128, kKB);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint32_t q_addr = smem_u32(sQ) + w * kQBytes;
      const uint32_t rel_addr = smem_u32(sRel);
      const uint32_t onehot_addr = smem_u32(sOneHot);
      const uint32_t tm = tmem_base + w * kColsPerWG;

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

      auto issue_s = [&](int kb) {
        const int st = kb % kStages;
        mbar_wait(&k_full[st], (kb / kStages) & 1);
        if (w == 0) ATTN_TRACE(2, kb, 0);  // K block in smem
        const uint32_t k_addr = smem_u32(sK + st * kKBytes);
        mbar_wait(&s_free[w], kb & 1);
        if (w == 0) ATTN_TRACE(2, kb, 1);  // S free -> issue
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                         k != 0);
          // + width bias (2 K-steps), + height bias and -m (1 K-step): fp16 A operands from TMEM, one-hot B from smem
#pragma unroll
          for (int k = 0; k < 3; ++k)
            umma_bf16_ts(tm, tm + kColEw + k * 8, umma_desc_sw128_kmajor(onehot_addr + k * 32), idesc_e, 1u);
          umma_commit(&s_full[w]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
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
          for (int k = 0; k < kKB / 16; ++k) {
            const uint32_t va = v_addr + (k >> 2) * 8192 + (k & 3) * 32;
            umma_bf16_ts(tm + kColO, tm + kColP + k * 8, umma_desc_sw128_kmajor(va), idesc_o, (kb | k) != 0);
          }
          umma_commit(&pv_done[w]);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
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
    uint32_t* bh_row = reinterpret_cast<uint32_t*>(smem + kOffBh + w * kBhBytes) + r * kBhStride;
    float* stage = reinterpret_cast<float*>(sRel + w * kBwBytes) + r * kBwStride;

    // ---- prologue: decomposed rel-pos bias of this query (log2 domain), as fp16 MMA operands ----
    mbar_wait(g_full, 0);  // every issuer's G MMAs have retired: the rel tables in smem are dead, G is in TMEM
    tc_fence_after();
    {
      const int off_h = 55 - qh;  // bh[kh] = G[off_h + kh]
      __half* bh_half = reinterpret_cast<__half*>(bh_row);
#pragma unroll
      for (int c = 0; c < 112; c += 16) {
        float v[16];
        tmem_ld16(lane_base + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int kh = c + i - off_h;
          if (kh >= 0 && kh < kGridH) bh_half[kh] = __float2half_rn(v[i]);
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
          if (kw >= 0 && kw < kGridW) stage[kw] = v[i];
        }
      }
    }
    {
      uint32_t ew[16];
#pragma unroll
      for (int i = 0; i < kGridW / 2; ++i) ew[i] = pack_f16x2(stage[2 * i], stage[2 * i + 1]);
      ew[14] = 0u;
      ew[15] = 0u;
      tmem_st16u(lane_base + kColEw, ew);
      const uint2 g0 = *reinterpret_cast<const uint2*>(bh_row);
      tmem_st4u(lane_base + kColEh, g0.x, g0.y, 0u, 0u);       // height bias of key block 0, -m = 0
      tmem_st4u(lane_base + kColEh + 4, 0u, 0u, 0u, 0u);
    }
    tmem_st_wait();
    // the staging area is about to be overwritten by TMA (it overlays the V stages): order this thread's generic-proxy
    // accesses to it before the async-proxy writes that follow the rel_free hand-off
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {  // G consumed and E written: the S region is free for S_0
      mbar_arrive(&s_free[w]);
      mbar_arrive(rel_free);
    }

    float m_run = 0.f;           // the row's softmax reference (a multiple of 16)
    float m_in_next = 0.f;       // the reference that is in Eh for the NEXT S block to be issued
    float l_run = 0.f;
    float alpha_pending = 1.0f;  // factor still to be applied to O (after the P*V that is in flight retires)
    uint32_t pk[kKB / 2];        // P of the current block: bf16 pairs

    // exponentiate columns [C0, C1) of the row (kAdjust adds the -- rare -- reference correction per element): P as bf16
    // pairs, partial row sums in ls[4]
    auto exp_cols = [&](auto adjust_tag, auto c0_tag, auto c1_tag, const float (&x)[kKB], float delta, float (&ls)[4]) {
      constexpr bool kAdjust = decltype(adjust_tag)::value;
      constexpr int C0 = decltype(c0_tag)::value, C1 = decltype(c1_tag)::value;
#pragma unroll
      for (int i = C0; i < C1; i += 4) {
        float x0 = x[i], x1 = x[i + 1], x2 = x[i + 2], x3 = x[i + 3];
        if constexpr (kAdjust) {
          x0 += delta; x1 += delta; x2 += delta; x3 += delta;
        }
        float p0, p1, p2, p3;
        // pairs (i/2) % 4 < BSEG_ATTN_POLY go to the FMA pipe
        if (((i >> 1) & 3) < BSEG_ATTN_POLY) {
          exp2_poly_x2(x0, x1, p0, p1);
        } else {
          p0 = BSEG_ATTN_SKIP_EXP ? x0 * 0.001f : ex2_approx(x0);
          p1 = BSEG_ATTN_SKIP_EXP ? x1 * 0.001f : ex2_approx(x1);
        }
        if ((((i >> 1) + 1) & 3) < BSEG_ATTN_POLY) {
          exp2_poly_x2(x2, x3, p2, p3);
        } else {
          p2 = BSEG_ATTN_SKIP_EXP ? x2 * 0.001f : ex2_approx(x2);
          p3 = BSEG_ATTN_SKIP_EXP ? x3 * 0.001f : ex2_approx(x3);
        }
        add_f32x2(ls[0], ls[1], p0, p1);
        add_f32x2(ls[2], ls[3], p2, p3);
        pk[i >> 1] = pack_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
      }
    };
    auto row_max = [&](const float (&x)[kKB]) {
      float mx = x[0];
#pragma unroll
      for (int i = 1; i < kKB; ++i) mx = fmaxf(mx, x[i]);
      return mx;
    };
    // raise the reference by `up` (>= 0, a multiple of 16): everything accumulated so far shrinks by 2^-up
    auto raise = [&](float up) {
      const float a = ex2_approx(-up);
      m_run += up;
      l_run *= a;
      alpha_pending *= a;
    };
    auto load_s_row = [&](float (&x)[kKB]) {  // asynchronous: tmem_ld_wait() before the first use
      tmem_ld32(lane_base, *reinterpret_cast<float(*)[32]>(&x[0]));
      tmem_ld32(lane_base + 32, *reinterpret_cast<float(*)[32]>(&x[32]));
      tmem_ld32(lane_base + 64, *reinterpret_cast<float(*)[32]>(&x[64]));
      tmem_ld16(lane_base + 96, *reinterpret_cast<float(*)[16]>(&x[96]));
    };
    using Fast = std::false_type;
    using Slow = std::true_type;
    using C0 = std::integral_constant<int, 0>;
    using C1 = std::integral_constant<int, 32>;   // the S region is handed back after this many exponentials
    using C2 = std::integral_constant<int, 80>;   // the barriers of the end of the block are probed here
    using C3 = std::integral_constant<int, kKB>;

#if BSEG_ATTN_PIPELINED
    // Software pipeline: every hand-off latency (mbarrier probe ~100-200 cycles, tcgen05.ld / st + wait ~100-150 cycles)
    // runs under exponentials of the same warp instead of in front of them -- the S row of block j+1 is fetched while
    // P_j is on its way to TMEM, Eh and P stores are waited for one stage later, barriers are probed ahead of time.
    float x[kKB];
    mbar_wait(&s_full[w], 0);
    tc_fence_after();
    load_s_row(x);
    tmem_ld_wait();

    for (int kb = 0; kb < kNumKB; ++kb) {
      const float m_in_s = m_in_next;  // the reference that was in Eh when THIS block's S was issued
      if (quarter == 0) ATTN_TRACE(w, kb, 0);  // block start: S row in registers
      if (kb == 0) m_run = ceil16(row_max(x));  // initial reference: row max over the first key block
      // ---------------- next block's bias row and reference into Eh (the store completes under the first exponentials) ----------------
      if (kb + 1 < kNumKB) {
        const uint2 gnext = *reinterpret_cast<const uint2*>(bh_row + 2 * (kb + 1));
        const float m_enc = fminf(fmaxf(m_run, -kMaxEncodedRef), kMaxEncodedRef);
        tmem_st4u(lane_base + kColEh, gnext.x, gnext.y, pack_f16x2(-m_enc, 0.f), 0u);
        m_in_next = m_enc;
      }
      float delta = m_in_s - m_run;
      const bool adjust = __any_sync(0xffffffffu, delta != 0.f);
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      if (adjust) exp_cols(Slow{}, C0{}, C1{}, x, delta, ls);
      else exp_cols(Fast{}, C0{}, C1{}, x, 0.f, ls);
      // ---------------- S region (with Eh) back to the tensor core: S_{j+1} runs under the rest of this block ----------------
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[w]);
      if (quarter == 0) ATTN_TRACE(w, kb, 1);  // S handed back
      if (adjust) exp_cols(Slow{}, C1{}, C2{}, x, delta, ls);
      else exp_cols(Fast{}, C1{}, C2{}, x, 0.f, ls);
      // probe the barriers the end of the block needs (results are consumed ~30 exponentials later)
      const bool pv_ready = kb > 0 ? mbar_test(&pv_done[w], (kb - 1) & 1) : true;
      if (adjust) exp_cols(Slow{}, C2{}, C3{}, x, delta, ls);
      else exp_cols(Fast{}, C2{}, C3{}, x, 0.f, ls);
      float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (__any_sync(0xffffffffu, !(lsum < kOverflowGuard))) {  // (practically never) redo against a safe reference
        raise(fmaxf(ceil16(row_max(x) + delta), 0.f));
        delta = m_in_s - m_run;
        ls[0] = ls[1] = ls[2] = ls[3] = 0.f;
        exp_cols(Slow{}, C0{}, C3{}, x, delta, ls);
        lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      }
      if (quarter == 0) ATTN_TRACE(w, kb, 2);  // exponentials done

      // ---------------- hand P to the tensor core ----------------
      if (kb > 0) {
        // the P region and O are ours again once the previous P*V has retired
        if (!__all_sync(0xffffffffu, pv_ready)) mbar_wait(&pv_done[w], (kb - 1) & 1);
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
      tmem_st16u(lane_base + kColP, &pk[0]);
      tmem_st16u(lane_base + kColP + 16, &pk[16]);
      tmem_st16u(lane_base + kColP + 32, &pk[32]);
      tmem_st8u(lane_base + kColP + 48, &pk[48]);
      if (quarter == 0) ATTN_TRACE(w, kb, 3);  // P stores issued
      // ---------------- fetch the next S row while the P stores complete ----------------
      if (kb + 1 < kNumKB) {
        mbar_wait(&s_full[w], (kb + 1) & 1);  // issued ~80 exponentials ago
        tc_fence_after();
        load_s_row(x);
      }
      if (quarter == 0) ATTN_TRACE(w, kb, 4);  // next S row requested
      l_run += lsum;
      // lazily raise the reference for the following blocks: row sum < 2^e and max P >= row sum / 112
      if (lsum > kRaiseThreshold) {
        const int e = ((__float_as_int(lsum) >> 23) & 0xff) - 126;
        raise(static_cast<float>((e + 15) & ~15));  // applied to O once this block's P*V has retired
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w]);
      tmem_ld_wait();
      if (quarter == 0) ATTN_TRACE(w, kb, 5);  // P handed over, next S row in registers
    }

#else
    float x[kKB];
    for (int kb = 0; kb < kNumKB; ++kb) {
      const float m_in_s = m_in_next;  // the reference that was in Eh when THIS block's S was issued

      // ---------------- S row -> registers, S region (with the next block's Eh) straight back to the tensor core ----------------
      if (quarter == 0) ATTN_TRACE(w, kb, 0);  // block start
      mbar_wait(&s_full[w], kb & 1);
      if (quarter == 0) ATTN_TRACE(w, kb, 1);  // S ready
      tc_fence_after();
      load_s_row(x);
      tmem_ld_wait();
      if (quarter == 0) ATTN_TRACE(w, kb, 6);  // S row in registers
      if (kb == 0) m_run = ceil16(row_max(x));  // initial reference: row max over the first key block
      if (kb + 1 < kNumKB) {
        const uint2 gnext = *reinterpret_cast<const uint2*>(bh_row + 2 * (kb + 1));
        const float m_enc = fminf(fmaxf(m_run, -kMaxEncodedRef), kMaxEncodedRef);
        tmem_st4u(lane_base + kColEh, gnext.x, gnext.y, pack_f16x2(-m_enc, 0.f), 0u);
        m_in_next = m_enc;
        tmem_st_wait();
      }
      if (quarter == 0) ATTN_TRACE(w, kb, 7);  // Eh stored
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[w]);  // the next block's S may be issued: it runs under this block's exponentials
      if (quarter == 0) ATTN_TRACE(w, kb, 2);  // S handed back

      // ---------------- exponentials ----------------
      float delta = m_in_s - m_run;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      if (__any_sync(0xffffffffu, delta != 0.f)) exp_cols(Slow{}, C0{}, C3{}, x, delta, ls);
      else exp_cols(Fast{}, C0{}, C3{}, x, 0.f, ls);
      float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (__any_sync(0xffffffffu, !(lsum < kOverflowGuard))) {  // (practically never) redo against a safe reference
        raise(fmaxf(ceil16(row_max(x) + delta), 0.f));
        delta = m_in_s - m_run;
        ls[0] = ls[1] = ls[2] = ls[3] = 0.f;
        exp_cols(Slow{}, C0{}, C3{}, x, delta, ls);
        lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      }
      if (quarter == 0) ATTN_TRACE(w, kb, 3);  // exponentials done

      // ---------------- hand P to the tensor core ----------------
      if (kb > 0) {
        // the P region and O are ours again once the previous P*V has retired
        mbar_wait(&pv_done[w], (kb - 1) & 1);
        if (quarter == 0) ATTN_TRACE(w, kb, 4);  // previous PV retired
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
      tmem_st16u(lane_base + kColP, &pk[0]);
      tmem_st16u(lane_base + kColP + 16, &pk[16]);
      tmem_st16u(lane_base + kColP + 32, &pk[32]);
      tmem_st8u(lane_base + kColP + 48, &pk[48]);
      l_run += lsum;
      if (quarter == 0) ATTN_TRACE(w, kb, 8);  // P stores issued
      tmem_st_wait();
      if (quarter == 0) ATTN_TRACE(w, kb, 9);  // P stores complete
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w]);
      if (quarter == 0) ATTN_TRACE(w, kb, 5);  // P handed over

      // lazily raise the reference for the following blocks: row sum < 2^e and max P >= row sum / 112
      if (lsum > kRaiseThreshold) {
        const int e = ((__float_as_int(lsum) >> 23) & 0xff) - 126;
        raise(static_cast<float>((e + 15) & ~15));  // applied to O once this block's P*V has retired
      }
    }

#endif
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
    }  // w < n_active
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
