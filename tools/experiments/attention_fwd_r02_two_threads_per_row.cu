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
// The score tile arrives from the tensor core COMPLETE -- scaled, biased and relative to the running softmax reference:
//     S = Qs K^T                       4 K-steps, bf16 operands from shared memory
//       + Ew Bw                        2 K-steps, fp16: Ew[q, kw] = width bias of query q (TMEM, A operand),
//                                                       Bw[kw, key] = [key % 28 == kw]   (constant one-hot, smem)
//       + Eh Bh                        1 K-step,  fp16: Eh[q, 0..3] = height bias of q for the block's 4 token rows,
//                                                       Eh[q, 4]    = -m (the row's softmax reference, a multiple of
//                                                       16, exact in fp16), Bh[j, key] = [key / 28 == j], Bh[4, :] = 1
// Key blocks are 112 keys = 4 rows of the 28-wide token grid (1568 = 14 * 112: no key masking).  The bias operands
// are fp16 (11 significant bits; they are products q.rel of bf16 factors) and the one-hot factors are exact.  What is
// left per score element is one MUFU.EX2, half a maximum, half a pack and half a packed add.
//
//   warp 0      TMA producer (Q tile + rel tables once; K blocks and V^T blocks through two 2-stage rings)
//   warp 1      tcgen05 issuer (G = Qs relcat8^T once; per key block S, then O += P V)
//   warps 2-3   idle (complete the control warpgroup, which gives its registers away with setmaxnreg)
//   warps 4-11  softmax: TWO threads per query row (TMEM lane), warp 4+q -> columns [0,56) and warp 8+q -> columns
//               [56,112) of the rows of lane quarter q
//
// Why two threads per row (profiles/r02_attn_fwd_*): the kernel is bound by MUFU.EX2 (16 /clk/SM: >= 896 cycles per
// tile and key block), but a key block also carries ~1000 cycles of hand-off latency per softmax warp (mbarrier probes,
// tcgen05.ld / st + wait, fences), TMEM only has room for two tiles per SM, and ONE warp per scheduler cannot keep the
// MUFU busy on its own (12.8 instead of 8 cycles per element).  With one thread per row (round 1 and the first round-2
// kernel, tools/experiments/attention_fwd_r02_row_per_thread.cu) that left the MUFU 63-68 % busy.  Two threads per row
// put four softmax warps on every scheduler, so the unit stays busy while one tile is in its hand-offs.
//
// Exact online softmax: the two threads of a row exchange their half-row maxima through shared memory (one 64-thread
// named barrier per lane quarter and key block), so the reference m (a multiple of 16, never below the running row
// maximum: P <= 1) is raised BEFORE the block is exponentiated; both threads derive the same m from the same data.  m is
// folded into the next S by the tensor core (Eh column 4); only the block in which m moved adds the difference per
// element.  O in TMEM is rescaled (by the columns-[0,56) thread) only in those blocks.
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

// experiment hook of tools/micro (how sensitive is the kernel to the exponentials?); default = full work
#ifndef BSEG_ATTN_SKIP_EXP
#define BSEG_ATTN_SKIP_EXP 0
#endif

namespace attn {
constexpr int kQTile = 128;            // queries per CTA
constexpr int kGridW = 28;
constexpr int kGridH = 56;
constexpr int kRowsPerKB = 4;          // token rows per key block
constexpr int kKB = kRowsPerKB * kGridW;  // 112 keys per block
constexpr int kHalf = kKB / 2;         // 56 score columns per softmax thread
constexpr int kT = kGridW * kGridH;    // 1568
constexpr int kNumKB = kT / kKB;       // 14
constexpr int kStages = 2;
constexpr int kThreads = 128 + 256;    // control warpgroup + two softmax warpgroups
constexpr int kCtasPerSm = 2;
constexpr int kRegsControl = 24;
constexpr int kRegsSoftmax = 104;      // 128*24 + 256*104 = 29696 <= 384 * 80 (the CTA's register pool at launch: 80 per thread)
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
constexpr int kXchgBytes = 2 * kQTile * 4;         // half-row maxima / row sums exchanged between the two threads of a row

constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffOneHot = kOffV + kStages * kVBytes;
constexpr int kOffBh = kOffOneHot + kOneHotBytes;
constexpr int kOffXchg = kOffBh + kBhBytes;
// the rel tables, then the bw staging, overlay the (not yet used) V stages
constexpr int kOffRel = kOffV;
static_assert(kRelBytes <= kStages * kVBytes && kBwBytes <= kStages * kVBytes, "rel overlay does not fit in the V stages");
constexpr int kOffBar = (kOffXchg + kXchgBytes + 1023) / 1024 * 1024;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kOffOneHot % 1024 == 0 && kKBytes % 1024 == 0, "swizzle alignment");
static_assert(kSmemBytes * kCtasPerSm + 1024 * kCtasPerSm <= 228 * 1024, "shared memory budget");

// TMEM columns (256 per CTA): S [0,112)  O [112,176)  P [176,232) (bf16 pairs)  Ew [232,248)  Eh [248,256) (fp16 pairs)
// G = Qs relcat8^T (176 columns) overlays S and O in the prologue
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColO = 112;
constexpr uint32_t kColP = 176;
constexpr uint32_t kColEw = 232;
constexpr uint32_t kColEh = 248;

constexpr float kMaxEncodedRef = 32768.0f;      // |m| that fp16 holds exactly in steps of 16
}  // namespace attn

// Optional timeline instrumentation (tools/micro/attn_trace.cu defines BSEG_ATTN_TRACE): clock64 stamps of one CTA.
#ifdef BSEG_ATTN_TRACE
__device__ long long g_attn_trace[3][16][16];  // [actor: softmax columns [0,56), columns [56,112), mma issuer][block][event]
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
// Instruction descriptor: A = B = fp16, D = fp32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// smallest multiple of 16 that is >= x (as a float)
__device__ __forceinline__ float ceil16(float x) { return 16.0f * ceilf(x * 0.0625f); }
// named barrier among the two softmax warps of a lane quarter (ids 1..4)
__device__ __forceinline__ void pair_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
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
  uint8_t* sOneHot = smem + kOffOneHot;
  float* sXchg = reinterpret_cast<float*>(smem + kOffXchg);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
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
  uint64_t* rel_free = bars + 14;  // softmax -> TMA: the rel / bw staging region is dead (it overlays the V stages)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

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
    mbar_init(rel_free, 4);   // the four warps that stage the width bias
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);     // one arrive per softmax warp
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
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
        mbar_arrive_expect_tx(q_full, kQBytes + kRelBytes);
        tma_load_3d(sQ, &tmap_q, q_full, 0, q0, sh);
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
    } else if (warp == 1) {
      // ============================ MMA issuer ============================
      // The whole warp runs the (warp-uniform) loop and one elected lane issues, which keeps every tcgen05.mma operand
      // in uniform registers.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB);
      constexpr uint32_t idesc_e = umma_idesc_f16(128, kKB);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
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
          for (int k = 0; k < 3; ++k)
            umma_bf16_ts(tm, tm + kColEw + k * 8, umma_desc_sw128_kmajor(onehot_addr + k * 32), idesc_e, 1u);
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
    // ============================ softmax warps: two threads per query row ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    const int quarter = warp & 3;
    const int h = (warp - 4) >> 2;      // 0: score columns [0,56), also owns Eh / m / the O rescale;  1: columns [56,112)
    const int r = quarter * 32 + lane;  // query row in the tile == TMEM lane
    const int qi_raw = q0 + r;
    const bool valid = qi_raw < kT;
    const int qi = valid ? qi_raw : kT - 1;
    const int qh = qi / kGridW, qw = qi % kGridW;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_cols = lane_base + h * kHalf;
    uint32_t* bh_row = reinterpret_cast<uint32_t*>(smem + kOffBh) + r * kBhStride;
    float* my_slot = sXchg + h * kQTile + r;
    const float* peer_slot = sXchg + (1 - h) * kQTile + r;

    // ---- prologue: decomposed rel-pos bias of this query (log2 domain), as fp16 MMA operands.  The columns-[0,56)
    // thread of a row extracts the height bias (table in smem, one block at a time into Eh), its partner the width
    // bias (Ew, written once) ----
    mbar_wait(g_full, 0);  // the G MMAs have retired: the rel tables in smem are dead, G is in TMEM
    tc_fence_after();
    if (h == 0) {
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
      const uint2 g0 = *reinterpret_cast<const uint2*>(bh_row);
      tmem_st4u(lane_base + kColEh, g0.x, g0.y, 0u, 0u);       // height bias of key block 0, -m = 0
      tmem_st4u(lane_base + kColEh + 4, 0u, 0u, 0u, 0u);
    } else {
      float* stage = reinterpret_cast<float*>(sRel) + r * kBwStride;
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
      uint32_t ew[16];
#pragma unroll
      for (int i = 0; i < kGridW / 2; ++i) ew[i] = pack_f16x2(stage[2 * i], stage[2 * i + 1]);
      ew[14] = 0u;
      ew[15] = 0u;
      tmem_st16u(lane_base + kColEw, ew);
      // the staging area is about to be overwritten by TMA (it overlays the V stages): order this thread's generic-proxy
      // accesses to it before the async-proxy writes that follow the rel_free hand-off
      fence_proxy_async_smem();
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {  // G consumed and E written (both threads of every row): the S region is free for S_0
      mbar_arrive(s_free);
      if (h == 1) mbar_arrive(rel_free);
    }

    float m_run = 0.f;           // the row's softmax reference (a multiple of 16); both threads of a row keep the same value
    float m_in_next = 0.f;       // the reference that is in Eh for the NEXT S block to be issued
    float l_run = 0.f;           // this thread's share of the row sum
    float alpha_pending = 1.0f;  // (h == 0) factor still to be applied to O

    for (int kb = 0; kb < kNumKB; ++kb) {
      const float m_in_s = m_in_next;  // the reference that was in Eh when THIS block's S was issued

      // ---------------- my half of the S row -> registers; half-row maximum to my partner ----------------
      if (quarter == 0) ATTN_TRACE(h, kb, 0);  // block start
      mbar_wait(s_full, kb & 1);
      if (quarter == 0) ATTN_TRACE(h, kb, 1);  // S ready
      tc_fence_after();
      float x[kHalf];
      tmem_ld32(s_cols, *reinterpret_cast<float(*)[32]>(&x[0]));
      tmem_ld16(s_cols + 32, *reinterpret_cast<float(*)[16]>(&x[32]));
      tmem_ld8(s_cols + 48, *reinterpret_cast<float(*)[8]>(&x[48]));
      tmem_ld_wait();
      float mx = x[0];
#pragma unroll
      for (int i = 1; i < kHalf; ++i) mx = fmaxf(mx, x[i]);
      *my_slot = mx;
      pair_sync(quarter);
      mx = fmaxf(mx, *peer_slot);             // row maximum of this block, relative to m_in_s
      pair_sync(quarter);                     // (the slot is reused in the next block)
      // raise the reference when the block exceeds it: P <= 1 always, no overflow guard needed
      const float over = mx + (m_in_s - m_run);
      if (kb == 0) {
        m_run = ceil16(mx);                   // (m_in_s = 0 for the first block)
      } else if (over > 0.f) {
        const float up = ceil16(over);
        const float a = ex2_approx(-up);
        m_run += up;
        l_run *= a;
        alpha_pending *= a;
      }
      // ---------------- next block's bias row and the reference into Eh; S region back to the tensor core ----------------
      if (h == 0 && kb + 1 < kNumKB) {
        const uint2 gnext = *reinterpret_cast<const uint2*>(bh_row + 2 * (kb + 1));
        const float m_enc = fminf(fmaxf(m_run, -kMaxEncodedRef), kMaxEncodedRef);
        tmem_st4u(lane_base + kColEh, gnext.x, gnext.y, pack_f16x2(-m_enc, 0.f), 0u);
        tmem_st_wait();
      }
      m_in_next = fminf(fmaxf(m_run, -kMaxEncodedRef), kMaxEncodedRef);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);  // the next block's S may be issued: it runs under this block's exponentials
      if (quarter == 0) ATTN_TRACE(h, kb, 2);  // S handed back

      // ---------------- exponentials ----------------
      const float delta = m_in_s - m_run;  // != 0 only in a block that moved the reference
      uint32_t pk[kHalf / 2];
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
      if (__any_sync(0xffffffffu, delta != 0.f)) {
#pragma unroll
        for (int i = 0; i < kHalf; ++i) x[i] += delta;
      }
#pragma unroll
      for (int i = 0; i < kHalf; i += 4) {
        const float p0 = BSEG_ATTN_SKIP_EXP ? x[i] * 0.001f : ex2_approx(x[i]);
        const float p1 = BSEG_ATTN_SKIP_EXP ? x[i + 1] * 0.001f : ex2_approx(x[i + 1]);
        const float p2 = BSEG_ATTN_SKIP_EXP ? x[i + 2] * 0.001f : ex2_approx(x[i + 2]);
        const float p3 = BSEG_ATTN_SKIP_EXP ? x[i + 3] * 0.001f : ex2_approx(x[i + 3]);
        add_f32x2(ls0, ls1, p0, p1);
        add_f32x2(ls2, ls3, p2, p3);
        pk[i >> 1] = pack_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
      }
      l_run += (ls0 + ls1) + (ls2 + ls3);
      if (quarter == 0) ATTN_TRACE(h, kb, 3);  // exponentials done

      // ---------------- hand P to the tensor core ----------------
      if (kb > 0) {
        // the P region and O are ours again once the previous P*V has retired
        mbar_wait(pv_done, (kb - 1) & 1);
        if (quarter == 0) ATTN_TRACE(h, kb, 4);  // previous PV retired
        tc_fence_after();
        if (h == 0 && __any_sync(0xffffffffu, alpha_pending != 1.0f)) {
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
      const uint32_t p_cols = lane_base + kColP + h * (kHalf / 2);
      tmem_st16u(p_cols, &pk[0]);
      tmem_st8u(p_cols + 16, &pk[16]);
      tmem_st4u(p_cols + 24, pk[24], pk[25], pk[26], pk[27]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (quarter == 0) ATTN_TRACE(h, kb, 5);  // P handed over
    }

    // ---- epilogue: O / l -> bf16, token-major [seq, t, heads*64]; each thread of a row writes 32 of its 64 columns ----
    *my_slot = l_run;
    pair_sync(quarter);
    const float l_row = l_run + *peer_slot;
    mbar_wait(pv_done, (kNumKB - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_row;  // (alpha_pending is 1: a reference raised in the last block was applied before its P V)
    // log2-domain log-sum-exp of the row (saved for the backward pass): P = exp2(x - lse)
    if (lse_out != nullptr && valid && h == 0) lse_out[static_cast<long long>(sh) * kT + qi] = m_run + log2f(l_row);
    __nv_bfloat16* dst = out + (static_cast<long long>(seq) * kT + qi) * (heads * 64) + head * 64 + h * 32;
    float v[32];
    tmem_ld32(lane_base + kColO + h * 32, v);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int c = 0; c < 32; c += 8)
        *reinterpret_cast<uint4*>(dst + c) =
            make_uint4(pack_bf16x2(v[c] * inv, v[c + 1] * inv), pack_bf16x2(v[c + 2] * inv, v[c + 3] * inv),
                       pack_bf16x2(v[c + 4] * inv, v[c + 5] * inv), pack_bf16x2(v[c + 6] * inv, v[c + 7] * inv));
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
  dim3 grid((kT + kQTile - 1) / kQTile, heads, nseq);
  ProfScope prof(CAT_ATTENTION, static_cast<double>(nseq) * heads * (4.0 * kT * kT * 64 + 2.0 * kT * 84 * 64),
                 static_cast<double>(nseq) * heads * kT * 64 * 2 * 4, stream);
  attention_fwd_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, tr, out, lse_out, heads);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bseg
