// Fused SegGPT attention for one (sequence, head, 128-query tile) per CTA:
//     out = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v
// with the decomposed relative-position bias of modeling_seggpt.py:268-311 computed from the UNSCALED q
// (modeling_seggpt.py:324-329) and an fp32 softmax (:331).  The reference materialises a
// (16n,1568,1568) fp32 score tensor; here S tiles live in TMEM and never touch HBM.
//
//   warp 0      TMA producer   (Q tile + rel tables once, then 112-key K / V^T blocks through a 3-stage ring)
//   warp 1      tcgen05 issuer (G = Q*Rel^T once; per key block S = Q*K^T (N=112) and O_part = P*V (N=64))
//   warps 2-5   softmax        (thread <-> query row == TMEM lane; P goes back through swizzled smem)
//
// Key blocks are 112 keys = 4 rows of the 28-wide token grid, so a score column maps to (kh, kw) at compile
// time and 1568 = 14 * 112 needs no key masking.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

namespace attn {
constexpr int kQTile = 128;
constexpr int kKB = 112;         // keys per block
constexpr int kGridW = 28;       // token grid width
constexpr int kGridH = 56;
constexpr int kT = kGridW * kGridH;  // 1568
constexpr int kNumKB = kT / kKB;     // 14
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kRelRows = 176;  // 112 (reversed rel_pos_h, 111 used) + 64 (reversed rel_pos_w, 55 used)

constexpr int kQBytes = kQTile * 128;        // 16384
constexpr int kKBytes = kKB * 128;           // 14336
constexpr int kVBytes = 2 * 64 * 128;        // 16384 (two 64-key halves)
constexpr int kPBytes = 2 * kQTile * 128;    // 32768 (two 64-key atoms; second uses 48 keys)
constexpr int kRelBytes = kRelRows * 128;    // 22528 (lives in the P buffer before the main loop)
constexpr int kBhStride = 57;                // fp32 words per row (odd -> conflict free)
constexpr int kBhBytes = kQTile * kBhStride * 4;

constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffP = kOffV + kStages * kVBytes;
constexpr int kOffBh = kOffP + kPBytes;
constexpr int kOffBar = kOffBh + kBhBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;

// TMEM columns
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0;     // 2 x 128
constexpr uint32_t kColO = 256;   // 2 x 64
constexpr uint32_t kColG = 256;   // 176 columns, dead before the first P*V

constexpr float kLog2e = 1.4426950408889634f;
}  // namespace attn

__global__ void __launch_bounds__(attn::kThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_rel,
                     __nv_bfloat16* __restrict__ out, int heads) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  float* sBh = reinterpret_cast<float*>(smem + kOffBh);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* g_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* kv_full = bars + 3;             // [3]
  uint64_t* kv_empty = bars + 6;            // [3]
  uint64_t* s_full = bars + 9;              // [2]
  uint64_t* o_full = bars + 11;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x;
  const int head = blockIdx.y;
  const int seq = blockIdx.z;
  const int sh = seq * heads + head;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_rel);
    mbar_init(q_full, 1);
    mbar_init(g_full, 1);
    mbar_init(p_full, 128);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes + kRelBytes);
      tma_load_3d(sQ, &tmap_q, q_full, 0, qt * kQTile, sh);
      tma_load_2d(sP, &tmap_rel, q_full, 0, 0);
      for (int kb = 0; kb < kNumKB; ++kb) {
        const int st = kb % kStages;
        if (kb >= kStages) mbar_wait(&kv_empty[st], ((kb / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(sK + st * kKBytes, &tmap_k, &kv_full[st], 0, kb * kKB, sh);
        tma_load_3d(sV + st * kVBytes, &tmap_vt, &kv_full[st], kb * kKB, 0, sh);
        tma_load_3d(sV + st * kVBytes + 8192, &tmap_vt, &kv_full[st], kb * kKB + 64, 0, sh);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, kRelRows);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);

      auto issue_s = [&](int kb) {
        const int st = kb % kStages;
        mbar_wait(&kv_full[st], (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * kKBytes);
        const uint32_t d = tmem_base + kColS + (kb & 1) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(d, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit(&s_full[kb & 1]);
      };

      mbar_wait(q_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem_base + kColG, umma_desc_sw128_kmajor(q_addr + k * 32),
                     umma_desc_sw128_kmajor(p_addr + k * 32), idesc_g, k != 0);
      umma_commit(g_full);
      issue_s(0);
      issue_s(1);
      for (int kb = 0; kb < kNumKB; ++kb) {
        const int st = kb % kStages;
        mbar_wait(p_full, kb & 1);  // P_kb is in smem, S[kb&1] and O[kb&1] are free
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + st * kVBytes);
        const uint32_t d = tmem_base + kColO + (kb & 1) * 64;
#pragma unroll
        for (int k = 0; k < kKB / 16; ++k) {
          const uint32_t pa = p_addr + (k >> 2) * (kQTile * 128) + (k & 3) * 32;
          const uint32_t va = v_addr + (k >> 2) * 8192 + (k & 3) * 32;
          umma_bf16_ss(d, umma_desc_sw128_kmajor(pa), umma_desc_sw128_kmajor(va), idesc_o, k != 0);
        }
        umma_commit(&o_full[kb & 1]);
        umma_commit(&kv_empty[st]);
        if (kb + 2 < kNumKB) issue_s(kb + 2);
      }
    }
  } else {
    // ============================ softmax warps ============================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row in tile == TMEM lane
    const int qi_raw = qt * kQTile + r;
    const bool valid = qi_raw < kT;
    const int qi = valid ? qi_raw : kT - 1;
    const int qh = qi / kGridW, qw = qi % kGridW;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* bh_row = sBh + r * kBhStride;
    float* stage = reinterpret_cast<float*>(sP + r * 128);  // this thread's own P row (atom 0)

    // ---- prologue: decomposed rel-pos bias for this query, pre-multiplied by log2(e) ----
    mbar_wait(g_full, 0);
    tc_fence_after();
    {
      const int off_h = 55 - qh;  // bh[kh] = G[off_h + kh]
#pragma unroll
      for (int c = 0; c < 112; c += 16) {
        float v[16];
        tmem_ld16(lane_base + kColG + c, v);
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
        tmem_ld16(lane_base + kColG + 112 + c, v);
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

    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    const float sc = 0.125f * kLog2e;  // head_dim^-0.5 * log2(e)

    for (int kb = 0; kb < kNumKB; ++kb) {
      const uint32_t s_addr = lane_base + kColS + (kb & 1) * 128;
      mbar_wait(&s_full[kb & 1], (kb >> 1) & 1);
      tc_fence_after();
      float bh4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) bh4[i] = bh_row[kb * 4 + i];

      // pass 1: block max of y = s*scale*log2e + (bh + bw)*log2e
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKB; c += 16) {
        float v[16];
        tmem_ld16(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = c + i;
          mx = fmaxf(mx, fmaf(v[i], sc, bh4[col / kGridW] + bw[col % kGridW]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f(m_run - m_new);  // first block: exp2(-inf) = 0
      m_run = m_new;

      // the previous P*V must have finished reading the P buffer before it is overwritten
      if (kb > 0) {
        mbar_wait(&o_full[(kb - 1) & 1], ((kb - 1) >> 1) & 1);
        tc_fence_after();
      }

      // pass 2: p = exp2(y - m), row sum, bf16 P tile into 128B-swizzled smem
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < kKB; c += 16) {
        float v[16];
        tmem_ld16(s_addr + c, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int c0 = c + i, c1 = c + i + 1;
          const float p0 = exp2f(fmaf(v[i], sc, bh4[c0 / kGridW] + bw[c0 % kGridW]) - m_new);
          const float p1 = exp2f(fmaf(v[i + 1], sc, bh4[c1 / kGridW] + bw[c1 % kGridW]) - m_new);
          lsum += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // columns c..c+15 -> two 16B chunks of this row
        const int atom = c >> 6;
        const int ch = (c & 63) >> 3;
        uint8_t* rowp = sP + atom * (kQTile * 128) + r * 128;
        *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      l_run = l_run * alpha + lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);

      // fold the previous block's P*V into the running output while the tensor core works on this block
      if (kb > 0) {
        const uint32_t o_addr = lane_base + kColO + ((kb - 1) & 1) * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          float v[32];
          tmem_ld32(o_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha_prev, v[i]);
        }
      }
      alpha_prev = alpha;
    }
    {
      constexpr int last = kNumKB - 1;
      mbar_wait(&o_full[last & 1], (last >> 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = lane_base + kColO + (last & 1) * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(o_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha_prev, v[i]);
      }
    }
    if (valid) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = out + (static_cast<long long>(seq) * kT + qi) * (heads * 64) + head * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        uint4 pk = make_uint4(pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv),
                              pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv),
                              pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv),
                              pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv));
        *reinterpret_cast<uint4*>(dst + i) = pk;
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
                     const __nv_bfloat16* relcat, __nv_bfloat16* out, int nseq, int heads, int grid_h, int grid_w,
                     cudaStream_t stream) {
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
  static bool attr_set = false;
  if (!attr_set) {
    BSEG_CHECK_CUDA(
        cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  dim3 grid((kT + kQTile - 1) / kQTile, heads, nseq);
  ProfScope prof(CAT_ATTENTION, static_cast<double>(nseq) * heads * (4.0 * kT * kT * 64 + 2.0 * kT * 84 * 64),
                 static_cast<double>(nseq) * heads * kT * 64 * 2 * 4, stream);
  attention_fwd_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, tr, out, heads);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bseg
