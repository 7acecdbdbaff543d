// C ABI of libbseg.so (declared in include/bseg.h): weight packing, the SegGPT forward schedule and thin
// wrappers around the kernel launchers.
#include <list>
#include <new>
#include <vector>

#include "../../include/bseg.h"
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

using namespace bseg;

namespace {

constexpr int kT = BSEG_T;
constexpr int kD = BSEG_HIDDEN;
constexpr int kMlp = 4096;
constexpr int kDecN = 16384;
constexpr float kQScale = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e), see attention.cu

struct LayerPack {
  __nv_bfloat16 *qkv_w, *proj_w, *lin1_w, *lin2_w, *relcat;
  float *qkv_b, *proj_b, *lin1_b, *lin2_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  // transposed copies ([in, out], the B operand of the dgrad GEMMs) and relcat^T; packed by bseg_train_prepare
  __nv_bfloat16 *qkv_wt = nullptr, *proj_wt = nullptr, *lin1_wt = nullptr, *lin2_wt = nullptr;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// rows of output per CTA of the ingest / preprocess kernels: a full 16-row patch band when the source rows fit in smem
void ingest_geometry(int crop, int* band_out, int* max_rows_out) {
  const double scale = static_cast<double>(crop) / 448.0;
  const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
  int band = 16, max_rows = 0;
  for (;;) {
    max_rows = static_cast<int>(band * scale + 2 * support + 4);
    if (3ull * max_rows * (crop + 448) <= 160 * 1024 || band == 1) break;
    band /= 2;
  }
  *band_out = band;
  *max_rows_out = max_rows;
}

// reversed + concatenated rel-pos tables, times 8 (see attention.cu): rows 0..110 = 8 rel_pos_h[110-i], 111 = 0,
// rows 112..166 = 8 rel_pos_w[54-(i-112)], rest 0.  8 = 1 / head_dim^-0.5 (exact in bf16): the attention kernels take
// qs = q * head_dim^-0.5 * log2(e), so qs . (8 rel) is the bias q . rel in the log2 domain.
// General token grid gh x gw: rel_pos_h has 2 gh - 1 rows, padded to a multiple of 16 (rh_pad), rel_pos_w 2 gw - 1.
__global__ void pack_relcat_kernel(const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                                   __nv_bfloat16* __restrict__ out, int gh, int gw, int rh_pad) {
  const int i = blockIdx.x, d = threadIdx.x;  // rows x 64
  const int nh = 2 * gh - 1, nw = 2 * gw - 1;
  float v = 0.f;
  if (i < nh) v = rel_h[(nh - 1 - i) * 64 + d];
  else if (i >= rh_pad && i < rh_pad + nw) v = rel_w[(nw - 1 - (i - rh_pad)) * 64 + d];
  out[i * 64 + d] = __float2bfloat16_rn(8.0f * v);
}

// conv3x3 dgrad taps: w9b[tap'][ci][co] = w9[8 - tap'][co][ci]   (tap' = (2-ky)*3 + (2-kx): flipped kernel)
__global__ void pack_conv_w9_dgrad_kernel(const __nv_bfloat16* __restrict__ w9, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * 64 * 64) return;
  const int co = idx & 63, ci = (idx >> 6) & 63, t = idx >> 12;
  out[idx] = w9[((8 - t) * 64 + co) * 64 + ci];
}

// conv weight [out,in,3,3] -> [tap][out][in] bf16
__global__ void pack_conv_w9_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * 64 * 64) return;
  const int i = idx & 63, o = (idx >> 6) & 63, t = idx >> 12;
  out[idx] = __float2bfloat16_rn(w[((o * 64 + i) * 3 + t / 3) * 3 + (t % 3)]);
}

// Additive table of the patch-embedding GEMM epilogue (modeling_seggpt.py:163-206):
//   tab[s][t][d] = (s==1 && t in bottom half ? mask_token : conv bias) + segment_token_{input|prompt}
//                  + bicubic(pos_embed 14x14 -> 56x28, align_corners=False)[t] + type_token
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__global__ void embed_table_kernel(const float* __restrict__ pos /*[197,1024]*/, const float* __restrict__ patch_b,
                                   const float* __restrict__ mask_token, const float* __restrict__ seg_in,
                                   const float* __restrict__ seg_pr, const float* __restrict__ type_tok,
                                   float* __restrict__ tab, int gh, int gw) {
  const int t = blockIdx.x;  // 0..T-1
  const int T = gh * gw;
  const int ph = t / gw, pw = t % gw;
  const float A = -0.75f;
  // F.interpolate(..., size=(gh, gw), mode="bicubic", align_corners=False) from the 14 x 14 pre-training grid:
  // source coordinate = (dst + 0.5) * (14 / size) - 0.5   (0.25 and 0.5 for the 56 x 28 grid)
  const float sy = 14.0f / static_cast<float>(gh), sx = 14.0f / static_cast<float>(gw);
  const float ry = sy * (ph + 0.5f) - 0.5f, rx = sx * (pw + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
  const float ty = ry - fy, tx = rx - fx;
  const float cy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
  const float cx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
  for (int d = threadIdx.x; d < 1024; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), 13);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), 13);
        row += cx[b] * pos[(1 + yy * 14 + xx) * 1024 + d];
      }
      acc += cy[a] * row;
    }
    const float ty_ = type_tok[d];
    tab[(0ll * T + t) * 1024 + d] = ((patch_b[d] + seg_in[d]) + acc) + ty_;
    const float base1 = (ph < gh / 2) ? patch_b[d] : mask_token[d];
    tab[(1ll * T + t) * 1024 + d] = ((base1 + seg_pr[d]) + acc) + ty_;
  }
}

}  // namespace

struct bseg_handle {
  // token grid of the stacked (prompt over query) image: 56 x 28 for the 448-px path, 64 x 32 for native 512-px tiles
  int img = BSEG_IMG, gh = 56, gw = 28, T = BSEG_T, relcat_rows = 176;
  int num_layers = 0, merge_index = 0;
  int inter[4] = {0, 0, 0, 0};
  float eps = 1e-6f;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  __nv_bfloat16* patch_w = nullptr;
  float* embed_tab[2] = {nullptr, nullptr};  // instance, semantic
  std::vector<LayerPack> layers;
  float *enc_ln_w = nullptr, *enc_ln_b = nullptr;
  __nv_bfloat16* dec_embed_w = nullptr;
  float* dec_embed_b = nullptr;
  __nv_bfloat16* conv_w9 = nullptr;
  float *conv_b = nullptr, *dec_ln_w = nullptr, *dec_ln_b = nullptr, *head_w = nullptr, *head_b = nullptr;
  // training-only packs (bseg_train_prepare)
  void* train_arena = nullptr;
  __nv_bfloat16 *patch_wt = nullptr, *dec_embed_wt = nullptr, *conv_w9b = nullptr;
  // CUDA graphs of whole forwards (bseg_set_graph_batch_limit): a forward is ~190 launches whose CUtensorMaps are
  // re-encoded on the host every time; at small batch that host work is as long as the device work.  The second call
  // with the same arguments captures the launch sequence once, later calls replay it.
  struct GraphKey {
    const void *px, *ppx, *pm, *ws, *pred;
    int batch, emb, prompts, query_half, pairs;
    bool operator==(const GraphKey& o) const {
      return px == o.px && ppx == o.ppx && pm == o.pm && ws == o.ws && pred == o.pred && batch == o.batch &&
             emb == o.emb && prompts == o.prompts && query_half == o.query_half && pairs == o.pairs;
    }
  };
  struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;  // nullptr: seen once (eagerly), capture on the next call
    int launches;          // kernels in the graph (bseg_launch_count keeps counting replays)
  };
  std::list<GraphEntry> graphs;  // most recently used first
  int graph_batch_limit = 0;     // 0: never use graphs
  // fp32 accuracy mode (bseg_enable_fp32): the handle's own fp32 copy of the matrices
  void* f32_arena = nullptr;
  std::vector<F32Layer> f32_layers;
  F32Weights f32;
};

extern "C" {

const char* bseg_last_error(void) { return last_error_buf(); }
int bseg_version(void) { return 1; }
long long bseg_launch_count(void) { return launch_count(); }

int bseg_destroy(bseg_handle* h) {
  if (!h) return 0;
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->arena) cudaFree(h->arena);
  if (h->train_arena) cudaFree(h->train_arena);
  if (h->f32_arena) cudaFree(h->f32_arena);
  delete h;
  return 0;
}

int bseg_create(const bseg_weights* w, bseg_handle** out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(w != nullptr && out != nullptr, "bseg_create: null argument");
  BSEG_REQUIRE(w->num_layers > 0 && w->num_layers <= BSEG_MAX_LAYERS, "bseg_create: num_layers=%d", w->num_layers);
  BSEG_REQUIRE(w->merge_index >= 0 && w->merge_index < w->num_layers, "bseg_create: merge_index=%d", w->merge_index);
  for (int j = 0; j < 4; ++j)
    BSEG_REQUIRE(w->intermediate_indices[j] >= w->merge_index && w->intermediate_indices[j] < w->num_layers,
                 "bseg_create: intermediate index %d out of range", w->intermediate_indices[j]);
  const int img = w->image_size > 0 ? w->image_size : BSEG_IMG;
  BSEG_REQUIRE(img == 448 || img == 512 || img == 1024,
               "bseg_create: image_size=%d is not built (448 = the resized path, 512 / 1024 = native tiles)", img);
  bseg_handle* h = new (std::nothrow) bseg_handle();
  BSEG_REQUIRE(h != nullptr, "bseg_create: out of host memory");
  h->img = img;
  h->gw = img / 16;
  h->gh = 2 * h->gw;
  h->T = h->gh * h->gw;
  h->relcat_rows = attention_relcat_rows(h->gh, h->gw);
  const int kT = h->T;  // (shadows the 448-path constant in this function)
  h->num_layers = w->num_layers;
  h->merge_index = w->merge_index;
  for (int j = 0; j < 4; ++j) h->inter[j] = w->intermediate_indices[j];
  h->eps = w->layer_norm_eps;
  h->layers.resize(w->num_layers);

  // ---- carve one arena ----
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_patch = carve(1024ull * 768 * 2);
  const size_t o_tab0 = carve(2ull * kT * kD * 4), o_tab1 = carve(2ull * kT * kD * 4);
  struct LOff { size_t qkv_w, proj_w, lin1_w, lin2_w, relcat, qkv_b, proj_b, lin1_b, lin2_b, ln1_w, ln1_b, ln2_w, ln2_b; };
  std::vector<LOff> lo(w->num_layers);
  for (int i = 0; i < w->num_layers; ++i) {
    lo[i].qkv_w = carve(3072ull * 1024 * 2);
    lo[i].proj_w = carve(1024ull * 1024 * 2);
    lo[i].lin1_w = carve(4096ull * 1024 * 2);
    lo[i].lin2_w = carve(1024ull * 4096 * 2);
    lo[i].relcat = carve(static_cast<size_t>(h->relcat_rows) * 64 * 2);
    lo[i].qkv_b = carve(3072 * 4);
    lo[i].proj_b = carve(1024 * 4);
    lo[i].lin1_b = carve(4096 * 4);
    lo[i].lin2_b = carve(1024 * 4);
    lo[i].ln1_w = carve(1024 * 4);
    lo[i].ln1_b = carve(1024 * 4);
    lo[i].ln2_w = carve(1024 * 4);
    lo[i].ln2_b = carve(1024 * 4);
  }
  const size_t o_eln_w = carve(1024 * 4), o_eln_b = carve(1024 * 4);
  const size_t o_dew = carve(16384ull * 4096 * 2), o_deb = carve(16384 * 4);
  const size_t o_w9 = carve(9 * 64 * 64 * 2), o_cb = carve(64 * 4), o_dlw = carve(64 * 4), o_dlb = carve(64 * 4);
  const size_t o_hw = carve(192 * 4), o_hb = carve(16);
  h->arena_bytes = off;
  cudaError_t e = cudaMalloc(&h->arena, off);
  if (e != cudaSuccess) {
    set_error("bseg_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e));
    delete h;
    return -static_cast<int>(e);
  }
  uint8_t* base = static_cast<uint8_t*>(h->arena);
  auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(base + o); };
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(base + o); };
  int rc = 0;
  auto cvt = [&](const float* src, __nv_bfloat16* dst, long long n) {
    if (!rc) rc = launch_f32_to_bf16(src, dst, n, stream);
  };
  auto cpy = [&](const float* src, float* dst, size_t n) {
    if (!rc) {
      cudaError_t ce = cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, stream);
      if (ce != cudaSuccess) {
        set_error("bseg_create: copy failed: %s", cudaGetErrorString(ce));
        rc = -static_cast<int>(ce);
      }
    }
  };
  h->patch_w = bf(o_patch);
  cvt(w->patch_w, h->patch_w, 1024ll * 768);
  h->embed_tab[0] = fp(o_tab0);
  h->embed_tab[1] = fp(o_tab1);
  embed_table_kernel<<<kT, 256, 0, stream>>>(w->position_embeddings, w->patch_b, w->mask_token,
                                             w->segment_token_input, w->segment_token_prompt,
                                             w->type_token_instance, h->embed_tab[0], h->gh, h->gw);
  embed_table_kernel<<<kT, 256, 0, stream>>>(w->position_embeddings, w->patch_b, w->mask_token,
                                             w->segment_token_input, w->segment_token_prompt,
                                             w->type_token_semantic, h->embed_tab[1], h->gh, h->gw);
  for (int i = 0; i < w->num_layers; ++i) {
    const bseg_layer_weights& lw = w->layers[i];
    LayerPack& lp = h->layers[i];
    lp.qkv_w = bf(lo[i].qkv_w);   cvt(lw.qkv_w, lp.qkv_w, 3072ll * 1024);
    lp.proj_w = bf(lo[i].proj_w); cvt(lw.proj_w, lp.proj_w, 1024ll * 1024);
    lp.lin1_w = bf(lo[i].lin1_w); cvt(lw.lin1_w, lp.lin1_w, 4096ll * 1024);
    lp.lin2_w = bf(lo[i].lin2_w); cvt(lw.lin2_w, lp.lin2_w, 1024ll * 4096);
    lp.relcat = bf(lo[i].relcat);
    pack_relcat_kernel<<<h->relcat_rows, 64, 0, stream>>>(lw.rel_pos_h, lw.rel_pos_w, lp.relcat, h->gh, h->gw,
                                                          (2 * h->gh - 1 + 15) / 16 * 16);
    lp.qkv_b = fp(lo[i].qkv_b);   cpy(lw.qkv_b, lp.qkv_b, 3072);
    lp.proj_b = fp(lo[i].proj_b); cpy(lw.proj_b, lp.proj_b, 1024);
    lp.lin1_b = fp(lo[i].lin1_b); cpy(lw.lin1_b, lp.lin1_b, 4096);
    lp.lin2_b = fp(lo[i].lin2_b); cpy(lw.lin2_b, lp.lin2_b, 1024);
    lp.ln1_w = fp(lo[i].ln1_w);   cpy(lw.ln1_w, lp.ln1_w, 1024);
    lp.ln1_b = fp(lo[i].ln1_b);   cpy(lw.ln1_b, lp.ln1_b, 1024);
    lp.ln2_w = fp(lo[i].ln2_w);   cpy(lw.ln2_w, lp.ln2_w, 1024);
    lp.ln2_b = fp(lo[i].ln2_b);   cpy(lw.ln2_b, lp.ln2_b, 1024);
  }
  h->enc_ln_w = fp(o_eln_w); cpy(w->enc_ln_w, h->enc_ln_w, 1024);
  h->enc_ln_b = fp(o_eln_b); cpy(w->enc_ln_b, h->enc_ln_b, 1024);
  h->dec_embed_w = bf(o_dew); cvt(w->dec_embed_w, h->dec_embed_w, 16384ll * 4096);
  h->dec_embed_b = fp(o_deb); cpy(w->dec_embed_b, h->dec_embed_b, 16384);
  h->conv_w9 = bf(o_w9);
  pack_conv_w9_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, stream>>>(w->dec_conv_w, h->conv_w9);
  h->conv_b = fp(o_cb);   cpy(w->dec_conv_b, h->conv_b, 64);
  h->dec_ln_w = fp(o_dlw); cpy(w->dec_ln_w, h->dec_ln_w, 64);
  h->dec_ln_b = fp(o_dlb); cpy(w->dec_ln_b, h->dec_ln_b, 64);
  h->head_w = fp(o_hw);   cpy(w->dec_head_w, h->head_w, 192);
  h->head_b = fp(o_hb);   cpy(w->dec_head_b, h->head_b, 3);
  if (!rc) {
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
      set_error("bseg_create: pack kernels failed: %s", cudaGetErrorString(ce));
      rc = -static_cast<int>(ce);
    }
  }
  if (rc) {
    bseg_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

// ---- workspace layout (all offsets 1024-aligned) ----
namespace {
struct WsLayout {
  size_t h, xn, att, q, k, vt, mlp, inter, dec, ln_stats, total;
};
// Exchange buffer of the residual+LayerNorm GEMM epilogue (EPI_RESID_LN): eight tagged 64-bit (mean, M2) partials per
// row; every fused launch of a forward pass uses its own tag (launch index, < 255), the buffer is set to 0xFF bytes
// ("never written") at the start of the pass.
size_t ln_stats_bytes(size_t rows) { return rows * 8 * sizeof(unsigned long long); }
WsLayout ws_layout(int B, int T = kT) {
  WsLayout L;
  const size_t rows2 = 2ull * B * T, rows1 = 1ull * B * T;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  L.h = carve(rows2 * kD * 4);
  L.xn = carve(rows2 * kD * 2);
  L.att = carve(rows2 * kD * 2);
  L.q = carve(rows2 * kD * 2);
  L.k = carve(rows2 * kD * 2);
  L.vt = carve(rows2 * kD * 2);
  L.mlp = carve(rows2 * kMlp * 2);  // also: patch-embedding operand [rows2,768] bf16, ensemble scratch fp32 [rows2,1024]
  L.inter = carve(rows1 * 4096 * 2);
  L.dec = carve(rows1 * kDecN * 2);
  L.ln_stats = carve(ln_stats_bytes(rows2));
  L.total = off;
  return L;
}

// Training workspace: forward transients + everything the backward needs (saved activations) + backward scratch.
// Saved per layer (rows = 2*B*T for the two-stream layers, B*T afterwards): the fp32 residual stream after the
// attention block (h_mid) and after the MLP (h_out), q / k / v^T / attention output (bf16), the MLP pre-activation z
// (bf16) and the softmax log-sum-exp.  DESIGN.md section 3 has the byte counts.
struct TrainLayer {
  size_t h_mid, h_out, q, k, vt, att, z, lse;
};
struct TrainLayout {
  // forward transients
  size_t xn, mlp, inter, ln_stats;
  // saved
  size_t h_emb, dec;
  std::vector<TrainLayer> layers;
  // backward scratch
  size_t dh, dhb, dxn, dz, datt, v, dvec, bias, dqkv, dinter, ddec, dconv, dpatch;
  size_t total;
};
TrainLayout train_layout(const bseg_handle* h, int B) {
  TrainLayout L;
  const size_t rows2 = 2ull * B * kT, rows1 = 1ull * B * kT;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  L.xn = carve(rows2 * kD * 2);
  L.mlp = carve(rows2 * kMlp * 2);
  L.inter = carve(rows1 * 4096 * 2);
  L.ln_stats = carve(ln_stats_bytes(rows2));
  L.h_emb = carve(rows2 * kD * 4);
  L.dec = carve(rows1 * kDecN * 2);
  L.layers.resize(h->num_layers);
  for (int i = 0; i < h->num_layers; ++i) {
    const size_t rows = (i <= h->merge_index) ? rows2 : rows1;
    TrainLayer& t = L.layers[i];
    t.h_mid = carve(rows * kD * 4);
    t.h_out = carve(rows * kD * 4);
    t.q = carve(rows * kD * 2);
    t.k = carve(rows * kD * 2);
    t.vt = carve(rows * kD * 2);
    t.att = carve(rows * kD * 2);
    t.z = carve(rows * kMlp * 2);
    t.lse = carve(rows * BSEG_HEADS * 4);
  }
  L.dh = carve(rows1 * kD * 4);
  L.dhb = carve(rows1 * kD * 2);
  L.dxn = carve(rows1 * kD * 4);
  L.dz = carve(rows1 * kMlp * 2);
  L.datt = carve(rows1 * kD * 2);
  L.v = carve(rows1 * kD * 2);
  L.dvec = carve(rows1 * BSEG_HEADS * 4);
  L.bias = carve(rows1 * BSEG_HEADS * 84 * 4);
  L.dqkv = carve(rows1 * 3 * kD * 2);
  L.dinter = carve(rows1 * 4096 * 4);
  L.ddec = carve(rows1 * kDecN * 2);
  L.dconv = carve(static_cast<size_t>(B) * 448 * 448 * 64 * 2);
  L.dpatch = carve(rows1 * 768 * 4);
  L.total = off;
  return L;
}

// Pointers one forward pass works with.  Inference: one in-place residual stream and per-layer-reused buffers;
// training: every buffer the backward needs lives in its own slot of the training workspace.
struct FwdBufs {
  float* h_emb;                    // embeddings output == input of layer 0
  __nv_bfloat16 *xn, *mlp, *inter, *dec;
  unsigned long long* ln_stats = nullptr;  // EPI_RESID_LN exchange (see ws_layout)
  struct PerLayer {
    float *h_mid, *h_out;
    __nv_bfloat16 *q, *k, *vt, *att, *z;
    float* lse;
  };
  std::vector<PerLayer> layers;
};

int forward_impl(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                 const float* prompt_masks, int B, int embedding_type, int P, const FwdBufs& fb, float* pred_masks,
                 cudaStream_t stream, bool query_half_only = false) {
  int rc;
  const PdlScope pdl_scope(B <= kPdlMaxBatch);  // programmatic dependent launch for small forwards (host_utils.h)
  const int kT = h->T;  // tokens of the stacked image (shadows the 448-path constant: 1568, or 2048 for native 512-px tiles)
  // ---- embeddings: patchify + GEMM (modeling_seggpt.py:713-737, 163-206) ----
  __nv_bfloat16* a_patch = fb.mlp;
  if ((rc = launch_patchify(pixel_values, prompt_pixel_values, prompt_masks, nullptr, a_patch, B, h->img, stream)))
    return rc;
  {
    GemmEpiParams ep;
    ep.out = fb.h_emb;
    ep.ldc = kD;
    ep.tab = h->embed_tab[embedding_type == 0 ? 0 : 1];
    ep.rows_per_stream = B * kT;
    ep.T = kT;
    if ((rc = launch_gemm(EPI_EMBED, a_patch, 768, h->patch_w, 2ll * B * kT, kD, 768, ep, stream))) return rc;
  }

  // ---- encoder (modeling_seggpt.py:453-501) ----
  // Fused residual + LayerNorm (gemm_set_fused_ln): the lin2 GEMM (K = 4096: ~32000 MMA cycles per tile hide the longer
  // epilogue) also emits the NEXT layer's norm1 when the stream is not merged / ensembled in between (level >= 1, the
  // default: 22 of the 52 layernorm1024 launches of a 24-layer forward and their fp32 re-read of the residual stream
  // disappear); at level 2 the proj GEMM (K = 1024) also emits norm2 of its rows -- measured slower than the separate
  // launch (its epilogue is the bottleneck already), kept as a switch.
  const bool fused_ln = gemm_set_fused_ln(-1) != 0 && fb.ln_stats != nullptr && 2 * h->num_layers < 255;
  const bool fused_ln_proj = fused_ln && gemm_set_fused_ln(-1) >= 2;
  if (fused_ln) {
    cudaError_t ce = cudaMemsetAsync(fb.ln_stats, 0xFF, ln_stats_bytes(2ull * B * kT), stream);
    if (ce != cudaSuccess) {
      set_error("bseg_forward: memset failed: %s", cudaGetErrorString(ce));
      return -static_cast<int>(ce);
    }
  }
  auto fuse_ln = [&](GemmEpiParams& ep, const float* gamma, const float* beta, int launch_idx) {
    ep.ln_gamma = gamma; ep.ln_beta = beta; ep.ln_out = fb.xn; ep.ld_ln = kD; ep.ln_eps = h->eps;
    ep.ln_stats = fb.ln_stats;
    ep.ln_tag = static_cast<unsigned int>(launch_idx);
  };
  float* h_in = fb.h_emb;
  bool ln1_done = false;  // norm1 of this layer was written by the previous layer's lin2 epilogue
  for (int i = 0; i < h->num_layers; ++i) {
    const LayerPack& lp = h->layers[i];
    const FwdBufs::PerLayer& pl = fb.layers[i];
    const int nstreams = (i <= h->merge_index) ? 2 : 1;
    const int nseq = nstreams * B;
    const long long M = static_cast<long long>(nseq) * kT;
    if (!ln1_done)
      if ((rc = launch_layernorm1024(h_in, kD, lp.ln1_w, lp.ln1_b, fb.xn, kD, M, h->eps, stream))) return rc;
    ln1_done = false;
    {
      GemmEpiParams ep;
      ep.bias = lp.qkv_b;
      ep.q = pl.q; ep.k = pl.k; ep.vt = pl.vt;
      ep.T = kT; ep.heads = BSEG_HEADS;
      ep.q_scale = kQScale;  // q is kept as bf16(q * head_dim^-0.5 * log2 e): the score lands in the log2 domain
      if ((rc = launch_gemm(EPI_QKV, fb.xn, kD, lp.qkv_w, M, 3 * kD, kD, ep, stream))) return rc;
    }
    if ((rc = launch_attention(pl.q, pl.k, pl.vt, lp.relcat, pl.att, pl.lse, nseq, BSEG_HEADS, h->gh, h->gw, stream)))
      return rc;
    bool ens = false;
    if (P > 0) ens = (i == h->merge_index) ? true : (P >= 2);
    bool ln2_done = false;
    if (!ens) {
      GemmEpiParams ep;
      ep.out = pl.h_mid; ep.ldc = kD; ep.bias = lp.proj_b; ep.resid = h_in; ep.ldr = kD;
      if (fused_ln_proj) fuse_ln(ep, lp.ln2_w, lp.ln2_b, 2 * i);
      if ((rc = launch_gemm(fused_ln_proj ? EPI_RESID_LN : EPI_RESID_F32, pl.att, kD, lp.proj_w, M, kD, kD, ep, stream)))
        return rc;
      ln2_done = fused_ln_proj;
    } else {
      BSEG_REQUIRE(pl.h_mid == h_in, "feature ensemble is an inference-only path");
      float* tmp = reinterpret_cast<float*>(fb.mlp);
      GemmEpiParams ep;
      ep.out = tmp; ep.ldc = kD; ep.bias = lp.proj_b;
      if ((rc = launch_gemm(EPI_F32, pl.att, kD, lp.proj_w, M, kD, kD, ep, stream))) return rc;
      if ((rc = launch_ensemble_residual(h_in, tmp, nstreams, B / P, P, i == h->merge_index ? 1 : 0, kT, kD, stream)))
        return rc;
    }
    if (!ln2_done)
      if ((rc = launch_layernorm1024(pl.h_mid, kD, lp.ln2_w, lp.ln2_b, fb.xn, kD, M, h->eps, stream))) return rc;
    {
      GemmEpiParams ep;
      ep.out = fb.mlp; ep.ldc = kMlp; ep.bias = lp.lin1_b; ep.aux = pl.z;
      if ((rc = launch_gemm(EPI_BF16_GELU, fb.xn, kD, lp.lin1_w, M, kMlp, kD, ep, stream))) return rc;
    }
    {
      GemmEpiParams ep;
      ep.out = pl.h_out; ep.ldc = kD; ep.bias = lp.lin2_b; ep.resid = pl.h_mid; ep.ldr = kD;
      // the next layer's norm1 reads exactly these rows unless the two streams are merged first
      const bool fuse_next = fused_ln && i + 1 < h->num_layers && i != h->merge_index;
      if (fuse_next) fuse_ln(ep, h->layers[i + 1].ln1_w, h->layers[i + 1].ln1_b, 2 * i + 1);
      if ((rc = launch_gemm(fuse_next ? EPI_RESID_LN : EPI_RESID_F32, fb.mlp, kMlp, lp.lin2_w, M, kD, kMlp, ep, stream)))
        return rc;
      ln1_done = fuse_next;
    }
    if (i == h->merge_index)
      if ((rc = launch_merge_streams(pl.h_out, static_cast<long long>(B) * kT * kD, stream))) return rc;
    for (int j = 0; j < 4; ++j)
      if (h->inter[j] == i)
        if ((rc = launch_layernorm1024(pl.h_out, kD, h->enc_ln_w, h->enc_ln_b, fb.inter + j * kD, 4 * kD,
                                       static_cast<long long>(B) * kT, h->eps, stream)))
          return rc;
    h_in = pl.h_out;
  }

  // ---- decoder (modeling_seggpt.py:555-585) ----
  // query_half_only: the reference's predict / loss code reads pred_masks[:, :, 448:] only (src/model.py:158-160,48-57).
  // The conv3x3 at image row 448 needs row 447, so decoder_embed runs from token row 27 (tokens 756..1567) and the head
  // from image row 448; pred_masks rows < 448 are zero-filled.
  const int kTok0 = query_half_only ? (h->gh / 2 - 1) * h->gw : 0;
  {
    GemmEpiParams ep;
    ep.out = fb.dec; ep.bias = h->dec_embed_b; ep.T = kT; ep.grid_w = h->gw;
    GemmRows gr{kT, B, kTok0, kT - kTok0};
    if ((rc = launch_gemm_rows(EPI_PIXSHUF, fb.inter, 4 * kD, h->dec_embed_w, gr, kDecN, 4 * kD, ep, stream))) return rc;
  }
  if (query_half_only) {
    const size_t plane = 2ull * h->img * h->img, half = static_cast<size_t>(h->img) * h->img;
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < 3; ++c) {
        cudaError_t ce = cudaMemsetAsync(pred_masks + (static_cast<size_t>(b) * 3 + c) * plane, 0, half * 4, stream);
        if (ce != cudaSuccess) {
          set_error("bseg_forward: memset failed: %s", cudaGetErrorString(ce));
          return -static_cast<int>(ce);
        }
      }
  }
  return launch_decoder_head(fb.dec, h->conv_w9, h->conv_b, h->dec_ln_w, h->dec_ln_b, h->head_w, h->head_b, pred_masks,
                             B, 2 * h->img, h->img, h->eps, query_half_only ? h->img : 0, stream);
}

int check_forward_args(bseg_handle* h, int batch, int embedding_type, const void* workspace, const char* who) {
  BSEG_REQUIRE(h != nullptr, "%s: null handle", who);
  BSEG_REQUIRE(batch > 0, "%s: batch=%d", who, batch);
  BSEG_REQUIRE(embedding_type == 0 || embedding_type == 1,
               "Embedding type should be either 'semantic' or 'instance', but got code %d", embedding_type);
  BSEG_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "%s: workspace must be non-null and 256B aligned", who);
  return 0;
}

FwdBufs train_bufs(bseg_handle* h, const TrainLayout& L, uint8_t* ws) {
  FwdBufs fb;
  auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(ws + o); };
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  fb.h_emb = fp(L.h_emb);
  fb.xn = bf(L.xn); fb.mlp = bf(L.mlp); fb.inter = bf(L.inter); fb.dec = bf(L.dec);
  fb.ln_stats = reinterpret_cast<unsigned long long*>(ws + L.ln_stats);
  fb.layers.resize(h->num_layers);
  for (int i = 0; i < h->num_layers; ++i) {
    const TrainLayer& t = L.layers[i];
    fb.layers[i] = {fp(t.h_mid), fp(t.h_out), bf(t.q), bf(t.k), bf(t.vt), bf(t.att), bf(t.z), fp(t.lse)};
  }
  return fb;
}
}  // namespace

size_t bseg_workspace_bytes(const bseg_handle* h, int batch) {
  if (batch <= 0) return 0;
  return ws_layout(batch, h ? h->T : kT).total;
}

size_t bseg_train_workspace_bytes(const bseg_handle* h, int batch) {
  if (h == nullptr || batch <= 0) return 0;
  return train_layout(h, batch).total;
}

static int forward_eager(bool query_half_only, bseg_handle* h, const float* pixel_values,
                         const float* prompt_pixel_values, const float* prompt_masks, int batch, int embedding_type,
                         int ensemble_prompts, void* workspace, float* pred_masks, cudaStream_t stream) {
  const WsLayout L = ws_layout(batch, h->T);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(ws + o); };
  FwdBufs fb;
  fb.h_emb = reinterpret_cast<float*>(ws + L.h);
  fb.xn = bf(L.xn); fb.mlp = bf(L.mlp); fb.inter = bf(L.inter); fb.dec = bf(L.dec);
  fb.ln_stats = reinterpret_cast<unsigned long long*>(ws + L.ln_stats);
  fb.layers.assign(h->num_layers, {fb.h_emb, fb.h_emb, bf(L.q), bf(L.k), bf(L.vt), bf(L.att), nullptr, nullptr});
  return forward_impl(h, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type, ensemble_prompts, fb,
                      pred_masks, stream, query_half_only);
}

static int forward_entry(bool query_half_only, bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                 const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts, void* workspace,
                 size_t workspace_bytes, float* pred_masks, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_forward_args(h, batch, embedding_type, workspace, "bseg_forward");
  if (rc) return rc;
  BSEG_REQUIRE(ensemble_prompts >= 0 && (ensemble_prompts == 0 || batch % ensemble_prompts == 0),
               "bseg_forward: batch=%d is not a multiple of ensemble_prompts=%d", batch, ensemble_prompts);
  const WsLayout L = ws_layout(batch, h->T);
  BSEG_REQUIRE(workspace_bytes >= L.total, "bseg_forward: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  auto eager = [&]() {
    return forward_eager(query_half_only, h, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type,
                         ensemble_prompts, workspace, pred_masks, stream);
  };
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (batch > h->graph_batch_limit || prof_enabled() ||
      cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return eager();  // (per-launch event timing and a caller's own capture both need the plain launch sequence)
  const bseg_handle::GraphKey key{pixel_values, prompt_pixel_values, prompt_masks, workspace, pred_masks, batch,
                                  embedding_type, ensemble_prompts, query_half_only ? 1 : 0, gemm_set_cta_pairs(-1) | (gemm_set_small_tiles(-1) << 1) | (gemm_set_fused_ln(-1) << 2) | (pdl_set(-1) << 4)};
  for (auto it = h->graphs.begin(); it != h->graphs.end(); ++it) {
    if (!(it->key == key)) continue;
    h->graphs.splice(h->graphs.begin(), h->graphs, it);  // most recently used first
    bseg_handle::GraphEntry& e = h->graphs.front();
    if (e.exec == nullptr) {  // second call with these arguments: capture (one-time attributes were set by the first)
      const long long launches0 = launch_count();
      if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return eager();
      }
      rc = eager();
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
      if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->graphs.pop_front();
        return rc != 0 ? rc : eager();
      }
      const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
      cudaGraphDestroy(graph);
      e.launches = static_cast<int>(launch_count() - launches0);
      count_launch(-e.launches);  // nothing ran yet: the replay below accounts for the kernels
      if (ie != cudaSuccess) {
        cudaGetLastError();
        h->graphs.pop_front();
        return eager();
      }
    }
    BSEG_CHECK_CUDA(cudaGraphLaunch(e.exec, stream));
    count_launch(e.launches);
    return 0;
  }
  // first call with these arguments: run it, remember the key (bounded cache)
  h->graphs.push_front({key, nullptr, 0});
  if (h->graphs.size() > 16) {
    if (h->graphs.back().exec) cudaGraphExecDestroy(h->graphs.back().exec);
    h->graphs.pop_back();
  }
  return eager();
}

int bseg_set_graph_batch_limit(bseg_handle* h, int max_batch) {
  BSEG_REQUIRE(h != nullptr && max_batch >= 0, "bseg_set_graph_batch_limit: bad argument");
  const int prev = h->graph_batch_limit;
  h->graph_batch_limit = max_batch;
  return prev;
}

int bseg_forward(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                 const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts, void* workspace,
                 size_t workspace_bytes, float* pred_masks, void* stream) {
  return forward_entry(false, h, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type,
                       ensemble_prompts, workspace, workspace_bytes, pred_masks, stream);
}

int bseg_forward_query_half(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                            const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts,
                            void* workspace, size_t workspace_bytes, float* pred_masks, void* stream) {
  return forward_entry(true, h, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type,
                       ensemble_prompts, workspace, workspace_bytes, pred_masks, stream);
}

// ---- fp32 accuracy mode (precise.cu) ----
int bseg_enable_fp32(bseg_handle* h, const bseg_weights* w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(h != nullptr && w != nullptr, "bseg_enable_fp32: null argument");
  BSEG_REQUIRE(h->img == BSEG_IMG, "bseg_enable_fp32: the fp32 accuracy mode is built for the 448-px path only");
  BSEG_REQUIRE(w->num_layers == h->num_layers, "bseg_enable_fp32: weights have %d layers, the handle %d", w->num_layers,
               h->num_layers);
  if (h->f32_arena != nullptr) return 0;
  size_t off = 0;
  auto carve = [&](size_t floats) {
    size_t o = off;
    off = align_up(off + floats * 4, 256);
    return o;
  };
  struct LOff { size_t qkv, proj, lin1, lin2, rh, rw; };
  std::vector<LOff> lo(h->num_layers);
  for (int i = 0; i < h->num_layers; ++i)
    lo[i] = {carve(3072ull * 1024), carve(1024ull * 1024), carve(4096ull * 1024), carve(4096ull * 1024),
             carve(111 * 64), carve(55 * 64)};
  const size_t o_patch = carve(1024ull * 768), o_dec = carve(16384ull * 4096), o_conv = carve(64 * 64 * 9);
  void* arena = nullptr;
  BSEG_CHECK_CUDA(cudaMalloc(&arena, off));
  uint8_t* base = static_cast<uint8_t*>(arena);
  cudaError_t ce = cudaSuccess;
  auto cpy = [&](const float* src, size_t o, size_t floats) -> const float* {
    float* dst = reinterpret_cast<float*>(base + o);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, floats * 4, cudaMemcpyDeviceToDevice, stream);
    return dst;
  };
  h->f32_layers.resize(h->num_layers);
  for (int i = 0; i < h->num_layers; ++i) {
    const bseg_layer_weights& lw = w->layers[i];
    const LayerPack& lp = h->layers[i];
    F32Layer& f = h->f32_layers[i];
    f.qkv_w = cpy(lw.qkv_w, lo[i].qkv, 3072ull * 1024);
    f.proj_w = cpy(lw.proj_w, lo[i].proj, 1024ull * 1024);
    f.lin1_w = cpy(lw.lin1_w, lo[i].lin1, 4096ull * 1024);
    f.lin2_w = cpy(lw.lin2_w, lo[i].lin2, 4096ull * 1024);
    f.rel_pos_h = cpy(lw.rel_pos_h, lo[i].rh, 111 * 64);
    f.rel_pos_w = cpy(lw.rel_pos_w, lo[i].rw, 55 * 64);
    f.ln1_w = lp.ln1_w; f.ln1_b = lp.ln1_b; f.ln2_w = lp.ln2_w; f.ln2_b = lp.ln2_b;   // already fp32 in the arena
    f.qkv_b = lp.qkv_b; f.proj_b = lp.proj_b; f.lin1_b = lp.lin1_b; f.lin2_b = lp.lin2_b;
  }
  F32Weights& f = h->f32;
  f.num_layers = h->num_layers;
  f.merge_index = h->merge_index;
  for (int j = 0; j < 4; ++j) f.inter[j] = h->inter[j];
  f.eps = h->eps;
  f.patch_w = cpy(w->patch_w, o_patch, 1024ull * 768);
  f.embed_tab[0] = h->embed_tab[0];
  f.embed_tab[1] = h->embed_tab[1];
  f.layers = h->f32_layers.data();
  f.enc_ln_w = h->enc_ln_w; f.enc_ln_b = h->enc_ln_b;
  f.dec_embed_w = cpy(w->dec_embed_w, o_dec, 16384ull * 4096);
  f.dec_embed_b = h->dec_embed_b;
  f.dec_conv_w = cpy(w->dec_conv_w, o_conv, 64 * 64 * 9);
  f.dec_conv_b = h->conv_b; f.dec_ln_w = h->dec_ln_w; f.dec_ln_b = h->dec_ln_b;
  f.dec_head_w = h->head_w; f.dec_head_b = h->head_b;
  if (ce != cudaSuccess) {
    set_error("bseg_enable_fp32: copy failed: %s", cudaGetErrorString(ce));
    cudaFree(arena);
    return -static_cast<int>(ce);
  }
  h->f32_arena = arena;
  return 0;
}

size_t bseg_workspace_bytes_f32(const bseg_handle* /*h*/, int batch) {
  return batch > 0 ? f32_workspace_bytes(batch) : 0;
}

int bseg_forward_f32(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts, void* workspace,
                     size_t workspace_bytes, float* pred_masks, void* stream_) {
  int rc = check_forward_args(h, batch, embedding_type, workspace, "bseg_forward_f32");
  if (rc) return rc;
  BSEG_REQUIRE(h->f32_arena != nullptr, "bseg_forward_f32: call bseg_enable_fp32 first");
  BSEG_REQUIRE(ensemble_prompts >= 0 && (ensemble_prompts == 0 || batch % ensemble_prompts == 0),
               "bseg_forward_f32: batch=%d is not a multiple of ensemble_prompts=%d", batch, ensemble_prompts);
  BSEG_REQUIRE(workspace_bytes >= f32_workspace_bytes(batch), "bseg_forward_f32: workspace too small (%zu < %zu)",
               workspace_bytes, f32_workspace_bytes(batch));
  return forward_f32_impl(h->f32, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type,
                          ensemble_prompts, workspace, pred_masks, static_cast<cudaStream_t>(stream_));
}

size_t bseg_train_workspace_bytes_f32(const bseg_handle* h, int batch) {
  if (h == nullptr || batch <= 0 || h->f32_arena == nullptr) return 0;
  return f32_train_workspace_bytes(h->f32, batch);
}

int bseg_forward_train_f32(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                           const float* prompt_masks, int batch, int embedding_type, void* workspace,
                           size_t workspace_bytes, float* pred_masks, void* stream_) {
  int rc = check_forward_args(h, batch, embedding_type, workspace, "bseg_forward_train_f32");
  if (rc) return rc;
  BSEG_REQUIRE(h->f32_arena != nullptr, "bseg_forward_train_f32: call bseg_enable_fp32 first");
  const size_t need = f32_train_workspace_bytes(h->f32, batch);
  BSEG_REQUIRE(workspace_bytes >= need, "bseg_forward_train_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
  return forward_f32_train_impl(h->f32, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type, workspace,
                                pred_masks, static_cast<cudaStream_t>(stream_));
}

int bseg_backward_to_prompt_f32(bseg_handle* h, const float* d_pred_masks, int batch, void* workspace,
                                size_t workspace_bytes, float* d_prompt_pixel_values, void* stream_) {
  int rc = check_forward_args(h, batch, 0, workspace, "bseg_backward_to_prompt_f32");
  if (rc) return rc;
  BSEG_REQUIRE(h->f32_arena != nullptr, "bseg_backward_to_prompt_f32: call bseg_enable_fp32 first");
  BSEG_REQUIRE(d_pred_masks != nullptr && d_prompt_pixel_values != nullptr, "bseg_backward_to_prompt_f32: null argument");
  const size_t need = f32_train_workspace_bytes(h->f32, batch);
  BSEG_REQUIRE(workspace_bytes >= need, "bseg_backward_to_prompt_f32: workspace too small (%zu < %zu)", workspace_bytes,
               need);
  return backward_f32_impl(h->f32, d_pred_masks, batch, workspace, d_prompt_pixel_values,
                           static_cast<cudaStream_t>(stream_));
}

// ---- training: transposed weight packs, forward that keeps what the backward needs, backward to the prompt ----
int bseg_train_prepare(bseg_handle* h, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(h != nullptr, "bseg_train_prepare: null handle");
  BSEG_REQUIRE(h->img == BSEG_IMG, "bseg_train_prepare: the train step is built for the 448-px path only (the native-"
               "resolution mode is inference-only)");
  if (h->train_arena != nullptr) return 0;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  struct LOff { size_t qkv, proj, lin1, lin2, rel; };
  std::vector<LOff> lo(h->num_layers);
  for (int i = 0; i < h->num_layers; ++i) {
    lo[i].qkv = carve(3072ull * 1024 * 2);
    lo[i].proj = carve(1024ull * 1024 * 2);
    lo[i].lin1 = carve(4096ull * 1024 * 2);
    lo[i].lin2 = carve(4096ull * 1024 * 2);
  }
  const size_t o_patch = carve(768ull * 1024 * 2), o_dec = carve(16384ull * 4096 * 2), o_w9b = carve(9 * 64 * 64 * 2);
  void* arena = nullptr;
  BSEG_CHECK_CUDA(cudaMalloc(&arena, off));
  uint8_t* base = static_cast<uint8_t*>(arena);
  auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(base + o); };
  int rc = 0;
  // W [out, in] -> W^T [in, out]
  auto tr = [&](const __nv_bfloat16* w, __nv_bfloat16* wt, int out, int in) {
    if (!rc) rc = launch_transpose_bf16(w, wt, out, in, 1, 0, 0, in, out, stream);
  };
  for (int i = 0; i < h->num_layers && !rc; ++i) {
    LayerPack& lp = h->layers[i];
    lp.qkv_wt = bf(lo[i].qkv);   tr(lp.qkv_w, lp.qkv_wt, 3072, 1024);
    lp.proj_wt = bf(lo[i].proj); tr(lp.proj_w, lp.proj_wt, 1024, 1024);
    lp.lin1_wt = bf(lo[i].lin1); tr(lp.lin1_w, lp.lin1_wt, 4096, 1024);
    lp.lin2_wt = bf(lo[i].lin2); tr(lp.lin2_w, lp.lin2_wt, 1024, 4096);
  }
  h->patch_wt = bf(o_patch);   tr(h->patch_w, h->patch_wt, 1024, 768);
  h->dec_embed_wt = bf(o_dec); tr(h->dec_embed_w, h->dec_embed_wt, 16384, 4096);
  h->conv_w9b = bf(o_w9b);
  pack_conv_w9_dgrad_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, stream>>>(h->conv_w9, h->conv_w9b);
  if (!rc) {
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
      set_error("bseg_train_prepare: pack kernels failed: %s", cudaGetErrorString(ce));
      rc = -static_cast<int>(ce);
    }
  }
  if (rc) {
    cudaFree(arena);
    return rc;
  }
  h->train_arena = arena;
  return 0;
}

int bseg_forward_train(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                       const float* prompt_masks, int batch, int embedding_type, void* workspace,
                       size_t workspace_bytes, float* pred_masks, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_forward_args(h, batch, embedding_type, workspace, "bseg_forward_train");
  if (rc) return rc;
  const TrainLayout L = train_layout(h, batch);
  BSEG_REQUIRE(workspace_bytes >= L.total, "bseg_forward_train: workspace too small (%zu < %zu)", workspace_bytes,
               L.total);
  const FwdBufs fb = train_bufs(h, L, static_cast<uint8_t*>(workspace));
  return forward_impl(h, pixel_values, prompt_pixel_values, prompt_masks, batch, embedding_type, 0, fb, pred_masks,
                      stream);
}

namespace {
// dq, dk, dv of one layer: D = rowsum(dO * O) and v = vt^T, then the two attention-backward kernels
int attention_backward(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                       const __nv_bfloat16* att, const __nv_bfloat16* datt, const float* lse,
                       const __nv_bfloat16* relcat, __nv_bfloat16* v, float* dvec, float* bias, __nv_bfloat16* dqkv,
                       int nseq, cudaStream_t stream) {
  int rc = launch_attn_bwd_prep(datt, att, vt, dvec, v, nseq, BSEG_HEADS, kT, stream);
  if (rc) return rc;
  return launch_attention_bwd(q, k, v, datt, lse, dvec, relcat, bias, dqkv, nseq, BSEG_HEADS, stream);
}
}  // namespace

int bseg_backward_to_prompt(bseg_handle* h, const float* d_pred_masks, int batch, void* workspace,
                            size_t workspace_bytes, float* d_prompt_pixel_values, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_forward_args(h, batch, 0, workspace, "bseg_backward_to_prompt");
  if (rc) return rc;
  BSEG_REQUIRE(h->train_arena != nullptr, "bseg_backward_to_prompt: call bseg_train_prepare first");
  const TrainLayout L = train_layout(h, batch);
  BSEG_REQUIRE(workspace_bytes >= L.total, "bseg_backward_to_prompt: workspace too small (%zu < %zu)", workspace_bytes,
               L.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const FwdBufs fb = train_bufs(h, L, ws);
  auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(ws + o); };
  auto fp = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  float *dh = fp(L.dh), *dxn = fp(L.dxn), *dinter = fp(L.dinter), *dpatch = fp(L.dpatch);
  __nv_bfloat16 *dhb = bf(L.dhb), *dz = bf(L.dz), *datt = bf(L.datt), *dqkv = bf(L.dqkv), *ddec = bf(L.ddec),
                *dconv = bf(L.dconv);
  const int B = batch;
  const long long M = static_cast<long long>(B) * kT;
  // the loss only sees the query half (src/model.py:48-57): d(pred_masks) is zero for image rows < 448, so the decoder
  // backward touches image rows >= 447 only == token rows 27..55 == tokens [756, 1568) of every sample
  const int kY0 = 448, kYFirst = 432, kTok0 = (kYFirst / 16) * 28, kToks = kT - kTok0;

  // ---- decoder head (modeling_seggpt.py:546-552) and conv3x3 dgrad ----
  if ((rc = launch_decoder_head_bwd(fb.dec, h->conv_w9, h->conv_b, h->dec_ln_w, h->dec_ln_b, h->head_w, h->head_b,
                                    d_pred_masks, dconv, B, 896, 448, kY0, h->eps, stream)))
    return rc;
  if ((rc = launch_decoder_conv_dgrad(dconv, h->conv_w9b, ddec, B, 896, 448, kY0, kYFirst, stream))) return rc;
  // ---- decoder_embed dgrad (pixel shuffle is the row layout of ddec) ----
  {
    GemmEpiParams ep;
    ep.out = dinter; ep.ldc = 4 * kD;
    GemmRows gr{kT, B, kTok0, kToks};
    if ((rc = launch_gemm_rows(EPI_F32, ddec, kDecN, h->dec_embed_wt, gr, 4 * kD, kDecN, ep, stream))) return rc;
  }
  BSEG_CHECK_CUDA(cudaMemsetAsync(dh, 0, static_cast<size_t>(M) * kD * 4, stream));
  BSEG_CHECK_CUDA(cudaMemsetAsync(dhb, 0, static_cast<size_t>(M) * kD * 2, stream));

  // ---- encoder, last layer first ----
  for (int i = h->num_layers - 1; i >= 0; --i) {
    const LayerPack& lp = h->layers[i];
    const FwdBufs::PerLayer& pl = fb.layers[i];
    // the four intermediates went through encoder.layernorm (modeling_seggpt.py:481-482)
    for (int j = 0; j < 4; ++j)
      if (h->inter[j] == i)
        if ((rc = launch_layernorm1024_bwd(pl.h_out, dinter + j * kD, 4 * kD, h->enc_ln_w, dh, dh, dhb, kT, B, kTok0,
                                           kToks, h->eps, stream)))
          return rc;
    // merge (modeling_seggpt.py:476-479): the image stream receives half of the merged gradient
    if (i == h->merge_index)
      if ((rc = launch_scale_f32_bf16(dh, dh, dhb, 0.5f, M * kD, stream))) return rc;
    // MLP: h_out = h_mid + lin2(gelu(lin1(LN2(h_mid))))
    {
      GemmEpiParams ep;
      ep.out = dz; ep.ldc = kMlp; ep.aux = pl.z;
      if ((rc = launch_gemm(EPI_DGELU, dhb, kD, lp.lin2_wt, M, kMlp, kD, ep, stream))) return rc;
    }
    {
      GemmEpiParams ep;
      ep.out = dxn; ep.ldc = kD;
      if ((rc = launch_gemm(EPI_F32, dz, kMlp, lp.lin1_wt, M, kD, kMlp, ep, stream))) return rc;
    }
    if ((rc = launch_layernorm1024_bwd(pl.h_mid, dxn, kD, lp.ln2_w, dh, dh, dhb, M, 1, 0, static_cast<int>(M), h->eps,
                                       stream)))
      return rc;
    // attention: h_mid = h_in + proj(attn(LN1(h_in)))
    {
      GemmEpiParams ep;
      ep.out = datt; ep.ldc = kD;
      if ((rc = launch_gemm(EPI_BF16, dhb, kD, lp.proj_wt, M, kD, kD, ep, stream))) return rc;
    }
    if ((rc = attention_backward(pl.q, pl.k, pl.vt, pl.att, datt, pl.lse, lp.relcat, bf(L.v), fp(L.dvec), fp(L.bias),
                                 dqkv, B, stream)))
      return rc;
    {
      GemmEpiParams ep;
      ep.out = dxn; ep.ldc = kD;
      if ((rc = launch_gemm(EPI_F32, dqkv, 3 * kD, lp.qkv_wt, M, kD, 3 * kD, ep, stream))) return rc;
    }
    const float* h_in = (i == 0) ? fb.h_emb : fb.layers[i - 1].h_out;
    if ((rc = launch_layernorm1024_bwd(h_in, dxn, kD, lp.ln1_w, dh, dh, dhb, M, 1, 0, static_cast<int>(M), h->eps,
                                       stream)))
      return rc;
  }

  // ---- patch embedding dgrad for the prompt half (token rows 0..27) and un-patchify ----
  {
    GemmEpiParams ep;
    ep.out = dpatch; ep.ldc = 768;
    GemmRows gr{kT, B, 0, kT / 2};
    if ((rc = launch_gemm_rows(EPI_F32, dhb, kD, h->patch_wt, gr, 768, kD, ep, stream))) return rc;
  }
  return launch_unpatchify_prompt_grad(dpatch, d_prompt_pixel_values, B, stream);
}

int bseg_scene_stats(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, float* stats, uint32_t* scratch,
                     void* stream) {
  BSEG_REQUIRE(Hs > 0 && Ws > 0, "scene_stats: empty scene");
  return launch_scene_stats(scene, nodata, Hs, Ws, stats, scratch, static_cast<cudaStream_t>(stream));
}

int bseg_ingest_u16x4(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                      const int32_t* boxes, int n_tiles, int crop, const int32_t* coef, const int32_t* bounds,
                      int ksize, const float* mean, const float* stdv, float* out_nchw, void* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, void* stream) {
  BSEG_REQUIRE(n_tiles >= 0 && crop > 0 && ksize > 0, "ingest: bad arguments");
  int band, max_rows;
  ingest_geometry(crop, &band, &max_rows);
  return launch_ingest(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, band, max_rows, mean,
                       stdv, out_nchw, static_cast<__nv_bfloat16*>(out_patch), patch_tile_stride, out_u8, out_nodata,
                       static_cast<cudaStream_t>(stream));
}

int bseg_ingest_native_u16x4(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int32_t* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, void* stream) {
  BSEG_REQUIRE(n_tiles >= 0 && crop > 0, "ingest_native: bad arguments");
  return launch_ingest_native(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, mean, stdv, out_nchw, out_u8,
                              out_nodata, static_cast<cudaStream_t>(stream));
}

int bseg_ingest_native_f32x4(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int32_t* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, void* stream) {
  BSEG_REQUIRE(n_tiles >= 0 && crop > 0, "ingest_native_f32: bad arguments");
  return launch_ingest_native_f32(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, mean, stdv, out_nchw, out_u8,
                                  out_nodata, static_cast<cudaStream_t>(stream));
}

int bseg_scene_stats_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, float* stats, uint32_t* scratch,
                         void* stream) {
  BSEG_REQUIRE(Hs > 0 && Ws > 0, "scene_stats_f32: empty scene");
  return launch_scene_stats_f32(scene, nodata, Hs, Ws, stats, scratch, static_cast<cudaStream_t>(stream));
}

int bseg_scene_stats_rows(const void* scene, int is_f32, const uint8_t* nodata, int Hs, int Ws, int row0, int row1,
                          uint32_t* keys, void* stream) {
  BSEG_REQUIRE(Hs > 0 && Ws > 0 && row0 >= 0 && row0 <= row1 && row1 <= Hs, "scene_stats_rows: rows [%d,%d) of %d",
               row0, row1, Hs);
  return launch_scene_stats_rows(scene, is_f32, nodata, Hs, Ws, row0, row1, keys, static_cast<cudaStream_t>(stream));
}

int bseg_scene_stats_finalize(const uint32_t* keys, float* stats, void* stream) {
  return launch_scene_stats_finalize(keys, stats, static_cast<cudaStream_t>(stream));
}

int bseg_ingest_f32x4(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                      const int32_t* boxes, int n_tiles, int crop, const int32_t* coef, const int32_t* bounds,
                      int ksize, const float* mean, const float* stdv, float* out_nchw, void* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, void* stream) {
  BSEG_REQUIRE(n_tiles >= 0 && crop > 0 && ksize > 0, "ingest_f32: bad arguments");
  int band, max_rows;
  ingest_geometry(crop, &band, &max_rows);
  return launch_ingest_f32(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, band, max_rows,
                           mean, stdv, out_nchw, static_cast<__nv_bfloat16*>(out_patch), patch_tile_stride, out_u8,
                           out_nodata, static_cast<cudaStream_t>(stream));
}

int bseg_merge_mosaic(const float* data, const uint8_t* yesdata, int n_rasters, int channels, int Hs, int Ws,
                      float* mean, uint8_t* nodata, void* stream) {
  BSEG_REQUIRE(n_rasters > 0 && channels > 0 && Hs > 0 && Ws > 0 && mean != nullptr && nodata != nullptr,
               "merge_mosaic: bad arguments");
  return launch_merge_mosaic(data, yesdata, n_rasters, channels, Hs, Ws, mean, nodata,
                             static_cast<cudaStream_t>(stream));
}

int bseg_preprocess_u8(const uint8_t* images, int layout_chw, int n, int crop, const int32_t* coef,
                       const int32_t* bounds, int ksize, int precision_bits, const float* mean255, const float* std255,
                       float* out_nchw, void* stream) {
  BSEG_REQUIRE(n >= 0 && crop > 0 && ksize > 0 && out_nchw != nullptr, "preprocess_u8: bad arguments");
  int band, max_rows;
  ingest_geometry(crop, &band, &max_rows);
  return launch_preprocess_u8(images, layout_chw, n, crop, coef, bounds, ksize, precision_bits, band, max_rows, mean255,
                              std255, out_nchw, static_cast<cudaStream_t>(stream));
}

int bseg_colorize_resize_norm255(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean255,
                                 const float* std255, const int32_t* resize_idx, float* out, int batch, int in_size,
                                 int out_size, void* stream) {
  BSEG_REQUIRE(batch > 0 && num_classes > 0 && in_size > 0 && out_size > 0, "colorize_resize_norm255: bad arguments");
  BSEG_REQUIRE(resize_idx != nullptr || in_size == out_size, "colorize_resize_norm255: resize needs an index table");
  return launch_colorize_resize_norm255(mask, palette, num_classes, mean255, std255, resize_idx, out, batch, in_size,
                                        out_size, static_cast<cudaStream_t>(stream));
}

int bseg_postprocess_semantic(const float* pred, const float* palette255, int num_classes, const float* mean,
                              const float* stdv, uint8_t* out_u8, int64_t* out_i64, const uint8_t* nodata,
                              const int32_t* resize_idx, int batch, int H, int W, int out_size, void* stream) {
  BSEG_REQUIRE(batch > 0 && num_classes > 0 && num_classes <= 256, "postprocess_semantic: bad arguments");
  BSEG_REQUIRE(resize_idx != nullptr || (out_size == H && out_size == W),
               "postprocess_semantic: out_size=%d needs a resize index table", out_size);
  return launch_postprocess_semantic(pred, palette255, num_classes, mean, stdv, out_u8,
                                     reinterpret_cast<long long*>(out_i64), nodata, resize_idx, batch, H, W, out_size,
                                     static_cast<cudaStream_t>(stream));
}

int bseg_colorize_norm(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean,
                       const float* stdv, float* out, int batch, int H, int W, void* stream) {
  BSEG_REQUIRE(batch > 0 && num_classes > 0, "colorize_norm: bad arguments");
  return launch_colorize_norm(mask, palette, num_classes, mean, stdv, out, batch, H, W,
                              static_cast<cudaStream_t>(stream));
}

int bseg_decode_palette(const float* pred, const float* palette_norm, int num_classes, uint8_t* out_u8,
                        int64_t* out_i64, const uint8_t* nodata, const int32_t* resize_idx, int batch, int H, int W,
                        int out_size, void* stream) {
  BSEG_REQUIRE(batch > 0 && num_classes > 0 && num_classes <= 256, "decode_palette: bad arguments");
  BSEG_REQUIRE(resize_idx != nullptr || (out_size == H && out_size == W),
               "decode_palette: out_size=%d needs a resize index table", out_size);
  return launch_decode_palette(pred, palette_norm, num_classes, out_u8, reinterpret_cast<long long*>(out_i64), nodata,
                               resize_idx, batch, H, W, out_size, static_cast<cudaStream_t>(stream));
}

int bseg_mean_over_prompts(const float* pred, float* out, int n_tiles, int prompts, long long elems_per_sample,
                           void* stream) {
  BSEG_REQUIRE(n_tiles > 0 && prompts > 0, "mean_over_prompts: bad arguments");
  return launch_mean_over_group(pred, out, n_tiles, prompts, elems_per_sample, static_cast<cudaStream_t>(stream));
}

int bseg_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int32_t* order4,
                       const float* noise, float noise_mean, float noise_std, const float* mean, const float* stdv,
                       float* out_image, uint8_t* out_mask, float* colour_out, int batch, int H, int W, void* stream) {
  BSEG_REQUIRE(batch >= 0 && H > 0 && W > 0, "train_aug_fwd: bad shape %d x %d x %d", batch, H, W);
  if (batch == 0) return 0;
  BSEG_REQUIRE(image && params && order4 && mean && stdv && out_image && colour_out, "train_aug_fwd: null argument");
  BSEG_REQUIRE((mask == nullptr) == (out_mask == nullptr), "train_aug_fwd: mask and out_mask go together");
  if (batch == 0) return 0;
  return launch_train_aug_fwd(image, mask, params, order4, noise, noise_mean, noise_std, mean, stdv, out_image,
                              out_mask, colour_out, batch, H, W, static_cast<cudaStream_t>(stream));
}

int bseg_train_aug_bwd(const float* image, const float* params, const int32_t* order4, const float* stdv,
                       const float* colour_out, const float* d_out, float* scratch, float* d_image, int batch, int H,
                       int W, void* stream) {
  BSEG_REQUIRE(batch >= 0 && H > 0 && W > 0, "train_aug_bwd: bad shape %d x %d x %d", batch, H, W);
  if (batch == 0) return 0;
  BSEG_REQUIRE(image && params && order4 && stdv && colour_out && d_out && scratch && d_image,
               "train_aug_bwd: null argument");
  if (batch == 0) return 0;
  return launch_train_aug_bwd(image, params, order4, stdv, colour_out, d_out, scratch, d_image, batch, H, W,
                              static_cast<cudaStream_t>(stream));
}

int bseg_vote_accumulate(uint32_t* counter, int Hs, int Ws, const uint8_t* cls, int n_tiles, int crop,
                         const int32_t* boxes, int use_atomics, void* stream) {
  BSEG_REQUIRE(n_tiles >= 0 && crop >= 0, "vote_accumulate: bad arguments");
  return launch_vote_accumulate(counter, Hs, Ws, cls, n_tiles, crop, boxes, use_atomics,
                                static_cast<cudaStream_t>(stream));
}

int bseg_vote_argmax(const uint32_t* counter, uint8_t* out, long long n_pixels, void* stream) {
  return launch_vote_argmax(counter, out, n_pixels, static_cast<cudaStream_t>(stream));
}

int bseg_paste_tiles_u8(uint8_t* canvas, int Hs, int Ws, const uint8_t* crops, int n_tiles, int crop,
                        const int32_t* boxes, void* stream) {
  BSEG_REQUIRE(Hs > 0 && Ws > 0 && n_tiles >= 0 && crop > 0, "paste_tiles: bad arguments");
  return launch_paste_tiles(canvas, Hs, Ws, crops, n_tiles, crop, boxes, static_cast<cudaStream_t>(stream));
}

int bseg_overlay_prediction(const uint8_t* img, const uint8_t* pred, const uint8_t* class_rgba, int n_classes,
                            long long n_pixels, uint8_t* out, void* stream) {
  BSEG_REQUIRE(n_classes > 0 && n_classes <= 256 && n_pixels >= 0, "overlay_prediction: bad arguments");
  return launch_overlay_prediction(img, pred, class_rgba, n_classes, n_pixels, out,
                                   static_cast<cudaStream_t>(stream));
}

int bseg_loss_smoothl1_fwd_bwd(const float* pred, const float* labels, const uint8_t* yesdata, float beta,
                               int per_sample, float* loss_out, float* grad_out, float* scratch, int batch, int H,
                               int W, void* stream) {
  BSEG_REQUIRE(batch > 0 && beta > 0.f, "loss: bad arguments");
  return launch_smooth_l1(pred, labels, yesdata, beta, per_sample, loss_out, grad_out, scratch, batch, H, W,
                          static_cast<cudaStream_t>(stream));
}

int bseg_gemm_set_cta_pairs(int on) { return gemm_set_cta_pairs(on); }
int bseg_gemm_set_small_tiles(int on) { return gemm_set_small_tiles(on); }
int bseg_gemm_set_fused_ln(int on) { return gemm_set_fused_ln(on); }
int bseg_set_pdl(int on) { return pdl_set(on); }

size_t bseg_gemm_resid_ln_scratch_bytes(long long M) {
  if (M <= 0) return 0;
  return ln_stats_bytes(static_cast<size_t>(M));
}
int bseg_gemm_bf16_resid_ln(const void* A, long long lda, const void* W, long long M, int K, const float* bias,
                            float* hres, const float* gamma, const float* beta, void* ln_out, float eps, void* scratch,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(A && W && hres && gamma && beta && ln_out && scratch && M > 0, "bseg_gemm_bf16_resid_ln: bad argument");
  BSEG_CHECK_CUDA(cudaMemsetAsync(scratch, 0xFF, ln_stats_bytes(static_cast<size_t>(M)), stream));
  GemmEpiParams ep;
  ep.out = hres; ep.ldc = 1024; ep.bias = bias; ep.resid = hres; ep.ldr = 1024;
  ep.ln_gamma = gamma; ep.ln_beta = beta; ep.ln_out = static_cast<__nv_bfloat16*>(ln_out); ep.ld_ln = 1024;
  ep.ln_eps = eps;
  ep.ln_stats = static_cast<unsigned long long*>(scratch);
  ep.ln_tag = 0;
  return launch_gemm(EPI_RESID_LN, static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), M,
                     1024, K, ep, stream);
}

int bseg_gemm_bf16(const void* A, long long lda, const void* W, long long M, int N, int K, const float* bias,
                   void* out, long long ldc, int out_is_bf16, int gelu, void* stream) {
  GemmEpiParams ep;
  ep.out = out;
  ep.ldc = ldc;
  ep.bias = bias;
  const int mode = out_is_bf16 ? (gelu ? EPI_BF16_GELU : EPI_BF16) : EPI_F32;
  return launch_gemm(mode, static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), M, N, K,
                     ep, static_cast<cudaStream_t>(stream));
}

int bseg_layernorm1024(const float* x, long long ldx, const float* gamma, const float* beta, void* out,
                       long long ldo, long long M, float eps, void* stream) {
  return launch_layernorm1024(x, ldx, gamma, beta, static_cast<__nv_bfloat16*>(out), ldo, M, eps,
                              static_cast<cudaStream_t>(stream));
}

int bseg_attention(const void* q, const void* k, const void* vt, const void* relcat, void* out, int nseq,
                   void* stream) {
  return launch_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                          static_cast<const __nv_bfloat16*>(vt), static_cast<const __nv_bfloat16*>(relcat),
                          static_cast<__nv_bfloat16*>(out), nullptr, nseq, BSEG_HEADS, 56, 28,
                          static_cast<cudaStream_t>(stream));
}

int bseg_attention_fwd_lse(const void* q, const void* k, const void* vt, const void* relcat, void* out, float* lse,
                           int nseq, void* stream) {
  return launch_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                          static_cast<const __nv_bfloat16*>(vt), static_cast<const __nv_bfloat16*>(relcat),
                          static_cast<__nv_bfloat16*>(out), lse, nseq, BSEG_HEADS, 56, 28,
                          static_cast<cudaStream_t>(stream));
}

int bseg_attention_grid(const void* q, const void* k, const void* vt, const void* relcat, void* out, float* lse,
                        int nseq, int grid_h, int grid_w, void* stream) {
  return launch_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                          static_cast<const __nv_bfloat16*>(vt), static_cast<const __nv_bfloat16*>(relcat),
                          static_cast<__nv_bfloat16*>(out), lse, nseq, BSEG_HEADS, grid_h, grid_w,
                          static_cast<cudaStream_t>(stream));
}

size_t bseg_attention_bwd_scratch_bytes(int nseq) {
  if (nseq <= 0) return 0;
  const size_t rows = static_cast<size_t>(nseq) * kT;
  // v (bf16), Dvec (fp32), the per-query 16-bit tables (sized as 84 floats per query and head)
  return align_up(rows * kD * 2, 1024) + align_up(rows * BSEG_HEADS * 4, 1024) +
         align_up(rows * BSEG_HEADS * 84 * 4, 1024);
}

int bseg_attention_bwd(const void* q, const void* k, const void* vt, const void* out, const void* d_out,
                       const float* lse, const void* relcat, void* dqkv, int nseq, void* scratch, size_t scratch_bytes,
                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(nseq > 0, "attention_bwd: nseq=%d", nseq);
  BSEG_REQUIRE(scratch != nullptr && scratch_bytes >= bseg_attention_bwd_scratch_bytes(nseq) &&
                   (reinterpret_cast<uintptr_t>(scratch) & 255) == 0,
               "attention_bwd: scratch too small or misaligned");
  const size_t rows = static_cast<size_t>(nseq) * kT;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  auto take = [&](size_t bytes) {
    uint8_t* r = p;
    p += align_up(bytes, 1024);
    return r;
  };
  auto* v = reinterpret_cast<__nv_bfloat16*>(take(rows * kD * 2));
  auto* dvec = reinterpret_cast<float*>(take(rows * BSEG_HEADS * 4));
  auto* bias = reinterpret_cast<float*>(take(rows * BSEG_HEADS * 84 * 4));
  return attention_backward(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                            static_cast<const __nv_bfloat16*>(vt), static_cast<const __nv_bfloat16*>(out),
                            static_cast<const __nv_bfloat16*>(d_out), lse, static_cast<const __nv_bfloat16*>(relcat),
                            v, dvec, bias, static_cast<__nv_bfloat16*>(dqkv), nseq, stream);
}

int bseg_layernorm1024_bwd(const float* x, const float* dy, long long lddy, const float* gamma, const float* dh_in,
                           float* dh_out, void* dh_bf16, long long M, float eps, void* stream) {
  BSEG_REQUIRE(M > 0 && M < (1ll << 31), "layernorm_bwd: M=%lld", M);
  return launch_layernorm1024_bwd(x, dy, lddy, gamma, dh_in, dh_out, static_cast<__nv_bfloat16*>(dh_bf16), M, 1, 0,
                                  static_cast<int>(M), eps, static_cast<cudaStream_t>(stream));
}

int bseg_gemm_bf16_dgelu(const void* A, long long lda, const void* W, long long M, int N, int K, const void* z,
                         void* out, long long ldc, void* stream) {
  GemmEpiParams ep;
  ep.out = out;
  ep.ldc = ldc;
  ep.aux = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(z));
  return launch_gemm(EPI_DGELU, static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), M, N,
                     K, ep, static_cast<cudaStream_t>(stream));
}

int bseg_pack_conv_w9_dgrad(const void* w9, void* w9b, void* stream) {
  pack_conv_w9_dgrad_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(w9), static_cast<__nv_bfloat16*>(w9b));
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int bseg_decoder_head_bwd(const void* x_nhwc, const void* w9, const void* w9b, const float* conv_b, const float* ln_w,
                          const float* ln_b, const float* head_w, const float* head_b, const float* d_pred,
                          void* d_conv, void* d_dec_rows, int batch, int H, int W, int y0, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BSEG_REQUIRE(y0 >= 16 && y0 % 16 == 0, "decoder_head_bwd: y0=%d must be a positive multiple of 16", y0);
  int rc = launch_decoder_head_bwd(static_cast<const __nv_bfloat16*>(x_nhwc), static_cast<const __nv_bfloat16*>(w9),
                                   conv_b, ln_w, ln_b, head_w, head_b, d_pred, static_cast<__nv_bfloat16*>(d_conv),
                                   batch, H, W, y0, eps, stream);
  if (rc) return rc;
  return launch_decoder_conv_dgrad(static_cast<const __nv_bfloat16*>(d_conv), static_cast<const __nv_bfloat16*>(w9b),
                                   static_cast<__nv_bfloat16*>(d_dec_rows), batch, H, W, y0, y0 - 16, stream);
}

int bseg_pack_relcat(const float* rel_pos_h, const float* rel_pos_w, void* relcat, void* stream) {
  return bseg_pack_relcat_grid(rel_pos_h, rel_pos_w, relcat, 56, 28, stream);
}

int bseg_relcat_rows(int grid_h, int grid_w) { return attention_relcat_rows(grid_h, grid_w); }

int bseg_pack_relcat_grid(const float* rel_pos_h, const float* rel_pos_w, void* relcat, int grid_h, int grid_w,
                          void* stream) {
  BSEG_REQUIRE(grid_h > 0 && grid_w > 0, "pack_relcat: empty token grid");
  pack_relcat_kernel<<<attention_relcat_rows(grid_h, grid_w), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      rel_pos_h, rel_pos_w, static_cast<__nv_bfloat16*>(relcat), grid_h, grid_w, (2 * grid_h - 1 + 15) / 16 * 16);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int bseg_decoder_head(const void* x_nhwc, const void* w9, const float* conv_b, const float* ln_w, const float* ln_b,
                      const float* head_w, const float* head_b, float* pred, int batch, int H, int W, float eps,
                      void* stream) {
  return launch_decoder_head(static_cast<const __nv_bfloat16*>(x_nhwc), static_cast<const __nv_bfloat16*>(w9),
                             conv_b, ln_w, ln_b, head_w, head_b, pred, batch, H, W, eps, 0,
                             static_cast<cudaStream_t>(stream));
}

int bseg_pack_conv_w9(const float* conv_w, void* w9, void* stream) {
  pack_conv_w9_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      conv_w, static_cast<__nv_bfloat16*>(w9));
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int bseg_profile_enable(int on) {
  prof_set_enabled(on != 0);
  return 0;
}

int bseg_profile_collect(double* ms, long long* launches, double* work, double* bytes) {
  prof_collect(ms, launches, work, bytes);
  return 0;
}

int bseg_profile_collect_gemm(double* ms, double* work) {
  prof_collect_sub(ms, work);
  return 0;
}

int bseg_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  return launch_f32_to_bf16(src, static_cast<__nv_bfloat16*>(dst), n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
