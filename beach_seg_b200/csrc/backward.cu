// Bandwidth-bound kernels of the train step's backward pass: the reference obtains
// d(loss)/d(prompt_pixel_values) from torch autograd through the frozen HF SegGPT (src/model.py:245-255, Lightning's
// backward).  Only data gradients exist (every backbone weight has requires_grad=False, src/util/ml_util.py:9-10),
// so the backward is a chain of dgrad GEMMs (gemm.cuh, transposed weight copies), the attention backward
// (attention_bwd.cu), the decoder-head backward (decoder_conv.cu) and the kernels below.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

static inline int blocks_for(long long n, int per_block, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<int>(b);
}

// ----------------------------------------------------------------------------------------------
// bf16 matrix transpose [R, C] -> [C, R], batched (weight packing for the dgrad GEMMs and the per-head operand
// transposes of the attention backward).  32x32 tiles through shared memory.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C,
                      long long src_batch_stride, long long dst_batch_stride, long long ld_src, long long ld_dst) {
  __shared__ __nv_bfloat16 tile[32][34];
  const __nv_bfloat16* s = src + blockIdx.z * src_batch_stride;
  __nv_bfloat16* d = dst + blockIdx.z * dst_batch_stride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < R && c < C) ? s[r * ld_src + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < C && r < R) d[c * ld_dst + r] = tile[tx][ty + i];
  }
}
int launch_transpose_bf16(const __nv_bfloat16* src, __nv_bfloat16* dst, int R, int C, int batch,
                          long long src_batch_stride, long long dst_batch_stride, long long ld_src, long long ld_dst,
                          cudaStream_t stream) {
  BSEG_REQUIRE(R > 0 && C > 0 && batch > 0 && batch <= 65535, "transpose: bad shape R=%d C=%d batch=%d", R, C, batch);
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  ProfScope prof(CAT_ELEMENTWISE, 0, 4.0 * R * C * batch, stream);
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(src, dst, R, C, src_batch_stride, dst_batch_stride, ld_src, ld_dst);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm backward over D = 1024 (nn.LayerNorm, eps 1e-6; forward: layernorm1024_kernel):
//   xhat = (x - mean) * rstd,  g = dy * gamma,  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
//   dh_out[row] = (dh_in ? dh_in[row] : 0) + dx        (fp32 residual-stream gradient; may alias dh_in)
//   dh_bf16[row] = bf16(dh_out[row])                   (A operand of the next dgrad GEMM; optional)
// One warp per row, the row in registers; statistics are recomputed from the saved fp32 input x.
// Rows: `nbatch` groups of `rows_per_batch`, rows [row_begin, row_begin + rows) of each (the decoder only sends
// gradient into the query half).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
layernorm1024_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long lddy,
                         const float* __restrict__ gamma, const float* dh_in, float* dh_out,
                         __nv_bfloat16* __restrict__ dh_bf16, long long rows_per_batch, int nbatch, int row_begin,
                         int rows, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)nbatch * rows;
  for (long long it = warp_global; it < total; it += nwarps) {
    const long long row = (it / rows) * rows_per_batch + row_begin + (it % rows);
    const float4* xr = reinterpret_cast<const float4*>(x + row * 1024);
    const float4* gr = reinterpret_cast<const float4*>(dy + row * lddy);
    float4 v[8], g[8], r[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = xr[lane + 32 * i];
      g[i] = gr[lane + 32 * i];
    }
    // the residual gradient is only needed at the very end: request it now, so that all three input streams of the
    // row are in flight together (it used to be fetched after the four warp reductions: 0.74 of the HBM rate)
    if (dh_in != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = reinterpret_cast<const float4*>(dh_in + row * 1024)[lane + 32 * i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / 1024.0f);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 1024.0f) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      g[i].x *= w.x; g[i].y *= w.y; g[i].z *= w.z; g[i].w *= w.w;
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float m1 = warp_sum(sg) * (1.0f / 1024.0f), m2 = warp_sum(sgx) * (1.0f / 1024.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 o;
      o.x = rstd * (g[i].x - m1 - v[i].x * m2);
      o.y = rstd * (g[i].y - m1 - v[i].y * m2);
      o.z = rstd * (g[i].z - m1 - v[i].z * m2);
      o.w = rstd * (g[i].w - m1 - v[i].w * m2);
      if (dh_in != nullptr) { o.x += r[i].x; o.y += r[i].y; o.z += r[i].z; o.w += r[i].w; }
      reinterpret_cast<float4*>(dh_out + row * 1024)[lane + 32 * i] = o;
      if (dh_bf16 != nullptr)
        reinterpret_cast<uint2*>(dh_bf16 + row * 1024)[lane + 32 * i] =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}
int launch_layernorm1024_bwd(const float* x, const float* dy, long long lddy, const float* gamma, const float* dh_in,
                             float* dh_out, __nv_bfloat16* dh_bf16, long long rows_per_batch, int nbatch,
                             int row_begin, int rows, float eps, cudaStream_t stream) {
  BSEG_REQUIRE(lddy % 4 == 0 && rows > 0 && nbatch > 0 && row_begin >= 0 && row_begin + rows <= rows_per_batch,
               "layernorm_bwd: bad arguments");
  const long long total = (long long)nbatch * rows;
  ProfScope prof(CAT_LAYERNORM, 0, static_cast<double>(total) * 1024 * (dh_in ? 18 : 14), stream);
  layernorm1024_bwd_kernel<<<blocks_for(total, 8, 148 * 8), 256, 0, stream>>>(x, dy, lddy, gamma, dh_in, dh_out, dh_bf16,
                                                                             rows_per_batch, nbatch, row_begin, rows,
                                                                             eps);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// dst_f32 = src * scale (+ bf16 copy): backward of the two-stream merge (modeling_seggpt.py:476-479): the image
// stream receives 0.5 * d(merged).  Also used as the plain fp32 -> (fp32, bf16) fan-out.
// ----------------------------------------------------------------------------------------------
__global__ void scale_f32_bf16_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                      uint2* __restrict__ dst_bf16, float scale, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = src[i];
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (dst) dst[i] = v;
    if (dst_bf16) dst_bf16[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}
int launch_scale_f32_bf16(const float* src, float* dst, __nv_bfloat16* dst_bf16, float scale, long long n,
                          cudaStream_t stream) {
  BSEG_REQUIRE(n % 4 == 0, "scale: size must be a multiple of 4");
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(n) * 10, stream);
  scale_f32_bf16_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), reinterpret_cast<uint2*>(dst_bf16), scale,
      n / 4);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Backward of patchify for the prompt image: the patch-embedding dgrad GEMM leaves d(A) [B*1568, 768] fp32 with the
// im2col column order c*256 + py*16 + px; the prompt occupies the top half (patch rows 0..27) of the image stream
// (modeling_seggpt.py:713), so d(prompt_pixel_values)[b,c,y,x] = dA[b*1568 + (y/16)*28 + x/16][c*256 + (y%16)*16 + x%16].
// ----------------------------------------------------------------------------------------------
__global__ void unpatchify_prompt_grad_kernel(const float* __restrict__ dA, float* __restrict__ dprompt, int B) {
  const long long total = (long long)B * 3 * 448 * 112;  // float4 along x
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x4 = static_cast<int>(idx % 112);
    long long r = idx / 112;
    const int y = static_cast<int>(r % 448);
    r /= 448;
    const int c = static_cast<int>(r % 3);
    const int b = static_cast<int>(r / 3);
    const int x = x4 * 4;
    const long long row = (long long)b * 1568 + (y >> 4) * 28 + (x >> 4);
    const float4 v = *reinterpret_cast<const float4*>(dA + row * 768 + c * 256 + (y & 15) * 16 + (x & 15));
    *reinterpret_cast<float4*>(dprompt + (((long long)b * 3 + c) * 448 + y) * 448 + x) = v;
  }
}
int launch_unpatchify_prompt_grad(const float* dA, float* dprompt, int B, cudaStream_t stream) {
  const long long total = (long long)B * 3 * 448 * 112;
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * 32, stream);
  unpatchify_prompt_grad_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(dA, dprompt, B);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Attention-backward operand preparation for one layer (all per (sequence, head)):
//   Dvec[sh][t] = sum_d dO[.][d] * O[.][d]      (softmax backward row term)
//   v[sh][t][d] = vt[sh][d][t]                  (K-major operand of dP = dO v^T)
// dO and O are token-major [nseq*T, heads*64] bf16; vt is [nseq*heads, 64, T].  One block handles 64 tokens of one
// (seq, head).  (The other transposed operands the backward once needed -- q^T, k^T, dO^T -- are gone: the kernels
// read the row-major tiles as MN-major MMA operands.)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                     const __nv_bfloat16* __restrict__ vt, float* __restrict__ Dvec, __nv_bfloat16* __restrict__ v,
                     int heads, int T) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int t0 = blockIdx.x * 64;
  const int sh = blockIdx.y;
  const int seq = sh / heads, head = sh % heads;
  const int D = heads * 64;
  const int tid = threadIdx.x;
  // ---- Dvec: thread (row = tid/4, 16 columns at (tid%4)*16) ----
  {
    const int r = tid >> 2, c0 = (tid & 3) * 16;
    float part = 0.f;
    if (t0 + r < T) {
      const long long off = ((long long)seq * T + t0 + r) * D + head * 64 + c0;
      const uint4 a0 = *reinterpret_cast<const uint4*>(dO + off), a1 = *reinterpret_cast<const uint4*>(dO + off + 8);
      const uint4 b0 = *reinterpret_cast<const uint4*>(O + off), b1 = *reinterpret_cast<const uint4*>(O + off + 8);
      const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
        part += __uint_as_float(aw[j] << 16) * __uint_as_float(bw[j] << 16) +
                __uint_as_float(aw[j] & 0xffff0000u) * __uint_as_float(bw[j] & 0xffff0000u);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if ((tid & 3) == 0 && t0 + r < T) Dvec[(long long)sh * T + t0 + r] = part;
  }
  // ---- vt -> v ----
  {
    const int tl = tid & 63, dg = (tid >> 6) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      tile[dg + j][tl] = (t0 + tl < T) ? vt[((long long)sh * 64 + dg + j) * T + t0 + tl] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  {
    const int r = tid >> 2, c0 = (tid & 3) * 16;
    if (t0 + r < T) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t lo = *reinterpret_cast<const uint16_t*>(&tile[c0 + 2 * j][r]);
        const uint32_t hi = *reinterpret_cast<const uint16_t*>(&tile[c0 + 2 * j + 1][r]);
        w[j] = lo | (hi << 16);
      }
      uint4* o = reinterpret_cast<uint4*>(v + ((long long)sh * T + t0 + r) * 64 + c0);
      o[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}
int launch_attn_bwd_prep(const __nv_bfloat16* dO, const __nv_bfloat16* O, const __nv_bfloat16* vt, float* Dvec,
                         __nv_bfloat16* v, int nseq, int heads, int T, cudaStream_t stream) {
  BSEG_REQUIRE(nseq > 0 && heads > 0 && nseq * heads <= 65535, "attn_bwd_prep: bad shape");
  dim3 grid((T + 63) / 64, nseq * heads);
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(nseq) * heads * T * (64 * 2 * 4 + 4), stream);
  attn_bwd_prep_kernel<<<grid, 256, 0, stream>>>(dO, O, vt, Dvec, v, heads, T);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bseg
