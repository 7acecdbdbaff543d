// Blackwell (sm_100a) device-side primitives used by every kernel in this library:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / TMEM load / commit)
// and the UMMA shared-memory / instruction descriptors.  Inline PTX only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace bseg {

// ----------------------------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Warp index as a value the compiler knows to be warp-uniform (so that role branches and everything derived from them
// stay on the uniform datapath), and single-lane election for tcgen05 / TMA issue.  A plain `if (lane == 0)` makes
// every operand look divergent: each tcgen05.mma then costs an ELECT + R2UR.BROADCAST waterfall loop (~100+ cycles).
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still draining.  pdl_launch_dependents() lets the NEXT kernel's CTAs be
// scheduled as soon as every CTA of this grid has started (they then sit in their own pdl_wait()); pdl_wait() blocks
// until the predecessor grid has completed and its memory operations are visible.  Both are no-ops for a normal launch.
// Rule of this library: every kernel that is ever launched with the attribute calls pdl_wait() before it touches global
// memory other than its parameters, so a chain of such kernels stays transitively ordered.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// exact-erf GELU (HF ACT2FN["gelu"] == F.gelu(approximate="none")): x * Phi(x).
// With h(a) = Phi(-a) = 0.5 * erfc(a / sqrt2) for a = |x|:  gelu(x) = max(x, 0) - |x| * h(|x|)  (no cancellation, no
// sign select).  h = exp2(q(a)) where q is a degree-6 polynomial fit of log2(h) on [0, 8] (weighted so that the
// absolute error of gelu is minimised): |gelu error| <= 5e-7 for all x, one MUFU and 9 FMA-pipe instructions per
// element instead of libm erff (the A&S 7.1.26 form used before needed 2 MUFU + 14 FMA-pipe instructions and made
// the lin1 epilogue the limiter of that GEMM).  For |x| > 8, h < 1e-15 and a is clamped.
__device__ __forceinline__ float gelu_tail_h(float ax) {
  const float a = fminf(ax, 8.0f);
  float q = fmaf(a, 2.980947283504065e-05f, -0.0007226605666801333f);
  q = fmaf(q, a, 0.007852795533835888f);
  q = fmaf(q, a, -0.052913740277290344f);
  q = fmaf(q, a, -0.45927515625953674f);
  q = fmaf(q, a, -1.150988221168518f);
  q = fmaf(q, a, -1.0000197887420654f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));
  return h;  // Phi(-|x|)
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  return fmaf(-ax, gelu_tail_h(ax), fmaxf(x, 0.0f));
}

// Two elements at a time with the polynomial on the packed fp32 pipe (fma.rn.f32x2 rounds each half exactly like the
// scalar fma, so the results are bit-identical to gelu_erf): 8 issue slots per element instead of 11.5 -- the lin1
// GEMM's epilogue was 13 % of its time.
__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1);
__device__ __forceinline__ void mul_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1);
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2_raw(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
  const float a0 = fminf(fabsf(x0), 8.0f), a1 = fminf(fabsf(x1), 8.0f);
  const uint64_t a = pack_f32x2(a0, a1);
  uint64_t q = fma_f32x2_raw(a, pack_f32x2(2.980947283504065e-05f, 2.980947283504065e-05f),
                             pack_f32x2(-0.0007226605666801333f, -0.0007226605666801333f));
  q = fma_f32x2_raw(q, a, pack_f32x2(0.007852795533835888f, 0.007852795533835888f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-0.052913740277290344f, -0.052913740277290344f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-0.45927515625953674f, -0.45927515625953674f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-1.150988221168518f, -1.150988221168518f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-1.0000197887420654f, -1.0000197887420654f));
  float q0, q1, h0, h1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(q1));
  x0 = fmaf(-fabsf(x0), h0, fmaxf(x0, 0.0f));
  x1 = fmaf(-fabsf(x1), h1, fmaxf(x1, 0.0f));
}

// d/dx of the exact-erf GELU: Phi(x) + x * phi(x), Phi from the same h, phi(x) = exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float h = gelu_tail_h(fabsf(x));
  const float cdf = x >= 0.f ? 1.0f - h : h;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));  // exp(-x^2/2)
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// two derivatives at a time, polynomial and the exponent's argument on the packed pipe (bit-identical to the scalar form)
__device__ __forceinline__ void gelu_erf_grad_x2(float x0, float x1, float& g0, float& g1) {
  const float a0 = fminf(fabsf(x0), 8.0f), a1 = fminf(fabsf(x1), 8.0f);
  const uint64_t a = pack_f32x2(a0, a1);
  uint64_t q = fma_f32x2_raw(a, pack_f32x2(2.980947283504065e-05f, 2.980947283504065e-05f),
                             pack_f32x2(-0.0007226605666801333f, -0.0007226605666801333f));
  q = fma_f32x2_raw(q, a, pack_f32x2(0.007852795533835888f, 0.007852795533835888f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-0.052913740277290344f, -0.052913740277290344f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-0.45927515625953674f, -0.45927515625953674f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-1.150988221168518f, -1.150988221168518f));
  q = fma_f32x2_raw(q, a, pack_f32x2(-1.0000197887420654f, -1.0000197887420654f));
  float q0, q1, h0, h1, e0, e1, s0, s1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(q1));
  // exp(-x^2 / 2): the scalar form evaluates (x * x) * c, so do the packed one
  mul_f32x2(s0, s1, x0, x1, x0, x1);
  mul_f32x2(s0, s1, s0, s1, -0.72134752044448170368f, -0.72134752044448170368f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
  const float c0 = x0 >= 0.f ? 1.0f - h0 : h0, c1 = x1 >= 0.f ? 1.0f - h1 : h1;
  g0 = fmaf(x0 * 0.39894228040143267794f, e0, c0);
  g1 = fmaf(x1 * 0.39894228040143267794f, e1, c1);
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA / UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Potentially blocking: the thread is suspended until the phase completes or the suspend-time hint (in ns) elapses, so
// a waiting warp costs (almost) no issue slots.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// non-blocking probe (no suspend-time hint): has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocking wait with a watchdog: a protocol bug traps instead of hanging the GPU box.  The clock is only read every
// 64th wake-up so that the spin itself stays a two-instruction loop.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      // ~2 s at 2 GHz for worker warps, twice that for the control warps 0-3 so that the workers report first
      if (now - t0 > (threadIdx.x < 128 ? 8000000000LL : 4000000000LL)) {
        printf("bseg: mbarrier watchdog block=(%d,%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
               blockIdx.z, threadIdx.x, smem_u32(bar), parity);
#ifdef BSEG_WATCHDOG_HOOK
        BSEG_WATCHDOG_HOOK
#endif
        __trap();
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}

// explicit shared-memory loads (a float* derived from the dynamic smem base otherwise compiles to generic LD)
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float lds32(const float* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ void red_shared_add_f32(float* p, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v) : "memory");
}

// named barrier among a subset of the CTA's warps (id 1..15; `count` threads, a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x 16 bf16 per instruction) lives in TMEM as 128 lanes x 8
// 32-bit columns, two consecutive K elements packed per column (even k in the low half).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of one TPC execute ONE tcgen05.mma of M = 256; each
// CTA stages its own 128 rows of A and HALF of the B tile (N/2 rows), holds its 128 accumulator lanes in its own TMEM
// and runs its own epilogue.  Only the even CTA (rank 0, the "leader") issues MMAs and owns the full / tmem_empty
// barriers; commits are multicast to both CTAs.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(cta_rank)
      : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's shared memory, the transaction bytes are credited to the
// LEADER CTA's mbarrier (bit 24 of a shared::cta address is the CTA's position in the pair).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {  // the same warp of BOTH CTAs
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// M = 256 MMA across the pair, issued by ONE thread of the leader CTA; the descriptors address the same offsets in
// both CTAs' shared memory
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once the pair's previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// TMEM -> registers: lane (32*(warp%4) + lane_id), 32 / 16 consecutive fp32 columns starting at taddr.col
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same lane/column mapping as tmem_ld16
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 2^x, MUFU.EX2 (approx, flush-to-zero); ex2(-inf) = 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp in the CUTLASS tree shipped with the image)
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major bf16 operand tile stored as rows of 128 bytes
// (64 bf16 along K) with the 128-byte swizzle that TMA's CU_TENSOR_MAP_SWIZZLE_128B produces.
// 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major; version=1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored)  [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version        [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor: A=bf16, B=bf16, D=fp32, both K-major, dense, no negate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) |         // c_format = F32
         (1u << 7) |         // a_format = BF16
         (1u << 10) |        // b_format = BF16
         ((N >> 3) << 17) |  // n_dim
         ((M >> 4) << 24);   // m_dim
}

// K-major is the layout of every TMA-loaded [rows][64 x 16-bit] tile used as "rows = M or N, columns = K".  The SAME bytes
// read as an MN-major operand are "rows = K, columns = M or N" (CUTLASS mma_traits_sm100.hpp, make_umma_desc<Major::MN>,
// SWIZZLE_128B: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 MN elements per 128-byte row, 8 K rows per
// 1024-byte swizzle atom, the next 8 K rows at SBO; LBO (the next 64 MN elements) is unused for MN extents of 64.
// One instruction consumes 16 K rows = 2048 bytes.  The instruction descriptor must carry the matching major bit.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// General 16-bit instruction descriptor (D = fp32): operand formats (0 = fp16, 1 = bf16) and majors (0 = K, 1 = MN).
__host__ __device__ constexpr uint32_t umma_idesc_16bit(uint32_t M, uint32_t N, uint32_t a_bf16, uint32_t b_bf16,
                                                        uint32_t a_mn = 0, uint32_t b_mn = 0) {
  return (1u << 4) | (a_bf16 << 7) | (b_bf16 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Instruction descriptor: A = B = fp16, D = fp32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return umma_idesc_16bit(M, N, 0, 0);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 adds / multiplies in one instruction (sm_100 packed fp32)
__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(d));
}
__device__ __forceinline__ void mul_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  uint64_t a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

}  // namespace bseg
