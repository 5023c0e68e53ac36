// common.cuh — error plumbing, PTX wrappers (mbarrier / bulk copy / tcgen05) and
// the shared "source loader" used by every convolution-family kernel.
// sm_100a only.  Test infrastructure lives elsewhere; this is product code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/atomai_b200.h"

// ---------------------------------------------------------------- host errors
void ab_set_error(const char* fmt, ...);
#define AB_CHECK(cond, ...)            \
  do {                                 \
    if (!(cond)) {                     \
      ab_set_error(__VA_ARGS__);       \
      return 1;                        \
    }                                  \
  } while (0)
#define AB_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ab_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                   __LINE__);                                                      \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define AB_LAUNCH_CHECK() AB_CUDA(cudaGetLastError())

int ab_num_sms();  // cached SM count of the current device
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: `done` is a per-kernel table
// indexed by the current device (set once per device, benign if two threads race)
int ab_optin_smem(const void* func, int bytes, unsigned char* done);

// ---------------------------------------------------------------- device utils
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float lrelu_f(float v, float a) { return v > 0.f ? v : v * a; }
// forward activation and its derivative expressed through the OUTPUT value a = act(pre)
// AB_ACT_RBF / AB_ACT_MATERN25: v is a squared scaled distance, slope the outputscale
// (gpytorch RBFKernel / MaternKernel(nu=2.5) under a ScaleKernel, atomai/nets/gp.py:41-46,100-111)
__device__ __forceinline__ float act_f(float v, int act, float slope) {
  if (act == AB_ACT_TANH) return tanhf(v);
  if (act == AB_ACT_RBF) return slope * expf(-0.5f * fmaxf(v, 0.f));
  if (act == AB_ACT_MATERN25) {
    const float d2 = fmaxf(v, 0.f);
    const float s5r = 2.2360679775f * sqrtf(d2);
    return slope * (1.f + s5r + 1.6666666667f * d2) * expf(-s5r);
  }
  return v > 0.f ? v : v * slope;
}
__device__ __forceinline__ float act_grad_from_out(float a, int act, float slope) {
  return act == AB_ACT_TANH ? 1.f - a * a : (a > 0.f ? 1.f : slope);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the
// hint expires) instead of returning at once, so a waiting warp does not hammer the shared-memory
// pipe the tensor core reads its operands through (ncu, round 2: mbarrier polling was 12 % of the
// L1 data-pipe wavefronts of conv_tc, next to 39 % of tcgen05 operand reads).
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(200000u)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never a hung GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    // wall-clock bound (20 s), not a spin count: under a profiler's instrumented replay a healthy
    // kernel can run two orders of magnitude slower, and the suspend hint is only an upper limit
    if ((++spins & 255u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 20000000000ull) {
        printf("atomai_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
               blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// One leader lane of a fully converged warp (the same lane every time): code around it stays
// warp-uniform, so descriptors live in uniform registers and no waterfall loops are generated.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the descriptors split into 32-bit halves: only the low word (start address field)
// changes between MMAs, so the issuing thread advances two 32-bit counters per MMA.
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                             uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (BF16 operands, fp32 accumulate), same split-descriptor form: the correction MMA of
// the tf32x3 mode accumulates into the same TMEM tile as the kind::tf32 MMA before it.
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                             uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread -> one arrive on `bar` when they retire
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns; warp w may touch lanes 32*(w%4)..+31 only
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, SWIZZLE_NONE ("interleave") canonical layouts:
//   K-major : ((8,m),(4,2)) : ((16B, SBO), (4B, LBO))   — 8-row core matrices of 16B rows
//   MN-major: ((4,m),(8,k)) : ((4B, SBO), (16B, LBO))
// Fields in 16-byte units; bit 46 = descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Same, with an explicit layout type (1 = SWIZZLE_128B_BASE32B, the only legal layout for
// MN-major TF32 operands) and swizzle base offset.
__device__ __forceinline__ uint64_t umma_desc_ex(uint32_t saddr, uint32_t lbo_bytes,
                                                 uint32_t sbo_bytes, uint32_t layout_type,
                                                 uint32_t base_offset) {
  uint64_t d = umma_desc(saddr, lbo_bytes, sbo_bytes);
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
// byte-address swizzle of SWIZZLE_128B_BASE32B: 32 B chunk index ^= (128 B row index mod 4)
__device__ __forceinline__ uint32_t swz128_32(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }
// Instruction descriptor, kind::tf32, fp32 accumulate.
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                        // D format: F32
  d |= 2u << 7;                        // A format: TF32
  d |= 2u << 10;                       // B format: TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// Instruction descriptor, kind::f16 with BF16 operands (K = 16 per MMA), fp32 accumulate.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                        // D format: F32
  d |= 1u << 7;                        // A format: BF16
  d |= 1u << 10;                       // B format: BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// two floats -> packed bf16x2 (lo half = first argument), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- source loader
struct SrcDev {
  const float* ptr;
  const float* scale;
  const float* shift;
  int C;
  int ld;
  int pool;
};
struct SrcSet {
  SrcDev s[2];
  int nsrc;
  int Ctot;
};

// ---- 2x upsampling on load (F.interpolate(scale_factor=2, mode), atomai/nets/blocks.py:130-131,
// fused into the consumer convolution's loader: source code pool = AB_SRC_UP_BILINEAR / _NEAREST,
// the source tensor is (H/2, W/2)).  For output coordinate o of an axis with n_lo low-res samples:
// the two contributing samples and the weight of the second (align_corners = False:
// src = (o + 0.5)/2 - 0.5 clamped at 0 -> even o = 2i: 0.25 x[i-1] + 0.75 x[i], odd: 0.75 x[i] +
// 0.25 x[i+1], edges clamped).
#define AB_SRC_POOL 1
#define AB_SRC_UP_BILINEAR 2
#define AB_SRC_UP_NEAREST 3
__device__ __forceinline__ void up2_coord(int o, int n_lo, bool bilinear, int& i0, int& i1,
                                          float& l1) {
  const int i = o >> 1;
  if (!bilinear) { i0 = i1 = i; l1 = 0.f; return; }
  if (o & 1) { i0 = i; i1 = min(i + 1, n_lo - 1); l1 = 0.25f; }
  else { i0 = max(i - 1, 0); i1 = i; l1 = 0.75f; }     // (o = 0: i0 = i1 = 0, any weight gives x[0])
}
__device__ __forceinline__ float lerp2(float v00, float v01, float v10, float v11, float lh,
                                       float lw) {
  // same association as ATen's upsample_bilinear2d kernel
  return (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
}
__device__ __forceinline__ float4 lerp2_4(float4 a, float4 b, float4 c, float4 d, float lh, float lw) {
  return make_float4(lerp2(a.x, b.x, c.x, d.x, lh, lw), lerp2(a.y, b.y, c.y, d.y, lh, lw),
                     lerp2(a.z, b.z, c.z, d.z, lh, lw), lerp2(a.w, b.w, c.w, d.w, lh, lw));
}
// upsampled value (before the affine) of 4 channels at hi-res pixel (n, h, w) of an H x W grid
__device__ __forceinline__ float4 load_up4(const float* __restrict__ ptr_c, int ld, int mode, int n,
                                           int h, int w, int H, int W) {
  const int Hl = H >> 1, Wl = W >> 1;
  int h0, h1, w0, w1;
  float lh, lw;
  up2_coord(h, Hl, mode == AB_SRC_UP_BILINEAR, h0, h1, lh);
  up2_coord(w, Wl, mode == AB_SRC_UP_BILINEAR, w0, w1, lw);
  const float* r0 = ptr_c + (size_t)(n * Hl + h0) * Wl * ld;
  const float* r1 = ptr_c + (size_t)(n * Hl + h1) * Wl * ld;
  const float4 a = __ldg(reinterpret_cast<const float4*>(r0 + (size_t)w0 * ld));
  if (mode != AB_SRC_UP_BILINEAR) return a;
  const float4 b = __ldg(reinterpret_cast<const float4*>(r0 + (size_t)w1 * ld));
  const float4 c = __ldg(reinterpret_cast<const float4*>(r1 + (size_t)w0 * ld));
  const float4 d = __ldg(reinterpret_cast<const float4*>(r1 + (size_t)w1 * ld));
  return lerp2_4(a, b, c, d, lh, lw);
}

static __device__ __noinline__ float4 load_up4_noinline(const float* __restrict__ ptr_c, int ld, int mode,
                                                 int n, int h, int w, int H, int W) {
  return load_up4(ptr_c, ld, mode, n, h, w, H, W);
}
static __device__ __noinline__ float load_up1_noinline(const float* __restrict__ ptr_c, int ld, int mode,
                                                int n, int h, int w, int H, int W) {
  const int Hl = H >> 1, Wl = W >> 1;
  int h0, h1, w0, w1;
  float lh, lw;
  up2_coord(h, Hl, mode == AB_SRC_UP_BILINEAR, h0, h1, lh);
  up2_coord(w, Wl, mode == AB_SRC_UP_BILINEAR, w0, w1, lw);
  const float* r0 = ptr_c + (size_t)(n * Hl + h0) * Wl * ld;
  const float* r1 = ptr_c + (size_t)(n * Hl + h1) * Wl * ld;
  return lerp2(__ldg(r0 + (size_t)w0 * ld), __ldg(r0 + (size_t)w1 * ld), __ldg(r1 + (size_t)w0 * ld),
               __ldg(r1 + (size_t)w1 * ld), lh, lw);
}

// 4 consecutive channels [c, c+4) of the logical (post-affine, post-pool, zero padded)
// input at pixel (n, h, w) of an H x W grid.  c must be a multiple of 4 and every source's
// C a multiple of 4 on this vector path.
__device__ __forceinline__ float4 load_src4(const SrcSet& S, int n, int h, int w, int H, int W,
                                            int c) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) return r;
  const SrcDev* s = &S.s[0];
  if (S.nsrc > 1 && c >= S.s[0].C) {
    s = &S.s[1];
    c -= S.s[0].C;
  }
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s->scale) {
    sc = __ldg(reinterpret_cast<const float4*>(s->scale + c));
    sh = __ldg(reinterpret_cast<const float4*>(s->shift + c));
  }
  if (s->pool >= AB_SRC_UP_BILINEAR) {
    // out of line: inlined, the four extra loads and the interpolation cost the thin-channel
    // kernels 12-20 registers (conv_pix<16>: 128 -> 148, one CTA per SM instead of two)
    const float4 v = load_up4_noinline(s->ptr + c, s->ld, s->pool, n, h, w, H, W);
    r.x = fmaf(v.x, sc.x, sh.x);
    r.y = fmaf(v.y, sc.y, sh.y);
    r.z = fmaf(v.z, sc.z, sh.z);
    r.w = fmaf(v.w, sc.w, sh.w);
  } else if (!s->pool) {
    const float4 v = __ldg(
        reinterpret_cast<const float4*>(s->ptr + ((size_t)(n * H + h) * W + w) * s->ld + c));
    r.x = fmaf(v.x, sc.x, sh.x);
    r.y = fmaf(v.y, sc.y, sh.y);
    r.z = fmaf(v.z, sc.z, sh.z);
    r.w = fmaf(v.w, sc.w, sh.w);
  } else {
    const int H2 = 2 * H, W2 = 2 * W;
    const float* base = s->ptr + ((size_t)(n * H2 + 2 * h) * W2 + 2 * w) * s->ld + c;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(base));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + s->ld));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + (size_t)W2 * s->ld));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + (size_t)W2 * s->ld + s->ld));
    r.x = fmaxf(fmaxf(fmaf(v0.x, sc.x, sh.x), fmaf(v1.x, sc.x, sh.x)),
                fmaxf(fmaf(v2.x, sc.x, sh.x), fmaf(v3.x, sc.x, sh.x)));
    r.y = fmaxf(fmaxf(fmaf(v0.y, sc.y, sh.y), fmaf(v1.y, sc.y, sh.y)),
                fmaxf(fmaf(v2.y, sc.y, sh.y), fmaf(v3.y, sc.y, sh.y)));
    r.z = fmaxf(fmaxf(fmaf(v0.z, sc.z, sh.z), fmaf(v1.z, sc.z, sh.z)),
                fmaxf(fmaf(v2.z, sc.z, sh.z), fmaf(v3.z, sc.z, sh.z)));
    r.w = fmaxf(fmaxf(fmaf(v0.w, sc.w, sh.w), fmaf(v1.w, sc.w, sh.w)),
                fmaxf(fmaf(v2.w, sc.w, sh.w), fmaf(v3.w, sc.w, sh.w)));
  }
  return r;
}

// scalar variant (any channel count; used by the exact-fp32 kernels for odd C)
__device__ __forceinline__ float load_src1(const SrcSet& S, int n, int h, int w, int H, int W,
                                           int c) {
  if ((unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) return 0.f;
  const SrcDev* s = &S.s[0];
  if (S.nsrc > 1 && c >= S.s[0].C) {
    s = &S.s[1];
    c -= S.s[0].C;
  }
  float sc = 1.f, sh = 0.f;
  if (s->scale) {
    sc = __ldg(s->scale + c);
    sh = __ldg(s->shift + c);
  }
  if (s->pool >= AB_SRC_UP_BILINEAR)
    return fmaf(load_up1_noinline(s->ptr + c, s->ld, s->pool, n, h, w, H, W), sc, sh);
  if (!s->pool) {
    return fmaf(__ldg(s->ptr + ((size_t)(n * H + h) * W + w) * s->ld + c), sc, sh);
  }
  const int H2 = 2 * H, W2 = 2 * W;
  const float* base = s->ptr + ((size_t)(n * H2 + 2 * h) * W2 + 2 * w) * s->ld + c;
  const float v0 = fmaf(__ldg(base), sc, sh), v1 = fmaf(__ldg(base + s->ld), sc, sh);
  const float v2 = fmaf(__ldg(base + (size_t)W2 * s->ld), sc, sh);
  const float v3 = fmaf(__ldg(base + (size_t)W2 * s->ld + s->ld), sc, sh);
  return fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
}

int ab_make_srcset(const ab_conv_t* d, SrcSet* out);  // validates + converts (host)
