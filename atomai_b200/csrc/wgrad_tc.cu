// wgrad_tc.cu — convolution weight gradient on tcgen05 (TF32 operands, fp32 accumulate in TMEM).
//
//   dW[co][ci][ty][tx] = sum_p dy[p][co] * x[p + (ty, tx) * dil][ci]        
//
// The reduction (GEMM K) runs over pixels, and both operands are contiguous along their channel
// axis in NHWC memory, i.e. "MN-major" for the tensor core.  For 32-bit operands the only legal
// MN-major shared-memory layout is SWIZZLE_128B_BASE32B: rows = k (pixels) at a 128 B pitch, 32
// channels per row, the four 32 B chunks of a row XOR-ed with (row index mod 4) — which is
// exactly a pixel-major NHWC tile.  The swizzle is a function of the absolute shared-memory
// address (pinned by the selftest variants >= 8 / >= 32), so a tap is only a different start
// address into the x halo tile, and one tcgen05.mma consumes 8 pixels (one row of a 16 x 8 tile).
//
// GEMM per kernel row ty:  D_ty[M = 128][N = co] += A[M][K = 8 px] * B[N][K]^T   with
//   B = dy tile           [32-channel chunk][128 px][128 B]       (LBO = chunk stride)
//   A = x halo tile, M = 4 chunks of 32 rows:
//        taps_w > 1 ("tap-stacked"): chunk c = the SAME 32 input channels shifted by c*dil pixels,
//                    i.e. LBO = dil * 128 B — the horizontal taps ride on the M axis, so a 3x3
//                    kernel costs 3 x 16 MMAs per tile instead of 9 x 16, and N = Cout is not
//                    padded to 128 (chunk 3 of a 3-wide kernel is a don't-care view);
//        taps_w == 1: chunk c = input channels 32c..32c+31 (LBO = chunk stride), CTA owns 128 ci.
// A CTA owns one input-channel block, one block of <= 128 output channels and a contiguous range
// of pixel tiles; its taps_h x N accumulators stay in TMEM for the whole kernel and are flushed
// once, with atomics, into dW (OIHW).  Warp roles: 0-3 epilogue, 4 MMA issue, 5 TMEM alloc, 8-15
// loaders (two groups of 4 warps, group g owns stage g; setmaxnreg moves registers to them).
//
// Replaces autograd's cuDNN bwd-filter behind loss.backward(), atomai/trainers/trainer.py:206.
#include "common.cuh"

namespace {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kNumEpiWarps = 4, kMmaWarp = 4, kAllocWarp = 5, kFirstLoadWarp = 8;
constexpr int kNumLoadWarps = 8, kGroupThreads = 128, kB = 8;
constexpr int kThreads = (kFirstLoadWarp + kNumLoadWarps) * 32;   // 512: 4 warpgroups
constexpr int kRegsEpi = 80, kRegsMma = 48, kRegsLoad = 192;      // setmaxnreg re-balancing
constexpr int kStages = 2;
constexpr int kChunk = 128 * 128;             // bytes of one [128 px][32 ch] chunk of the dy tile

struct WgradTcParams {
  SrcSet S;
  int N, H, W, Cout, Cin;
  int taps_h, taps_w, dil;
  const float* dy;
  int ld_dy;
  float* dw;
  int tiles_h, tiles_w, num_tiles;
  int cib;               // input channels per CTA: 32 (tap-stacked) or 128 (taps_w == 1)
  int n_cc;              // input-channel blocks
  int co_blocks;         // ceil(Cout / 128)
  int ranges;            // pixel-tile ranges per (cc, co_block)
  int TWp, THp, HP;
  int x_chunk;           // taps_w == 1: bytes between 32-channel chunks of the x tile
  int x_bytes, stage_bytes;
  int tmem_cols;
};

struct __align__(8) Ctl {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                        uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
// RN-to-TF32: the tensor core reads the upper 19 bits only (see conv_tc.cu tf32_bits)
__device__ __forceinline__ uint32_t tf32b(float x) { return __float_as_uint(x) + 0x1000u; }
// exact e / P for e < 65536 (mul = ceil(2^32 / P), P >= 2)
__device__ __forceinline__ uint32_t fdiv(uint32_t e, uint32_t P, uint32_t mul) {
  return P == 1 ? e : __umulhi(e, mul);
}
__host__ __device__ inline uint32_t fdiv_mul(uint32_t P) {
  return P <= 1 ? 0u : (uint32_t)(((1ull << 32) + P - 1) / P);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t base = (smem_u32(smem) + 128 + 1023) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool stacked = p.taps_w > 1;

  // work assignment: blockIdx.x -> (range r, cout block cb, channel block cc); cc fastest so the
  // CTAs that share a pixel range (and therefore the dy tiles) run at the same time.
  const int cc = blockIdx.x % p.n_cc;
  const int cb = (blockIdx.x / p.n_cc) % p.co_blocks;
  const int r = blockIdx.x / (p.n_cc * p.co_blocks);
  const int per = (p.num_tiles + p.ranges - 1) / p.ranges;
  const int t_begin = r * per;
  const int t_end = min(p.num_tiles, t_begin + per);
  const int co0 = cb * 128;
  const int co_n = min(128, p.Cout - co0);          // valid output channels in this block
  const int ci0 = cc * p.cib;
  const int ci_n = min(p.cib, p.Cin - ci0);         // valid input channels in this block
  const int PD = co_n >> 2, PX = ci_n >> 2;         // 16 B pieces per pixel
  const int Npad = (co_n + 31) & ~31;               // GEMM N (zero padded to the 32-wide atom)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), kNumLoadWarps / 2);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->done), 1);
    fence_barrier_init();
  }
  if (warp == kAllocWarp) tmem_alloc(smem_u32(&ctl->tmem_base), p.tmem_cols);
  // zero both stages once: channel padding (co >= co_n, ci >= ci_n) must contribute exact zeros,
  // and the don't-care rows behind the halo tile must at least be finite
  for (int i = threadIdx.x; i < kStages * p.stage_bytes / 16; i += kThreads)
    sts128u(base + i * 16, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp >= kFirstLoadWarp) {
    // ===================== loaders: dy tile + x halo tile =====================
    // Group g stages the tiles with (local index & 1) == g into stage g, so two tiles are in
    // flight per SM.  Loads are issued in branch-free batches of kB from clamped addresses.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsLoad));
    const int grp = (warp - kFirstLoadWarp) >> 2;
    const int gt = threadIdx.x - (kFirstLoadWarp + grp * 4) * 32;   // 0..127
    const int n_d = PD * 128;     // 16 B pieces of the dy tile
    const int n_x = PX * p.HP;    // 16 B pieces of the x halo block
    const int H = p.H, W = p.W;
    const uint32_t mulD = fdiv_mul(PD), mulX = fdiv_mul(PX), mulT = fdiv_mul(p.TWp);
    const int tpi = p.tiles_w * p.tiles_h;
    const uint32_t mulTpi = fdiv_mul(tpi), mulTw = fdiv_mul(p.tiles_w);
    // a thread's channel piece is the same for every element when P divides the group size:
    // hoist the source selection and the BN affine of the x operand
    const bool x_const = (kGroupThreads % PX) == 0;
    const uint32_t x0 = base + grp * p.stage_bytes, d0 = x0 + p.x_bytes;
    uint32_t use = 0;
    for (int tile = t_begin + grp; tile < t_end; tile += 2, ++use) {
      const int n = (int)fdiv(tile, tpi, mulTpi);
      const int rem = tile - n * tpi;
      const int th_i = (int)fdiv(rem, p.tiles_w, mulTw);
      const int tw_i = rem - th_i * p.tiles_w;
      const int h0 = th_i * kTileH, w0 = tw_i * kTileW;
      const int h_org = h0 - p.dil * (p.taps_h >> 1), w_org = w0 - p.dil * (p.taps_w >> 1);
      mbar_wait(smem_u32(&ctl->empty[grp]), (use & 1) ^ 1);
      const size_t img = (size_t)n * H;
      const bool full_tile = h0 + kTileH <= H && w0 + kTileW <= W;
      // ---- dy tile: piece e -> (pixel q = e / PD, channel piece j = e % PD)
      const float* dyb = p.dy + ((img + h0) * W + w0) * p.ld_dy + co0;
      for (int e0 = gt; e0 < n_d; e0 += kB * kGroupThreads) {
        float4 v[kB];
        uint32_t ok = 0;
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const uint32_t e = min(e0 + k * kGroupThreads, n_d - 1);
          const uint32_t q = fdiv(e, PD, mulD), j = e - q * PD;
          int rr = q >> 3, cw = q & 7;
          bool m = true;
          if (!full_tile) {
            m = h0 + rr < H && w0 + cw < W;
            rr = min(rr, H - 1 - h0);
            cw = min(cw, W - 1 - w0);
          }
          v[k] = __ldg(reinterpret_cast<const float4*>(
              dyb + ((size_t)rr * W + cw) * p.ld_dy + j * 4));
          ok |= (m ? 1u : 0u) << k;
        }
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const uint32_t e = e0 + k * kGroupThreads;
          if (e < (uint32_t)n_d) {
            const uint32_t q = fdiv(e, PD, mulD), j = e - q * PD;
            const uint32_t msk = (ok >> k) & 1u ? 0xFFFFFFFFu : 0u;
            sts128u(swz128_32(d0 + (j >> 3) * kChunk + q * 128 + (j & 7) * 16),
                    tf32b(v[k].x) & msk, tf32b(v[k].y) & msk, tf32b(v[k].z) & msk,
                    tf32b(v[k].w) & msk);
          }
        }
      }
      // ---- x halo block: piece e -> (halo pixel q = e / PX, channel piece j = e % PX), BN
      // affine / 2x2 max-pool / zero padding applied on load (normalise-on-load)
      const bool interior =
          h_org >= 0 && w_org >= 0 && h_org + p.THp <= H && w_org + p.TWp <= W;
      const SrcDev* spc = &p.S.s[0];
      int cch = ci0 + (gt % PX) * 4;
      if (p.S.nsrc > 1 && cch >= p.S.s[0].C) { spc = &p.S.s[1]; cch -= p.S.s[0].C; }
      const bool fast = x_const && interior && !spc->pool;
      if (fast) {
        const uint32_t ld = spc->ld;
        const float* xb = spc->ptr + ((img + h_org) * W + w_org) * ld + cch;
        const bool aff = spc->scale != nullptr;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (aff) {
          sc = __ldg(reinterpret_cast<const float4*>(spc->scale + cch));
          sh = __ldg(reinterpret_cast<const float4*>(spc->shift + cch));
        }
        for (int e0 = gt; e0 < n_x; e0 += kB * kGroupThreads) {
          float4 v[kB];
#pragma unroll
          for (int k = 0; k < kB; ++k) {
            const uint32_t e = min(e0 + k * kGroupThreads, n_x - 1);
            const uint32_t q = fdiv(e, PX, mulX);
            const uint32_t hh = fdiv(q, p.TWp, mulT), ww = q - hh * p.TWp;
            v[k] = __ldg(reinterpret_cast<const float4*>(xb + (hh * W + ww) * ld));
          }
#pragma unroll
          for (int k = 0; k < kB; ++k) {
            const uint32_t e = e0 + k * kGroupThreads;
            if (e < (uint32_t)n_x) {
              const uint32_t q = fdiv(e, PX, mulX), j = e - q * PX;
              const uint32_t dst = stacked ? x0 + q * 128 + j * 16
                                           : x0 + (j >> 3) * p.x_chunk + q * 128 + (j & 7) * 16;
              sts128u(swz128_32(dst), tf32b(fmaf(v[k].x, sc.x, sh.x)),
                      tf32b(fmaf(v[k].y, sc.y, sh.y)), tf32b(fmaf(v[k].z, sc.z, sh.z)),
                      tf32b(fmaf(v[k].w, sc.w, sh.w)));
            }
          }
        }
      } else {
        for (int e0 = gt; e0 < n_x; e0 += kB * kGroupThreads) {
          float4 v[kB];
#pragma unroll
          for (int k = 0; k < kB; ++k) {
            const uint32_t e = min(e0 + k * kGroupThreads, n_x - 1);
            const uint32_t q = fdiv(e, PX, mulX), j = e - q * PX;
            const uint32_t hh = fdiv(q, p.TWp, mulT), ww = q - hh * p.TWp;
            v[k] = load_src4(p.S, n, h_org + (int)hh, w_org + (int)ww, H, W, ci0 + j * 4);
          }
#pragma unroll
          for (int k = 0; k < kB; ++k) {
            const uint32_t e = e0 + k * kGroupThreads;
            if (e < (uint32_t)n_x) {
              const uint32_t q = fdiv(e, PX, mulX), j = e - q * PX;
              const uint32_t dst = stacked ? x0 + q * 128 + j * 16
                                           : x0 + (j >> 3) * p.x_chunk + q * 128 + (j & 7) * 16;
              sts128u(swz128_32(dst), tf32b(v[k].x), tf32b(v[k].y), tf32b(v[k].z), tf32b(v[k].w));
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->full[grp]));
    }
  } else if (warp >= kNumEpiWarps) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMma));
   if (warp == kMmaWarp) {
    if (elect_one()) {   // one elected thread issues every MMA / commit
      const uint32_t idesc = umma_idesc_tf32(128, Npad, 1, 1);
      const uint64_t a_tmpl = umma_desc_ex(0, stacked ? p.dil * 128 : p.x_chunk, 512, 1, 0);
      const uint64_t b_tmpl = umma_desc_ex(0, kChunk, 512, 1, 0);
      const uint32_t a_row16 = stacked ? (uint32_t)p.TWp * 8 : 64u;   // next tile row of x
      const uint32_t a_ty16 = (uint32_t)(p.dil * p.TWp) * 8;          // next kernel row
      const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
      const int th = p.taps_h;
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const uint32_t st = it & 1;
        mbar_wait(smem_u32(&ctl->full[st]), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t x0 = base + st * p.stage_bytes, d0 = x0 + p.x_bytes;
        uint32_t a_ty = (uint32_t)a_tmpl + (x0 >> 4);
        const uint32_t b0 = (uint32_t)b_tmpl + (d0 >> 4);
        uint32_t dcol = tmem_base;
        for (int ty = 0; ty < th; ++ty, a_ty += a_ty16, dcol += Npad) {
          uint32_t ad = a_ty, bd = b0;
          uint32_t accum = it > 0 ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < kTileH; ++h) {
            umma_tf32_lh(dcol, ad, a_hi, bd, b_hi, idesc, accum);
            accum = 1u;
            ad += a_row16;         // next halo row (8 output pixels further down)
            bd += 64;              // next 8 pixels of the dy tile (8 x 128 B)
          }
        }
        umma_commit(smem_u32(&ctl->empty[st]));
      }
      umma_commit(smem_u32(&ctl->done));
    }
    __syncwarp();
   }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    // ===================== epilogue: TMEM -> atomics into dW (OIHW) =====================
    // accumulator row m = chunk * 32 + lane: chunk = warp = horizontal tap (stacked) or 32-channel
    // sub-block; column = ty * Npad + (co - co0)
    if (t_end > t_begin) {
      mbar_wait(smem_u32(&ctl->done), 0);
      tc_fence_after();
      const int taps = p.taps_h * p.taps_w;
      const int tx = stacked ? warp : 0;
      const int ci = stacked ? ci0 + lane : ci0 + warp * 32 + lane;
      const bool row_ok = ci < ci0 + ci_n && tx < p.taps_w;
      if (stacked ? warp < p.taps_w : warp * 32 < ci_n) {
        for (int ty = 0; ty < p.taps_h; ++ty) {
          for (int c0 = 0; c0 < co_n; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ty * Npad + c0 + ((uint32_t)(warp * 32) << 16), v);
            if (row_ok) {
              float* o = p.dw + ((size_t)(co0 + c0) * p.Cin + ci) * taps + ty * p.taps_w + tx;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < co_n) atomicAdd(o + (size_t)i * p.Cin * taps, v[i]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int wgrad_plan(const ab_conv_t* d, WgradTcParams* p, int* smem_bytes) {
  if (ab_make_srcset(d, &p->S)) return 1;
  p->N = d->N; p->H = d->H; p->W = d->W; p->Cout = d->Cout; p->Cin = p->S.Ctot;
  p->taps_h = d->ks_h; p->taps_w = d->ks_w; p->dil = d->dil;
  AB_CHECK(d->ks_w <= 4 && d->ks_h <= 4, "wgrad_tc: kernel %dx%d too large", d->ks_h, d->ks_w);
  p->tiles_h = (d->H + kTileH - 1) / kTileH;
  p->tiles_w = (d->W + kTileW - 1) / kTileW;
  p->num_tiles = d->N * p->tiles_h * p->tiles_w;
  AB_CHECK((uint64_t)p->num_tiles * (uint64_t)(p->tiles_h * p->tiles_w) < (1ull << 32),
           "wgrad_tc: too many tiles");
  p->THp = kTileH + d->dil * (d->ks_h - 1);
  p->TWp = kTileW + d->dil * (d->ks_w - 1);
  p->HP = p->THp * p->TWp;
  const bool stacked = d->ks_w > 1;
  p->co_blocks = (d->Cout + 127) / 128;
  const int co_max = d->Cout < 128 ? d->Cout : 128;
  const int npad = (co_max + 31) & ~31;
  const int d_bytes = (npad / 32) * kChunk;
  // A stage is [x tile][dy tile]: the don't-care rows of the M = 128 operand (chunk 3 of a 3-wide
  // kernel, chunks >= cib/32 of a 1x1 kernel) then alias the dy tile — finite and in bounds.
  if (stacked) {
    p->cib = 32;
    p->x_chunk = 0;
    p->x_bytes = (p->HP * 128 + 1023) & ~1023;
    AB_CHECK(3 * d->dil * 128 + 1024 <= d_bytes, "wgrad_tc: dilation %d too large", d->dil);
  } else {
    p->x_chunk = (p->HP * 128 + 1023) & ~1023;
    p->cib = 128;
    if (kStages * (4 * p->x_chunk + d_bytes) + 2048 > 225 * 1024) p->cib = 64;
    p->x_bytes = (p->cib / 32) * p->x_chunk;
    AB_CHECK(p->cib == 128 || 2 * p->x_chunk <= d_bytes, "wgrad_tc: no shared-memory plan");
  }
  p->stage_bytes = p->x_bytes + d_bytes;
  *smem_bytes = kStages * p->stage_bytes + 128 + 1024;
  AB_CHECK(*smem_bytes <= 225 * 1024, "wgrad_tc: halo tile too large for shared memory (dil=%d)",
           d->dil);
  p->n_cc = (p->Cin + p->cib - 1) / p->cib;
  int cols = 32;
  while (cols < d->ks_h * npad) cols <<= 1;
  AB_CHECK(cols <= 512, "wgrad_tc: too many accumulator columns");
  p->tmem_cols = cols;
  const int groups = p->n_cc * p->co_blocks;
  int ranges = ab_num_sms() / groups;
  if (ranges < 1) ranges = 1;
  if (ranges > p->num_tiles) ranges = p->num_tiles > 0 ? p->num_tiles : 1;
  p->ranges = ranges;
  return 0;
}

}  // namespace

int ab_wgrad_tc_supported(const ab_conv_t* d) {
  for (int i = 0; i < d->nsrc; ++i) {
    if (d->src[i].C % 4 != 0 || d->src[i].ld % 4 != 0) return 0;
    if (((uintptr_t)d->src[i].ptr & 15) != 0) return 0;
  }
  if (d->Cout % 4 != 0) return 0;
  if (d->ks_w > 4 || d->ks_h > 4) return 0;
  WgradTcParams p;
  int smem = 0;
  if (wgrad_plan(d, &p, &smem)) return 0;
  return 1;
}

int ab_conv_tc_wgrad(const ab_conv_t* d, const float* dy, int ld_dy, float* dw,
                     cudaStream_t stream) {
  WgradTcParams p;
  int smem = 0;
  if (wgrad_plan(d, &p, &smem)) return 1;
  AB_CHECK(((uintptr_t)dy & 15) == 0 && ld_dy % 4 == 0, "wgrad_tc: unaligned dy");
  p.dy = dy; p.ld_dy = ld_dy; p.dw = dw;
  if (p.num_tiles == 0) return 0;
  static int configured = 0;
  if (!configured) {
    AB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 226 * 1024));
    configured = 1;
  }
  const int grid = p.ranges * p.n_cc * p.co_blocks;
  wgrad_tc_kernel<<<grid, kThreads, smem, stream>>>(p);
  AB_LAUNCH_CHECK();
  return 0;
}
