// wgrad_tc.cu — convolution weight gradient on tcgen05 (TF32 operands, fp32 accumulate in TMEM).
//
//   dW[co][ci][tap] = sum_p dy[p][co] * x[p + tap][ci]        
//
// GEMM view per tap: D[M = co][N = ci] += A[M][K] * B[N][K]^T with K = pixels.  Both operands are
// contiguous along their channel (M / N) axis in NHWC memory, i.e. "MN-major" for the tensor
// core.  For 32-bit operands the only legal MN-major shared-memory layout is
// SWIZZLE_128B_BASE32B: rows = k (pixels) at a 128 B pitch, 32 channels per row, the four 32 B
// chunks of a row XOR-ed with (row index mod 4) — which is exactly a pixel-major NHWC tile.  The
// swizzle is a function of the absolute shared-memory address (pinned by tests/test_umma_layouts),
// so a convolution tap is again only a different start address into the halo tile, and one
// tcgen05.mma consumes 8 pixels (one 8-wide row of the 16 x 8 tile).
//
// A CTA owns one block of 32 input channels and a contiguous range of pixel tiles; its
// taps x 32 accumulators live in TMEM for the whole kernel and are flushed once, with atomics,
// into dW (OIHW).  Warp roles: 0-3 epilogue, 4 MMA issue, 5 TMEM alloc, 8-15 loaders (two groups
// of 4 warps on alternate tiles; setmaxnreg moves registers to them).
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
constexpr int kNB = 32;                       // input channels per CTA (one 128 B swizzle row)
constexpr int kDChunk = 128 * 128;            // bytes between 32-channel chunks of the dy tile
constexpr int kDBytes = 4 * kDChunk;          // M = 128 output channels (zero padded)

struct WgradTcParams {
  SrcSet S;
  int N, H, W, Cout, Cin;
  int taps_h, taps_w, dil;
  const float* dy;
  int ld_dy;
  float* dw;
  int tiles_h, tiles_w, num_tiles;
  int n_cc;              // input-channel blocks
  int co_blocks;         // ceil(Cout / 128)
  int ranges;            // pixel-tile ranges per (cc, co_block)
  int TWp, THp, HP;
  int x_bytes, stage_bytes;
  int tmem_cols;
};

struct __align__(8) Ctl {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t base = (smem_u32(smem) + 128 + 1023) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.taps_h * p.taps_w;

  // work assignment: blockIdx.x -> (range r, cout block cb, channel block cc); cc fastest so the
  // CTAs that share a pixel range (and therefore the dy tiles) run at the same time.
  const int cc = blockIdx.x % p.n_cc;
  const int cb = (blockIdx.x / p.n_cc) % p.co_blocks;
  const int r = blockIdx.x / (p.n_cc * p.co_blocks);
  const int per = (p.num_tiles + p.ranges - 1) / p.ranges;
  const int t_begin = r * per;
  const int t_end = min(p.num_tiles, t_begin + per);
  const int co0 = cb * 128;
  const int co_n = min(128, p.Cout - co0);      // valid output channels in this block
  const int ci_n = min(kNB, p.Cin - cc * kNB);  // valid input channels in this block
  const int PD = co_n >> 2, PX = ci_n >> 2;     // 16 B pieces per pixel

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), kNumLoadWarps / 2);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->done), 1);
    fence_barrier_init();
  }
  if (warp == kAllocWarp) tmem_alloc(smem_u32(&ctl->tmem_base), p.tmem_cols);
  // zero both stages once: channel padding (co >= co_n, ci >= ci_n) must contribute exact zeros
  for (int i = threadIdx.x; i < kStages * p.stage_bytes / 16; i += kThreads)
    sts128(base + i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp >= kFirstLoadWarp) {
    // ===================== loaders: dy tile + x halo tile =====================
    // Two groups of 4 warps stage alternate pixel tiles, so two tiles are in flight per SM.  Loads
    // are issued in branch-free batches of kB from clamped addresses and masked afterwards.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsLoad));
    const int grp = (warp - kFirstLoadWarp) >> 2;
    const int gt = threadIdx.x - (kFirstLoadWarp + grp * 4) * 32;   // 0..127
    const int n_d = PD * 128;     // 16 B pieces of the dy tile
    const int n_x = PX * p.HP;    // 16 B pieces of the x halo block
    const int H = p.H, W = p.W;
    uint32_t it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      if ((int)(it & 1) != grp) continue;
      const int tw_i = tile % p.tiles_w;
      const int th_i = (tile / p.tiles_w) % p.tiles_h;
      const int n = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th_i * kTileH, w0 = tw_i * kTileW;
      const int h_org = h0 - p.dil * (p.taps_h >> 1), w_org = w0 - p.dil * (p.taps_w >> 1);
      const uint32_t st = it % kStages;
      mbar_wait(smem_u32(&ctl->empty[st]), ((it / kStages) & 1) ^ 1);
      const uint32_t d0 = base + st * p.stage_bytes, x0 = d0 + kDBytes;
      const size_t img = (size_t)n * H;
      // ---- dy tile: piece e -> (pixel q = e / PD, channel piece j = e % PD)
      for (int e0 = gt; e0 < n_d; e0 += kB * kGroupThreads) {
        float4 v[kB];
        uint32_t ok = 0;
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const int e = min(e0 + k * kGroupThreads, n_d - 1);
          const int q = e / PD, j = e - q * PD;
          const int gh = h0 + (q >> 3), gw = w0 + (q & 7);
          const int ghc = min(gh, H - 1), gwc = min(gw, W - 1);
          v[k] = __ldg(reinterpret_cast<const float4*>(
              p.dy + ((img + ghc) * W + gwc) * p.ld_dy + co0 + j * 4));
          ok |= (gh < H && gw < W ? 1u : 0u) << k;
        }
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const int e = e0 + k * kGroupThreads;
          if (e < n_d) {
            const int q = e / PD, j = e - q * PD;
            const bool m = (ok >> k) & 1u;
            sts128(swz128_32(d0 + (j >> 3) * kDChunk + q * 128 + (j & 7) * 16),
                   make_float4(m ? to_tf32(v[k].x) : 0.f, m ? to_tf32(v[k].y) : 0.f,
                               m ? to_tf32(v[k].z) : 0.f, m ? to_tf32(v[k].w) : 0.f));
          }
        }
      }
      // ---- x halo block through the normalise-on-load source loader
      for (int e0 = gt; e0 < n_x; e0 += kB * kGroupThreads) {
        float4 v[kB];
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const int e = min(e0 + k * kGroupThreads, n_x - 1);
          const int q = e / PX, j = e - q * PX;
          const int hh = q / p.TWp, ww = q - hh * p.TWp;
          v[k] = load_src4(p.S, n, h_org + hh, w_org + ww, H, W, cc * kNB + j * 4);
        }
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const int e = e0 + k * kGroupThreads;
          if (e < n_x) {
            const int q = e / PX, j = e - q * PX;
            sts128(swz128_32(x0 + q * 128 + j * 16),
                   make_float4(to_tf32(v[k].x), to_tf32(v[k].y), to_tf32(v[k].z), to_tf32(v[k].w)));
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->full[st]));
    }
  } else if (warp >= kNumEpiWarps) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMma));
   if (warp == kMmaWarp) {
    if (elect_one()) {   // one elected thread issues every MMA / commit
      const uint32_t idesc = umma_idesc_tf32(128, kNB, 1, 1);
      const uint64_t a_tmpl = umma_desc_ex(0, kDChunk, 512, 1, 0);
      const uint64_t b_tmpl = umma_desc_ex(0, 1024, 512, 1, 0);
      const uint32_t b_row16 = (uint32_t)p.TWp * 8;          // one halo row = TWp * 128 B
      const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const uint32_t st = it % kStages;
        mbar_wait(smem_u32(&ctl->full[st]), (it / kStages) & 1);
        tc_fence_after();
        const uint32_t d0 = base + st * p.stage_bytes, x0 = d0 + kDBytes;
        const uint32_t a0 = (uint32_t)a_tmpl + (d0 >> 4), b0 = (uint32_t)b_tmpl + (x0 >> 4);
        for (int t = 0; t < taps; ++t) {
          const int ty = t / p.taps_w, tx = t - ty * p.taps_w;
          uint32_t ad = a0;
          uint32_t bd = b0 + (uint32_t)((ty * p.dil * p.TWp + tx * p.dil) * 8);
          const uint32_t dcol = tmem_base + t * kNB;
          uint32_t accum = it > 0 ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < kTileH; ++h) {
            umma_tf32_lh(dcol, ad, a_hi, bd, b_hi, idesc, accum);
            accum = 1u;
            ad += 64;              // next 8 pixels of the dy tile (8 x 128 B)
            bd += b_row16;         // next halo row
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
    if (t_end > t_begin) {
      mbar_wait(smem_u32(&ctl->done), 0);
      tc_fence_after();
      const int co = co0 + warp * 32 + lane;
      for (int t = 0; t < taps; ++t) {
        for (int c0 = 0; c0 < kNB; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + t * kNB + c0 + ((uint32_t)(warp * 32) << 16), v);
          if (co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = cc * kNB + c0 + i;
              if (ci < p.Cin) atomicAdd(p.dw + ((size_t)co * p.Cin + ci) * taps + t, v[i]);
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
  p->tiles_h = (d->H + kTileH - 1) / kTileH;
  p->tiles_w = (d->W + kTileW - 1) / kTileW;
  p->num_tiles = d->N * p->tiles_h * p->tiles_w;
  p->THp = kTileH + d->dil * (d->ks_h - 1);
  p->TWp = kTileW + d->dil * (d->ks_w - 1);
  p->HP = p->THp * p->TWp;
  const int taps = d->ks_h * d->ks_w;
  p->co_blocks = (d->Cout + 127) / 128;
  p->x_bytes = (p->HP * 128 + 1023) & ~1023;
  p->stage_bytes = kDBytes + p->x_bytes;
  *smem_bytes = kStages * p->stage_bytes + 128 + 1024;
  AB_CHECK(*smem_bytes <= 225 * 1024, "wgrad_tc: halo tile too large for shared memory (dil=%d)",
           d->dil);
  p->n_cc = (p->Cin + kNB - 1) / kNB;
  int cols = 32;
  while (cols < taps * kNB) cols <<= 1;
  AB_CHECK(cols <= 512, "wgrad_tc: too many taps");
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
