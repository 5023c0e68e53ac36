// elementwise.cu — the HBM-bound kernels around the convolutions: BatchNorm finalize / backward,
// LeakyReLU backward, 2x2 max-pool, bilinear/nearest x2 upsampling, gradient routing, the
// DilatedBlock sum, losses and the fused multi-tensor Adam step.  All are coalesced along the
// channel axis of the NHWC activations (float4 when C % 4 == 0) and grid-sized as a multiple of
// the SM count.  Per-channel reductions go warp/CTA-local first and finish with one double
// atomicAdd per (CTA, channel).
//
// Reference call sites: nn.BatchNorm2d atomai/nets/blocks.py:73; nn.LeakyReLU :70;
// F.max_pool2d atomai/nets/fcnn.py:123-127; F.interpolate atomai/nets/blocks.py:130-131;
// DilatedBlock.forward blocks.py:321-329; nn.CrossEntropyLoss / BCEWithLogits / MSE
// atomai/losses_metrics/losses.py:154-164; torch.optim.Adam atomai/trainers/trainer.py:539.
#include "common.cuh"

namespace {

constexpr int kT = 256;

inline int grid_for(int64_t work_items, int per_block = kT) {
  int64_t b = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)ab_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------- BatchNorm finalize
__global__ void bn_finalize_kernel(const double* stats, int C, double count, const float* gamma,
                                   const float* beta, float* rmean, float* rvar, float momentum,
                                   float eps, int training, float* scale, float* shift,
                                   float* mean_o, float* invstd_o) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var_b;
  if (training) {
    const double m = stats[c] / count;
    double v = stats[C + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var_b = (float)v;
    if (rmean) {
      const double unb = count > 1 ? v * count / (count - 1) : v;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  } else {
    mean = rmean[c];
    var_b = rvar[c];
  }
  const float invstd = rsqrtf(var_b + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - mean * g * invstd;
  if (mean_o) mean_o[c] = mean;
  if (invstd_o) invstd_o[c] = invstd;
}

// ---------------------------------------------------------------- y = a*scale + shift
__global__ void affine_kernel(const float* __restrict__ a, int ld_a, const float* scale,
                              const float* shift, float* __restrict__ y, int ld_y, int64_t npix,
                              int C, int hw_nchw) {
  const int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    float v = a[pix * ld_a + c];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    if (hw_nchw) {
      const int64_t n = pix / hw_nchw, r = pix % hw_nchw;
      y[(n * C + c) * hw_nchw + r] = v;
    } else {
      y[pix * ld_y + c] = v;
    }
  }
}

// ---------------------------------------------------------------- per-channel reductions
// Thread layout: CW = pow2 >= C (<= 256) channel lanes, R = 256/CW pixel rows per CTA step.
struct ChanLayout {
  int CW, R;
};
inline ChanLayout chan_layout(int C) {
  int cw = 1;
  while (cw < C) cw <<= 1;
  ChanLayout l;
  l.CW = cw;
  l.R = kT / cw;
  return l;
}

__device__ __forceinline__ void block_chan_reduce2(float v1, float v2, int cx, int py, int CW,
                                                   int R, int C, double* out1, double* out2) {
  __shared__ float s1[kT], s2[kT];
  s1[py * CW + cx] = v1;
  s2[py * CW + cx] = v2;
  __syncthreads();
  if (py == 0 && cx < C) {
    float a = 0.f, b = 0.f;
    for (int r = 0; r < R; ++r) {
      a += s1[r * CW + cx];
      b += s2[r * CW + cx];
    }
    if (out1) atomicAdd(out1 + cx, (double)a);
    if (out2) atomicAdd(out2 + cx, (double)b);
  }
}

__global__ void __launch_bounds__(kT) bn_bwd_reduce_kernel(const float* __restrict__ dy, int ld_dy,
                                                           const float* __restrict__ a, int ld_a,
                                                           const float* mean, const float* invstd,
                                                           int64_t npix, int C, int CW, int R,
                                                           double* sums) {
  const int cx = threadIdx.x % CW, py = threadIdx.x / CW;
  float s1 = 0.f, s2 = 0.f;
  if (cx < C) {
    const float m = mean[cx], is = invstd[cx];
    for (int64_t p = (int64_t)blockIdx.x * R + py; p < npix; p += (int64_t)gridDim.x * R) {
      const float g = dy[p * ld_dy + cx];
      const float xh = (a[p * ld_a + cx] - m) * is;
      s1 += g;
      s2 = fmaf(g, xh, s2);
    }
  }
  block_chan_reduce2(s1, s2, cx, py, CW, R, C, sums, sums + C);
}

__global__ void __launch_bounds__(kT)
    bn_lrelu_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ a,
                        int ld_a, const float* mean, const float* invstd, const float* scale,
                        const double* sums, double count, const float* __restrict__ extra,
                        int ld_extra, int act, float alpha, float* __restrict__ dpre, int ld_dpre,
                        double* dbias, int64_t npix, int C, int CW, int R) {
  const int cx = threadIdx.x % CW, py = threadIdx.x / CW;
  float sb = 0.f;
  if (cx < C) {
    float m = 0.f, is = 0.f, sc = 1.f, k1 = 0.f, k2 = 0.f;
    const bool bn = scale != nullptr;
    if (bn) {
      m = mean[cx];
      is = invstd[cx];
      sc = scale[cx];
      k1 = (float)(sums[cx] / count);
      k2 = (float)(sums[C + cx] / count);
    }
    for (int64_t p = (int64_t)blockIdx.x * R + py; p < npix; p += (int64_t)gridDim.x * R) {
      const float av = a[p * ld_a + cx];
      float g = dy ? dy[p * ld_dy + cx] : 0.f;
      if (bn) {
        const float xh = (av - m) * is;
        g = sc * (g - k1 - xh * k2);
      }
      float ex = 0.f;
      if (extra) {
        ex = extra[p * ld_extra + cx];
        g += ex;
      }
      const float d = g * act_grad_from_out(av, act, alpha) + ex;
      dpre[p * ld_dpre + cx] = d;
      sb += d;
    }
  }
  if (dbias) block_chan_reduce2(sb, 0.f, cx, py, CW, R, C, dbias, nullptr);
}

// ---------------------------------------------------------------- 2x2 max pool
__global__ void pool_fwd_kernel(const float* __restrict__ a, int ld_a, const float* scale,
                                const float* shift, float* __restrict__ y, int ld_y, int N, int Ho,
                                int Wo, int C) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  const int W2 = 2 * Wo, H2 = 2 * Ho;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
    const float* b = a + (((int64_t)n * H2 + 2 * ho) * W2 + 2 * wo) * ld_a + c;
    const float v0 = fmaf(b[0], sc, sh), v1 = fmaf(b[ld_a], sc, sh);
    const float v2 = fmaf(b[(int64_t)W2 * ld_a], sc, sh);
    const float v3 = fmaf(b[(int64_t)W2 * ld_a + ld_a], sc, sh);
    y[(((int64_t)n * Ho + ho) * Wo + wo) * ld_y + c] = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ dp, int ld_dp,
                                const float* __restrict__ a, int ld_a, const float* scale,
                                const float* shift, float* __restrict__ df, int ld_df,
                                int accumulate, int N, int Ho, int Wo, int C) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  const int W2 = 2 * Wo, H2 = 2 * Ho;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
    const int64_t base = ((int64_t)n * H2 + 2 * ho) * W2 + 2 * wo;
    const int64_t off[4] = {0, 1, W2, (int64_t)W2 + 1};
    int best = 0;
    float bv = fmaf(a[(base + off[0]) * ld_a + c], sc, sh);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = fmaf(a[(base + off[k]) * ld_a + c], sc, sh);
      if (v > bv) {  // first maximum wins, as in ATen's max_pool2d
        bv = v;
        best = k;
      }
    }
    const float g = dp[(((int64_t)n * Ho + ho) * Wo + wo) * ld_dp + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* o = df + (base + off[k]) * ld_df + c;
      const float v = (k == best) ? g : 0.f;
      *o = accumulate ? *o + v : v;
    }
  }
}

// ---------------------------------------------------------------- x2 upsampling
__device__ __forceinline__ void up_src(int o, int n, int& i0, int& i1, float& l1) {
  // align_corners=False, scale 2: src = max(o/2 - 0.25, 0)
  float s = 0.5f * o - 0.25f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  l1 = s - i0;
  i1 = i0 + 1 < n ? i0 + 1 : n - 1;
}

__global__ void upsample_fwd_kernel(const float* __restrict__ x, int ld_x, float* __restrict__ y,
                                    int ld_y, int N, int h, int w, int C, int bilinear) {
  const int H = 2 * h, W = 2 * w;
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ow = (int)(r % W); r /= W;
    const int oh = (int)(r % H);
    const int n = (int)(r / H);
    float v;
    const float* xb = x + (int64_t)n * h * w * ld_x + c;
    if (bilinear) {
      int h0, h1, w0, w1;
      float lh, lw;
      up_src(oh, h, h0, h1, lh);
      up_src(ow, w, w0, w1, lw);
      const float x00 = xb[((int64_t)h0 * w + w0) * ld_x], x01 = xb[((int64_t)h0 * w + w1) * ld_x];
      const float x10 = xb[((int64_t)h1 * w + w0) * ld_x], x11 = xb[((int64_t)h1 * w + w1) * ld_x];
      v = (1.f - lh) * ((1.f - lw) * x00 + lw * x01) + lh * ((1.f - lw) * x10 + lw * x11);
    } else {
      v = xb[((int64_t)(oh >> 1) * w + (ow >> 1)) * ld_x];
    }
    y[(((int64_t)n * H + oh) * W + ow) * ld_y + c] = v;
  }
}

// adjoint of the above: gather form, <= 4 taps per axis
__device__ __forceinline__ int up_adj(int i, int n, int* o, float* wt) {
  int k = 0;
  if (i >= 1) { o[k] = 2 * i - 1; wt[k++] = 0.25f; }
  o[k] = 2 * i; wt[k++] = (i == 0) ? 1.0f : 0.75f;
  o[k] = 2 * i + 1; wt[k++] = (i == n - 1) ? 1.0f : 0.75f;
  if (i <= n - 2) { o[k] = 2 * i + 2; wt[k++] = 0.25f; }
  return k;
}

__global__ void upsample_bwd_kernel(const float* __restrict__ dy, int ld_dy, float* __restrict__ dx,
                                    int ld_dx, int N, int h, int w, int C, int bilinear) {
  const int H = 2 * h, W = 2 * w;
  const int64_t total = (int64_t)N * h * w * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int iw = (int)(r % w); r /= w;
    const int ih = (int)(r % h);
    const int n = (int)(r / h);
    const float* db = dy + (int64_t)n * H * W * ld_dy + c;
    float acc = 0.f;
    if (bilinear) {
      int oh[4], ow[4];
      float wh[4], ww[4];
      const int nh = up_adj(ih, h, oh, wh), nw = up_adj(iw, w, ow, ww);
      for (int a = 0; a < nh; ++a) {
        float rowacc = 0.f;
        for (int b = 0; b < nw; ++b) rowacc = fmaf(ww[b], db[((int64_t)oh[a] * W + ow[b]) * ld_dy], rowacc);
        acc = fmaf(wh[a], rowacc, acc);
      }
    } else {
      const int64_t b0 = ((int64_t)(2 * ih) * W + 2 * iw) * ld_dy;
      acc = db[b0] + db[b0 + ld_dy] + db[b0 + (int64_t)W * ld_dy] + db[b0 + (int64_t)W * ld_dy + ld_dy];
    }
    dx[(((int64_t)n * h + ih) * w + iw) * ld_dx + c] = acc;
  }
}

__global__ void add_slice_kernel(const float* __restrict__ src, int ld_s, float* __restrict__ dst,
                                 int ld_d, int accumulate, int64_t npix, int C) {
  const int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    const float v = src[p * ld_s + c];
    float* o = dst + p * ld_d + c;
    *o = accumulate ? *o + v : v;
  }
}

constexpr int kMaxDil = 8;
struct DilPtrs {
  const float* a[kMaxDil];
  const float* scale[kMaxDil];
  const float* shift[kMaxDil];
};
__global__ void dilated_sum_kernel(DilPtrs P, int nl, float alpha, float* __restrict__ out,
                                   int64_t n, int C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float acc = 0.f;
    for (int l = 0; l < nl; ++l) {
      const float av = P.a[l][i];
      const float pre = av > 0.f ? av : av / alpha;
      acc += pre + av;
      if (P.scale[l]) acc += fmaf(av, P.scale[l][c], P.shift[l][c]);
    }
    out[i] = acc;
  }
}

// ---------------------------------------------------------------- losses
__global__ void __launch_bounds__(kT) ce_kernel(const float* __restrict__ logits, int ld,
                                                const int64_t* __restrict__ labels, int64_t npix,
                                                int C, double* loss_sum,
                                                float* __restrict__ dlogits, int ld_d,
                                                float gscale, const float* gscale_dev) {
  float local = 0.f;
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix;
       p += (int64_t)gridDim.x * blockDim.x) {
    const float* l = logits + p * ld;
    float m = l[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += __expf(l[c] - m);
    const float lse = m + __logf(s);
    const int y = (int)labels[p];
    local += lse - l[y];
    if (dlogits) {
      const float inv = 1.f / s;
      for (int c = 0; c < C; ++c) {
        const float pr = __expf(l[c] - m) * inv;
        dlogits[p * ld_d + c] = (pr - (c == y ? 1.f : 0.f)) * gscale;
      }
    }
  }
  __shared__ float red[kT / 32];
  const float ws = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kT / 32; ++i) t += red[i];
    if (loss_sum) atomicAdd(loss_sum, (double)t);
  }
}

__global__ void __launch_bounds__(kT) pointwise_loss_kernel(const float* __restrict__ pred,
                                                            const float* __restrict__ tgt,
                                                            int64_t n, int kind, double* loss_sum,
                                                            float* __restrict__ dpred,
                                                            float gscale,
                                                            const float* gscale_dev) {
  float local = 0.f;
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pred[i], t = tgt[i];
    if (kind == 0) {
      const float d = x - t;
      local = fmaf(d, d, local);
      if (dpred) dpred[i] = 2.f * d * gscale;
    } else {  // BCE with logits: max(x,0) - x*t + log(1 + exp(-|x|))
      local += fmaxf(x, 0.f) - x * t + log1pf(__expf(-fabsf(x)));
      if (dpred) dpred[i] = (1.f / (1.f + __expf(-x)) - t) * gscale;
    }
  }
  __shared__ float red[kT / 32];
  const float ws = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kT / 32; ++i) t += red[i];
    if (loss_sum) atomicAdd(loss_sum, (double)t);
  }
}

__global__ void __launch_bounds__(kT) sqerr_kernel(const float* __restrict__ x,
                                                   const float* __restrict__ xhat, int64_t n,
                                                   double* out, float* __restrict__ dxhat,
                                                   float gscale, const float* gscale_dev) {
  float local = 0.f;
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float d = xhat[i] - x[i];
    local = fmaf(d, d, local);
    if (dxhat) dxhat[i] = d * gscale;
  }
  __shared__ float red[kT / 32];
  const float ws = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kT / 32; ++i) t += red[i];
    if (out) atomicAdd(out, 0.5 * (double)t);
  }
}

// ---------------------------------------------------------------- Adam (multi-tensor)
__global__ void adam_multi_kernel(const int64_t* __restrict__ table, float lr, float b1, float b2,
                                  float eps, float wd, float bc1, float bc2_sqrt,
                                  float grad_scale) {
  const int64_t* row = table + (int64_t)blockIdx.y * 5;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* m = reinterpret_cast<float*>(row[2]);
  float* v = reinterpret_cast<float*>(row[3]);
  const int64_t n = row[4];
  const float step = lr / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (wd != 0.f) gi = fmaf(wd, p[i], gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step * (mi / denom);
  }
}

}  // namespace

#define STREAM ((cudaStream_t)stream)

extern "C" {

int atomai_b200_bn_finalize(const double* stats, int C, double count, const float* gamma,
                            const float* beta, float* running_mean, float* running_var,
                            float momentum, float eps, int training, float* scale, float* shift,
                            float* mean, float* invstd, void* stream) {
  AB_CHECK(C > 0, "bn_finalize: C=%d", C);
  AB_CHECK(training || (running_mean && running_var), "bn_finalize: eval needs running stats");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(stats, C, count, gamma, beta,
                                                          running_mean, running_var, momentum,
                                                          eps, training, scale, shift, mean,
                                                          invstd);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_affine(const float* a, int ld_a, const float* scale, const float* shift, float* y,
                       int ld_y, int64_t npix, int C, int out_nchw_hw, void* stream) {
  if (npix * C == 0) return 0;
  affine_kernel<<<grid_for(npix * C), kT, 0, STREAM>>>(a, ld_a, scale, shift, y, ld_y, npix, C,
                                                       out_nchw_hw);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_bn_bwd_reduce(const float* dy, int ld_dy, const float* a, int ld_a,
                              const float* mean, const float* invstd, int64_t npix, int C,
                              double* sums, void* stream) {
  AB_CHECK(C <= kT, "bn_bwd_reduce: C=%d > %d unsupported", C, kT);
  if (npix == 0) return 0;
  const ChanLayout l = chan_layout(C);
  bn_bwd_reduce_kernel<<<grid_for(npix, l.R * 8), kT, 0, STREAM>>>(dy, ld_dy, a, ld_a, mean,
                                                                  invstd, npix, C, l.CW, l.R,
                                                                  sums);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_bn_lrelu_bwd(const float* dy, int ld_dy, const float* a, int ld_a,
                             const float* mean, const float* invstd, const float* scale,
                             const double* sums, double count, const float* extra, int ld_extra,
                             int act, float lrelu, float* dpre, int ld_dpre, double* dbias,
                             int64_t npix, int C, void* stream) {
  AB_CHECK(C <= kT, "bn_lrelu_bwd: C=%d > %d unsupported", C, kT);
  if (npix == 0) return 0;
  const ChanLayout l = chan_layout(C);
  bn_lrelu_bwd_kernel<<<grid_for(npix, l.R * 8), kT, 0, STREAM>>>(
      dy, ld_dy, a, ld_a, mean, invstd, scale, sums, count, extra, ld_extra, act, lrelu, dpre,
      ld_dpre, dbias, npix, C, l.CW, l.R);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_pool2x2_fwd(const float* a, int ld_a, const float* scale, const float* shift,
                            float* y, int ld_y, int N, int Ho, int Wo, int C, void* stream) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  if (total == 0) return 0;
  pool_fwd_kernel<<<grid_for(total), kT, 0, STREAM>>>(a, ld_a, scale, shift, y, ld_y, N, Ho, Wo,
                                                      C);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_pool2x2_bwd(const float* dp, int ld_dp, const float* a, int ld_a,
                            const float* scale, const float* shift, float* dfull, int ld_df,
                            int accumulate, int N, int Ho, int Wo, int C, void* stream) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  if (total == 0) return 0;
  pool_bwd_kernel<<<grid_for(total), kT, 0, STREAM>>>(dp, ld_dp, a, ld_a, scale, shift, dfull,
                                                      ld_df, accumulate, N, Ho, Wo, C);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_upsample2x_fwd(const float* x, int ld_x, float* y, int ld_y, int N, int h, int w,
                               int C, int bilinear, void* stream) {
  const int64_t total = (int64_t)N * 4 * h * w * C;
  if (total == 0) return 0;
  upsample_fwd_kernel<<<grid_for(total), kT, 0, STREAM>>>(x, ld_x, y, ld_y, N, h, w, C, bilinear);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_upsample2x_bwd(const float* dy, int ld_dy, float* dx, int ld_dx, int N, int h,
                               int w, int C, int bilinear, void* stream) {
  const int64_t total = (int64_t)N * h * w * C;
  if (total == 0) return 0;
  upsample_bwd_kernel<<<grid_for(total), kT, 0, STREAM>>>(dy, ld_dy, dx, ld_dx, N, h, w, C,
                                                          bilinear);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_add_slice(const float* src, int ld_s, float* dst, int ld_d, int accumulate,
                          int64_t npix, int C, void* stream) {
  if (npix * C == 0) return 0;
  add_slice_kernel<<<grid_for(npix * C), kT, 0, STREAM>>>(src, ld_s, dst, ld_d, accumulate, npix,
                                                          C);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_dilated_sum(const float* const* a_ptrs, const float* const* scale_ptrs,
                            const float* const* shift_ptrs, int nlayers, float lrelu, float* out,
                            int64_t n_elems, int C, void* stream) {
  AB_CHECK(nlayers >= 1 && nlayers <= kMaxDil, "dilated_sum: nlayers=%d (max %d)", nlayers,
           kMaxDil);
  AB_CHECK(lrelu != 0.f, "dilated_sum: LeakyReLU slope 0 is not invertible");
  DilPtrs P;
  for (int l = 0; l < kMaxDil; ++l) {
    P.a[l] = l < nlayers ? a_ptrs[l] : nullptr;
    P.scale[l] = (l < nlayers && scale_ptrs) ? scale_ptrs[l] : nullptr;
    P.shift[l] = (l < nlayers && shift_ptrs) ? shift_ptrs[l] : nullptr;
  }
  if (n_elems == 0) return 0;
  dilated_sum_kernel<<<grid_for(n_elems), kT, 0, STREAM>>>(P, nlayers, lrelu, out, n_elems, C);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_ce_fwd_bwd(const float* logits, int ld, const int64_t* labels, int64_t npix, int C,
                           double* loss_sum, float* dlogits, int ld_d, float gscale,
                           const float* gscale_dev, void* stream) {
  if (npix == 0) return 0;
  ce_kernel<<<grid_for(npix), kT, 0, STREAM>>>(logits, ld, labels, npix, C, loss_sum, dlogits,
                                               ld_d, gscale, gscale_dev);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_pointwise_loss(const float* pred, const float* target, int64_t n, int kind,
                               double* loss_sum, float* dpred, float gscale,
                               const float* gscale_dev, void* stream) {
  AB_CHECK(kind == 0 || kind == 1, "pointwise_loss: kind=%d", kind);
  if (n == 0) return 0;
  pointwise_loss_kernel<<<grid_for(n), kT, 0, STREAM>>>(pred, target, n, kind, loss_sum, dpred,
                                                        gscale, gscale_dev);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_sqerr_reduce(const float* x, const float* xhat, int64_t n, double* out,
                             float* dxhat, float gscale, const float* gscale_dev, void* stream) {
  if (n == 0) return 0;
  sqerr_kernel<<<grid_for(n), kT, 0, STREAM>>>(x, xhat, n, out, dxhat, gscale, gscale_dev);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_adam_multi(const int64_t* table_dev, int n, int64_t max_numel, float lr,
                           float beta1, float beta2, float eps, float weight_decay, int step,
                           float grad_scale, void* stream) {
  if (n == 0 || max_numel == 0) return 0;
  AB_CHECK(step >= 1, "adam: step=%d", step);
  AB_CHECK(n <= 65535, "adam: too many tensors (%d)", n);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  int gx = (int)((max_numel + kT * 4 - 1) / (kT * 4));
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n);
  adam_multi_kernel<<<grid, kT, 0, STREAM>>>(table_dev, lr, beta1, beta2, eps, weight_decay, bc1,
                                             sqrtf(bc2), grad_scale);
  AB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
