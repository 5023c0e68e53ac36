// gram.cu — dense deep-kernel Gram matrix K[i][j] = os * k(|| (x1_i - x2_j) * inv_ls ||):
// tiled pairwise distances (||a||^2 + ||b||^2 - 2 a.b, row norms by warp shuffles) with the
// exponentiation fused into the epilogue so the n1 x n2 matrix is written exactly once.
// RBF / Matern-2.5 / outputscale follow gpytorch's public kernel definitions as configured at
// atomai/nets/gp.py:41-46 and :100-111 (gpytorch itself is not vendored by the reference).
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, GT = 256;

// scaled copy xs = x * inv_ls and squared row norms (one warp per row)
__global__ void scale_norm_kernel(const float* __restrict__ x, const float* __restrict__ inv_ls,
                                  int n, int d, float* __restrict__ xs, float* __restrict__ nrm) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = x[(int64_t)row * d + k] * inv_ls[k];
    xs[(int64_t)row * d + k] = v;
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) nrm[row] = acc;
}

__global__ void __launch_bounds__(GT)
    gram_kernel(const float* __restrict__ a, const float* __restrict__ b,
                const float* __restrict__ na, const float* __restrict__ nb, int n1, int n2, int d,
                float os, int kind, float* __restrict__ K, int64_t ldk) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += TK) {
    __syncthreads();
    for (int i = tid; i < TM * TK; i += GT) {
      const int k = i % TK, m = i / TK;
      sA[k][m] = (m0 + m < n1 && k0 + k < d) ? __ldg(a + (int64_t)(m0 + m) * d + k0 + k) : 0.f;
      sB[k][m] = (n0 + m < n2 && k0 + k < d) ? __ldg(b + (int64_t)(n0 + m) * d + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][tm]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][tn]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= n1) continue;
    const float nam = na[m];
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      float d2 = n < n2 ? nam + nb[n] - 2.f * acc[i][j] : 0.f;
      d2 = fmaxf(d2, 0.f);
      if (kind == 0) {
        out[j] = os * __expf(-0.5f * d2);
      } else {  // Matern nu = 2.5: (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)
        const float r = sqrtf(d2);
        const float s5r = 2.2360679775f * r;
        out[j] = os * (1.f + s5r + 1.6666666667f * d2) * __expf(-s5r);
      }
    }
    float* o = K + (int64_t)m * ldk + n0 + tn;
    if (n0 + tn + 3 < n2 && ((ldk & 3) == 0) && (((uintptr_t)K & 15) == 0)) {
      *reinterpret_cast<float4*>(o) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (n0 + tn + j < n2) o[j] = out[j];
    }
  }
}

}  // namespace

// scratch layout inside the caller's K is not possible (K is the product), so the scaled copies
// live in a small cached workspace owned by the library per device (n*(d+1) floats per side).
static float* g_ws = nullptr;
static size_t g_ws_bytes = 0;

extern "C" int atomai_b200_gram(const float* x1, const float* x2, const float* inv_ls,
                                float outputscale, int n1, int n2, int d, int kind, float* K,
                                int64_t ldk, void* stream) {
  AB_CHECK(x1 && x2 && inv_ls && K, "gram: null pointer");
  AB_CHECK(n1 >= 0 && n2 >= 0 && d > 0 && ldk >= n2, "gram: bad dims n1=%d n2=%d d=%d", n1, n2, d);
  AB_CHECK(kind == 0 || kind == 1, "gram: kind=%d", kind);
  if (n1 == 0 || n2 == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = ((size_t)(n1 + n2) * (d + 1) + 64) * sizeof(float);
  if (need > g_ws_bytes) {
    if (g_ws) cudaFree(g_ws);
    g_ws = nullptr; g_ws_bytes = 0;
    AB_CUDA(cudaMalloc(&g_ws, need));
    g_ws_bytes = need;
  }
  float* a = g_ws;
  float* b = a + (size_t)n1 * d;
  float* na = b + (size_t)n2 * d;
  float* nb = na + n1;
  scale_norm_kernel<<<(n1 + 7) / 8, 256, 0, st>>>(x1, inv_ls, n1, d, a, na);
  AB_LAUNCH_CHECK();
  scale_norm_kernel<<<(n2 + 7) / 8, 256, 0, st>>>(x2, inv_ls, n2, d, b, nb);
  AB_LAUNCH_CHECK();
  dim3 grid((n2 + TN - 1) / TN, (n1 + TM - 1) / TM);
  AB_CHECK(grid.y <= 65535, "gram: n1 too large for one launch");
  gram_kernel<<<grid, GT, 0, st>>>(a, b, na, nb, n1, n2, d, outputscale, kind, K, ldk);
  AB_LAUNCH_CHECK();
  return 0;
}
