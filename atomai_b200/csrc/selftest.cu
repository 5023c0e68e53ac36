// selftest.cu — a single-CTA tcgen05 GEMM used by tests/ to pin the UMMA shared-memory
// descriptor conventions the convolution kernels rely on (no-swizzle "interleave" core matrices,
// K-major and MN-major, shifted starts and non-128B group strides as used by the halo-tile tap
// addressing).  D[128][N] = A[128][K] * B[N][K]^T with TF32 operands.
//
// variant bit 0: swap the LBO/SBO fields (a mismatch with the documented convention shows up as a
//                wrong result here instead of inside a convolution);
// variant bit 1: MN-major operands (A given as A^T [K][128], B as B^T [K][N]);
// variant >= 8: MN-major operands in the SWIZZLE_128B_BASE32B layout (rows = k at a 128 B pitch,
//                32 m/n elements per row, 32 B chunks XOR-swizzled by row index mod 4):
//                bit 0 swaps LBO/SBO, bits 1-2 shift the tile start by 1..3 rows (tap addressing),
//                bit 4 (variant >= 16) additionally sets the descriptor base_offset to (start>>7)&7;
// variant bit 2: "halo" addressing — K-major A rows stored with 8-row groups at a 160 B stride and
//                a 48 B start offset (what a tap of a 10-pixel-wide halo tile looks like).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
    selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                         float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (variant >= 8) {
    // ---------------- MN-major, SWIZZLE_128B_BASE32B ----------------
    const bool swap8 = variant < 32 && (variant & 1);
    const uint32_t shift = variant < 32 ? (variant >> 1) & 3 : 0u;
    const uint32_t use_bo = variant < 32 ? (variant >> 4) & 1 : 0u;
    const uint32_t base0 = (smem_u32(smem) + 1023) & ~1023u;
    const uint32_t chunk = ((uint32_t)K + 8) * 128;               // stride between 32-wide m/n chunks
    const uint32_t lboA = (chunk + 1023) & ~1023u;
    const uint32_t a0 = base0 + shift * 128;
    const uint32_t b0 = base0 + 4 * lboA + 1024 + shift * 128;
    uint8_t* gen0 = smem - smem_u32(smem);                         // generic pointer of smem offset 0
    // variant >= 32: "stacked taps" — A is ONE 32-channel tile X[K+16][32] (given in A) and the
    // four 32-row chunks of the M = 128 operand are row-shifted views of it: LBO = stack * 128 B,
    // i.e. A[m = c*32 + i][k] = X[k + c*stack][i] (what wgrad_tc uses to put the tx taps on M).
    const uint32_t stack = variant >= 32 ? (uint32_t)(variant - 32) : 0u;
    if (stack) {
      for (int i = tid; i < (K + 16) * 32; i += 128) {
        const int k = i / 32, m = i % 32;
        const uint32_t la = a0 + k * 128 + m * 4;
        *reinterpret_cast<float*>(gen0 + swz128_32(la)) = to_tf32(A[i]);
      }
    } else {
      for (int i = tid; i < 128 * K; i += 128) {                   // A^T[k][m]
        const int k = i / 128, m = i % 128;
        const uint32_t la = a0 + (m / 32) * lboA + k * 128 + (m % 32) * 4;
        *reinterpret_cast<float*>(gen0 + swz128_32(la)) = to_tf32(A[i]);
      }
    }
    for (int i = tid; i < N * K; i += 128) {                       // B^T[k][n]
      const int k = i / N, n = i % N;
      const uint32_t la = b0 + (n / 32) * lboA + k * 128 + (n % 32) * 4;
      *reinterpret_cast<float*>(gen0 + swz128_32(la)) = to_tf32(B[i]);
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, N, 1, 1);
      for (int ks = 0; ks < K / 8; ++ks) {
        uint32_t lbo = lboA, sbo = 512;
        if (swap8) { lbo = 512; sbo = lboA; }
        const uint32_t sa = a0 + ks * 1024, sb = b0 + ks * 1024;
        const uint64_t ad = umma_desc_ex(sa, stack ? stack * 128 : lbo, sbo, 1, use_bo ? (sa >> 7) & 7 : 0);
        const uint64_t bd = umma_desc_ex(sb, lbo, sbo, 1, use_bo ? (sb >> 7) & 7 : 0);
        umma_tf32(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + c0 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
    return;
  }
  const bool swap = variant & 1, mn = variant & 2, halo = variant & 4;

  uint32_t a_plane, a_group, a_off, b_plane;
  const uint32_t a_base = smem_u32(smem);
  if (!mn) {
    a_group = halo ? 160 : 128;                                  // bytes between 8-row groups
    a_off = halo ? 48 : 0;
    a_plane = 16 * a_group + 16 + (halo ? 64 : 0);               // plane = 4 k-columns
    b_plane = N * 16 + 16;
  } else {
    a_group = 128; a_off = 0;
    a_plane = K * 16 + 16;                                       // plane = 4 m-rows, K rows of 16B
    b_plane = K * 16 + 16;
  }
  const uint32_t a_bytes = (mn ? 32u : (uint32_t)(K / 4)) * a_plane + 256;
  const uint32_t b_base = a_base + ((a_bytes + 127) & ~127u);

  if (!mn) {
    for (int i = tid; i < 128 * K; i += 128) {  // A[r][k]
      const int r = i / K, k = i % K;
      const uint32_t addr = (k / 4) * a_plane + a_off + (r / 8) * a_group + (r % 8) * 16 + (k % 4) * 4;
      *reinterpret_cast<float*>(smem + addr) = to_tf32(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {  // B[n][k]
      const int n = i / K, k = i % K;
      const uint32_t addr = (k / 4) * b_plane + n * 16 + (k % 4) * 4;
      *reinterpret_cast<float*>(smem + (b_base - a_base) + addr) = to_tf32(B[i]);
    }
  } else {
    for (int i = tid; i < 128 * K; i += 128) {  // A^T[k][m]
      const int k = i / 128, m = i % 128;
      const uint32_t addr = (m / 4) * a_plane + k * 16 + (m % 4) * 4;
      *reinterpret_cast<float*>(smem + addr) = to_tf32(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {  // B^T[k][n]
      const int k = i / N, n = i % N;
      const uint32_t addr = (n / 4) * b_plane + k * 16 + (n % 4) * 4;
      *reinterpret_cast<float*>(smem + (b_base - a_base) + addr) = to_tf32(B[i]);
    }
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N, mn ? 1 : 0, mn ? 1 : 0);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t ad, bd;
      if (!mn) {
        uint32_t lbo_a = a_plane, sbo_a = a_group, lbo_b = b_plane, sbo_b = 128;
        if (swap) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
        ad = umma_desc(a_base + a_off + ks * 2 * a_plane, lbo_a, sbo_a);
        bd = umma_desc(b_base + ks * 2 * b_plane, lbo_b, sbo_b);
      } else {
        uint32_t lbo_a = 128, sbo_a = a_plane, lbo_b = 128, sbo_b = b_plane;
        if (swap) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
        ad = umma_desc(a_base + ks * 128, lbo_a, sbo_a);
        bd = umma_desc(b_base + ks * 128, lbo_b, sbo_b);
      }
      umma_tf32(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + c0 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// Feasibility probe for the TMA-fed convolution planned in DESIGN.md §6.1 (not used by the hot
// path, not part of the tested selftest group): K-major operands in the SWIZZLE_128B layout —
// rows (pixels) at a 128 B pitch holding 32 TF32 channels, 16 B chunks XOR-ed with (row mod 8),
// which is what cp.async.bulk.tensor writes for a {32 ch, W, H} box with CU_TENSOR_MAP_SWIZZLE_128B.
// variant bit 0: "halo" addressing — the 8-row groups of A sit TWp = 18 rows apart (SBO = 2304 B)
// and the start is shifted by (variant >> 1) & 7 rows, i.e. a convolution tap of a 16 x 8 tile.
__device__ __forceinline__ uint32_t swz128_16(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }

__global__ void __launch_bounds__(128, 1)
    selftest_sw128_kernel(const float* __restrict__ A, const float* __restrict__ B,
                          float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool halo = variant & 1;
  const uint32_t shift = halo ? (variant >> 1) & 7 : 0;
  const uint32_t TWp = 18;
  const uint32_t base0 = (smem_u32(smem) + 1023) & ~1023u;
  uint8_t* gen0 = smem - smem_u32(smem);                         // generic pointer of smem offset 0
  const uint32_t a_blk = 48 * 1024;                              // bytes between 32-wide k blocks of A
  const uint32_t b0 = base0 + 2 * a_blk, b_blk = 32 * 1024;
  const uint32_t a_sbo = halo ? TWp * 128 : 1024;
  for (int i = tid; i < 128 * K; i += 128) {                     // A[m][k]
    const int m = i / K, k = i % K;
    const uint32_t row = halo ? (uint32_t)(m / 8) * TWp + (m % 8) + shift : (uint32_t)m;
    const uint32_t la = base0 + (k / 32) * a_blk + row * 128 + (k % 32) * 4;
    *reinterpret_cast<float*>(gen0 + swz128_16(la)) = to_tf32(A[i]);
  }
  for (int i = tid; i < N * K; i += 128) {                       // B[n][k]
    const int n = i / K, k = i % K;
    const uint32_t la = b0 + (k / 32) * b_blk + n * 128 + (k % 32) * 4;
    *reinterpret_cast<float*>(gen0 + swz128_16(la)) = to_tf32(B[i]);
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
    for (int ks = 0; ks < K / 8; ++ks) {
      // one MMA = 8 k = 32 B of every row: advance the start address inside the 128 B row
      const uint32_t sa = base0 + (ks / 4) * a_blk + shift * 128 + (ks % 4) * 32;
      const uint32_t sb = b0 + (ks / 4) * b_blk + (ks % 4) * 32;
      const uint64_t ad = umma_desc_ex(sa, 16, a_sbo, 2, 0);     // layout_type 2 = SWIZZLE_128B
      const uint64_t bd = umma_desc_ex(sb, 16, 1024, 2, 0);
      umma_tf32(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + c0 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// Issue-rate probe: `issuers` elected threads (one per warp) each issue `iters` back-to-back
// tcgen05.mma (M = 128, N, K = 8, TF32) on resident shared-memory operands and commit; reports
// the elapsed SM clocks.  layout 0 = K-major no-swizzle (conv_tc), 1 = MN-major SW128_32B
// (wgrad_tc).  Used to place the thin-layer floor (shared-memory operand reads) in DESIGN.md.
__global__ void __launch_bounds__(128, 1)
    umma_rate_kernel(int N, int layout, int issuers, int iters, int unroll_b, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base0 = (smem_u32(smem) + 1023) & ~1023u;
  for (int i = tid; i < 96 * 1024 / 16; i += 128)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base0 + i * 16), "r"(0) : "memory");
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (warp < issuers) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_tf32(128, N, layout, layout);
      uint64_t a_t, b_t;
      if (layout == 0) {
        a_t = umma_desc(base0, 180 * 16 + 16, 18 * 16);          // halo-tile style planes
        b_t = umma_desc(base0 + 48 * 1024, N * 16, 128);
      } else {
        a_t = umma_desc_ex(base0, 128, 512, 1, 0);
        b_t = umma_desc_ex(base0 + 48 * 1024, 2048, 512, 1, 0);   // overlapping chunks: timing only
      }
      const uint32_t a_hi = (uint32_t)(a_t >> 32), b_hi = (uint32_t)(b_t >> 32);
      uint32_t a_lo = (uint32_t)a_t, b_lo = (uint32_t)b_t;
      const uint32_t d = tmem_base + warp * 128;
      t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // step the start addresses like the kernels do (16 B: next pixel / 128 B: next row)
          const uint32_t stp = layout == 0 ? k : 8 * k;
          umma_tf32_lh(d, a_lo + stp, a_hi, b_lo + (unroll_b ? stp : 0), b_hi, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar[warp]));
      t1 = clock64();
      out[warp * 2] = t1 - t0;                                    // issue time
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    if ((tid & 31) == 0) out[warp * 2 + 1] = clock64() - t0;      // until the last MMA retired
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

extern "C" int atomai_b200_selftest_umma(const float* A, const float* B, float* D, int N, int K,
                                         int variant, void* stream) {
  AB_CHECK(N % 16 == 0 && N >= 16 && N <= 256 && K % 8 == 0 && K >= 8 && K <= 64,
           "selftest_umma: N=%d K=%d out of range", N, K);
  AB_CHECK(variant < 8 || N % 32 == 0, "selftest_umma: MN-major SW128_32B needs N %% 32 == 0");
  const int smem = 200 * 1024;
  AB_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               smem));
  selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant);
  AB_LAUNCH_CHECK();
  return 0;
}

extern "C" int atomai_b200_umma_rate(int N, int layout, int issuers, int iters, long long* out,
                                     void* stream) {
  AB_CHECK(N % 16 == 0 && N >= 16 && N <= 128 && issuers >= 1 && issuers <= 4 && iters % 4 == 0,
           "umma_rate: bad arguments");
  const int smem = 100 * 1024;
  AB_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, layout, issuers, iters, 1, out);
  AB_LAUNCH_CHECK();
  return 0;
}

extern "C" int atomai_b200_selftest_sw128(const float* A, const float* B, float* D, int N, int K,
                                          int variant, void* stream) {
  AB_CHECK(N % 16 == 0 && N >= 16 && N <= 256 && K % 32 == 0 && K >= 32 && K <= 64,
           "selftest_sw128: N=%d K=%d out of range", N, K);
  const int smem = 200 * 1024;
  AB_CUDA(cudaFuncSetAttribute(selftest_sw128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               smem));
  selftest_sw128_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant);
  AB_LAUNCH_CHECK();
  return 0;
}
