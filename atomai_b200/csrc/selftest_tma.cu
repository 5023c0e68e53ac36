// selftest_tma.cu — feasibility probe for the TMA-fed activation loader planned in DESIGN.md §6.1
// (dev export, not on the hot path): one CTA issues a single cp.async.bulk.tensor of a
// {32 channels, TWp, THp, 1} box out of an NHWC fp32 tensor — negative / out-of-range
// coordinates included, which the hardware zero-fills like the convolution's padding — into
// shared memory with the requested swizzle mode, then dumps the raw shared-memory bytes so that
// the host can check (a) the zero fill and (b) that the swizzle is the one the UMMA descriptor
// layouts of selftest.cu expect (a function of the absolute shared-memory address).
#include <cuda.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
    selftest_tma_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ out, int c0,
                        int w0, int h0, int n0, int box_bytes, int smem_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t base = ((smem_u32(smem) + 1023) & ~1023u) + smem_offset;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar), box_bytes);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(base),
        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(w0), "r"(h0), "r"(n0),
        "r"(smem_u32(&bar))
        : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  uint8_t* gen0 = smem - smem_u32(smem);
  for (int i = threadIdx.x; i < box_bytes / 4; i += blockDim.x)
    out[i] = *reinterpret_cast<const float*>(gen0 + base + i * 4);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// swizzle_mode: 0 none, 1 32B, 2 64B, 3 128B (K-major conv tiles), 4 128B_ATOM_32B (the
// MN-major TF32 layout of wgrad_tc).  out receives THp*TWp*32 floats = the raw smem image.
extern "C" int atomai_b200_selftest_tma(const float* x, int N, int H, int W, int C, int c0, int w0,
                                        int h0, int n0, int TWp, int THp, int swizzle_mode,
                                        int smem_offset, float* out, void* stream) {
  AB_CHECK(x && out && C % 4 == 0 && TWp >= 1 && THp >= 1 && TWp <= 256 && THp <= 256,
           "selftest_tma: bad arguments");
  AB_CHECK(((uintptr_t)x & 15) == 0, "selftest_tma: tensor must be 16 B aligned");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  AB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  AB_CHECK(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)TWp, (cuuint32_t)THp, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_mode == 1) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle_mode == 2) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle_mode == 3) sw = CU_TENSOR_MAP_SWIZZLE_128B;
#if CUDA_VERSION >= 12080
  if (swizzle_mode == 4) sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
#else
  AB_CHECK(swizzle_mode != 4, "SWIZZLE_128B_ATOM_32B needs CUDA >= 12.8 headers");
#endif
  const CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(
      &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  const int box_bytes = TWp * THp * 128;
  const int smem = box_bytes + 2048 + smem_offset;
  AB_CHECK(smem <= 200 * 1024, "selftest_tma: box too large");
  AB_CUDA(cudaFuncSetAttribute(selftest_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               200 * 1024));
  selftest_tma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmap, out, c0, w0, h0, n0, box_bytes,
                                                             smem_offset);
  AB_LAUNCH_CHECK();
  return 0;
}
