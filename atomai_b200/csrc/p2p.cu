// p2p.cu — small all-reduce over NVLink peer memory, fused with the BatchNorm finalisation.
//
// Synchronised BatchNorm needs the per-channel (sum, sum^2) of EVERY rank between a convolution and
// the next layer's loader: 13 layers forward + 13 backward per step, 2C doubles each.  As NCCL
// calls these are 26 latency-bound launches on the compute stream.  Here each rank owns a small
// mailbox buffer that all peers have mapped (CUDA IPC over NVLink 5 / NVSwitch): one kernel
//   1. stores its 2C doubles into its slot of every peer's mailbox (remote st.global),
//   2. publishes a per-(epoch, rank) flag with system-scope release semantics,
//   3. waits (bounded) until the flags of all peers for this epoch have arrived locally,
//   4. sums the slots in rank order — bit-identical on every rank — and, in the fused variant,
//      goes straight on to the BatchNorm scale / shift / running-statistics update.
// Two slot sets alternate by epoch parity: a rank can be at most one collective ahead of a peer
// that is still reading the previous one.  New functionality (the reference is single-device,
// SURVEY.md §2.3); semantics = torch.distributed.all_reduce(SUM) on 2C doubles.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kSlotDoubles = 1024;       // per (set, rank): up to 2*512 channels
constexpr int kSets = 2;

struct Peers {
  double* data[kMaxWorld];               // peer r's mailbox: [kSets][world][kSlotDoubles]
  unsigned long long* flag[kMaxWorld];   // peer r's flags:   [kSets][world]
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all ranks launch this with the same (n, epoch); on return `vals[0:n]` holds the sum over ranks
__device__ void exchange_and_sum(const Peers& P, int world, int rank, unsigned long long epoch,
                                 double* vals, int n) {
  const int set = (int)(epoch & (kSets - 1));
  const size_t slot = ((size_t)set * world + rank) * kSlotDoubles;
  for (int r = 0; r < world; ++r)
    for (int i = threadIdx.x; i < n; i += blockDim.x) P.data[r][slot + i] = vals[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys(P.flag[threadIdx.x] + (size_t)set * world + rank, epoch);
  if (threadIdx.x < world) {
    const unsigned long long* f = P.flag[rank] + (size_t)set * world + threadIdx.x;
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > (1u << 25)) {
        printf("atomai_b200: p2p all-reduce timed out (rank %d waiting for rank %d, epoch %llu)\n",
               rank, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  const double* mine = P.data[rank] + (size_t)set * world * kSlotDoubles;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += mine[(size_t)r * kSlotDoubles + i];
    vals[i] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(Peers P, int world, int rank,
                                                            unsigned long long epoch, double* vals,
                                                            int n) {
  exchange_and_sum(P, world, rank, epoch, vals, n);
}

// fused: all-reduce of the statistics + BatchNorm finalisation (same maths as bn_finalize_kernel)
__global__ void __launch_bounds__(256) p2p_bn_finalize_kernel(
    Peers P, int world, int rank, unsigned long long epoch, double* stats, int C, double count,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
    float* running_var, float momentum, float eps, float* scale, float* shift, float* mean,
    float* invstd) {
  exchange_and_sum(P, world, rank, epoch, stats, 2 * C);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {     // exactly bn_finalize_kernel's arithmetic
    const double m = stats[c] / count;
    double v = stats[C + c] / count - m * m;
    if (v < 0) v = 0;
    const float mean_f = (float)m, var_b = (float)v;
    if (running_mean) {
      const double unb = count > 1 ? v * count / (count - 1) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean_f;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
    const float is = rsqrtf(var_b + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * is;
    shift[c] = b - mean_f * g * is;
    if (mean) mean[c] = mean_f;
    if (invstd) invstd[c] = is;
  }
}

int fill_peers(Peers* P, void* const* data_ptrs, void* const* flag_ptrs, int world) {
  AB_CHECK(world >= 1 && world <= kMaxWorld, "p2p: world=%d", world);
  for (int r = 0; r < kMaxWorld; ++r) {
    P->data[r] = r < world ? static_cast<double*>(data_ptrs[r]) : nullptr;
    P->flag[r] = r < world ? static_cast<unsigned long long*>(flag_ptrs[r]) : nullptr;
  }
  return 0;
}

}  // namespace

extern "C" {

// CUDA IPC plumbing for the mailboxes: export = (handle of the enclosing cudaMalloc block, offset
// of `ptr` inside it); import maps the block into THIS process for the current device with
// cudaIpcMemLazyEnablePeerAccess, i.e. kernels of this rank may store to the peer's HBM over NVLink.
int atomai_b200_ipc_export(const void* ptr, unsigned char* handle64, int64_t* offset) {
  AB_CHECK(ptr && handle64 && offset, "ipc_export: null pointer");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  AB_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
  AB_CHECK(fn && q == cudaDriverEntryPointSuccess, "cuMemGetAddressRange not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  AB_CHECK(reinterpret_cast<RangeFn>(fn)(&base, &size, (CUdeviceptr)ptr) == CUDA_SUCCESS,
           "cuMemGetAddressRange failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  AB_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((CUdeviceptr)ptr - base);
  return 0;
}

int atomai_b200_ipc_import(const unsigned char* handle64, int64_t offset, void** ptr_out) {
  AB_CHECK(handle64 && ptr_out, "ipc_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  AB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = static_cast<char*>(base) + offset;
  return 0;
}

// bytes of the mailbox (data) and flag buffers every rank must expose to its peers (zeroed)
int64_t atomai_b200_p2p_data_bytes(int world) { return (int64_t)kSets * world * kSlotDoubles * 8; }
int64_t atomai_b200_p2p_flag_bytes(int world) { return (int64_t)kSets * world * 8; }

int atomai_b200_p2p_allreduce(void* const* data_ptrs_host, void* const* flag_ptrs_host, int world,
                              int rank, uint64_t epoch, double* vals, int n, void* stream) {
  AB_CHECK(vals && n > 0 && n <= kSlotDoubles && epoch > 0, "p2p_allreduce: bad arguments (n=%d)", n);
  Peers P;
  if (fill_peers(&P, data_ptrs_host, flag_ptrs_host, world)) return 1;
  p2p_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(P, world, rank, epoch, vals, n);
  AB_LAUNCH_CHECK();
  return 0;
}

int atomai_b200_p2p_bn_finalize(void* const* data_ptrs_host, void* const* flag_ptrs_host, int world,
                                int rank, uint64_t epoch, double* stats, int C, double count,
                                const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, float* scale,
                                float* shift, float* mean, float* invstd, void* stream) {
  AB_CHECK(stats && scale && shift && C > 0 && 2 * C <= kSlotDoubles && epoch > 0 && count > 0,
           "p2p_bn_finalize: bad arguments (C=%d)", C);
  Peers P;
  if (fill_peers(&P, data_ptrs_host, flag_ptrs_host, world)) return 1;
  p2p_bn_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(
      P, world, rank, epoch, stats, C, count, gamma, beta, running_mean, running_var, momentum, eps,
      scale, shift, mean, invstd);
  AB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
