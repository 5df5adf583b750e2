// Peer-memory plumbing for the fused compute + gather (SURVEY.md 8(e) phase 2; replaces nn.DataParallel's gather of the
// reference, train.py:33 / synthesize.py:62).  One process per GPU: rank 0 owns a receive region, every other rank maps it
// through CUDA IPC and points the fp32 outputs of its LAST kernels (mel_linear and the final PostNet convolution: the
// `out_f32` / `out2_f32` stores in the conv1d_tc epilogue) straight at its slice, so the mel tensors cross NVLink as they
// are produced by the tensor-core kernel that computes them -- no staging copy, no separate collective.  What is left is
// a completion protocol, two tiny kernels:
//   signal: after the rank's last kernel (stream order), one thread publishes a monotonically increasing counter with a
//           system-scope release store into rank 0's flag array;
//   wait  : rank 0 spins (system-scope acquire loads, watchdog) until every rank's counter has reached the step.
// Flow control for buffer reuse runs the same way in the other direction (rank 0 -> per-rank ack flags).
#include "common.cuh"

namespace sb {
namespace {

__global__ void peer_signal_kernel(unsigned long long* flag, unsigned long long value) {
  __threadfence_system();          // everything this stream wrote before (previous kernels) is ordered before the flag
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// flags[i * stride] for i in [0, n): wait until all are >= value.  One warp; lane i polls flags i, i+32, ...
__global__ void peer_wait_kernel(const unsigned long long* flags, int n, long long stride, unsigned long long value) {
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long* f = flags + static_cast<long long>(i) * stride;
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= value) break;
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL) {    // ~10 s: a peer died or the protocol is broken -- fail loudly, do not hang the GPU
        printf("styler_b200: peer wait watchdog (flag %d holds %llu, waiting for %llu)\n", i, v, value);
        __trap();
      }
    } while (true);
  }
  __threadfence_system();
}

}  // namespace
}  // namespace sb

using namespace sb;

extern "C" int styler_peer_alloc(int64_t bytes, void** dptr, void* handle64) {
  SB_REQUIRE(bytes > 0 && dptr != nullptr && handle64 != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  SB_CUDA_OK(cudaMalloc(&p, static_cast<size_t>(bytes)));
  SB_CUDA_OK(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  memcpy(handle64, &h, sizeof(h));
  *dptr = p;
  return 0;
}

extern "C" int styler_peer_open(const void* handle64, void** dptr) {
  SB_REQUIRE(handle64 != nullptr && dptr != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  SB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr = p;
  return 0;
}

extern "C" int styler_peer_close(void* dptr) {
  SB_REQUIRE(dptr != nullptr, "peer_close: null pointer");
  SB_CUDA_OK(cudaIpcCloseMemHandle(dptr));
  return 0;
}

extern "C" int styler_peer_free(void* dptr) {
  SB_REQUIRE(dptr != nullptr, "peer_free: null pointer");
  SB_CUDA_OK(cudaFree(dptr));
  return 0;
}

extern "C" int styler_peer_signal(void* flag, uint64_t value, void* stream) {
  SB_REQUIRE(flag != nullptr && (reinterpret_cast<uintptr_t>(flag) & 7) == 0, "peer_signal: flag must be 8-byte aligned");
  peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long*>(flag), value);
  SB_LAUNCH_OK();
  return 0;
}

extern "C" int styler_peer_wait(const void* flags, int32_t n, int64_t stride, uint64_t value, void* stream) {
  SB_REQUIRE(flags != nullptr && n > 0 && (reinterpret_cast<uintptr_t>(flags) & 7) == 0, "peer_wait: bad arguments");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned long long*>(flags), n, stride, value);
  SB_LAUNCH_OK();
  return 0;
}
