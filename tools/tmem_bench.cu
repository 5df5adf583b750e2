// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput on sm_100a as a function of the number of warps, the
// instruction shape (.x16 / .x32 / .x64) and the number of co-resident CTAs.  Prints cycles per 64-column (x 32 lanes
// x 4 B = 8 KB) row-block load per warp, and the implied bytes/clk/SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_bench tmem_bench.cu && ./tmem_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int N> struct Ld;
template <> struct Ld<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
  }
};

template <int SHAPE>
__global__ void __launch_bounds__(512) tmem_ld_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 64;
  uint32_t acc = 0;
  uint32_t r[64];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 64; c += SHAPE) Ld<SHAPE>::ld(base + c, r + c);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc ^= r[0] ^ r[21] ^ r[42] ^ r[63];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256) : "memory");
}

template <int SHAPE>
void run(int warps, int ctas_per_sm, int sms) {
  const int iters = 2000;
  long long* out;
  uint32_t* sink;
  const int grid = sms * ctas_per_sm;
  cudaMalloc(&out, sizeof(long long) * 16 * grid);
  cudaMalloc(&sink, 4);
  cudaMemset(out, 0, sizeof(long long) * 16 * grid);
  tmem_ld_kernel<SHAPE><<<grid, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  const double cyc = mx / iters;                          // cycles per 8 KB per warp (all warps concurrently)
  printf("shape x%-3d warps/CTA %2d  CTAs/SM %d : %7.1f clk per 64-col block per warp -> %6.1f B/clk/SM\n", SHAPE, warps,
         ctas_per_sm, cyc, 8192.0 * warps * ctas_per_sm / cyc);
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int cps = 1; cps <= 2; ++cps)
    for (int warps : {1, 4, 8, 16}) {
      run<16>(warps, cps, sms);
      run<32>(warps, cps, sms);
    }
  return 0;
}
