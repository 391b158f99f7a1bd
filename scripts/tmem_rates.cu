// Tensor-memory read / write rate probe (sm_100a): cycles per tcgen05.ld / tcgen05.st 32x32b.x32 (4 KB per warp
// instruction) with 1, 2, 4, 8 warps (warp w touches the lane quadrant w % 4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tmem_rates.bin scripts/tmem_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD32(taddr, r) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), \
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), \
        "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define ST32(taddr, r) asm volatile( \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), \
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

template <int MODE>       // 0: loads, one wait per 4 loads; 1: stores, one wait per 4; 2: load + wait each (latency)
__global__ void probe(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[4][32];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < 32; ++i) r[q][i] = threadIdx.x + i + q;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      LD32(base, r[0]); LD32(base + 32, r[1]); LD32(base + 64, r[2]); LD32(base + 96, r[3]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0][it & 31] + r[1][3] + r[2][5] + r[3][7];
    } else if (MODE == 1) {
      ST32(base, r[0]); ST32(base + 32, r[1]); ST32(base + 64, r[2]); ST32(base + 96, r[3]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      LD32(base, r[0]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0][it & 31];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int per_iter) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 1024 * 148 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  for (int warps = 1; warps <= 16; warps *= 2) {
    probe<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = double(h) / (double(iters) * per_iter);
    printf("%-22s warps %2d: %7.1f cycles per x32 instruction per warp, %7.1f B/clk/SM  (%s)\n", name, warps, per,
           4096.0 * warps / per, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tcgen05.ld x32 (x4)", 4); run<1>("tcgen05.st x32 (x4)", 4); run<2>("tcgen05.ld x32 + wait", 1);
  return 0;
}
