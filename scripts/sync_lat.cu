// Latency microbenchmarks for the GEMM pipeline handshakes (diagnostic).
#include <cstdio>
#include "hd_common.cuh"
using namespace hd;

// mode 0: same-thread  tcgen05.commit -> mbarrier wait
// mode 1: ping-pong between two warps through two mbarriers (plain arrive)
// mode 2: 4 x UMMA(128x256x16) + commit -> wait (same thread)
// mode 3: ping-pong where warp A does commit, warp B does plain arrive
__global__ void __launch_bounds__(128, 1) lat_kernel(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 65536 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tptr, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr;
  long long t0 = 0, t1 = 0;
  if (mode == 0 || mode == 2) {
    if (threadIdx.x == 0) {
      const uint32_t a = smem_u32(smem), b = a + 16384;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (mode == 2) {
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem, umma_smem_desc_sw128(a + kk * 32), umma_smem_desc_sw128(b + kk * 32), umma_idesc_f16(256), kk > 0);
        }
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], i & 1, nullptr, 1);
      }
      t1 = clock64();
      out[0] = (t1 - t0) / iters;
    }
  } else {
    if (warp == 0 && lane == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (mode == 3) umma_commit(&bar[0]); else mbar_arrive(&bar[0]);
        mbar_wait(&bar[1], i & 1, nullptr, 1);
      }
      t1 = clock64();
      out[0] = (t1 - t0) / iters;
    } else if (warp == 2 && lane == 0) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(&bar[0], i & 1, nullptr, 2);
        mbar_arrive(&bar[1]);
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[4] = {"commit -> wait (same thread)      ", "arrive ping-pong (2 warps, RTT)   ", "4xUMMA N256 + commit -> wait      ",
                          "commit / arrive ping-pong (RTT)   "};
  for (int mode = 0; mode < 4; ++mode) {
    lat_kernel<<<1, 128, 70 * 1024>>>(mode, 2000, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s: %lld cycles/iter  (%s)\n", names[mode], h, cudaGetErrorString(e));
  }
  return 0;
}
