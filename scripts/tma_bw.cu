// TMA feed microbenchmark (diagnostic, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hudiff_b200/csrc scripts/tma_bw.cu -o /tmp/tma_bw -lcuda
// Every CTA streams K-step "stages" (A box 128 rows x 128 B + B box BN rows x 128 B) through a ring
// of smem stages like gemm_tc.cu's producer, with a trivial consumer.  Reports bytes/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "hd_common.cuh"

using namespace hd;

struct Params {
  const CUtensorMap* amap_g; const CUtensorMap* bmap_g;     // descriptors in global memory
  int use_param_desc; int stages; int bn; int ksteps; int tiles; int a_rows_total; int k_total; int do_a; int do_b;
  int b_box; int same_b;
};

__global__ void __launch_bounds__(128, 1)
tma_stream_kernel(const __grid_constant__ CUtensorMap amap_p, const __grid_constant__ CUtensorMap bmap_p,
                  const __grid_constant__ Params P, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 16384 + P.bn * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + P.stages * stage_bytes);
  uint64_t* empty = full + P.stages;
  const CUtensorMap* amap = P.use_param_desc ? &amap_p : P.amap_g;
  const CUtensorMap* bmap = P.use_param_desc ? &bmap_p : P.bmap_g;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t it = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      const int row0 = (t * 128) % (P.a_rows_total - 128);
      for (int ks = 0; ks < P.ksteps; ++ks, ++it) {
        const int s = it % P.stages; const uint32_t ph = (it / P.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1, nullptr, 1);
        uint8_t* dst = smem + s * stage_bytes;
        const uint32_t bytes = (P.do_a ? 16384u : 0u) + (P.do_b ? static_cast<uint32_t>(P.bn) * 128u : 0u);
        mbar_arrive_expect_tx(&full[s], bytes);
        const int kc = (ks * 64) % P.k_total;
        if (P.do_a) tma_load_3d(dst, amap, &full[s], kc, row0, 0);
        if (P.do_b) {
          const int bk = P.same_b ? kc : ((ks + t) * 64) % P.k_total;
          for (int nb = 0; nb < P.bn; nb += P.b_box) tma_load_2d(dst + 16384 + nb * 128, bmap, &full[s], bk, nb);
        }
      }
    }
  } else if (threadIdx.x == 32) {
    uint32_t it = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      for (int ks = 0; ks < P.ksteps; ++ks, ++it) {
        const int s = it % P.stages; const uint32_t ph = (it / P.stages) & 1;
        mbar_wait(&full[s], ph, nullptr, 2);
        mbar_arrive(&empty[s]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  const int R = 74496, K = 768, WN = 384, WK = 2688;
  __half *A, *W;
  cudaMalloc(&A, size_t(R) * K * 2); cudaMemset(A, 0, size_t(R) * K * 2);
  cudaMalloc(&W, size_t(WN) * WK * 2); cudaMemset(W, 0, size_t(WN) * WK * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cuuint32_t estr[3] = {1, 1, 1};
  cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);

  struct Cfg { const char* name; int bn, stages, do_a, do_b, param, b_box, same_b, a_small; };
  std::vector<Cfg> cfgs = {
    {"A+B bn256 s3 global-desc", 256, 3, 1, 1, 0, 128, 1, 0},
    {"A+B bn256 s3 param-desc ", 256, 3, 1, 1, 1, 128, 1, 0},
    {"A only     s3           ", 256, 3, 1, 0, 0, 128, 1, 0},
    {"A only L2-resident s3   ", 256, 3, 1, 0, 0, 128, 1, 1},
    {"B only bn256 s3         ", 256, 3, 0, 1, 0, 128, 1, 0},
    {"B only bn256 s3 box256  ", 256, 3, 0, 1, 0, 256, 1, 0},
    {"B only bn256 s3 distinct", 256, 3, 0, 1, 0, 128, 0, 0},
    {"A+B bn128 s6            ", 128, 6, 1, 1, 0, 128, 1, 0},
    {"A+B bn384 s2            ", 384, 2, 1, 1, 0, 128, 1, 0},
    {"A+B bn64  s8            ", 64, 8, 1, 1, 0, 64, 1, 0},
  };
  for (const Cfg& c : cfgs) {
    const int a_rows = c.a_small ? 8192 : R;     // 8192 x 768 x 2 = 12.6 MB -> L2 resident
    CUtensorMap am, bm;
    cuuint64_t adims[3] = {cuuint64_t(K), cuuint64_t(a_rows), 1}; cuuint64_t astr[2] = {cuuint64_t(K) * 2, cuuint64_t(a_rows) * K * 2};
    cuuint32_t abox[3] = {64, 128, 1};
    CUresult r1 = enc(&am, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, A, adims, astr, abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t bdims[2] = {cuuint64_t(WK), cuuint64_t(WN)}; cuuint64_t bstr[1] = {cuuint64_t(WK) * 2};
    cuuint32_t bbox[2] = {64, cuuint32_t(c.b_box)};
    CUresult r2 = enc(&bm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, W, bdims, bstr, bbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("encode failed %d %d\n", r1, r2); return 1; }
    CUtensorMap* dmaps; cudaMalloc(&dmaps, 2 * sizeof(CUtensorMap));
    cudaMemcpy(dmaps, &am, sizeof(am), cudaMemcpyHostToDevice); cudaMemcpy(dmaps + 1, &bm, sizeof(bm), cudaMemcpyHostToDevice);
    Params P{dmaps, dmaps + 1, c.param, c.stages, c.bn, 12, 582 * 3, a_rows, K, c.do_a, c.do_b, c.b_box, c.same_b};
    const size_t smem = size_t(c.stages) * (16384 + c.bn * 128) + 2 * c.stages * 8 + 1024;
    for (int rep = 0; rep < 2; ++rep) tma_stream_kernel<<<sms, 128, smem>>>(am, bm, P, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tma_stream_kernel<<<sms, 128, smem>>>(am, bm, P, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms); cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = std::max(mx, v);
    const double bytes = double(P.tiles) * P.ksteps * ((c.do_a ? 16384.0 : 0.0) + (c.do_b ? c.bn * 128.0 : 0.0));
    printf("%s : %8.1f us  %6.2f TB/s  %6.1f B/clk/SM (max cycles %lld)  %s\n", c.name, ms * 1e3, bytes / (ms * 1e-3) / 1e12,
           bytes / sms / double(mx), mx, cudaGetErrorString(err));
    cudaFree(dmaps);
  }
  return 0;
}
