// SIMT reference implementation of the GEMM / dilated-convolution contract of gemm_tc.cu.
// Debug and validation only (HUDIFF_GEMM=simt): it reads the same fp16 operand planes and
// packed weights, accumulates in fp32 on CUDA cores into a scratch buffer, and then runs the
// SAME epilogue code (run_epilogue) with one thread per row.  It is how the tcgen05 path is
// checked at sizes the CPU oracle cannot reach.
#include "hd_common.cuh"

namespace hd {

// scratch[tile_row, n] = sum_k A[row(tile_row) shifted][k] * W[n][k]  (all problems of one launch
// use disjoint scratch regions: problem 1 starts after problem 0's tile rows)
__global__ void gemm_ref_accum_kernel(const __grid_constant__ GemmParams G, float* scratch, int ld_scratch) {
  int tile_m = blockIdx.x;
  int pi = 0;
  int row_base = 0;
  if (G.nprob > 1 && tile_m >= G.p[0].m_tiles) { pi = 1; tile_m -= G.p[0].m_tiles; row_base = G.p[0].m_tiles * 128; }
  const GemmProblem& P = G.p[pi];
  const int Kp = P.kchunks * 64;
  const int half = P.taps / 2;
  for (int idx = threadIdx.x; idx < 128 * P.N; idx += blockDim.x) {
    const int r = idx / P.N, n = idx - r * P.N;
    const int tile_row = tile_m * 128 + r;
    const RowRef rr = row_lookup(P.rs, tile_row);
    float acc = 0.f, acc_lo = 0.f;
    if (rr.seq < P.rs.nseq) {
      for (int tap = 0; tap < P.taps; ++tap) {
        const int l = rr.l + (tap - half) * P.dil;
        if (l < 0 || l >= P.rs.Lc) continue;            // zero padding / ragged tail
        const long long g = static_cast<long long>(rr.seq) * P.rs.seq_stride + P.rs.row_off + l;
        const __half* a = P.a_hi + g * P.lda;
        const __half* al = P.a_lo ? P.a_lo + g * P.lda : nullptr;
        const __half* w = P.w_hi + static_cast<long long>(n) * P.ldw + tap * Kp;
        const __half* wl = P.w_lo ? P.w_lo + static_cast<long long>(n) * P.ldw + tap * Kp : nullptr;
        for (int k = 0; k < Kp; ++k) {
          const float av = __half2float(a[k]), wv = __half2float(w[k]);
          acc = fmaf(av, wv, acc);
          if (G.x3) {
            acc_lo = fmaf(av, __half2float(wl[k]), acc_lo);
            acc_lo = fmaf(__half2float(al[k]), wv, acc_lo);
          }
        }
      }
    }
    scratch[static_cast<long long>(row_base + tile_row) * ld_scratch + n] = acc + acc_lo;
  }
}

struct ScratchLoader {
  const float* row;   // scratch row, offset to this CTA's first column
  __device__ __forceinline__ void load32(int col, float* v) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = row[col + j];
  }
};

__global__ void gemm_ref_epilogue_kernel(const __grid_constant__ GemmParams G, const float* scratch, int ld_scratch) {
  int tile_m = blockIdx.x;
  int pi = 0;
  int row_base = 0;
  if (G.nprob > 1 && tile_m >= G.p[0].m_tiles) { pi = 1; tile_m -= G.p[0].m_tiles; row_base = G.p[0].m_tiles * 128; }
  const GemmProblem& P = G.p[pi];
  const int n0 = blockIdx.y * P.BN;
  if (n0 >= P.N) return;
  const int tile_row = tile_m * 128 + threadIdx.x;
  ScratchLoader ld{scratch + static_cast<long long>(row_base + tile_row) * ld_scratch + n0};
  run_epilogue(P, G, tile_row, n0, ld);
}

cudaError_t gemm_ref_launch(const GemmParams& G, float* scratch, size_t scratch_floats, cudaStream_t stream) {
  int m_tiles = 0, n_tiles = 1, ldn = 0;
  for (int i = 0; i < G.nprob; ++i) {
    m_tiles += G.p[i].m_tiles;
    n_tiles = max(n_tiles, (G.p[i].N + G.p[i].BN - 1) / G.p[i].BN);
    ldn = max(ldn, G.p[i].N);
  }
  if (static_cast<size_t>(m_tiles) * 128 * ldn > scratch_floats) return cudaErrorMemoryAllocation;
  gemm_ref_accum_kernel<<<m_tiles, 256, 0, stream>>>(G, scratch, ldn);
  gemm_ref_epilogue_kernel<<<dim3(m_tiles, n_tiles), 128, 0, stream>>>(G, scratch, ldn);
  return cudaGetLastError();
}

}  // namespace hd
