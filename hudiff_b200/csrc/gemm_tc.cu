// tcgen05 GEMM / implicit-GEMM dilated convolution for sm_100a.
//
//   D[128 x BN] (fp32, TMEM) = sum over K-steps  A_tile[128 x 64] (fp16, smem)  *  W_tile[BN x 64]^T
//
// One CTA per (M tile, N tile).  Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM
// allocator + MMA issuer (one lane), warps 2..5 = epilogue (one thread per accumulator row,
// TMEM lane quadrant = warp_idx % 4).  A K-step is one (tap, 64-channel chunk) pair; the A tile
// of tap t is the activation tile shifted by (t - taps/2) * dilation rows, fetched through a
// 3-D tensor map (channels, rows-in-chain, sequence) whose out-of-bounds fill supplies the
// convolution's zero padding and the ragged chain tails.  In fp16x3 mode the K loop runs three
// passes (hi*hi, hi*lo, lo*hi) into the same accumulator.
#include "hd_common.cuh"
#include "hd_kernels.h"

namespace hd {

constexpr int BM = 128;
constexpr int BK = 64;               // fp16 elements per K-step = one 128-byte swizzle atom
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 192;

struct TmemLoader {
  uint32_t taddr;                    // TMEM address of (lane quadrant base, column 0)
  __device__ __forceinline__ void load32(int col, float* v) { tmem_ld32(taddr + static_cast<uint32_t>(col), v); }
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams G, int stages, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // which problem (heavy / light chain) and which tile
  int tile_m = blockIdx.x;
  int pi = 0;
  if (G.nprob > 1 && tile_m >= G.p[0].m_tiles) { pi = 1; tile_m -= G.p[0].m_tiles; }
  const GemmProblem& P = G.p[pi];
  const int n0 = blockIdx.y * P.BN;
  if (n0 >= P.N) return;             // grid.y is sized for the wider problem

  const int stage_bytes = A_TILE_BYTES + P.BN * BK * 2;
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int nk1 = P.taps * P.kchunks;
  const int nk = G.x3 ? 3 * nk1 : nk1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      // decompose the 128 tile rows into per-sequence parts (box heights are multiples of 8)
      const RowSpace rs = P.rs;
      const int total_rows = rs.nseq * rs.Lcp;
      const int r_begin = tile_m * BM;
      uint32_t a_bytes = 0;
      {
        int r = r_begin, rem = BM;
        while (rem > 0 && r < total_rows) {
          int l0 = r % rs.Lcp;
          int h = min(rem, rs.Lcp - l0);
          a_bytes += static_cast<uint32_t>(h) * BK * 2;
          r += h; rem -= h;
        }
      }
      const uint32_t b_bytes = static_cast<uint32_t>(P.BN) * BK * 2;
      const int half_taps = P.taps / 2;
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, G.err_flag, 101);
        const int pass = ks / nk1;                     // 0: hi*hi  1: hi*lo  2: lo*hi
        const int k1 = ks - pass * nk1;
        const int tap = k1 / P.kchunks;
        const int kc = k1 - tap * P.kchunks;
        const CUtensorMap* amaps = (pass == 2) ? P.a_maps_lo : P.a_maps_hi;
        const CUtensorMap* bmap = (pass == 1) ? P.b_map_lo : P.b_map_hi;
        uint8_t* a_dst = tiles + s * stage_bytes;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes);
        const int shift = (tap - half_taps) * P.dil;
        int r = r_begin, rem = BM;
        while (rem > 0 && r < total_rows) {
          int seq = r / rs.Lcp;
          int l0 = r - seq * rs.Lcp;
          int h = min(rem, rs.Lcp - l0);
          tma_load_3d(a_dst + (BM - rem) * (BK * 2), amaps + (h / 8 - 1), &full_bar[s], kc * BK, l0 + shift, seq);
          r += h; rem -= h;
        }
        const int bk = (tap * P.kchunks + kc) * BK;
        for (int nb = 0; nb < P.BN; nb += P.b_box_n)
          tma_load_2d(b_dst + nb * (BK * 2), bmap, &full_bar[s], bk, n0 + nb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % stages;
        const uint32_t ph = (ks / stages) & 1;
        mbar_wait(&full_bar[s], ph, G.err_flag, 102);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * stage_bytes);
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t adesc = umma_smem_desc_sw128(a_addr + kk * 32);
          for (int nc = 0; nc < P.BN; nc += 256) {
            const int n_instr = min(256, P.BN - nc);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + nc * (BK * 2) + kk * 32);
            umma_f16(tmem_base + static_cast<uint32_t>(nc), adesc, bdesc, umma_idesc_f16(n_instr),
                     (ks > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);          // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);            // accumulator complete
    }
  } else {
    // ===================== epilogue (4 warps, one row per thread) =====================
    mbar_wait(tmem_full_bar, 0, G.err_flag, 103);
    tc_fence_after();
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    TmemLoader ld{tmem_base + (static_cast<uint32_t>(quad * 32) << 16)};
    run_epilogue(P, G, tile_m * BM + row_in_tile, n0, ld);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

// ----------------------------------------------------------------------------------------
cudaError_t gemm_tc_init() {
  return cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
cudaError_t kernels_init() {
  cudaError_t e = gemm_tc_init();
  if (e != cudaSuccess) return e;
  return attention_init();
}

size_t gemm_tc_smem_bytes(int BN, int stages) {
  return static_cast<size_t>(stages) * (A_TILE_BYTES + BN * BK * 2) + (2 * stages + 1) * 8 + 16 + 1024;
}

int gemm_tc_pick_stages(int BN, int nk) {
  int s = 6;
  while (s > 1 && gemm_tc_smem_bytes(BN, s) > 227 * 1024) --s;
  if (s > nk) s = nk > 0 ? nk : 1;
  return s;
}

cudaError_t gemm_tc_launch(const GemmParams& G, cudaStream_t stream) {
  int m_tiles = 0, n_tiles = 1, BN = 0, nk = 1 << 30;
  for (int i = 0; i < G.nprob; ++i) {
    const GemmProblem& P = G.p[i];
    m_tiles += P.m_tiles;
    n_tiles = max(n_tiles, (P.N + P.BN - 1) / P.BN);
    BN = max(BN, P.BN);
    nk = min(nk, P.taps * P.kchunks * (G.x3 ? 3 : 1));
  }
  int tmem_cols = 32;
  while (tmem_cols < BN) tmem_cols <<= 1;
  const int stages = gemm_tc_pick_stages(BN, nk);
  const size_t smem = gemm_tc_smem_bytes(BN, stages);
  dim3 grid(m_tiles, n_tiles, 1);
  gemm_tc_kernel<<<grid, GEMM_THREADS, smem, stream>>>(G, stages, tmem_cols);
  return cudaGetLastError();
}

}  // namespace hd
