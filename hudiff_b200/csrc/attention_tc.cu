// tcgen05 attention core for sm_100a (fp16-operand policy):  O = softmax(Q K^T) V  per (sequence, head).
//
// One CTA per (sequence, head).  K and V of the whole sequence (<= 320 keys x 64) are brought into
// shared memory once by TMA (128B-swizzled, rows past the sequence end zero-filled by the tensor
// map); the query tiles (128 rows) are processed one after the other:
//   warp 0 (one lane)  TMA loads, then  S[128 x Lk] = Q K^T   (tcgen05.mma, fp32 accumulators in TMEM)
//   warps 1..8         thread = (query row, half of the key chunks): row max and exp2 straight from
//                      TMEM, P written as fp16 into the K-major swizzled A-operand layout in smem
//   warp 0             O[128 x 64] = P V   (V consumed in its natural [key][d] layout as an MN-major
//                      B operand, accumulators in TMEM)
//   warps 1..8         O / rowsum -> fp16 operand rows of the output projection
// The [L, L] score matrix never leaves TMEM / shared memory; there is no mask (pad slots are
// attended exactly like the reference, cross_attention.py:168-170).  q already carries 1/sqrt(dh)
// and the rotary phase (QKV GEMM epilogue).
#include "hd_common.cuh"
#include "hd_kernels.h"

namespace hd {

constexpr int ATC_THREADS = 288;          // warp 0: TMA + MMA issue; warps 1..8: softmax (2 per TMEM lane quadrant)
constexpr int ATC_BQ = 128;
constexpr int ATC_D = 64;
constexpr int ATC_O_COL = 320;          // TMEM column of the O accumulator (S occupies [0, 304))

struct AttnTcParams {
  const CUtensorMap* map_q;             // box 64 x 128 rows over the qkv operand plane
  const CUtensorMap* map_kv;            // box 64 x (LK / 2) rows
  __half* o_hi; int ld_o;
  int L, att_dim, LK, LN;               // LK = keys held in smem (multiple of 64), LN = S width (multiple of 16)
  int* err_flag;
};

// kind::f16 instruction descriptor with an MN-major B operand (V in [key][d] layout)
__device__ __forceinline__ uint32_t umma_idesc_f16_bmn(int n) { return umma_idesc_f16(n) | (1u << 16); }

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttnTcParams A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int L = A.L, LK = A.LK;
  const int kv_bytes = LK * 128;
  uint8_t* sQ = smem;                                  // [128][64] fp16
  uint8_t* sK = sQ + ATC_BQ * 128;                     // [LK][64]
  uint8_t* sV = sK + kv_bytes;                         // [LK][64]
  uint8_t* sP = sV + kv_bytes;                         // [LK/64][128][64]  (A operand of P V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (LK / 64) * (ATC_BQ * 128));
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_q = bars + 1;
  uint64_t* bar_s = bars + 2;                          // S accumulator complete
  uint64_t* bar_p = bars + 3;                          // P written (128 arrivals); S region free again
  uint64_t* bar_o = bars + 4;                          // O accumulator complete
  uint64_t* bar_ofree = bars + 5;                      // O read back (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);
  float* pmax = reinterpret_cast<float*>(bars + 8);    // [2][128] partial row maxima of the two key-chunk groups
  float* psum = pmax + 2 * ATC_BQ;                     // [2][128] partial row sums

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_s, 1);
    mbar_init(bar_p, 256); mbar_init(bar_o, 1); mbar_init(bar_ofree, 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int ntiles = (L + ATC_BQ - 1) / ATC_BQ;
  const int nchunks = LK / 64;

  if (warp == 0) {
    if (lane == 0) {
      const int qc = h * ATC_D, kc = A.att_dim + h * ATC_D, vc = 2 * A.att_dim + h * ATC_D;
      mbar_arrive_expect_tx(bar_kv, 2u * static_cast<uint32_t>(kv_bytes));
      for (int half = 0; half < 2; ++half) {
        tma_load_3d(sK + half * (kv_bytes / 2), A.map_kv, bar_kv, kc, half * (LK / 2), b);
        tma_load_3d(sV + half * (kv_bytes / 2), A.map_kv, bar_kv, vc, half * (LK / 2), b);
      }
      mbar_arrive_expect_tx(bar_q, ATC_BQ * 128);
      tma_load_3d(sQ, A.map_q, bar_q, qc, 0, b);
      mbar_wait(bar_kv, 0, A.err_flag, 201);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      for (int qt = 0; qt < ntiles; ++qt) {
        const uint32_t ph = qt & 1;
        mbar_wait(bar_q, ph, A.err_flag, 202);
        tc_fence_after();
        // S = Q K^T : M = 128, N = LN (<= 256 per instruction), K = 64
#pragma unroll
        for (int kk = 0; kk < ATC_D / 16; ++kk) {
          const uint64_t adesc = umma_smem_desc_sw128(q_addr + kk * 32);
          for (int nc = 0; nc < A.LN; nc += 256) {
            const int n_instr = min(256, A.LN - nc);
            umma_f16(tmem_base + static_cast<uint32_t>(nc), adesc, umma_smem_desc_sw128(k_addr + nc * 128 + kk * 32),
                     umma_idesc_f16(n_instr), kk > 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_s);
        // Q buffer is free once S is complete: fetch the next query tile behind it
        mbar_wait(bar_s, ph, A.err_flag, 203);
        if (qt + 1 < ntiles) {
          mbar_arrive_expect_tx(bar_q, ATC_BQ * 128);
          tma_load_3d(sQ, A.map_q, bar_q, qc, (qt + 1) * ATC_BQ, b);
        }
        // O = P V once the softmax warps have written P (and the previous O has been read back)
        mbar_wait(bar_p, ph, A.err_flag, 204);
        if (qt > 0) mbar_wait(bar_ofree, (qt - 1) & 1, A.err_flag, 205);
        tc_fence_after();
        for (int c = 0; c < nchunks; ++c) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adesc = umma_smem_desc_sw128(p_addr + c * (ATC_BQ * 128) + kk * 32);
            const uint64_t bdesc = umma_smem_desc_sw128(v_addr + (c * 64 + kk * 16) * 128);
            umma_f16(tmem_base + ATC_O_COL, adesc, bdesc, umma_idesc_f16_bmn(ATC_D), (c > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ===================== softmax / output warps =====================
    // thread = (query row, group): group g owns the 64-key chunks c with (c & 1) == g
    const int quad = warp & 3;
    const int grp = (warp - 1) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float v[32];
    for (int qt = 0; qt < ntiles; ++qt) {
      const uint32_t ph = qt & 1;
      const int row = qt * ATC_BQ + row_in_tile;
      mbar_wait(bar_s, ph, A.err_flag, 206);
      tc_fence_after();
      const bool warp_live = (qt * ATC_BQ + quad * 32) < L;     // uniform: some row of this warp exists
      // pass 1: partial row maximum over this group's real keys
      float mx = -INFINITY;
      for (int c = grp; warp_live && c < nchunks; c += 2) {
#pragma unroll
        for (int hch = 0; hch < 2; ++hch) {
          const int k0 = c * 64 + hch * 32;
          if (k0 < L) {
            tmem_ld32(taddr + static_cast<uint32_t>(k0), v);
            if (k0 + 32 <= L) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (k0 + j < L) mx = fmaxf(mx, v[j]);
            }
          }
        }
      }
      pmax[grp * ATC_BQ + row_in_tile] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(pmax[row_in_tile], pmax[ATC_BQ + row_in_tile]);
      const float mxs = mx * LOG2E;
      // pass 2: p = exp(s - max) as fp16 into the swizzled A-operand tiles; partial sum of the rounded values
      float sum = 0.f;
      for (int c = grp; warp_live && c < nchunks; c += 2) {
        uint8_t* prow = sP + c * (ATC_BQ * 128) + row_in_tile * 128;
#pragma unroll
        for (int hch = 0; hch < 2; ++hch) {
          const int k0 = c * 64 + hch * 32;
          uint32_t pk[16];
          if (k0 < L) {
            tmem_ld32(taddr + static_cast<uint32_t>(k0), v);
            const bool full = (k0 + 32 <= L);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = exp2f(fmaf(v[j], LOG2E, -mxs));
              float p1 = exp2f(fmaf(v[j + 1], LOG2E, -mxs));
              if (!full) {
                if (k0 + j >= L) p0 = 0.f;
                if (k0 + j + 1 >= L) p1 = 0.f;
              }
              const __half2 hp = __floats2half2_rn(p0, p1);
              const float2 fp = __half22float2(hp);
              sum += fp.x + fp.y;
              pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int jc = hch * 4 + q;                      // 16-byte chunk inside the 128-byte row
            *reinterpret_cast<uint4*>(prow + ((jc ^ (row_in_tile & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
      }
      psum[grp * ATC_BQ + row_in_tile] = sum;
      // generic-proxy writes of P must be visible to the tensor core (async proxy) before the MMA
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar_p);
      // O / rowsum -> fp16 operand row; each group reads back half of the 64 output columns
      mbar_wait(bar_o, ph, A.err_flag, 207);
      tc_fence_after();
      const float inv = warp_live ? 1.0f / (psum[row_in_tile] + psum[ATC_BQ + row_in_tile]) : 0.f;
      uint32_t ok[16];
      if (warp_live) tmem_ld32(taddr + static_cast<uint32_t>(ATC_O_COL + grp * 32), v);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const __half2 ho = __floats2half2_rn(v[j] * inv, v[j + 1] * inv);
        ok[j >> 1] = *reinterpret_cast<const uint32_t*>(&ho);
      }
      tc_fence_before();
      mbar_arrive(bar_ofree);
      if (row < L) {
        uint4* dst = reinterpret_cast<uint4*>(A.o_hi + (static_cast<long long>(b) * L + row) * A.ld_o + h * ATC_D + grp * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(ok[4 * q], ok[4 * q + 1], ok[4 * q + 2], ok[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t attn_tc_smem(int LK) { return static_cast<size_t>(ATC_BQ) * 128 + 2 * LK * 128 + (LK / 64) * (ATC_BQ * 128) + 64 + 4 * ATC_BQ * 4; }

cudaError_t attention_tc_init() {
  return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t attention_tc_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream) {
  AttnTcParams P;
  P.map_q = map_q; P.map_kv = map_kv; P.o_hi = o_hi; P.ld_o = ld_o; P.L = L; P.att_dim = att_dim;
  P.LK = ((L + 63) / 64) * 64;
  P.LN = ((L + 15) / 16) * 16;
  P.err_flag = err_flag;
  attention_tc_kernel<<<dim3(nhead, B), ATC_THREADS, attn_tc_smem(P.LK), stream>>>(P);
  return cudaGetLastError();
}

}  // namespace hd
