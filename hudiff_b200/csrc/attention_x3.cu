// tcgen05 attention core for the split-fp16 ("fp16x3", ~fp32) operand policy:  O = softmax(Q K^T) V with every
// operand carried as hi + lo fp16 planes and every product evaluated as hi*hi + hi*lo + lo*hi into fp32 TMEM
// accumulators (the policy under which the sampling loop is token-for-token equal to the reference).
//
// Same persistent schedule as attention_tc2.cu (one CTA per SM walks the (sequence, head) items; K and V of the
// item sit in shared memory once; 128-row query tiles), with two differences forced by the doubled operands:
//   * the probabilities never go to shared memory: a softmax thread writes P (hi and lo, packed fp16 pairs) with
//     tcgen05.st over the TMEM columns of the 16 scores it has just consumed -- [hi x 8 | lo x 8] per 16-key chunk --
//     and the P V product takes its A operand FROM TENSOR MEMORY (tcgen05.mma, TS form).  That frees the 80 KB the
//     fp16 kernel spends on its P tiles, which is what makes room for the lo planes of K and V;
//   * Q is single-buffered, and because P(g) lives in the S columns the issue order is P V (g-1) before S(g).
//
//   warp 0 (one lane)   TMA producer + tcgen05.mma issuer
//   warps 4..19         softmax: thread = (query row, one of four key ranges), two passes over S in TMEM
//
// Shared memory (Ab, L = 291): Q hi/lo 32 KB + K hi/lo 80 KB + V hi/lo 80 KB = 192 KB.
// Tensor memory: S / P in columns [0, 320), O double-buffered at 320 and 384.
// No mask (pad slots are attended exactly like the reference, cross_attention.py:168-170); q already carries
// 1/sqrt(dh) and the rotary phase (QKV GEMM epilogue).
#include "hd_common.cuh"
#include "hd_kernels.h"

namespace hd {

constexpr int X3_THREADS = 640;
constexpr int X3_BQ = 128;
constexpr int X3_D = 64;
constexpr int X3_MAXC = 5;
constexpr int X3_O_COL0 = 320;
constexpr int X3_O_STRIDE = 64;

struct AttnX3Params {
  const CUtensorMap* maps;              // [0] q hi, [1] k|v hi, [2] q lo, [3] k|v lo  (3-D: cols, rows-in-seq, seq)
  __half* o_hi; __half* o_lo; int ld_o;
  int L, att_dim, nhead, n_items;
  int LK, LN;
  int reverse;
  int* err_flag;
};

__device__ __forceinline__ uint32_t x3_idesc_bmn(int n) { return umma_idesc_f16(n) | (1u << 16); }

// D[tmem] (+)= A[tmem] * B[smem]^T: A = packed fp16 pairs, row m in lane m, elements (2c, 2c+1) of the K slice in column c
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void x3_tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void x3_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void x3_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void x3_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float x3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA pipe (Cody-Waite split + degree-4 polynomial, relative error 4e-5 -- the same offload as
// attention_tc2.cu): 4 of every 16 exponentials, the MUFU unit (49 % busy in the ncu capture) is what bounds the exp pass
__device__ __forceinline__ float x3_ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.00961812911f, f, 0.0555041087f);
  p = fmaf(p, f, 0.240226507f);
  p = fmaf(p, f, 0.693147182f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__global__ void __launch_bounds__(X3_THREADS, 1)
attention_x3_kernel(const __grid_constant__ AttnX3Params A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int L = A.L, LK = A.LK;
  const int kv_bytes = LK * 128;
  uint8_t* sQh = smem;                                 // [128][64] fp16
  uint8_t* sQl = sQh + X3_BQ * 128;
  uint8_t* sKh = sQl + X3_BQ * 128;                    // [LK][64]
  uint8_t* sKl = sKh + kv_bytes;
  uint8_t* sVh = sKl + kv_bytes;
  uint8_t* sVl = sVh + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sVl + kv_bytes);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_q = bars + 2;
  uint64_t* bar_s = bars + 3;                          // S accumulator complete
  uint64_t* bar_p = bars + 4;                          // P written to TMEM (one arrival per softmax warp)
  uint64_t* bar_o = bars + 5;                          // [2] O accumulator complete
  uint64_t* bar_ofree = bars + 7;                      // [2] O read back (one arrival per softmax warp)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);
  float* pmax = reinterpret_cast<float*>(bars + 10);   // [4 groups][128]
  float* psum = pmax + 4 * X3_BQ;                      // [2 tiles][4 groups][128]

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_q, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 16);
    mbar_init(&bar_o[0], 1); mbar_init(&bar_o[1], 1);
    mbar_init(&bar_ofree[0], 16); mbar_init(&bar_ofree[1], 16);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();                                          // prologue above overlaps the previous kernel's tail (HUDIFF_PDL)
  const int nt = (L + X3_BQ - 1) / X3_BQ;
  const int my_items = (A.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int G = my_items * nt;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");       // registers go to the softmax warps
    if (warp == 0 && G > 0) {
      // ===================== TMA producer + MMA issuer =====================
      const CUtensorMap* mq_h = A.maps + 0; const CUtensorMap* mkv_h = A.maps + 1;
      const CUtensorMap* mq_l = A.maps + 2; const CUtensorMap* mkv_l = A.maps + 3;
      auto item_bh = [&](int i, int& b, int& h) {
        int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        if (A.reverse) item = A.n_items - 1 - item;
        b = item / A.nhead; h = item - b * A.nhead;
      };
      auto load_q = [&](int g) {
        int b, h; item_bh(g / nt, b, h);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_q, 2 * X3_BQ * 128);
          tma_load_3d(sQh, mq_h, bar_q, h * X3_D, (g % nt) * X3_BQ, b);
          tma_load_3d(sQl, mq_l, bar_q, h * X3_D, (g % nt) * X3_BQ, b);
        }
        __syncwarp();
      };
      auto load_kv = [&](int i, bool v) {
        int b, h; item_bh(i, b, h);
        uint64_t* bar = v ? bar_v : bar_k;
        uint8_t* dh = v ? sVh : sKh;
        uint8_t* dl = v ? sVl : sKl;
        const int col = (v ? 2 : 1) * A.att_dim + h * X3_D;
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(2 * kv_bytes));
          tma_load_3d(dh, mkv_h, bar, col, 0, b);
          tma_load_3d(dh + kv_bytes / 2, mkv_h, bar, col, LK / 2, b);
          tma_load_3d(dl, mkv_l, bar, col, 0, b);
          tma_load_3d(dl + kv_bytes / 2, mkv_l, bar, col, LK / 2, b);
        }
        __syncwarp();
      };
      const uint64_t qh_desc = umma_smem_desc_sw128(smem_u32(sQh)), ql_desc = umma_smem_desc_sw128(smem_u32(sQl));
      const uint64_t kh_desc = umma_smem_desc_sw128(smem_u32(sKh)), kl_desc = umma_smem_desc_sw128(smem_u32(sKl));
      const uint64_t vh_desc = umma_smem_desc_sw128(smem_u32(sVh)), vl_desc = umma_smem_desc_sw128(smem_u32(sVl));
      // O[gp & 1] = P_hi V_hi + P_hi V_lo + P_lo V_hi, A operand (P) from tensor memory
      auto issue_pv = [&](int gp) {
        const int ip = gp / nt, qtp = gp - ip * nt;
        if (gp >= 2) mbar_wait(&bar_ofree[gp & 1], ((gp - 2) >> 1) & 1, A.err_flag, 405);
        if (qtp == 0) mbar_wait(bar_v, ip & 1, A.err_flag, 406);
        tc_fence_after();
        const uint32_t d = tmem_base + X3_O_COL0 + (gp & 1) * X3_O_STRIDE;
        const int nks = A.LN >> 4;
        if (elect_one()) {
          const uint32_t idesc = x3_idesc_bmn(X3_D);
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t p_hi = tmem_base + static_cast<uint32_t>(ks * 16), p_lo = p_hi + 8u;
            const uint64_t voff = static_cast<uint64_t>((ks * 16 * 128) >> 4);
            umma_f16_ts(d, p_hi, vh_desc + voff, idesc, ks > 0 ? 1u : 0u);
            umma_f16_ts(d, p_hi, vl_desc + voff, idesc, 1u);
            umma_f16_ts(d, p_lo, vh_desc + voff, idesc, 1u);
          }
          umma_commit(&bar_o[gp & 1]);
        }
        __syncwarp();
      };
      load_kv(0, false);
      load_q(0);
      load_kv(0, true);
      for (int g = 0; g < G; ++g) {
        const int i = g / nt, qt = g - i * nt;
        if (g > 0) {
          mbar_wait(bar_p, (g - 1) & 1, A.err_flag, 401);                     // P(g-1) is in TMEM (over the S columns)
          issue_pv(g - 1);                                                    // must precede S(g), which overwrites it
        }
        mbar_wait(bar_q, g & 1, A.err_flag, 402);
        if (qt == 0) mbar_wait(bar_k, i & 1, A.err_flag, 403);
        tc_fence_after();
        if (elect_one()) {
          const int n_first = min(256, A.LN), n_second = A.LN - n_first;
          const uint64_t k2 = static_cast<uint64_t>((256 * 128) >> 4);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {                              // hi*hi, hi*lo, lo*hi
            const uint64_t qd = (pass == 2) ? ql_desc : qh_desc;
            const uint64_t kd = (pass == 1) ? kl_desc : kh_desc;
#pragma unroll
            for (int kk = 0; kk < X3_D / 16; ++kk) {
              const uint32_t acc = (pass > 0 || kk > 0) ? 1u : 0u;
              umma_f16(tmem_base, qd + 2 * kk, kd + 2 * kk, umma_idesc_f16(n_first), acc);
              if (n_second > 0) umma_f16(tmem_base + 256u, qd + 2 * kk, kd + k2 + 2 * kk, umma_idesc_f16(n_second), acc);
            }
          }
          umma_commit(bar_s);
        }
        __syncwarp();
        // S(g) complete (and, the tensor pipe being in order, P V (g-1) too): Q is free; after the item's last tile K
        // is free; after the previous item's last P V its V is free
        mbar_wait(bar_s, g & 1, A.err_flag, 404);
        if (g + 1 < G) load_q(g + 1);
        if (qt == nt - 1 && i + 1 < my_items) load_kv(i + 1, false);
        if (g > 0) {
          const int gp = g - 1, ip = gp / nt, qtp = gp - ip * nt;
          if (qtp == nt - 1 && ip + 1 < my_items) {
            mbar_wait(&bar_o[gp & 1], (gp >> 1) & 1, A.err_flag, 407);
            load_kv(ip + 1, true);
          }
        }
      }
      mbar_wait(bar_p, (G - 1) & 1, A.err_flag, 408);
      issue_pv(G - 1);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    // ===================== softmax / output warps =====================
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const int nc_total = (L + 15) >> 4;
    const int nc_base = nc_total >> 2, nc_rem = nc_total & 3;
    const int nc = nc_base + (grp < nc_rem ? 1 : 0);
    const int ch_begin = grp * nc_base + min(grp, nc_rem);
    const int k_begin = ch_begin * 16;
    float inv_prev = 0.f;

    auto readback = [&](int gp, float inv) {
      const int ip = gp / nt, qtp = gp - ip * nt;
      int item = static_cast<int>(blockIdx.x) + ip * static_cast<int>(gridDim.x);
      if (A.reverse) item = A.n_items - 1 - item;
      const int b = item / A.nhead, h = item - b * A.nhead;
      const int row = qtp * X3_BQ + row_in_tile;
      const bool live = (qtp * X3_BQ + quad * 32) < L;
      mbar_wait(&bar_o[gp & 1], (gp >> 1) & 1, A.err_flag, 410);
      tc_fence_after();
      uint32_t o[16];
      if (live) {
        x3_tmem_ld16_nowait(taddr + static_cast<uint32_t>(X3_O_COL0 + (gp & 1) * X3_O_STRIDE + grp * 16), o);
        x3_tmem_wait_ld();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ofree[gp & 1]);
      if (live && row < L) {
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          __half h0, l0, h1, l1;
          split_f16(__uint_as_float(o[j]) * inv, h0, l0);
          split_f16(__uint_as_float(o[j + 1]) * inv, h1, l1);
          oh[j >> 1] = pack_h2(h0, h1);
          ol[j >> 1] = pack_h2(l0, l1);
        }
        const long long off = (static_cast<long long>(b) * L + row) * A.ld_o + h * X3_D + grp * 16;
        uint4* dh = reinterpret_cast<uint4*>(A.o_hi + off);
        uint4* dl = reinterpret_cast<uint4*>(A.o_lo + off);
        dh[0] = make_uint4(oh[0], oh[1], oh[2], oh[3]); dh[1] = make_uint4(oh[4], oh[5], oh[6], oh[7]);
        dl[0] = make_uint4(ol[0], ol[1], ol[2], ol[3]); dl[1] = make_uint4(ol[4], ol[5], ol[6], ol[7]);
      }
    };

    for (int g = 0; g < G; ++g) {
      const int qt = g % nt;
      const bool live = (qt * X3_BQ + quad * 32) < L;
      mbar_wait(bar_s, g & 1, A.err_flag, 409);
      tc_fence_after();
      // ---- pass 1: row maximum over this thread's keys, chunks fetched three at a time: a tcgen05.ld / wait::ld
      // round trip occupies the SM's tensor-memory read path for ~72 cycles whatever its size (scripts/tmem_rates.cu),
      // and this kernel used to spend nine of them per warp and tile (now six).
      float mx = -INFINITY;
      if (live && nc > 0) {
#pragma unroll
        for (int c0 = 0; c0 < X3_MAXC; c0 += 3) {
          if (c0 < nc) {
            uint32_t vv[3][16];
#pragma unroll
            for (int u = 0; u < 3; ++u)
              if (c0 + u < X3_MAXC && c0 + u < nc) x3_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin + (c0 + u) * 16), vv[u]);
            x3_tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              if (c0 + u < X3_MAXC && c0 + u < nc) {
                const int k0 = k_begin + (c0 + u) * 16;
                if (k0 + 16 <= L) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(vv[u][j]));
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (k0 + j < L) mx = fmaxf(mx, __uint_as_float(vv[u][j]));
                }
              }
            }
          }
        }
      }
      pmax[grp * X3_BQ + row_in_tile] = mx;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
      mx = fmaxf(fmaxf(pmax[row_in_tile], pmax[X3_BQ + row_in_tile]),
                 fmaxf(pmax[2 * X3_BQ + row_in_tile], pmax[3 * X3_BQ + row_in_tile]));
      if (g > 0) {
        const float* ps = psum + ((g - 1) & 1) * (4 * X3_BQ);
        inv_prev = 1.0f / ((ps[row_in_tile] + ps[X3_BQ + row_in_tile]) + (ps[2 * X3_BQ + row_in_tile] + ps[3 * X3_BQ + row_in_tile]));
      }
      const float mxs = mx * LOG2E;
      float sum = 0.f;
      // ---- pass 2: p = exp2(s*log2e - max*log2e), split into hi + lo fp16, written over the consumed S columns of
      // the chunk: [hi pairs x 8 | lo pairs x 8].  The row sum uses the fp32 values.
      if (live && nc > 0) {
        // load groups of two chunks (P of a chunk overwrites exactly that chunk's 16 columns, so the stores of one
        // group never touch the scores of the next); three chunks per group spill at the 104 registers of this kernel
#pragma unroll
        for (int c0 = 0; c0 < X3_MAXC; c0 += 2) {
          if (c0 < nc) {
            uint32_t vv[2][16];
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (c0 + u < X3_MAXC && c0 + u < nc) x3_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin + (c0 + u) * 16), vv[u]);
            x3_tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (c0 + u < X3_MAXC && c0 + u < nc) {
                const uint32_t* cur = vv[u];
                const int k0 = k_begin + (c0 + u) * 16;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float p0 = x3_ex2(fmaf(__uint_as_float(cur[j]), LOG2E, -mxs));
                  const float x1 = fmaf(__uint_as_float(cur[j + 1]), LOG2E, -mxs);
                  float p1 = ((j & 2) != 0) ? x3_ex2_poly(x1) : x3_ex2(x1);          // elements 3, 7, 11, 15 of the chunk
                  if (k0 + j >= L) p0 = 0.f;
                  if (k0 + j + 1 >= L) p1 = 0.f;
                  sum += p0 + p1;
                  __half h0, l0, h1, l1;
                  split_f16(p0, h0, l0);
                  split_f16(p1, h1, l1);
                  pk[j >> 1] = pack_h2(h0, h1);
                  pk[8 + (j >> 1)] = pack_h2(l0, l1);
                }
                x3_tmem_st16(taddr + static_cast<uint32_t>(k0), pk);
              }
            }
          }
        }
        x3_tmem_wait_st();
      }
      psum[(g & 1) * (4 * X3_BQ) + grp * X3_BQ + row_in_tile] = sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (g > 0) readback(g - 1, inv_prev);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
    {
      const float* ps = psum + ((G - 1) & 1) * (4 * X3_BQ);
      readback(G - 1, 1.0f / ((ps[row_in_tile] + ps[X3_BQ + row_in_tile]) + (ps[2 * X3_BQ + row_in_tile] + ps[3 * X3_BQ + row_in_tile])));
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

static size_t attn_x3_smem(int LK) {
  return static_cast<size_t>(2 * X3_BQ) * 128 + 4 * static_cast<size_t>(LK) * 128 + 10 * 8 + (4 + 8) * X3_BQ * 4 + 64;
}

static int g_x3_sms = 0;
cudaError_t attention_x3_init() {
  cudaError_t e = cudaFuncSetAttribute(attention_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_x3_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t attention_x3_launch(const CUtensorMap* maps, __half* o_hi, __half* o_lo, int ld_o, int L, int att_dim, int nhead,
                                int B, int* err_flag, cudaStream_t stream, int max_sms, int reverse, int pdl) {
  if (L > X3_MAXC * 64 || !o_lo) return cudaErrorInvalidValue;
  AttnX3Params P;
  P.maps = maps; P.o_hi = o_hi; P.o_lo = o_lo; P.ld_o = ld_o; P.L = L; P.att_dim = att_dim; P.nhead = nhead;
  P.n_items = B * nhead;
  P.LK = ((L + 63) / 64) * 64;
  P.LN = ((L + 15) / 16) * 16;
  P.reverse = reverse;
  P.err_flag = err_flag;
  if (attn_x3_smem(P.LK) > 227 * 1024) return cudaErrorInvalidValue;
  if (g_x3_sms <= 0) {
    cudaError_t e = attention_x3_init();
    if (e != cudaSuccess) return e;
  }
  int sms = g_x3_sms;
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  const int grid = P.n_items < sms ? P.n_items : sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(X3_THREADS, 1, 1);
  cfg.dynamicSmemBytes = attn_x3_smem(P.LK); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attention_x3_kernel, P);
}

}  // namespace hd
