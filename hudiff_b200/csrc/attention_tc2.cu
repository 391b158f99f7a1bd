// tcgen05 attention core, second generation (fp16-operand policy):  O = softmax(Q K^T) V.
//
// Persistent: one CTA per SM walks the (sequence, head) work items.  Per item K and V (<= 320 keys
// x 64) sit in shared memory once; the 128-row query tiles stream through a 2-deep Q ring.  The
// loads of the NEXT item are issued the moment the tensor core has finished with the buffer they
// replace (K after the item's last S = Q K^T, V after its last P V), so they land underneath the
// softmax of the current tile and no TMA latency is exposed.
//
//   warp 0 (one lane)   TMA producer + tcgen05.mma issuer
//   warps 4..19         softmax: thread = (query row, one of four key ranges); four warps per
//                       scheduler hide the MUFU / TMEM latencies
//
// Per tile a thread reads its S row segment (<= 80 fp32 columns) from TMEM twice in 16-column
// chunks (max pass, exp pass; the next chunk is in flight while the current one is processed), the
// row maximum is exchanged between the four key ranges through shared memory,
// p = exp2(s*log2e - max*log2e) is written as fp16 into
// the K-major 128B-swizzled A-operand tiles of the P V MMA, and the O tile of the PREVIOUS query
// tile (double-buffered in TMEM) is read back, scaled by 1/rowsum and stored while the tensor core
// already works on S of the next tile.  The [L, L] score matrix never leaves TMEM / shared memory;
// there is no mask (pad slots are attended exactly like the reference, cross_attention.py:168-170).
// q already carries 1/sqrt(dh) and the rotary phase (QKV GEMM epilogue).
#include "hd_common.cuh"
#include "hd_kernels.h"

#include <cstdlib>

namespace hd {

constexpr int A2_THREADS = 640;          // warps 0-3: warp 0 = TMA + MMA issuer (1-3 idle); warps 4..19: softmax
constexpr int A2_BQ = 128;
constexpr int A2_D = 64;
constexpr int A2_MAXC = 5;              // 16-key chunks per softmax thread (4 key groups: L <= 320)
constexpr int A2_O_COL0 = 320;          // TMEM: S in [0, 320), O buffers at 320 and 384
constexpr int A2_O_STRIDE = 64;

struct Attn2Params {
  const CUtensorMap* map_q;             // box 64 x 128 rows over the q|k|v operand plane (3-D: cols, rows-in-seq, seq)
  const CUtensorMap* map_kv;            // box 64 x (LK / 2) rows
  __half* o_hi; int ld_o;
  int L, att_dim, nhead, n_items;       // n_items = B * nhead
  int LK, LN;                           // LK = key rows held in smem (multiple of 64); LN = S width (multiple of 16)
  int pdl;
  int l2hint;                           // 1: q / k / v tiles are fetched with an L2 evict-first policy (read once per item)
  int reverse;                          // 1: items walk from the last sequence to the first (L2 reuse, GemmParams::reverse)
  int* err_flag;
};

__device__ __forceinline__ uint32_t a2_idesc_bmn(int n) { return umma_idesc_f16(n) | (1u << 16); }

// 32 consecutive fp32 TMEM columns of this thread's lane, WITHOUT waiting (pair with a2_tmem_wait)
__device__ __forceinline__ void a2_tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void a2_tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void a2_tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// phase clocks of CTA 0 (diagnostic: softmax warp of quadrant 0 / group 0 and the issuer), see hudiff_debug_attn_clocks
__device__ long long g_a2_clk[16];

// 2^x for x <= 0 on the FMA pipe: x = n + f, n = round(x), |f| <= 0.5; 2^f by a degree-4 polynomial,
// 2^n by adding n to the exponent field.  Inputs below -120 are clamped (result < 1e-36, rounds to 0 in fp16).
__device__ __forceinline__ float a2_ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.00961812911f, f, 0.0555041087f);
  p = fmaf(p, f, 0.240226507f);
  p = fmaf(p, f, 0.693147182f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// NPOLY of every 16 exponentials are evaluated by the FMA-pipe polynomial instead of the MUFU unit (HUDIFF_ATTN_POLY,
// measured on B200: 0 / 2 / 4 / 6 within noise of each other, 8 slower; 4 is the default).  A packed variant
// (ex2.approx.f16x2 on fp16-rounded exponents, fp16 pair sums) was also tried: SASS splits it into two MUFU.EX2.F16
// per pair, so it saves no MUFU work and measured 2 % slower end to end.
template <int NPOLY>
__global__ void __launch_bounds__(A2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ Attn2Params A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int L = A.L, LK = A.LK;
  const int kv_bytes = LK * 128;
  uint8_t* sQ = smem;                                  // [2][128][64] fp16
  uint8_t* sK = sQ + 2 * A2_BQ * 128;                  // [LK][64]
  uint8_t* sV = sK + kv_bytes;                         // [LK][64]
  uint8_t* sP = sV + kv_bytes;                         // [LK/64][128][64]  (A operand of P V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (LK / 64) * (A2_BQ * 128));
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_q = bars + 2;                          // [2]
  uint64_t* bar_s = bars + 4;                          // S accumulator complete
  uint64_t* bar_p = bars + 5;                          // P written (256 arrivals); S region free again
  uint64_t* bar_o = bars + 6;                          // [2] O accumulator complete
  uint64_t* bar_ofree = bars + 8;                      // [2] O read back (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);
  float* pmax = reinterpret_cast<float*>(bars + 12);   // [4 groups][128]
  float* psum = pmax + 4 * A2_BQ;                      // [2 tiles][4 groups][128]

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1);
    // one arrival per softmax warp (lane 0 after __syncwarp): 512 per-thread arrivals on one barrier word
    // serialise in the shared-memory atomic unit
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
  const int nt = (L + A2_BQ - 1) / A2_BQ;                                   // query tiles per item
  const int my_items = (A.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int G = my_items * nt;                                              // tiles this CTA processes

  // Register rebalancing: the issuer warpgroup (warps 0-3, three of them idle) hands most of its registers to the
  // sixteen softmax warps, whose exp loop spilled at the 96 registers that 640 threads allow.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 0 && G > 0) {
      // ===================== TMA producer + MMA issuer =====================
      // The whole warp walks the schedule in uniform control flow (barrier waits included); one
      // elected lane issues the copies and the MMAs.
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      const uint64_t pol = A.l2hint ? l2_policy_evict_first() : 0ull;
      auto load3 = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        if (pol) tma_load_3d_hint(dst, m, bar, c0, c1, c2, pol);
        else tma_load_3d(dst, m, bar, c0, c1, c2);
      };
      auto item_bh = [&](int i, int& b, int& h) {
        int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        if (A.reverse) item = A.n_items - 1 - item;
        b = item / A.nhead; h = item - b * A.nhead;
      };
      auto load_q = [&](int g) {
        int b, h; item_bh(g / nt, b, h);
        uint64_t* bar = &bar_q[g & 1];
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, A2_BQ * 128);
          load3(sQ + (g & 1) * (A2_BQ * 128), A.map_q, bar, h * A2_D, (g % nt) * A2_BQ, b);
        }
        __syncwarp();
      };
      auto load_kv = [&](int i, bool v) {
        int b, h; item_bh(i, b, h);
        uint64_t* bar = v ? bar_v : bar_k;
        uint8_t* dst = v ? sV : sK;
        const int col = (v ? 2 : 1) * A.att_dim + h * A2_D;
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kv_bytes));
          load3(dst, A.map_kv, bar, col, 0, b);
          load3(dst + kv_bytes / 2, A.map_kv, bar, col, LK / 2, b);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int gp) {
        const int ip = gp / nt, qtp = gp - ip * nt;
        if (gp >= 2) mbar_wait(&bar_ofree[gp & 1], ((gp - 2) >> 1) & 1, A.err_flag, 305);
        if (qtp == 0) mbar_wait(bar_v, ip & 1, A.err_flag, 306);
        tc_fence_after();
        const uint32_t d = tmem_base + A2_O_COL0 + (gp & 1) * A2_O_STRIDE;
        const int nks = A.LN >> 4;
        if (elect_one()) {
          const uint64_t pdesc = umma_smem_desc_sw128(p_addr), vdesc = umma_smem_desc_sw128(v_addr);
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t adesc = pdesc + static_cast<uint64_t>(((ks >> 2) * (A2_BQ * 128) + (ks & 3) * 32) >> 4);
            const uint64_t bdesc = vdesc + static_cast<uint64_t>((ks * 16 * 128) >> 4);
            umma_f16(d, adesc, bdesc, a2_idesc_bmn(A2_D), ks > 0 ? 1u : 0u);
          }
          umma_commit(&bar_o[gp & 1]);
        }
        __syncwarp();
      };
      load_kv(0, false);
      load_q(0);
      load_kv(0, true);
      if (G > 1) load_q(1);
      for (int g = 0; g < G; ++g) {
        const int i = g / nt, qt = g - i * nt;
        if (g > 0) mbar_wait(bar_p, (g - 1) & 1, A.err_flag, 301);          // S region free, P(g-1) ready
        mbar_wait(&bar_q[g & 1], (g >> 1) & 1, A.err_flag, 302);
        if (qt == 0) mbar_wait(bar_k, i & 1, A.err_flag, 303);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + (g & 1) * (A2_BQ * 128)));
          const uint64_t kdesc = umma_smem_desc_sw128(k_addr);
          const int n_first = min(256, A.LN), n_second = A.LN - n_first;
#pragma unroll
          for (int kk = 0; kk < A2_D / 16; ++kk) {
            umma_f16(tmem_base, qdesc + 2 * kk, kdesc + 2 * kk, umma_idesc_f16(n_first), kk > 0 ? 1u : 0u);
            if (n_second > 0)
              umma_f16(tmem_base + 256u, qdesc + 2 * kk, kdesc + ((256 * 128) >> 4) + 2 * kk, umma_idesc_f16(n_second),
                       kk > 0 ? 1u : 0u);
          }
          umma_commit(bar_s);
        }
        __syncwarp();
        if (g > 0) issue_pv(g - 1);
        // S(g) complete: its Q buffer is free; after the item's last tile K is free too
        mbar_wait(bar_s, g & 1, A.err_flag, 304);
        if (g + 2 < G) load_q(g + 2);
        if (qt == nt - 1 && i + 1 < my_items) load_kv(i + 1, false);
        if (g > 0) {
          const int gp = g - 1, ip = gp / nt, qtp = gp - ip * nt;
          if (qtp == nt - 1 && ip + 1 < my_items) {                         // V is free once the item's last P V is done
            mbar_wait(&bar_o[gp & 1], (gp >> 1) & 1, A.err_flag, 307);
            load_kv(ip + 1, true);
          }
        }
      }
      mbar_wait(bar_p, (G - 1) & 1, A.err_flag, 308);
      issue_pv(G - 1);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    // ===================== softmax / output warps =====================
    // 16 warps: quadrant = warp & 3 (the TMEM lanes a warp may touch), group = one of 4 key ranges.
    // Two passes over the thread's S row segment straight from TMEM (TMEM reads are cheap; four
    // warps per scheduler hide the MUFU / TMEM latencies that two register-heavy warps cannot).
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;                       // 0..3
    const int row_in_tile = quad * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const int nc_total = (L + 15) >> 4;                    // 16-key chunks with at least one real key
    const int nc_base = nc_total >> 2, nc_rem = nc_total & 3;
    const int nc = nc_base + (grp < nc_rem ? 1 : 0);       // chunks of this thread (<= A2_MAXC)
    const int ch_begin = grp * nc_base + min(grp, nc_rem);
    const int k_begin = ch_begin * 16;
    float inv_prev = 0.f;
    const bool clk_on = (blockIdx.x == 0 && warp == 4 && lane == 0);
    unsigned c_wait = 0, c_ld = 0, c_max = 0, c_exp = 0, c_rb = 0, c_sync = 0, t0 = 0, t1 = 0;

    // O tile of tile gp (complete in TMEM) -> fp16 rows of the out-projection operand (16 columns per group)
    auto readback = [&](int gp, float inv) {
      const int ip = gp / nt, qtp = gp - ip * nt;
      int item = static_cast<int>(blockIdx.x) + ip * static_cast<int>(gridDim.x);
      if (A.reverse) item = A.n_items - 1 - item;
      const int b = item / A.nhead, h = item - b * A.nhead;
      const int row = qtp * A2_BQ + row_in_tile;
      const bool live = (qtp * A2_BQ + quad * 32) < L;
      mbar_wait(&bar_o[gp & 1], (gp >> 1) & 1, A.err_flag, 310);
      tc_fence_after();
      uint32_t o[16];
      if (live) {
        a2_tmem_ld16_nowait(taddr + static_cast<uint32_t>(A2_O_COL0 + (gp & 1) * A2_O_STRIDE + grp * 16), o);
        a2_tmem_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ofree[gp & 1]);
      if (live && row < L) {
        uint32_t ok[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const __half2 ho = __floats2half2_rn(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
          ok[j >> 1] = *reinterpret_cast<const uint32_t*>(&ho);
        }
        uint4* dst = reinterpret_cast<uint4*>(A.o_hi + (static_cast<long long>(b) * L + row) * A.ld_o + h * A2_D + grp * 16);
        dst[0] = make_uint4(ok[0], ok[1], ok[2], ok[3]);
        dst[1] = make_uint4(ok[4], ok[5], ok[6], ok[7]);
      }
    };

    for (int g = 0; g < G; ++g) {
      const int i = g / nt, qt = g - i * nt;
      const bool live = (qt * A2_BQ + quad * 32) < L;           // warp-uniform: some row of this warp exists
      if (clk_on) t0 = clock();
      mbar_wait(bar_s, g & 1, A.err_flag, 309);
      tc_fence_after();
      if (clk_on) { t1 = clock(); c_wait += t1 - t0; t0 = t1; }
      // ---- pass 1: row maximum over this thread's keys.  tcgen05.wait::ld drains every outstanding
      // load, so the chunks are fetched in pairs with one wait per pair.
      float mx = -INFINITY;
      if (live && nc > 0) {
        uint32_t vv[2][16];
        auto chunk_max = [&](const uint32_t* cur, int c) {
          const int k0 = k_begin + c * 16;
          if (k0 + 16 <= L) {
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(cur[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (k0 + j < L) mx = fmaxf(mx, __uint_as_float(cur[j]));
          }
        };
#pragma unroll
        for (int c0 = 0; c0 < A2_MAXC; c0 += 2) {
          if (c0 < nc) {
            a2_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin + c0 * 16), vv[0]);
            if (c0 + 1 < nc) a2_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin + (c0 + 1) * 16), vv[1]);
            a2_tmem_wait();
            chunk_max(vv[0], c0);
            if (c0 + 1 < nc) chunk_max(vv[1], c0 + 1);
          }
        }
      }
      if (clk_on) { t1 = clock(); c_ld += t1 - t0; t0 = t1; }
      pmax[grp * A2_BQ + row_in_tile] = mx;
      // the four key-range warps of a TMEM lane quadrant exchange their maxima (128-thread named barrier)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
      mx = fmaxf(fmaxf(pmax[row_in_tile], pmax[A2_BQ + row_in_tile]),
                 fmaxf(pmax[2 * A2_BQ + row_in_tile], pmax[3 * A2_BQ + row_in_tile]));
      if (g > 0) {                                           // the same barrier published the previous tile's partial sums
        const float* ps = psum + ((g - 1) & 1) * (4 * A2_BQ);
        inv_prev = 1.0f / ((ps[row_in_tile] + ps[A2_BQ + row_in_tile]) + (ps[2 * A2_BQ + row_in_tile] + ps[3 * A2_BQ + row_in_tile]));
      }
      const float mxs = mx * LOG2E;
      float sum = 0.f;
      // P(g) overwrites the shared-memory tiles the P V product of the PREVIOUS tile reads: that MMA is queued behind
      // S(g) and, on the first tile of an item, additionally waits for the item's V load, so "S(g) complete" does not
      // imply "P V (g-1) complete".  (Round 1 had no such wait: with short rows -- Nb, L = 152 -- and a slow V load the
      // probabilities of tile g-1 were overwritten before they were consumed; 2 of 1024 sequences at B = 1024.)
      if (g > 0) {
        mbar_wait(&bar_o[(g - 1) & 1], ((g - 1) >> 1) & 1, A.err_flag, 311);
        tc_fence_after();
      }
      if (clk_on) { t1 = clock(); c_max += t1 - t0; t0 = t1; }
      // ---- pass 2: p = exp2(s*log2e - max*log2e) -> fp16 into the swizzled A-operand tiles of P V.
      // The row sum uses the fp32 values (the fp16 rounding of p is unbiased; its effect on the
      // normaliser is ~2^-12 / sqrt(L)).  (Row sums on the tensor core -- P x ones, 19 extra N = 16 MMAs per
      // tile -- were tried in round 2: each small MMA costs far more than its 8-cycle data floor, the P V
      // product finished ~950 cycles later and the tile got 10 % slower.)
      if (live && nc > 0) {
        uint32_t va[16], vb[16];
        a2_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin), va);
        a2_tmem_wait();
#pragma unroll
        for (int c = 0; c < A2_MAXC; ++c) {
          if (c < nc) {
            uint32_t* cur = (c & 1) ? vb : va;
            uint32_t* nxt = (c & 1) ? va : vb;
            if (c + 1 < nc) a2_tmem_ld16_nowait(taddr + static_cast<uint32_t>(k_begin + (c + 1) * 16), nxt);
            const int k0 = k_begin + c * 16;
            // 12 of 16 exponentials on the MUFU unit, 4 as a degree-4 polynomial on the FMA pipe
            // (Cody-Waite split; relative error 4e-5, below the fp16 rounding of p): the MUFU rate
            // (16 / clk / SM) is what bounds this loop
            uint32_t pk[8];
            {
              float p[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float x = fmaf(__uint_as_float(cur[j]), LOG2E, -mxs);
                p[j] = (NPOLY > 0 && ((j + 1) * NPOLY) / 16 != (j * NPOLY) / 16) ? a2_ex2_poly(x) : a2_ex2(x);
              }
              if (k0 + 16 > L) {                                   // warp-uniform: chunk with masked keys
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (k0 + j >= L) p[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                sum += p[j] + p[j + 1];
                const __half2 hp = __floats2half2_rn(p[j], p[j + 1]);
                pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
              }
            }
            uint8_t* prow = sP + (k0 >> 6) * (A2_BQ * 128) + row_in_tile * 128;
            const int jc = (k0 & 63) >> 3;                         // first of two 16-byte pieces inside the 128-byte row
            *reinterpret_cast<uint4*>(prow + ((jc ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(prow + (((jc + 1) ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (c + 1 < nc) a2_tmem_wait();
          }
        }
      }
      psum[(g & 1) * (4 * A2_BQ) + grp * A2_BQ + row_in_tile] = sum;
      if (clk_on) { t1 = clock(); c_exp += t1 - t0; t0 = t1; }
      // generic-proxy writes of P must be visible to the tensor core (async proxy) before the MMA
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (clk_on) { t1 = clock(); c_sync += t1 - t0; t0 = t1; }
      if (g > 0) readback(g - 1, inv_prev);
      if (clk_on) { t1 = clock(); c_rb += t1 - t0; t0 = t1; }
    }
    if (clk_on) {
      g_a2_clk[0] = G; g_a2_clk[1] = c_wait; g_a2_clk[2] = c_ld; g_a2_clk[3] = c_max; g_a2_clk[4] = c_exp; g_a2_clk[5] = c_rb; g_a2_clk[6] = c_sync;
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
    {
      const float* ps = psum + ((G - 1) & 1) * (4 * A2_BQ);
      readback(G - 1, 1.0f / ((ps[row_in_tile] + ps[A2_BQ + row_in_tile]) + (ps[2 * A2_BQ + row_in_tile] + ps[3 * A2_BQ + row_in_tile])));
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

static size_t attn2_smem(int LK) {
  return static_cast<size_t>(2 * A2_BQ) * 128 + 2 * LK * 128 + (LK / 64) * (A2_BQ * 128) + 12 * 8 + (4 + 8) * A2_BQ * 4;
}

static int g_a2_sms = 0;
typedef void (*Attn2Kernel)(Attn2Params);
static Attn2Kernel attn2_kernel(int npoly) {
  switch (npoly) {
    case 0: return attention_tc2_kernel<0>;
    case 2: return attention_tc2_kernel<2>;
    case 6: return attention_tc2_kernel<6>;
    case 8: return attention_tc2_kernel<8>;
    default: return attention_tc2_kernel<4>;
  }
}
static int g_a2_npoly = 4;            // HUDIFF_ATTN_POLY (0, 2, 4, 6, 8): exponentials per 16 on the FMA pipe
cudaError_t attention_tc2_init() {
  if (const char* e = std::getenv("HUDIFF_ATTN_POLY")) g_a2_npoly = std::atoi(e);
  cudaError_t e = cudaFuncSetAttribute(attn2_kernel(g_a2_npoly), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_a2_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t attention_tc2_clocks(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_a2_clk, sizeof(long long) * 16);
}

cudaError_t attention_tc2_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms, int reverse, int l2hint, int pdl) {
  if (L > A2_MAXC * 64) return cudaErrorInvalidValue;
  Attn2Params P;
  P.map_q = map_q; P.map_kv = map_kv; P.o_hi = o_hi; P.ld_o = ld_o; P.L = L; P.att_dim = att_dim; P.nhead = nhead;
  P.n_items = B * nhead;
  P.LK = ((L + 63) / 64) * 64;
  P.LN = ((L + 15) / 16) * 16;
  P.err_flag = err_flag;
  P.reverse = reverse;
  P.l2hint = l2hint;
  if (g_a2_sms <= 0) {
    cudaError_t e = attention_tc2_init();
    if (e != cudaSuccess) return e;
  }
  int sms = g_a2_sms;
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  const int grid = P.n_items < sms ? P.n_items : sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(A2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = attn2_smem(P.LK); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn2_kernel(g_a2_npoly), P);
}

}  // namespace hd
