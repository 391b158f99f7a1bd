// tcgen05 attention core, third generation (fp16-operand policy):  O = softmax(Q K^T) V.
//
// attention_tc2.cu is bound by its softmax warps, and those run in lock step: four warps share a query row (one key
// range each), exchange the row maximum through shared memory and read S from tensor memory twice, so every
// scheduler alternates between "all four warps wait on TMEM / a barrier" and "all four warps want the MUFU unit"
// (issue slots 56 % busy, 6255 cycles per 128-query tile).  This kernel removes the coupling:
//
//   * one THREAD owns one query row.  The keys of a sequence (L <= 320) are walked in blocks of 128 columns; a thread
//     pulls its S row segment of a block into registers with ONE pass over tensor memory, so the row maximum and
//     the row sum are private (no exchange, no named barrier, no second TMEM read);
//   * two query tiles are in flight ("slots"), each with its own softmax warpgroup and its own quarter-kilobyte of
//     tensor memory: S block [0,128), P block [128,192), O [192,256).  S is free again the moment the warpgroup has
//     loaded it, so the tensor core computes the NEXT S block of the slot underneath the exponentials of the current
//     one; P is written as packed fp16 pairs into its own columns (tcgen05.st) and the P V product takes its A
//     operand from tensor memory (TS form): no shared memory is spent on P, and K / V are DOUBLE-buffered -- the next
//     item's keys and values land while the current item is processed;
//   * online softmax across the blocks of a row: a later block keeps the running maximum as long as its own
//     maximum exceeds it by less than 2^tau (p <= 256 in fp16); otherwise the running sum and the O accumulator in
//     tensor memory are rescaled (warp-uniform correction branch);
//   * the O tile of a finished query tile is read back AFTER the first block of the slot's next tile has been
//     exponentiated, i.e. when its last P V has long completed;
//   * exponentials: p = exp2(s*log2e - max*log2e) with packed FFMA2 / FADD2 (two fp32 lanes per instruction); NPOLY of
//     every 16 are evaluated by a Cody-Waite + degree-4 polynomial on the FMA pipe so that the MUFU unit
//     (8 cycles per warp instruction) and the FMA pipe (2 cycles per packed instruction) are about evenly loaded
//     (measured rates: scripts/pipe_rates.cu);
//   * the 35-row (Ab) / 24-row (Nb) tail tile of a sequence is placed on a different TMEM lane quadrant -- i.e. a
//     different scheduler -- from item to item, and the tile order alternates between items, so the four schedulers
//     and the two slots carry the same load.
//
//   warp 0   S = Q K^T issuer      warp 1   TMA producer (Q ring, K / V double buffer)      warp 2   P V issuer
//   warps 4-7 / 8-11   softmax + output warpgroup of slot 0 / 1  (thread = TMEM lane = query row)
//
// Shared memory (Ab): Q 4 x 16 KB (ring over the tile sequence), K 2 x 40 KB, V 2 x 40 KB = 224 KB.
// No mask (pad slots are attended exactly like the reference, cross_attention.py:168-170); q already carries
// 1/sqrt(dh) and the rotary phase (QKV GEMM epilogue).  Reference: cross_attention.py:131-173.
#include "hd_common.cuh"
#include "hd_kernels.h"

#include <cstdlib>

namespace hd {

constexpr int A3_CTRL_THREADS = 128;    // warps 0-3
constexpr int A3_BQ = 128;
constexpr int A3_D = 64;
constexpr int A3_BK = 128;              // key columns per S block
constexpr int A3_SLOT = 256;            // TMEM columns per slot: S [0,128)  P [128,192)  O [192,256)
constexpr int A3_P_COL = 128;
constexpr int A3_O_COL = 192;

struct Attn3Params {
  const CUtensorMap* map_q;             // box 64 x 128 rows over the q|k|v operand plane (3-D: cols, rows-in-seq, seq)
  const CUtensorMap* map_kv;            // box 64 x (LK / 2) rows
  __half* o_hi; int ld_o;
  int L, att_dim, nhead, n_items;       // n_items = B * nhead
  int LK;                               // key rows per K / V buffer (multiple of 64)
  int LN, nb;                           // S width of a row (multiple of 16) and its 128-column blocks
  int nt;                               // query tiles per item
  int tail_rows, tail_quads, tail_places;   // rows of an item's last tile, lane quadrants they need, distinct placements
  float tau;                            // log2 of the largest probability tolerated before the maximum is refreshed
  int pdl, l2hint, reverse;
  int* err_flag;
};


__device__ __forceinline__ uint32_t a3_idesc_bmn(int n) { return umma_idesc_f16(n) | (1u << 16); }

// D[tmem] (+)= A[tmem] * B[smem]^T: A = packed fp16 pairs, row m in lane m, elements (2c, 2c+1) of the K slice in column c
__device__ __forceinline__ void a3_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void a3_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void a3_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void a3_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void a3_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void a3_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void a3_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float a3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two fp32 lanes per FMA-pipe instruction (sm_100: FFMA2 / FADD2 on aligned register pairs)
__device__ __forceinline__ void a3_ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void a3_fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 2^x for a pair, x <= ~8, on the FMA pipe: x = n + f, n = round(x), |f| <= 0.5; 2^f by a degree-4 polynomial, 2^n by
// adding n to the exponent field.  Inputs below -120 are clamped (result < 1e-36, rounds to 0 in fp16).
__device__ __forceinline__ void a3_ex2_poly2(float& p0, float& p1, float x0, float x1) {
  x0 = fmaxf(x0, -120.0f); x1 = fmaxf(x1, -120.0f);
  const float MAGIC = 12582912.0f;                  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  float t0, t1, u0, u1, f0, f1;
  a3_fadd2(t0, t1, x0, x1, MAGIC, MAGIC);
  a3_fadd2(u0, u1, t0, t1, -MAGIC, -MAGIC);
  a3_ffma2(f0, f1, u0, u1, -1.0f, -1.0f, x0, x1);
  float q0, q1;
  a3_ffma2(q0, q1, f0, f1, 0.00961812911f, 0.00961812911f, 0.0555041087f, 0.0555041087f);
  a3_ffma2(q0, q1, q0, q1, f0, f1, 0.240226507f, 0.240226507f);
  a3_ffma2(q0, q1, q0, q1, f0, f1, 0.693147182f, 0.693147182f);
  a3_ffma2(q0, q1, q0, q1, f0, f1, 1.0f, 1.0f);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// diagnostic: [0] tiles of CTA 0 / slot 0, [1] its total cycles, [2] cycles waiting for S, [3] cycles waiting for O
__device__ long long g_a3_clk[16];

// 16 scores -> 16 probabilities (8 packed fp16 pairs) + row-sum contribution.  NPOLY of the 16 go to the FMA pipe.
template <int NPOLY>
__device__ __forceinline__ void a3_exp_chunk(const uint32_t* s, float c, float mc, uint32_t* pk, float& a0, float& a1,
                                             float& b0, float& b1) {
#pragma unroll
  for (int jp = 0; jp < 8; ++jp) {
    float x0, x1, p0, p1;
    a3_ffma2(x0, x1, __uint_as_float(s[2 * jp]), __uint_as_float(s[2 * jp + 1]), c, c, mc, mc);
    const bool poly = ((jp + 1) * (NPOLY / 2)) / 8 != (jp * (NPOLY / 2)) / 8;
    if (poly) {
      a3_ex2_poly2(p0, p1, x0, x1);
    } else {
      p0 = a3_ex2(x0); p1 = a3_ex2(x1);
    }
    if (jp & 1) a3_fadd2(b0, b1, b0, b1, p0, p1);
    else a3_fadd2(a0, a1, a0, a1, p0, p1);
    const __half2 hp = __floats2half2_rn(p0, p1);
    pk[jp] = *reinterpret_cast<const uint32_t*>(&hp);
  }
}


// warp-uniform cursor over the tile sequence of this CTA: position n -> (CTA-local item i, position r inside the item)
template <int NQ>
struct A3CursorT {
  int n, i, r;
  int qb, qph;                                          // Q ring buffer n % NQ and the parity of its use count
  __device__ __forceinline__ void advance(int step, int nt) {
    n += step; r += step; qb += step;
    while (r >= nt) { r -= nt; ++i; }
    while (qb >= NQ) { qb -= NQ; qph ^= 1; }
  }
};

// SP = softmax warps per (slot, lane quadrant): 1 -> a thread owns a whole row segment of a block (128 columns in
// registers); 2 -> two warps split the block's columns (64 each) and exchange the block maximum / the row sum through
// shared memory on a 64-thread named barrier: four instead of two softmax warps per scheduler.
template <int NPOLY, int SP>
__global__ void __launch_bounds__(A3_CTRL_THREADS + SP * 256, 1)
attention_tc3_kernel(const __grid_constant__ Attn3Params A) {
  constexpr int CW = A3_BK / SP;                        // S columns of a block per softmax thread
  constexpr int A3_NQ = (SP == 1) ? 4 : 3;              // query tiles in shared memory (ring over the tile sequence)
  typedef A3CursorT<A3_NQ> A3Cursor;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int L = A.L, LK = A.LK, nt = A.nt, nb = A.nb;
  const int kv_bytes = LK * 128;
  uint8_t* sQ = smem;                                  // [4][128][64] fp16: ring over the tile sequence (two per slot)
  uint8_t* sK = sQ + A3_NQ * A3_BQ * 128;              // [2 items][LK][64]
  uint8_t* sV = sK + 2 * kv_bytes;                     // [2 items][LK][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kv_bytes);
  uint64_t* bar_k = bars + 0;                          // [2] keys of item parity landed
  uint64_t* bar_v = bars + 2;                          // [2] values landed
  uint64_t* bar_kfree = bars + 4;                      // [2] every S MMA of the item complete (nt commits)
  uint64_t* bar_vfree = bars + 6;                      // [2] every P V MMA of the item complete (nt commits)
  uint64_t* bar_s = bars + 8;                          // [2 slots] S block complete
  uint64_t* bar_sfree = bars + 10;                     // [2 slots] S block is in registers (one arrival per softmax warp)
  uint64_t* bar_p = bars + 12;                         // [2 slots] P block written (one arrival per softmax warp)
  uint64_t* bar_pfree = bars + 14;                     // [2 slots] P V of the block complete: P columns free, O current
  uint64_t* bar_q = bars + 16;                         // [A3_NQ] query tile landed
  uint64_t* bar_qfree = bars + 20;                     // [A3_NQ] last S MMA of the tile complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);
  float* xmax = reinterpret_cast<float*>(bars + 26);   // [2 parities][2 slots][SP][128] block maxima of the column halves (SP == 2)
  float* xsum = xmax + 2 * 2 * SP * A3_BQ;             // [2 parities][2 slots][SP][128] row sums of the column halves

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      // K / V of an item are free once EACH of its nt tiles has committed its last S block / last P V (the two slots
      // are served independently, so "the item's last tile" says nothing about the tile before it on the other slot)
      mbar_init(&bar_k[b], 1); mbar_init(&bar_v[b], 1); mbar_init(&bar_kfree[b], nt); mbar_init(&bar_vfree[b], nt);
      mbar_init(&bar_s[b], 1); mbar_init(&bar_sfree[b], 4 * SP); mbar_init(&bar_p[b], 4 * SP); mbar_init(&bar_pfree[b], 1);
    }
    for (int b = 0; b < A3_NQ; ++b) { mbar_init(&bar_q[b], 1); mbar_init(&bar_qfree[b], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (A.pdl != 2) pdl_launch_dependents();
  pdl_wait();                                          // prologue above overlaps the previous kernel's tail (HUDIFF_PDL)
  const int my_items = (A.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int G = my_items * nt;                         // tiles of this CTA, walked in sequence order n = 0 .. G-1

  // odd items walk their query tiles backwards, so that the cheap tail tiles alternate between the two slots
  auto tile_qt = [&](int i, int r) -> int { return (i & 1) ? nt - 1 - r : r; };
  // tail tiles start `off` rows early: their live rows sit on TMEM lanes [off, off + tail_rows)
  auto tile_off = [&](int i, int qt) -> int {
    return (nt > 1 && qt == nt - 1) ? (i % A.tail_places) * A.tail_quads * 32 : 0;
  };
  auto item_bh = [&](int i, int& b, int& h) {
    int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
    if (A.reverse) item = A.n_items - 1 - item;
    b = item / A.nhead; h = item - b * A.nhead;
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");       // registers go to the softmax warpgroups
    if ((warp == 0 || warp == 2) && G > (warp >> 1)) {
      // ===================== MMA issuer of slot warp / 2 =====================
      // Whole warp in uniform control flow (barrier waits included); one elected lane issues.  The events of ONE slot
      // come in a fixed order -- S block j in registers, P block j written, S block j + 1 in registers, ... -- so an
      // issuer per slot can block on them in that order and a slow warpgroup never holds back the other slot
      // (a single issuer polling both slots spent a quarter of the kernel's instructions spinning).
      const int w = warp >> 1;
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(w * A3_SLOT);
      auto issue_s = [&](const A3Cursor& c, int k) {
        if (k == 0) mbar_wait(&bar_q[c.qb], static_cast<uint32_t>(c.qph), A.err_flag, 501);
        mbar_wait(&bar_k[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 502);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qdesc = umma_smem_desc_sw128(q_addr + c.qb * (A3_BQ * 128));
          const uint64_t kdesc = umma_smem_desc_sw128(k_addr + (c.i & 1) * kv_bytes + k * (A3_BK * 128));
          const uint32_t idesc = umma_idesc_f16(min(A3_BK, A.LN - k * A3_BK));
#pragma unroll
          for (int kk = 0; kk < A3_D / 16; ++kk) umma_f16(s_tmem, qdesc + 2 * kk, kdesc + 2 * kk, idesc, kk > 0 ? 1u : 0u);
          umma_commit(&bar_s[w]);
          if (k == nb - 1) {
            umma_commit(&bar_qfree[c.qb]);
            umma_commit(&bar_kfree[c.i & 1]);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](const A3Cursor& c, int k) {
        mbar_wait(&bar_v[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 503);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vdesc = umma_smem_desc_sw128(v_addr + (c.i & 1) * kv_bytes + k * (A3_BK * 128));
          const uint32_t idesc = a3_idesc_bmn(A3_D);
          const uint32_t d = s_tmem + static_cast<uint32_t>(A3_O_COL);
          const uint32_t p = s_tmem + static_cast<uint32_t>(A3_P_COL);
          const int nks = min(A3_BK, A.LN - k * A3_BK) >> 4;
          for (int ks = 0; ks < nks; ++ks)
            a3_umma_ts(d, p + static_cast<uint32_t>(8 * ks), vdesc + static_cast<uint64_t>((ks * 16 * 128) >> 4), idesc,
                       (k > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&bar_pfree[w]);
          if (k == nb - 1) umma_commit(&bar_vfree[c.i & 1]);
        }
        __syncwarp();
      };
      A3Cursor c{0, 0, 0, 0, 0};
      c.advance(w, nt);
      issue_s(c, 0);
      uint32_t ev = 0;                                  // blocks of this slot so far
      while (c.n < G) {
        for (int k = 0; k < nb; ++k, ++ev) {
          // S block (tile, k) is in the warpgroup's registers: the slot's next S block may overwrite it
          mbar_wait(&bar_sfree[w], ev & 1u, A.err_flag, 509);
          if (k + 1 < nb) {
            issue_s(c, k + 1);
          } else {
            A3Cursor cn = c;
            cn.advance(2, nt);
            if (cn.n < G) issue_s(cn, 0);
          }
          // P block (tile, k) is in tensor memory
          mbar_wait(&bar_p[w], ev & 1u, A.err_flag, 504);
          tc_fence_after();
          issue_pv(c, k);
        }
        c.advance(2, nt);
      }
    } else if (warp == 1 && G > 0) {
      // ===================== TMA producer =====================
      const uint64_t pol = A.l2hint ? l2_policy_evict_first() : 0ull;
      auto load3 = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        if (pol) tma_load_3d_hint(dst, m, bar, c0, c1, c2, pol);
        else tma_load_3d(dst, m, bar, c0, c1, c2);
      };
      auto load_q = [&](const A3Cursor& c) {
        int b, h; item_bh(c.i, b, h);
        const int qb = c.qb;
        const int qt = tile_qt(c.i, c.r);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_q[qb], A3_BQ * 128);
          load3(sQ + qb * (A3_BQ * 128), A.map_q, &bar_q[qb], h * A3_D, qt * A3_BQ - tile_off(c.i, qt), b);
        }
        __syncwarp();
      };
      auto load_kv = [&](int i, bool v) {
        int b, h; item_bh(i, b, h);
        const int s = i & 1;
        uint64_t* bar = v ? &bar_v[s] : &bar_k[s];
        uint8_t* dst = (v ? sV : sK) + s * kv_bytes;
        const int col = (v ? 2 : 1) * A.att_dim + h * A3_D;
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kv_bytes));
          load3(dst, A.map_kv, bar, col, 0, b);
          load3(dst + kv_bytes / 2, A.map_kv, bar, col, LK / 2, b);
        }
        __syncwarp();
      };
      A3Cursor c{0, 0, 0, 0, 0};
      load_kv(0, false);
      load_q(c);
      load_kv(0, true);
      A3Cursor c2 = c;
      for (int q = 1; q < A3_NQ; ++q) {
        c2.advance(1, nt);
        if (c2.n < G) load_q(c2);
        if (q == 1 && my_items > 1) { load_kv(1, false); load_kv(1, true); }
      }
      c2.advance(1, nt);                                   // c2 = position n + A3_NQ
      // resources are released in issue order: the Q slot of tile n after its last S block, the K buffer of an item after
      // its last S block, its V buffer after its last P V
      for (; c.n < G; c.advance(1, nt), c2.advance(1, nt)) {
        if (c2.n < G) {
          mbar_wait(&bar_qfree[c.qb], static_cast<uint32_t>(c.qph), A.err_flag, 505);
          load_q(c2);
        }
        if (c.r == nt - 1 && c.i + 2 < my_items) {
          mbar_wait(&bar_kfree[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 506);
          load_kv(c.i + 2, false);
          mbar_wait(&bar_vfree[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 510);
          load_kv(c.i + 2, true);
        }
      }
    }
  } else {
    if (SP == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    // ===================== softmax / output warpgroups =====================
    const int sw = warp - 4;
    const int w = sw / (4 * SP);                           // slot
    const int hf = (sw >> 2) % SP;                         // column half of the block this warp owns (SP == 2)
    const int quad = warp & 3;                             // TMEM lane quadrant of this warp (= its scheduler)
    const int lane_g = quad * 32 + lane;                   // TMEM lane = row of the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(w * A3_SLOT);
    const uint32_t s_addr = lane_addr + static_cast<uint32_t>(hf * CW);
    const uint32_t p_addr = lane_addr + static_cast<uint32_t>(A3_P_COL + hf * (CW / 2));
    const uint32_t o_addr = lane_addr + static_cast<uint32_t>(A3_O_COL + hf * (A3_D / SP));
    const int pair_bar = 1 + w * 4 + quad;                 // named barrier of the two column halves of (slot, quadrant)
    const float LOG2E = 1.4426950408889634f;
#ifdef HD_ATTN_CLOCKS
    const bool clk_on = (blockIdx.x == 0 && warp == 4 && lane == 0);
    long long c_ws = 0, c_wp = 0, c_start = clock64();
#endif
    // O of a finished tile -> fp16 rows of the out-projection operand (TMEM lanes [lo, hi) are real rows); a warp
    // converts A3_D / SP columns
    auto readback = [&](int i, int row0, int lo, int hi, float inv) {
      int b, hd; item_bh(i, b, hd);
      const bool row_live = lane_g >= lo && lane_g < hi;
      uint4* dst = reinterpret_cast<uint4*>(A.o_hi + (static_cast<long long>(b) * L + row0 + lane_g) * A.ld_o + hd * A3_D + hf * (A3_D / SP));
      // one load group for all my columns: every tcgen05.ld / wait::ld round trip occupies the SM's tensor-memory read
      // path for ~70 cycles whatever its size (scripts/tmem_rates.cu)
      uint32_t o[2 / SP][32];
#pragma unroll
      for (int part = 0; part < 2 / SP; ++part) a3_ld32(o_addr + static_cast<uint32_t>(part * 32), o[part]);
      a3_wait_ld();
#pragma unroll
      for (int part = 0; part < 2 / SP; ++part) {
        uint32_t ok[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const __half2 ho = __floats2half2_rn(__uint_as_float(o[part][j]) * inv, __uint_as_float(o[part][j + 1]) * inv);
          ok[j >> 1] = *reinterpret_cast<const uint32_t*>(&ho);
        }
        if (row_live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[part * 4 + q] = make_uint4(ok[4 * q], ok[4 * q + 1], ok[4 * q + 2], ok[4 * q + 3]);
        }
      }
    };
    uint32_t ev = 0;                                       // blocks of this slot so far
    uint32_t tcount = 0;                                   // tiles of this slot so far
    bool pend = false;                                     // a finished tile whose O has not been read back yet
    bool pend_live = false;
    int pend_i = 0, pend_row0 = 0, pend_lo = 0, pend_hi = 0;
    float pend_l = 0.f;
    A3Cursor c{0, 0, 0, 0, 0};
    c.advance(w, nt);
    for (; c.n < G; c.advance(2, nt), ++tcount) {
      const int i = c.i;
      const int qt = tile_qt(i, c.r);
      const int off = tile_off(i, qt);
      const int row0 = qt * A3_BQ - off;                   // sequence row of TMEM lane 0
      const int live_lo = off;                             // live lanes [live_lo, live_hi)
      const int live_hi = min(A3_BQ, L - row0);
      const bool warp_live = (quad * 32 < live_hi) && (quad * 32 + 32 > live_lo);      // warp-uniform
      float m = -INFINITY, l = 0.f;                        // running row maximum (natural units) and row sum (my columns)
      for (int k = 0; k < nb; ++k, ++ev) {
        const int bw_blk = min(A3_BK, A.LN - k * A3_BK);   // S width of the block
        const int bw = max(0, min(CW, bw_blk - hf * CW));  // my columns of it
#ifdef HD_ATTN_CLOCKS
        long long t0 = clk_on ? clock64() : 0;
#endif
        mbar_wait(&bar_s[w], ev & 1u, A.err_flag, 507);
        tc_fence_after();
#ifdef HD_ATTN_CLOCKS
        if (clk_on) c_ws += clock64() - t0;
#endif
        uint32_t s[CW];
        if (warp_live && bw > 0) {
          // ---- the whole row segment in one pass over tensor memory
#pragma unroll
          for (int b = 0; b < CW / 32; ++b) {
            if (b * 32 + 32 <= bw) a3_ld32(s_addr + static_cast<uint32_t>(b * 32), &s[b * 32]);
            else if (b * 32 + 16 <= bw) a3_ld16(s_addr + static_cast<uint32_t>(b * 32), &s[b * 32]);
          }
          a3_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(&bar_sfree[w]);       // the tensor core may overwrite S with the slot's next block
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        bool fix = false;
        float alpha = 1.0f;
        if (warp_live) {
          float m_loc = -INFINITY;
          if (bw > 0) {
            const int nvalid = min(bw, L - k * A3_BK - hf * CW);     // real keys among my columns (the rest is zero padding)
            if (nvalid < bw) {
              const int rem = max(nvalid, 0) & 15;
              switch (max(nvalid, 0) >> 4) {
#define A3_MASK_CASE(C) case C: if ((C) * 16 < CW) { _Pragma("unroll") for (int j = 0; j < 16; ++j) if (j >= rem) s[((C) * 16 + j) % CW] = 0xff800000u; } break;
                A3_MASK_CASE(0) A3_MASK_CASE(1) A3_MASK_CASE(2) A3_MASK_CASE(3)
                A3_MASK_CASE(4) A3_MASK_CASE(5) A3_MASK_CASE(6) A3_MASK_CASE(7)
#undef A3_MASK_CASE
                default: break;
              }
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c16 = 0; c16 < CW / 16; ++c16) {
              if (c16 * 16 < bw) {
#pragma unroll
                for (int j = 0; j < 16; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(s[c16 * 16 + j]));
              }
            }
            m_loc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          }
          if (SP == 2) {
            // block maximum over both column halves (the partner warp of the quadrant holds the other one)
            float* xm = xmax + ((ev & 1u) * 2 + w) * (SP * A3_BQ);
            xm[hf * A3_BQ + lane_g] = m_loc;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            m_loc = fmaxf(m_loc, xm[(hf ^ 1) * A3_BQ + lane_g]);
          }
          if (k == 0) {
            m = m_loc;
          } else {
            // online softmax: keep the running maximum unless a probability would exceed 2^tau
            const bool need = (m_loc - m) * LOG2E > A.tau;
            if (need) { alpha = a3_ex2((m - m_loc) * LOG2E); m = m_loc; l *= alpha; }
            fix = __any_sync(0xffffffffu, need);
          }
          // ---- p = exp2(s*log2e - max*log2e) as packed fp16
          const float mc = -m * LOG2E;
          if (k > 0 || pend) {
            // P columns: the previous P V of this slot has consumed them; it also completes O up to the previous block
#ifdef HD_ATTN_CLOCKS
            long long t1 = clk_on ? clock64() : 0;
#endif
            mbar_wait(&bar_pfree[w], (ev - 1u) & 1u, A.err_flag, 508);
            tc_fence_after();
#ifdef HD_ATTN_CLOCKS
            if (clk_on) c_wp += clock64() - t1;
#endif
          }
          if (fix) {
            // rescale the O accumulator of the earlier blocks in place (my A3_D / SP columns)
#pragma unroll
            for (int part = 0; part < 2 / SP; ++part) {
              uint32_t o[32];
              a3_ld32(o_addr + static_cast<uint32_t>(part * 32), o);
              a3_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
              a3_st32(o_addr + static_cast<uint32_t>(part * 32), o);
            }
          }
#pragma unroll
          for (int c16 = 0; c16 < CW / 16; ++c16) {
            if (c16 * 16 < bw) {
              uint32_t pk[8];
              a3_exp_chunk<NPOLY>(&s[c16 * 16], LOG2E, mc, pk, a0, a1, b0, b1);
              a3_st8(p_addr + static_cast<uint32_t>(c16 * 8), pk);
            }
          }
          l += (a0 + a1) + (b0 + b1);
        } else if (k > 0 || pend) {
          mbar_wait(&bar_pfree[w], (ev - 1u) & 1u, A.err_flag, 508);
          tc_fence_after();
        }
        // the previous tile of the slot: its last P V is complete (bar_pfree above); read its O back before this
        // tile's first P V (issued after the arrival below) overwrites the accumulator
        if (k == 0 && pend) {
          if (pend_live) {
            float lt = pend_l;
            // (SP == 2) the partner's row sum was published before the named barrier of this block; a quadrant that is
            // idle in THIS tile has not passed one, so it synchronises here (both warps of the pair take the same path)
            if (SP == 2 && !warp_live) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            if (SP == 2) lt += xsum[(((tcount - 1u) & 1u) * 2 + w) * (SP * A3_BQ) + (hf ^ 1) * A3_BQ + lane_g];
            readback(pend_i, pend_row0, pend_lo, pend_hi, 1.0f / lt);
          }
          pend = false;
        }
        if (warp_live) a3_wait_st();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(&bar_p[w]);
      }
      pend = true; pend_live = warp_live; pend_i = i; pend_row0 = row0; pend_lo = live_lo; pend_hi = live_hi;
      pend_l = l;
      if (SP == 2 && warp_live) xsum[((tcount & 1u) * 2 + w) * (SP * A3_BQ) + hf * A3_BQ + lane_g] = l;
#ifdef HD_ATTN_CLOCKS
      if (clk_on) { g_a3_clk[0] = (c.n >> 1) + 1; g_a3_clk[1] = clock64() - c_start; g_a3_clk[2] = c_ws; g_a3_clk[3] = c_wp; }
#endif
    }
    if (pend) {
      mbar_wait(&bar_pfree[w], (ev - 1u) & 1u, A.err_flag, 511);
      tc_fence_after();
      if (pend_live) {
        float lt = pend_l;
        if (SP == 2) {
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          lt += xsum[(((tcount - 1u) & 1u) * 2 + w) * (SP * A3_BQ) + (hf ^ 1) * A3_BQ + lane_g];
        }
        readback(pend_i, pend_row0, pend_lo, pend_hi, 1.0f / lt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (A.pdl == 2) pdl_launch_dependents();             // late trigger (HUDIFF_PDL=2)
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t attn3_smem(int LK, int sp) {
  const int nq = sp == 1 ? 4 : 3;
  const size_t xchg = sp == 1 ? 0 : 2 * (2 * 2 * sp * A3_BQ) * 4;
  return static_cast<size_t>(nq * A3_BQ) * 128 + 4 * static_cast<size_t>(LK) * 128 + 26 * 8 + xchg + 64;
}

static int g_a3_sms = 0;
typedef void (*Attn3Kernel)(Attn3Params);
static int g_a3_split = 1;            // HUDIFF_ATTN_SPLIT: softmax warps per (slot, lane quadrant), 1 or 2 (2: measured 7 % slower)
static Attn3Kernel attn3_kernel(int npoly) {
  if (g_a3_split == 1) {
    switch (npoly) {
      case 0: return attention_tc3_kernel<0, 1>;
      case 4: return attention_tc3_kernel<4, 1>;
      case 8: return attention_tc3_kernel<8, 1>;
      default: return attention_tc3_kernel<6, 1>;
    }
  }
  switch (npoly) {
    case 0: return attention_tc3_kernel<0, 2>;
    case 4: return attention_tc3_kernel<4, 2>;
    case 8: return attention_tc3_kernel<8, 2>;
    default: return attention_tc3_kernel<6, 2>;
  }
}
static int g_a3_npoly = 6;            // HUDIFF_ATTN_POLY (0, 4, 6, 8): exponentials per 16 on the FMA pipe (4 and 6 measure the same)
static float g_a3_tau = 8.0f;         // HUDIFF_ATTN_TAU: rescale threshold in log2 units (<= 0 forces the correction path)
cudaError_t attention_tc3_init() {
  if (const char* e = std::getenv("HUDIFF_ATTN_POLY")) g_a3_npoly = std::atoi(e);
  if (const char* e = std::getenv("HUDIFF_ATTN_TAU")) g_a3_tau = static_cast<float>(std::atof(e));
  if (const char* e = std::getenv("HUDIFF_ATTN_SPLIT")) g_a3_split = std::atoi(e) == 2 ? 2 : 1;
  cudaError_t e = cudaFuncSetAttribute(attn3_kernel(g_a3_npoly), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_a3_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t attention_tc3_clocks(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_a3_clk, sizeof(long long) * 16);
}

cudaError_t attention_tc3_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms, int reverse,
                                 int l2hint, int pdl) {
  if (L > 320 || L < 1) return cudaErrorInvalidValue;
  Attn3Params P;
  P.map_q = map_q; P.map_kv = map_kv; P.o_hi = o_hi; P.ld_o = ld_o; P.L = L; P.att_dim = att_dim; P.nhead = nhead;
  P.n_items = B * nhead;
  P.LK = ((L + 63) / 64) * 64;
  P.LN = ((L + 15) / 16) * 16;
  P.nb = (P.LN + A3_BK - 1) / A3_BK;
  P.nt = (L + A3_BQ - 1) / A3_BQ;
  P.tail_rows = L - (P.nt - 1) * A3_BQ;
  P.tail_quads = (P.tail_rows + 31) / 32;
  P.tail_places = P.nt > 1 ? 4 / P.tail_quads : 1;
  if (P.tail_places < 1) P.tail_places = 1;
  P.err_flag = err_flag;
  P.reverse = reverse; P.l2hint = l2hint; P.pdl = pdl;
  if (g_a3_sms <= 0) {
    cudaError_t e = attention_tc3_init();
    if (e != cudaSuccess) return e;
  }
  P.tau = g_a3_tau;
  const size_t smem = attn3_smem(P.LK, g_a3_split);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  int sms = g_a3_sms;
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  const int grid = P.n_items < sms ? P.n_items : sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(A3_CTRL_THREADS + g_a3_split * 256, 1, 1);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn3_kernel(g_a3_npoly), P);
}


// =====================================================================================================================
// Fourth schedule: ONE query tile in flight, its keys split in two halves that two softmax warpgroups exponentiate
// SIMULTANEOUSLY and independently (attention_tc4_kernel).
//
// The tc3 schedule above walks the key blocks of a row one after the other (online softmax), and every block costs a
// fixed round trip of ~1.5 k cycles (S ready -> tcgen05.ld -> max -> exp -> tcgen05.st -> P V -> next block) whatever
// its width; a 291-key row has three of them.  Here warpgroup A owns keys [0, hw0) and warpgroup B keys [hw0, LN) of
// the SAME tile: each has its own S buffer (P written in place over its first half), its own row maximum and row sum and
// its own O accumulator, so neither ever waits for the other's numbers while it exponentiates; the two partial results
//     O = (O_A e^(mA - m) + O_B e^(mB - m)) / (lA e^(mA - m) + lB e^(mB - m)),   m = max(mA, mB)
// are merged when the tile is written out (the two warps of a lane quadrant exchange (m, l) through shared memory on a
// 64-thread named barrier and convert 32 of the 64 output columns each).  One block per warpgroup and tile, one
// tcgen05.ld group per warp for S and one for O, no correction pass.
//
//   warp 0 / warp 2   MMA issuer of half A / B: P V (tile) then S (next tile), in that order on the tensor pipe, so S can
//                     overwrite the buffer P lives in        warp 1   TMA producer (Q ring of 3, K / V double buffer)
//   warps 4-7 / 8-11  softmax + output warpgroup of half A / B (thread = TMEM lane = query row)
// Tensor memory: S_A [0,160)  S_B [160,320)  O_A [320,384)  O_B [384,448).
constexpr int A4_THREADS = 384;
constexpr int A4_HW = 160;              // widest key half (S columns per buffer)
constexpr int A4_NQ = 3;
constexpr int A4_O_COL = 320;

struct Attn4Params {
  const CUtensorMap* map_q; const CUtensorMap* map_kv;
  __half* o_hi; int ld_o;
  int L, att_dim, nhead, n_items, LK;
  int hw0, hw1;                         // S widths of the two key halves (multiples of 16, hw0 >= hw1 >= 16)
  int nt, tail_rows, tail_quads, tail_places;
  int pdl, l2hint, reverse;
  int* err_flag;
};

template <int NPOLY>
__global__ void __launch_bounds__(A4_THREADS, 1)
attention_tc4_kernel(const __grid_constant__ Attn4Params A) {
  typedef A3CursorT<A4_NQ> Cursor;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int L = A.L, LK = A.LK, nt = A.nt;
  const int kv_bytes = LK * 128;
  uint8_t* sQ = smem;                                  // [3][128][64] fp16 ring over the tile sequence
  uint8_t* sK = sQ + A4_NQ * A3_BQ * 128;              // [2 items][LK][64]
  uint8_t* sV = sK + 2 * kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kv_bytes);
  uint64_t* bar_k = bars + 0;                          // [2] keys of item parity landed
  uint64_t* bar_v = bars + 2;                          // [2] values landed
  uint64_t* bar_kfree = bars + 4;                      // [2] every S MMA of the item complete (2 nt commits)
  uint64_t* bar_vfree = bars + 6;                      // [2] every P V MMA of the item complete (2 nt commits)
  uint64_t* bar_s = bars + 8;                          // [2 halves] S complete
  uint64_t* bar_p = bars + 10;                         // [2 halves] P written (one arrival per warp of the warpgroup)
  uint64_t* bar_o = bars + 12;                         // [2 halves] P V complete
  uint64_t* bar_ofree = bars + 14;                     // both O accumulators read back (one arrival per softmax warp)
  uint64_t* bar_q = bars + 16;                         // [3] query tile landed
  uint64_t* bar_qfree = bars + 20;                     // [3] both S MMAs of the tile complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);
  float2* xch = reinterpret_cast<float2*>(bars + 26);  // [2 parities][2 halves][128] (row maximum, row sum)

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_k[b], 1); mbar_init(&bar_v[b], 1); mbar_init(&bar_kfree[b], 2 * nt); mbar_init(&bar_vfree[b], 2 * nt);
      mbar_init(&bar_s[b], 1); mbar_init(&bar_p[b], 4); mbar_init(&bar_o[b], 1);
    }
    mbar_init(bar_ofree, 8);
    for (int b = 0; b < A4_NQ; ++b) { mbar_init(&bar_q[b], 1); mbar_init(&bar_qfree[b], 2); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();
  const int my_items = (A.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int G = my_items * nt;

  auto tile_off = [&](int i, int qt) -> int {          // tail tiles start `off` rows early (see attention_tc3_kernel)
    return (nt > 1 && qt == nt - 1) ? (i % A.tail_places) * A.tail_quads * 32 : 0;
  };
  auto item_bh = [&](int i, int& b, int& h) {
    int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
    if (A.reverse) item = A.n_items - 1 - item;
    b = item / A.nhead; h = item - b * A.nhead;
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if ((warp == 0 || warp == 2) && G > 0) {
      // ===================== MMA issuer of key half warp / 2 =====================
      const int hf = warp >> 1;
      const int hw = hf ? A.hw1 : A.hw0;
      const int k0 = hf ? A.hw0 : 0;                    // first key of the half
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t s_tmem = tmem_base + static_cast<uint32_t>(hf * A4_HW);
      const uint32_t o_tmem = tmem_base + static_cast<uint32_t>(A4_O_COL + hf * A3_D);
      auto issue_s = [&](const Cursor& c) {
        mbar_wait(&bar_q[c.qb], static_cast<uint32_t>(c.qph), A.err_flag, 601);
        mbar_wait(&bar_k[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 602);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qdesc = umma_smem_desc_sw128(q_addr + c.qb * (A3_BQ * 128));
          const uint64_t kdesc = umma_smem_desc_sw128(k_addr + (c.i & 1) * kv_bytes + k0 * 128);
          const uint32_t idesc = umma_idesc_f16(hw);
#pragma unroll
          for (int kk = 0; kk < A3_D / 16; ++kk) umma_f16(s_tmem, qdesc + 2 * kk, kdesc + 2 * kk, idesc, kk > 0 ? 1u : 0u);
          umma_commit(&bar_s[hf]);
          umma_commit(&bar_qfree[c.qb]);
          umma_commit(&bar_kfree[c.i & 1]);
        }
        __syncwarp();
      };
      Cursor c{0, 0, 0, 0, 0};
      issue_s(c);
      for (uint32_t t = 0; c.n < G; ++t) {
        mbar_wait(&bar_p[hf], t & 1u, A.err_flag, 604);                        // P of this half is in tensor memory
        if (t > 0) mbar_wait(bar_ofree, (t - 1u) & 1u, A.err_flag, 605);        // both accumulators of the previous tile were read
        mbar_wait(&bar_v[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 603);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vdesc = umma_smem_desc_sw128(v_addr + (c.i & 1) * kv_bytes + k0 * 128);
          const uint32_t idesc = a3_idesc_bmn(A3_D);
          const int nks = hw >> 4;
          for (int ks = 0; ks < nks; ++ks)
            a3_umma_ts(o_tmem, s_tmem + static_cast<uint32_t>(8 * ks), vdesc + static_cast<uint64_t>((ks * 16 * 128) >> 4), idesc,
                       ks > 0 ? 1u : 0u);
          umma_commit(&bar_o[hf]);
          umma_commit(&bar_vfree[c.i & 1]);
        }
        __syncwarp();
        c.advance(1, nt);
        if (c.n < G) issue_s(c);                       // in order behind the P V above: it may overwrite P
      }
    } else if (warp == 1 && G > 0) {
      // ===================== TMA producer =====================
      const uint64_t pol = A.l2hint ? l2_policy_evict_first() : 0ull;
      auto load3 = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        if (pol) tma_load_3d_hint(dst, m, bar, c0, c1, c2, pol);
        else tma_load_3d(dst, m, bar, c0, c1, c2);
      };
      auto load_q = [&](const Cursor& c) {
        int b, h; item_bh(c.i, b, h);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_q[c.qb], A3_BQ * 128);
          load3(sQ + c.qb * (A3_BQ * 128), A.map_q, &bar_q[c.qb], h * A3_D, c.r * A3_BQ - tile_off(c.i, c.r), b);
        }
        __syncwarp();
      };
      auto load_kv = [&](int i, bool v) {
        int b, h; item_bh(i, b, h);
        const int s = i & 1;
        uint64_t* bar = v ? &bar_v[s] : &bar_k[s];
        uint8_t* dst = (v ? sV : sK) + s * kv_bytes;
        const int col = (v ? 2 : 1) * A.att_dim + h * A3_D;
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kv_bytes));
          load3(dst, A.map_kv, bar, col, 0, b);
          load3(dst + kv_bytes / 2, A.map_kv, bar, col, LK / 2, b);
        }
        __syncwarp();
      };
      Cursor c{0, 0, 0, 0, 0};
      load_kv(0, false);
      load_q(c);
      load_kv(0, true);
      Cursor c2 = c;
      for (int q = 1; q < A4_NQ; ++q) {
        c2.advance(1, nt);
        if (c2.n < G) load_q(c2);
        if (q == 1 && my_items > 1) { load_kv(1, false); load_kv(1, true); }
      }
      c2.advance(1, nt);
      for (; c.n < G; c.advance(1, nt), c2.advance(1, nt)) {
        if (c2.n < G) {
          mbar_wait(&bar_qfree[c.qb], static_cast<uint32_t>(c.qph), A.err_flag, 606);
          load_q(c2);
        }
        if (c.r == nt - 1 && c.i + 2 < my_items) {
          mbar_wait(&bar_kfree[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 607);
          load_kv(c.i + 2, false);
          mbar_wait(&bar_vfree[c.i & 1], static_cast<uint32_t>(c.i >> 1) & 1u, A.err_flag, 608);
          load_kv(c.i + 2, true);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    // ===================== softmax / output warpgroups =====================
    const int hf = (warp - 4) >> 2;                        // key half
    const int quad = warp & 3;
    const int lane_g = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + static_cast<uint32_t>(hf * A4_HW);
    const uint32_t o_addr = lane_addr + static_cast<uint32_t>(A4_O_COL + hf * 32);      // my 32 output columns of O_A; O_B at + 64
    const int pair_bar = 1 + quad;
    const int hw = hf ? A.hw1 : A.hw0;
    const int nvalid = min(hw, L - (hf ? A.hw0 : 0));      // real keys of my half (the rest is zero padding)
    const float LOG2E = 1.4426950408889634f;
    Cursor c{0, 0, 0, 0, 0};
    for (uint32_t t = 0; c.n < G; c.advance(1, nt), ++t) {
      const int i = c.i, qt = c.r;
      const int off = tile_off(i, qt);
      const int row0 = qt * A3_BQ - off;
      const int live_lo = off, live_hi = min(A3_BQ, L - row0);
      const bool warp_live = (quad * 32 < live_hi) && (quad * 32 + 32 > live_lo);
      mbar_wait(&bar_s[hf], t & 1u, A.err_flag, 609);
      tc_fence_after();
      float m = 0.f, l = 1.f;
      if (warp_live) {
        uint32_t s[A4_HW];
#pragma unroll
        for (int b = 0; b < A4_HW / 32; ++b) {
          if (b * 32 + 32 <= hw) a3_ld32(s_addr + static_cast<uint32_t>(b * 32), &s[b * 32]);
          else if (b * 32 + 16 <= hw) a3_ld16(s_addr + static_cast<uint32_t>(b * 32), &s[b * 32]);
        }
        a3_wait_ld();
        if (nvalid < hw) {
          const int rem = nvalid & 15;
          switch (nvalid >> 4) {
#define A4_MASK_CASE(C) case C: { _Pragma("unroll") for (int j = 0; j < 16; ++j) if (j >= rem) s[(C) * 16 + j] = 0xff800000u; } break;
            A4_MASK_CASE(0) A4_MASK_CASE(1) A4_MASK_CASE(2) A4_MASK_CASE(3) A4_MASK_CASE(4)
            A4_MASK_CASE(5) A4_MASK_CASE(6) A4_MASK_CASE(7) A4_MASK_CASE(8) A4_MASK_CASE(9)
#undef A4_MASK_CASE
            default: break;
          }
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c16 = 0; c16 < A4_HW / 16; ++c16) {
          if (c16 * 16 < hw) {
#pragma unroll
            for (int j = 0; j < 16; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(s[c16 * 16 + j]));
          }
        }
        m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const float mc = -m * LOG2E;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int c16 = 0; c16 < A4_HW / 16; ++c16) {
          if (c16 * 16 < hw) {
            uint32_t pk[8];
            a3_exp_chunk<NPOLY>(&s[c16 * 16], LOG2E, mc, pk, a0, a1, b0, b1);
            a3_st8(s_addr + static_cast<uint32_t>(c16 * 8), pk);          // in place: the whole S row is in registers
          }
        }
        l = (a0 + a1) + (b0 + b1);
        xch[((t & 1u) * 2 + hf) * A3_BQ + lane_g] = make_float2(m, l);
        a3_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&bar_p[hf]);
      float fa = 0.f, fb = 0.f;                            // factors of O_A and O_B in the merged, normalised output
      if (warp_live) {
        // the partner warp of this lane quadrant (other key half) has published its (maximum, sum)
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float2 o2 = xch[((t & 1u) * 2 + (hf ^ 1)) * A3_BQ + lane_g];
        const float mm = fmaxf(m, o2.x);
        const float e_own = a3_ex2((m - mm) * LOG2E), e_oth = a3_ex2((o2.x - mm) * LOG2E);
        const float inv = 1.0f / (l * e_own + o2.y * e_oth);
        fa = (hf ? e_oth : e_own) * inv;
        fb = (hf ? e_own : e_oth) * inv;
      }
      mbar_wait(&bar_o[0], t & 1u, A.err_flag, 610);
      mbar_wait(&bar_o[1], t & 1u, A.err_flag, 611);
      tc_fence_after();
      if (warp_live) {
        int b, hd; item_bh(i, b, hd);
        const bool row_live = lane_g >= live_lo && lane_g < live_hi;
        uint32_t oa[32], ob[32];
        a3_ld32(o_addr, oa);
        a3_ld32(o_addr + static_cast<uint32_t>(A3_D), ob);
        a3_wait_ld();
        uint32_t ok[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float v0 = fmaf(__uint_as_float(oa[j]), fa, __uint_as_float(ob[j]) * fb);
          const float v1 = fmaf(__uint_as_float(oa[j + 1]), fa, __uint_as_float(ob[j + 1]) * fb);
          const __half2 ho = __floats2half2_rn(v0, v1);
          ok[j >> 1] = *reinterpret_cast<const uint32_t*>(&ho);
        }
        if (row_live) {
          uint4* dst = reinterpret_cast<uint4*>(A.o_hi + (static_cast<long long>(b) * L + row0 + lane_g) * A.ld_o + hd * A3_D + hf * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(ok[4 * q], ok[4 * q + 1], ok[4 * q + 2], ok[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(bar_ofree);
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

static size_t attn4_smem(int LK) {
  return static_cast<size_t>(A4_NQ * A3_BQ) * 128 + 4 * static_cast<size_t>(LK) * 128 + 26 * 8 + 2 * 2 * A3_BQ * 8 + 64;
}
typedef void (*Attn4Kernel)(Attn4Params);
static Attn4Kernel attn4_kernel(int npoly) {
  switch (npoly) {
    case 0: return attention_tc4_kernel<0>;
    case 4: return attention_tc4_kernel<4>;
    case 8: return attention_tc4_kernel<8>;
    default: return attention_tc4_kernel<6>;
  }
}
cudaError_t attention_tc4_init() {
  cudaError_t e = attention_tc3_init();
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn4_kernel(g_a3_npoly), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t attention_tc4_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms, int reverse,
                                 int l2hint, int pdl) {
  if (L > 2 * A4_HW || L < 17) return cudaErrorInvalidValue;
  Attn4Params P;
  P.map_q = map_q; P.map_kv = map_kv; P.o_hi = o_hi; P.ld_o = ld_o; P.L = L; P.att_dim = att_dim; P.nhead = nhead;
  P.n_items = B * nhead;
  P.LK = ((L + 63) / 64) * 64;
  const int LN = ((L + 15) / 16) * 16;
  P.hw0 = ((LN / 2 + 15) / 16) * 16;
  P.hw1 = LN - P.hw0;
  P.nt = (L + A3_BQ - 1) / A3_BQ;
  P.tail_rows = L - (P.nt - 1) * A3_BQ;
  P.tail_quads = (P.tail_rows + 31) / 32;
  P.tail_places = P.nt > 1 ? 4 / P.tail_quads : 1;
  if (P.tail_places < 1) P.tail_places = 1;
  P.err_flag = err_flag;
  P.reverse = reverse; P.l2hint = l2hint; P.pdl = pdl;
  if (g_a3_sms <= 0) {
    cudaError_t e = attention_tc4_init();
    if (e != cudaSuccess) return e;
  }
  if (P.hw1 < 16 || P.hw0 > A4_HW || attn4_smem(P.LK) > 227 * 1024) return cudaErrorInvalidValue;
  int sms = g_a3_sms;
  if (max_sms > 0 && max_sms < sms) sms = max_sms;
  const int grid = P.n_items < sms ? P.n_items : sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(A4_THREADS, 1, 1);
  cfg.dynamicSmemBytes = attn4_smem(P.LK); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn4_kernel(g_a3_npoly), P);
}

}  // namespace hd
