// Fused RoPE'd multi-head self-attention core: O = softmax(Q K^T) V per (sequence, head).
// (q already carries 1/sqrt(dh) and the rotary phase from the QKV GEMM epilogue.)
// No mask: pad slots are attended exactly like the reference (cross_attention.py:168-170);
// the [L, L] score matrix never leaves registers (flash-style online softmax).
//
// v1 data path: K, V (whole sequence, <= 320 keys) and a 64-query tile are staged in shared
// memory with cp.async in an XOR-swizzled layout; QK^T and PV run on mma.sync m16n8k16 fp16
// tensor-core fragments fed by ldmatrix; fp16x3 mode issues hi*hi + hi*lo + lo*hi.
#include "hd_common.cuh"
#include "hd_kernels.h"

namespace hd {

constexpr int ATT_BQ = 64;      // queries per CTA (4 warps x 16 rows)
constexpr int ATT_D = 64;       // head dim (fixed by the checkpoints: att_model 512 / 8 heads)
constexpr int ATT_THREADS = 128;


__device__ __forceinline__ uint32_t swz(int row, int col) {     // byte offset inside a [rows][64] fp16 tile
  return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7))) << 4) + ((col & 7) << 1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool X3>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __grid_constant__ AttnParams A, int LP) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int L = A.L;
  const int kv_bytes = LP * 128;
  uint8_t* sK = smem;                          // [LP][64] hi
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sQ = sV + kv_bytes;                 // [64][64] hi
  uint8_t* sKl = sQ + ATT_BQ * 128;            // lo planes (X3 only)
  uint8_t* sVl = sKl + kv_bytes;
  uint8_t* sQl = sVl + kv_bytes;

  const long long row0 = static_cast<long long>(b) * L;
  const int qcol = h * ATT_D, kcol = A.att_dim + h * ATT_D, vcol = 2 * A.att_dim + h * ATT_D;

  // ---- stage K, V (all keys) and the Q tile ------------------------------------------------
  for (int idx = tid; idx < LP * 8; idx += ATT_THREADS) {
    const int r = idx >> 3, ch = idx & 7;
    const uint32_t off = swz(r, ch * 8);
    if (r < L) {
      const long long g = (row0 + r) * A.ld_qkv + ch * 8;
      cp_async16(smem_u32(sK + off), A.qkv_hi + g + kcol);
      cp_async16(smem_u32(sV + off), A.qkv_hi + g + vcol);
      if (X3) {
        cp_async16(smem_u32(sKl + off), A.qkv_lo + g + kcol);
        cp_async16(smem_u32(sVl + off), A.qkv_lo + g + vcol);
      }
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sK + off) = z;
      *reinterpret_cast<uint4*>(sV + off) = z;
      if (X3) { *reinterpret_cast<uint4*>(sKl + off) = z; *reinterpret_cast<uint4*>(sVl + off) = z; }
    }
  }
  for (int idx = tid; idx < ATT_BQ * 8; idx += ATT_THREADS) {
    const int r = idx >> 3, ch = idx & 7;
    const uint32_t off = swz(r, ch * 8);
    const int qr = qt * ATT_BQ + r;
    if (qr < L) {
      const long long g = (row0 + qr) * A.ld_qkv + qcol + ch * 8;
      cp_async16(smem_u32(sQ + off), A.qkv_hi + g);
      if (X3) cp_async16(smem_u32(sQl + off), A.qkv_lo + g);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sQ + off) = z;
      if (X3) *reinterpret_cast<uint4*>(sQl + off) = z;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- Q fragments (16 rows x 64 d per warp) -----------------------------------------------
  uint32_t qf[4][4], qfl[4][4];
  {
    const int r = warp * 16 + (lane & 15);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int c = kk * 16 + (lane >> 4) * 8;
      ldsm_x4(smem_u32(sQ + swz(r, c)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      if (X3) ldsm_x4(smem_u32(sQl + swz(r, c)), qfl[kk][0], qfl[kk][1], qfl[kk][2], qfl[kk][3]);
    }
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;    // rows g and g+8
  const float LOG2E = 1.4426950408889634f;
  const int g = lane >> 2, t = lane & 3;

  const int nkb = (L + 63) / 64;
  for (int kb = 0; kb < nkb; ++kb) {
    const int key0 = kb * 64;
    const int valid = min(64, L - key0);              // keys of this block that exist
    const int npairs = (valid + 15) / 16;             // 16-key groups to process (uniform)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    // S = Q K^T
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      if (jp < npairs) {
        const int krow = key0 + jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int c = kk * 16 + ((lane >> 3) & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(sK + swz(krow, c)), b0, b1, b2, b3);
          mma16816(s[2 * jp], qf[kk], b0, b1);
          mma16816(s[2 * jp + 1], qf[kk], b2, b3);
          if (X3) {
            mma16816(s[2 * jp], qfl[kk], b0, b1);
            mma16816(s[2 * jp + 1], qfl[kk], b2, b3);
            uint32_t c0, c1, c2, c3;
            ldsm_x4(smem_u32(sKl + swz(krow, c)), c0, c1, c2, c3);
            mma16816(s[2 * jp], qf[kk], c0, c1);
            mma16816(s[2 * jp + 1], qf[kk], c2, c3);
          }
        }
      }
    }
    // mask keys beyond L, block row max
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kc = j * 8 + 2 * t;
      if (kc >= valid) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (kc + 1 >= valid) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
      bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float sc0 = exp2f((m0 - nm0) * LOG2E), sc1 = exp2f((m1 - nm1) * LOG2E);
    m0 = nm0; m1 = nm1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f((s[j][0] - m0) * LOG2E);
      s[j][1] = exp2f((s[j][1] - m0) * LOG2E);
      s[j][2] = exp2f((s[j][2] - m1) * LOG2E);
      s[j][3] = exp2f((s[j][3] - m1) * LOG2E);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * sc0 + rs0;
    l1 = l1 * sc1 + rs1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1; }
    // O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < npairs) {
        uint32_t pa[4], pl[4];
        {
          __half h0, h1, q0, q1;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = 2 * ks + (u >> 1);
            const float x0 = s[j][(u & 1) * 2], x1 = s[j][(u & 1) * 2 + 1];
            split_f16(x0, h0, q0);
            split_f16(x1, h1, q1);
            pa[u] = pack_h2(h0, h1);
            pl[u] = pack_h2(q0, q1);
          }
        }
        const int vrow = key0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          const int c = jd * 16 + ((lane >> 4) & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(sV + swz(vrow, c)), b0, b1, b2, b3);
          mma16816(o[2 * jd], pa, b0, b1);
          mma16816(o[2 * jd + 1], pa, b2, b3);
          if (X3) {
            mma16816(o[2 * jd], pl, b0, b1);
            mma16816(o[2 * jd + 1], pl, b2, b3);
            uint32_t c0, c1, c2, c3;
            ldsm_x4_t(smem_u32(sVl + swz(vrow, c)), c0, c1, c2, c3);
            mma16816(o[2 * jd], pa, c0, c1);
            mma16816(o[2 * jd + 1], pa, c2, c3);
          }
        }
      }
    }
  }
  // ---- finalize: O / rowsum -> fp16 operand planes -------------------------------------------
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const int qr0 = qt * ATT_BQ + warp * 16 + g, qr1 = qr0 + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = h * ATT_D + j * 8 + 2 * t;
    __half h0, h1, q0, q1;
    if (qr0 < L) {
      split_f16(o[j][0] * inv0, h0, q0);
      split_f16(o[j][1] * inv0, h1, q1);
      const long long off = (row0 + qr0) * A.ld_o + col;
      *reinterpret_cast<uint32_t*>(A.o_hi + off) = pack_h2(h0, h1);
      if (X3) *reinterpret_cast<uint32_t*>(A.o_lo + off) = pack_h2(q0, q1);
    }
    if (qr1 < L) {
      split_f16(o[j][2] * inv1, h0, q0);
      split_f16(o[j][3] * inv1, h1, q1);
      const long long off = (row0 + qr1) * A.ld_o + col;
      *reinterpret_cast<uint32_t*>(A.o_hi + off) = pack_h2(h0, h1);
      if (X3) *reinterpret_cast<uint32_t*>(A.o_lo + off) = pack_h2(q0, q1);
    }
  }
}

cudaError_t attention_init() {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t attention_launch(const AttnParams& A, bool x3, cudaStream_t stream) {
  const int LP = ((A.L + 63) / 64) * 64;
  const size_t smem = static_cast<size_t>(x3 ? 2 : 1) * (2 * LP * 128 + ATT_BQ * 128);
  dim3 grid((A.L + ATT_BQ - 1) / ATT_BQ, A.nhead, A.B);
  if (x3) attention_kernel<true><<<grid, ATT_THREADS, smem, stream>>>(A, LP);
  else attention_kernel<false><<<grid, ATT_THREADS, smem, stream>>>(A, LP);
  return cudaGetLastError();
}

}  // namespace hd
