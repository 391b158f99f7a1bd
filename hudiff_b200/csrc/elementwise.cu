// Row-wise kernels around the GEMMs: embedding gather, LayerNorm -> activation -> fp16 operand
// cast, feature assembly, the logits head, the per-step categorical resample, weight packing
// and the small exact-fp32 kernels used once at load time for constant folding.
// All of them are HBM/L2-bound: one warp per row, 128-bit accesses, warp-shuffle reductions.
#include "hd_common.cuh"
#include "hd_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace hd {

constexpr int ROWS_PER_BLOCK = 8;      // 256 threads, one warp per row

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Row of C floats (C % 128 == 0, C <= 1024) held as float4 x[i] = elements (i*32 + lane)*4 .. +3.
struct RowRegs {
  float4 x[8];
  int nv;
  __device__ __forceinline__ void load(const float* p, int C, int lane) {
    nv = C >> 7;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) x[i] = reinterpret_cast<const float4*>(p)[i * 32 + lane];
  }
  // two-pass LayerNorm statistics (eps 1e-5, biased variance) like torch.nn.LayerNorm
  __device__ __forceinline__ void stats(int C, float& mean, float& rstd) const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(C) + 1e-5f);
  }
  __device__ __forceinline__ void normalize(const float* gamma, const float* beta, float mean, float rstd, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
        const float4 b = reinterpret_cast<const float4*>(beta)[i * 32 + lane];
        x[i].x = (x[i].x - mean) * rstd * g.x + b.x;
        x[i].y = (x[i].y - mean) * rstd * g.y + b.y;
        x[i].z = (x[i].z - mean) * rstd * g.z + b.z;
        x[i].w = (x[i].w - mean) * rstd * g.w + b.w;
      }
  }
};

// LayerNorm (two-pass statistics, eps 1e-5) -> act -> fp16 planes of one row held by a warp as NV float4 per lane
// (x[i] = elements (i*32 + lane)*4 .. +3): the arithmetic of ln_cast_kernel.
template <int NV>
__device__ __forceinline__ void warp_row_ln_store(const float4* x, const RowLnOut& ln, long long r, int lane) {
  constexpr int C = NV * 128;
  const int set = (ln.seq_len > 0 && static_cast<int>(r % ln.seq_len) >= ln.split) ? 1 : 0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(C) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(ln.gamma[set]) + lane;
  const float4* b4 = reinterpret_cast<const float4*>(ln.beta[set]) + lane;
  __half* oh = ln.out_hi + r * ln.ldo + lane * 4;
  __half* ol = ln.out_lo ? ln.out_lo + r * ln.ldo + lane * 4 : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + i * 32), b = __ldg(b4 + i * 32);
    __half h[4], l[4];
    split_f16(act_apply((x[i].x - mean) * rstd * g.x + b.x, ln.act), h[0], l[0]);
    split_f16(act_apply((x[i].y - mean) * rstd * g.y + b.y, ln.act), h[1], l[1]);
    split_f16(act_apply((x[i].z - mean) * rstd * g.z + b.z, ln.act), h[2], l[2]);
    split_f16(act_apply((x[i].w - mean) * rstd * g.w + b.w, ln.act), h[3], l[3]);
    *reinterpret_cast<uint2*>(oh + i * 128) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    if (ol) *reinterpret_cast<uint2*>(ol + i * 128) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
  }
}

// ---- embedding gather: x[r, :] = table[tok[r], :]  (+ optional LN -> act -> fp16 operand of the first block) -----
template <int NV>
__global__ void embed_kernel(const int* __restrict__ tok, const float* __restrict__ table, float* __restrict__ x,
                             long long rows, int n_tokens, const __grid_constant__ RowLnOut ln) {
  constexpr int C = NV * 128;
  const long long r = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  int t = tok[r];
  t = min(max(t, 0), n_tokens - 1);
  const float4* src = reinterpret_cast<const float4*>(table + static_cast<long long>(t) * C) + lane;
  float4* dst = reinterpret_cast<float4*>(x + r * C) + lane;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i] = src[i * 32]; dst[i * 32] = v[i]; }
  if (ln.out_hi) warp_row_ln_store<NV>(v, ln, r, lane);
}

// ---- LayerNorm -> act -> fp16 operand planes -------------------------------------------------
// LPR lanes per row (32 / LPR rows per warp), NV = C / (4 LPR) float4 per lane: every load of the row
// is in flight before the first use, statistics are two-pass on registers (torch.nn.LayerNorm
// arithmetic) reduced over the LPR lanes of the row.  Narrow rows (C = 256) use 8 lanes per row so
// that four rows share one instruction stream and the short rows do not serialise on the shuffles.
// HBM-bound: 4 B in + 2 B (fp16) / 4 B (fp16x3) out per element.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NV, int LPR>
__global__ void __launch_bounds__(1024) ln_cast_kernel(const __grid_constant__ LnCastParams P) {
  constexpr int RPW = 32 / LPR;                  // rows per warp
  constexpr int C = NV * 4 * LPR;
  if (P.pdl != 2) pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int li = lane % LPR;                     // lane inside the row group
  const unsigned blk = P.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  long long r = (static_cast<long long>(blk) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = r < P.rows;
  if (!live) r = P.rows - 1;                     // keep the warp converged for the shuffles
  const float4* src = reinterpret_cast<const float4*>(P.x + r * P.ldx) + li;
  float4 x[NV];
  if (P.l2hint) {
    const uint64_t pol = l2_policy_evict_first();
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = ld_global_f4_hint(src + i * LPR, pol);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = src[i * LPR];
  }
  const int set = (P.seq_len > 0 && static_cast<int>(r % P.seq_len) >= P.split) ? 1 : 0;
  const float* gamma = P.gamma[set];
  if (P.gamma[0]) {                              // (both sets exist or none)
    const float4* g4 = reinterpret_cast<const float4*>(gamma) + li;
    const float4* b4 = reinterpret_cast<const float4*>(P.beta[set]) + li;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = group_sum<LPR>(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(group_sum<LPR>(q) / static_cast<float>(C) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(g4 + i * LPR), b = __ldg(b4 + i * LPR);
      x[i].x = (x[i].x - mean) * rstd * g.x + b.x;
      x[i].y = (x[i].y - mean) * rstd * g.y + b.y;
      x[i].z = (x[i].z - mean) * rstd * g.z + b.z;
      x[i].w = (x[i].w - mean) * rstd * g.w + b.w;
    }
  }
  if (!live) return;
  __half* oh = P.out_hi + r * P.ldo + li * 4;
  __half* ol = P.out_lo ? P.out_lo + r * P.ldo + li * 4 : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    __half h[4], l[4];
    split_f16(act_apply(x[i].x, P.act), h[0], l[0]);
    split_f16(act_apply(x[i].y, P.act), h[1], l[1]);
    split_f16(act_apply(x[i].z, P.act), h[2], l[2]);
    split_f16(act_apply(x[i].w, P.act), h[3], l[3]);
    *reinterpret_cast<uint2*>(oh + i * (4 * LPR)) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
    if (ol) *reinterpret_cast<uint2*>(ol + i * (4 * LPR)) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
  }
}

// ---- row statistics of an fp16 operand (folded LayerNorm): one warp per 4 rows, two-pass on registers ----
constexpr int STAT_ROWS = 1;      // rows per warp (4 measured slower than 1 on B200: fewer warps in flight)
template <int NV>      // C = NV * 256 halfs: lane holds NV x 8 consecutive elements of each row
__global__ void __launch_bounds__(256) row_stats_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int ld,
                                                        long long rows, float2* __restrict__ out) {
  const long long r0 = (static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5)) * STAT_ROWS;
  if (r0 >= rows) return;
  const int lane = threadIdx.x & 31;
  uint4 u[STAT_ROWS][NV], ul[STAT_ROWS][NV];
#pragma unroll
  for (int k = 0; k < STAT_ROWS; ++k) {
    const long long r = min(r0 + k, rows - 1);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      u[k][i] = reinterpret_cast<const uint4*>(hi + r * ld)[i * 32 + lane];
      ul[k][i] = lo ? reinterpret_cast<const uint4*>(lo + r * ld)[i * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
#pragma unroll
  for (int k = 0; k < STAT_ROWS; ++k) {
    float x[NV * 8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&u[k][i]);
      const __half2* l2 = reinterpret_cast<const __half2*>(&ul[k][i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]), g = __half22float2(l2[e]);
        x[i * 8 + 2 * e] = f.x + g.x; x[i * 8 + 2 * e + 1] = f.y + g.y;
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) s += x[i];
    const float mean = warp_sum(s) / static_cast<float>(NV * 256);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(NV * 256) + 1e-5f);
    if (lane == 0 && r0 + k < rows) out[r0 + k] = make_float2(mean, rstd);
  }
}

__global__ void weight_colsum_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int N, int ldw,
                                     float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < ldw; k += 32) {
    s += __half2float(hi[static_cast<long long>(n) * ldw + k]);
    if (lo) s += __half2float(lo[static_cast<long long>(n) * ldw + k]);
  }
  s = warp_sum(s);
  if (lane == 0) out[n] = s;
}

// ---- feature assembly: X = cat(E + P + C, P, C)  (Ab)   /   cat(E + P, P)  (Nb) -----------------
// (+ optional LN -> act -> fp16 operand of the first block of the wide stack)
template <int ND, bool SIDE>      // ND = d / 128 float4 per lane and section; SIDE: chain-type section present (Ab)
__global__ void assemble_kernel(const __grid_constant__ AssembleParams P) {
  constexpr int NSEC = SIDE ? 3 : 2;
  const unsigned blk = P.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const long long r = static_cast<long long>(blk) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= P.rows) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(r / P.L), l = static_cast<int>(r - static_cast<long long>(b) * P.L);
  int reg = P.region[r];
  reg = min(max(reg, 0), P.n_region - 1);
  const float4* pp = reinterpret_cast<const float4*>(P.pos_table + (static_cast<long long>(reg) * P.L + l) * P.d) + lane;
  const float4* cp = nullptr;
  if (SIDE) {
    int c = (l < P.split) ? P.chain[b] : P.chain[P.B + b];
    c = min(max(c, 0), P.n_side - 1);
    cp = reinterpret_cast<const float4*>(P.side_table + static_cast<long long>(c) * P.d) + lane;
  }
  const float4* ep = reinterpret_cast<const float4*>(P.e + r * P.d) + lane;
  float4* xp = reinterpret_cast<float4*>(P.x + r * P.ldx) + lane;
  float4 v[NSEC * ND];              // row layout: section s, float4 i -> v[s * ND + i] = elements s*d + (i*32 + lane)*4 ..
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    float4 e = ep[i * 32];
    const float4 p = pp[i * 32];
    e.x += p.x; e.y += p.y; e.z += p.z; e.w += p.w;         // (E + P) + C, reference order model.py:357
    if (SIDE) {
      const float4 c = cp[i * 32];
      e.x += c.x; e.y += c.y; e.z += c.z; e.w += c.w;
      v[2 * ND + i] = c;
    }
    v[i] = e;
    v[ND + i] = p;
  }
#pragma unroll
  for (int k = 0; k < NSEC * ND; ++k) xp[k * 32] = v[k];   // section s starts at column s*d = float4 index s*ND*32
  if (P.ln.out_hi) warp_row_ln_store<NSEC * ND>(v, P.ln, r, lane);
}

// ---- logits head: LayerNorm + Linear(C -> n_out) -------------------------------------------
// mode 0: every row -> logits[r, :]            (hudiff_forward)
// mode 1: one row per sequence (the slot sampled this step) -> categorical resample (hudiff_sample)
__global__ void head_kernel(const __grid_constant__ HeadParams P) {
  const long long w = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int step = P.step_ptr ? *P.step_ptr : 0;
  long long r;
  int pos = 0;
  if (P.mode == 0) {
    if (w >= P.rows) return;
    r = w;
  } else {
    if (w >= P.B) return;
    pos = P.order[w * P.T + step];
    if (pos < 0 || pos >= P.L) {                       // never write outside the token matrix (validate_kernel reports it)
      if (lane == 0 && P.err_flag) atomicCAS(P.err_flag, 0, 401);
      return;
    }
    r = w * P.L + pos;
  }
  RowRegs row;
  row.load(P.x + ((P.mode == 1 && P.compact) ? w : r) * P.ldx, P.C, lane);
  float mean, rstd;
  row.stats(P.C, mean, rstd);
  row.normalize(P.gamma, P.beta, mean, rstd, lane);
  float mine = 0.f;                                   // lane o keeps logit o (n_out <= 32)
#pragma unroll 4
  for (int o = 0; o < P.n_out; ++o) {
    const float* wr = P.w + static_cast<long long>(o) * P.C;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < row.nv) {
        const float4 wv = reinterpret_cast<const float4*>(wr)[i * 32 + lane];
        acc = fmaf(row.x[i].x, wv.x, acc);
        acc = fmaf(row.x[i].y, wv.y, acc);
        acc = fmaf(row.x[i].z, wv.z, acc);
        acc = fmaf(row.x[i].w, wv.w, acc);
      }
    acc = warp_sum(acc) + P.bias[o];
    if (lane == o) mine = acc;
  }
  if (P.mode == 0) {
    if (lane < P.n_out) P.logits[r * P.n_out + lane] = mine;
    return;
  }
  if (P.sel_logits && lane < P.n_out)
    P.sel_logits[(static_cast<long long>(step) * P.B + w) * P.n_out + lane] = mine;
  // softmax over the first n_sample classes, then torch.multinomial(., 1) == argmax(p / q), q ~ Exp(1)
  const bool act = lane < P.n_sample;
  const float mx = warp_max(act ? mine : -INFINITY);
  const float e = act ? expf(mine - mx) : 0.f;
  const float inv = 1.0f / warp_sum(e);
  const float p = e * inv;
  float q = 1.f;
  if (act) {
    if (P.exp_draws) {
      q = P.exp_draws[(static_cast<long long>(step) * P.B + w) * P.n_sample + lane];
    } else {
      const unsigned long long seed = P.rng_ptr ? P.rng_ptr[0] : 0ull;
      const uint32_t seq_offset = P.rng_ptr ? static_cast<uint32_t>(P.rng_ptr[1]) : 0u;
      const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      const uint4 rn = philox4x32(make_uint4(static_cast<uint32_t>(w) + seq_offset, static_cast<uint32_t>(lane),
                                             static_cast<uint32_t>(step), 0xE4Bu), key);
      const float u = (static_cast<float>(rn.x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
      q = -logf(u);
    }
  }
  float best = act ? __fdiv_rn(p, q) : -INFINITY;
  int idx = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane == 0) P.tokens[r] = idx;
}

__global__ void advance_step_kernel(int* step) { *step += 1; }
__global__ void set_rng_kernel(unsigned long long* rng, unsigned long long seed, unsigned long long seq_offset) {
  rng[0] = seed;
  rng[1] = seq_offset;
}
__global__ void validate_kernel(const int* __restrict__ tokens, const int* __restrict__ region, const int* __restrict__ chain,
                                const int* __restrict__ order, long long M, int B, int T, int L, int n_tokens, int n_region,
                                int n_side, int* err_flag) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  int bad = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < M; i += stride) {
    if (tokens) { const int t = tokens[i]; if (t < 0 || t >= n_tokens) bad = bad ? bad : 402; }
    if (region) { const int r = region[i]; if (r < 0 || r >= n_region) bad = bad ? bad : 403; }
  }
  if (chain)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 2LL * B; i += stride) {
      const int c = chain[i];
      if (c < 0 || c >= n_side) bad = bad ? bad : 404;
    }
  if (order)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<long long>(B) * T; i += stride) {
      const int o = order[i];
      if (o < 0 || o >= L) bad = 401;
    }
  if (bad) atomicCAS(err_flag, 0, bad);
}

// ---- last-block shortcut: gather the sampled slot's rows into compact buffers -----------------------
__global__ void tail_gather_kernel(const __grid_constant__ TailGatherParams P) {
  const int b = blockIdx.x;
  const int step = P.step_ptr ? *P.step_ptr : 0;
  int pos = P.order[static_cast<long long>(b) * P.T + step];
  pos = min(max(pos, 0), P.L - 1);                     // out-of-range orders are reported by validate_kernel / head_kernel
  const long long r = static_cast<long long>(b) * P.L + pos;
  if (threadIdx.x == 0) P.row_pos[b] = pos;
  if (P.stats && static_cast<int>(threadIdx.x) < P.stat_slots)
    P.gstats[static_cast<long long>(b) * P.stat_slots + threadIdx.x] = P.stats[r * P.stat_slots + threadIdx.x];
  for (int i = threadIdx.x; i < P.D / 8; i += blockDim.x) {          // 8 fp16 = 16 bytes
    reinterpret_cast<uint4*>(P.g_hi + static_cast<long long>(b) * P.ld_g)[i] =
        reinterpret_cast<const uint4*>(P.op_hi + r * P.ld_op)[i];
    if (P.g_lo) reinterpret_cast<uint4*>(P.g_lo + static_cast<long long>(b) * P.ld_g)[i] =
        reinterpret_cast<const uint4*>(P.op_lo + r * P.ld_op)[i];
  }
  for (int i = threadIdx.x; i < P.D / 4; i += blockDim.x) {
    reinterpret_cast<float4*>(P.ga32 + static_cast<long long>(b) * P.D)[i] = reinterpret_cast<const float4*>(P.a32 + r * P.ld32)[i];
    reinterpret_cast<float4*>(P.gx32 + static_cast<long long>(b) * P.D)[i] = reinterpret_cast<const float4*>(P.x32 + r * P.ld32)[i];
  }
}

// ---- last-block shortcut: softmax(q K^T) V for ONE query row per (sequence, head) ---------------------
// grid = B, one warp per head (head dim 64).  Keys are split over lanes for the scores, the output
// dims over lanes for P V.  fp32 arithmetic on hi (+ lo) fp16 planes; no mask (cross_attention.py:168-170).
constexpr int ATTN_ROW_MAXL = 320;
__global__ void __launch_bounds__(256) attn_row_kernel(const __grid_constant__ AttnRowParams P) {
  __shared__ float s_p[8][ATTN_ROW_MAXL];
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= P.nhead) return;
  const int A = P.att_dim, L = P.L;
  float q[64];
  {
    const __half* qh = P.q_hi + static_cast<long long>(b) * P.ld_q + h * 64;
    const __half* ql = P.q_lo ? P.q_lo + static_cast<long long>(b) * P.ld_q + h * 64 : nullptr;
#pragma unroll
    for (int d = 0; d < 64; ++d) q[d] = __half2float(qh[d]) + (ql ? __half2float(ql[d]) : 0.f);
  }
  const long long row0 = static_cast<long long>(b) * L;
  float mx = -INFINITY;
#pragma unroll 2
  for (int j = lane; j < L; j += 32) {
    const uint4* kh = reinterpret_cast<const uint4*>(P.kv_hi + (row0 + j) * P.ld_kv + A + h * 64);
    const uint4* kl = P.kv_lo ? reinterpret_cast<const uint4*>(P.kv_lo + (row0 + j) * P.ld_kv + A + h * 64) : nullptr;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = kh[c];
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      uint4 ul = make_uint4(0u, 0u, 0u, 0u);
      if (kl) ul = kl[c];
      const __half2* l2 = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]), g = __half22float2(l2[e]);
        acc = fmaf(q[c * 8 + 2 * e], f.x + g.x, acc);
        acc = fmaf(q[c * 8 + 2 * e + 1], f.y + g.y, acc);
      }
    }
    s_p[h][j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float p = expf(s_p[h][j] - mx);
    s_p[h][j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // output dims 2*lane, 2*lane+1
  float o0 = 0.f, o1 = 0.f;
  const __half* vbase_h = P.kv_hi + row0 * P.ld_kv + 2 * A + h * 64 + 2 * lane;
  const __half* vbase_l = P.kv_lo ? P.kv_lo + row0 * P.ld_kv + 2 * A + h * 64 + 2 * lane : nullptr;
  for (int j0 = 0; j0 < L; j0 += 8) {                      // 8 independent loads in flight per lane
    __half2 vh[8], vl[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = min(j0 + u, L - 1);
      vh[u] = *reinterpret_cast<const __half2*>(vbase_h + static_cast<long long>(j) * P.ld_kv);
      vl[u] = vbase_l ? *reinterpret_cast<const __half2*>(vbase_l + static_cast<long long>(j) * P.ld_kv) : __floats2half2_rn(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float p = (j0 + u < L) ? s_p[h][j0 + u] : 0.f;
      const float2 f = __half22float2(vh[u]), g = __half22float2(vl[u]);
      o0 = fmaf(p, f.x + g.x, o0);
      o1 = fmaf(p, f.y + g.y, o1);
    }
  }
  const float inv = 1.0f / sum;
  __half h0, l0, h1, l1;
  split_f16(o0 * inv, h0, l0);
  split_f16(o1 * inv, h1, l1);
  const long long oo = static_cast<long long>(b) * P.ld_o + h * 64 + 2 * lane;
  *reinterpret_cast<uint32_t*>(P.o_hi + oo) = pack_h2(h0, h1);
  if (P.o_lo) *reinterpret_cast<uint32_t*>(P.o_lo + oo) = pack_h2(l0, l1);
}

// ---- weight packing: fp32 [N, K] or conv [N, K, taps] -> fp16 planes [N, taps*Kp], scaled ----------
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int K, int taps, int Kp, float scale,
                                   __half* __restrict__ hi, __half* __restrict__ lo) {
  const long long total = static_cast<long long>(N) * taps * Kp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kp);
    const int tap = static_cast<int>((i / Kp) % taps);
    const int n = static_cast<int>(i / (static_cast<long long>(Kp) * taps));
    float v = 0.f;
    if (k < K) v = w[(static_cast<long long>(n) * K + k) * taps + tap] * scale;
    __half h, l;
    split_f16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ---- exact fp32 helpers for load-time constant folding (tiny shapes) ------------------------------
__global__ void f32_linear_kernel(const float* x, const float* w, const float* b, float* y, int rows, int K, int N,
                                  int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  const int r = i / N, n = i - r * N;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(x[r * K + k], w[n * K + k], acc);
  y[i] = act_apply(acc + (b ? b[n] : 0.f), act);
}
__global__ void f32_layernorm_kernel(const float* x, const float* g, const float* b, float* y, int rows, int C,
                                     int act) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += x[r * C + c];
  const float mean = s / C;
  float q = 0.f;
  for (int c = 0; c < C; ++c) { const float d = x[r * C + c] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(q / C + 1e-5f);
  for (int c = 0; c < C; ++c) y[r * C + c] = act_apply((x[r * C + c] - mean) * rstd * g[c] + b[c], act);
}
// y[(a*nb + b), :] = xa[a, :] + xb[b, :]
__global__ void f32_outer_add_kernel(const float* xa, const float* xb, float* y, int na, int nb, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * nb * C) return;
  const int c = i % C, b = (i / C) % nb, a = i / (C * nb);
  y[i] = xa[a * C + c] + xb[b * C + c];
}
__global__ void f32_add_kernel(const float* a, const float* b, float* y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// ---- launchers ----------------------------------------------------------------------------------
static inline int row_blocks(long long rows) { return static_cast<int>((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK); }

cudaError_t embed_launch(const int* tok, const float* table, float* x, long long rows, int C, int n_tokens,
                         const RowLnOut& ln, cudaStream_t s) {
  switch (C >> 7) {
    case 1: embed_kernel<1><<<row_blocks(rows), 256, 0, s>>>(tok, table, x, rows, n_tokens, ln); break;
    case 2: embed_kernel<2><<<row_blocks(rows), 256, 0, s>>>(tok, table, x, rows, n_tokens, ln); break;
    case 3: embed_kernel<3><<<row_blocks(rows), 256, 0, s>>>(tok, table, x, rows, n_tokens, ln); break;
    case 4: embed_kernel<4><<<row_blocks(rows), 256, 0, s>>>(tok, table, x, rows, n_tokens, ln); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
template <int NV, int LPR>
static cudaError_t ln_cast_go(const LnCastParams& P, int bs, cudaStream_t s) {
  const long long rows_per_block = static_cast<long long>(bs >> 5) * (32 / LPR);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((P.rows + rows_per_block - 1) / rows_per_block), 1, 1);
  cfg.blockDim = dim3(bs, 1, 1); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = P.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, ln_cast_kernel<NV, LPR>, P);
}
cudaError_t ln_cast_launch(const LnCastParams& P, cudaStream_t s) {
  static int bs = 0;                              // 64-thread blocks measured best on B200 (HUDIFF_LNCAST_BLOCK overrides)
  if (bs == 0) { const char* e = std::getenv("HUDIFF_LNCAST_BLOCK"); bs = e ? std::atoi(e) : 64; if (bs < 32 || bs > 1024 || (bs & 31)) bs = 64; }
  if ((P.gamma[0] == nullptr) != (P.gamma[1] == nullptr) && P.seq_len > 0) return cudaErrorInvalidValue;
  switch (P.C >> 7) {
    case 1: return ln_cast_go<1, 32>(P, bs, s);
    case 2: return ln_cast_go<8, 8>(P, bs, s);      // C = 256: 8 lanes per row, 4 rows per warp
    case 3: return ln_cast_go<3, 32>(P, bs, s);
    case 4: return ln_cast_go<8, 16>(P, bs, s);     // C = 512: 16 lanes per row, 2 rows per warp
    case 5: return ln_cast_go<5, 32>(P, bs, s);
    case 6: return ln_cast_go<6, 32>(P, bs, s);
    case 7: return ln_cast_go<7, 32>(P, bs, s);
    case 8: return ln_cast_go<8, 32>(P, bs, s);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t assemble_launch(const AssembleParams& P, cudaStream_t s) {
  const bool side = P.side_table != nullptr;
  if ((P.d & 127) || P.d > 512 || (side ? 3 : 2) * P.d != P.ldx) return cudaErrorInvalidValue;
  const int nb = row_blocks(P.rows);
  switch ((P.d >> 7) * 2 + (side ? 1 : 0)) {
    case 2: assemble_kernel<1, false><<<nb, 256, 0, s>>>(P); break;
    case 3: assemble_kernel<1, true><<<nb, 256, 0, s>>>(P); break;
    case 4: assemble_kernel<2, false><<<nb, 256, 0, s>>>(P); break;
    case 5: assemble_kernel<2, true><<<nb, 256, 0, s>>>(P); break;
    case 6: assemble_kernel<3, false><<<nb, 256, 0, s>>>(P); break;
    case 7: assemble_kernel<3, true><<<nb, 256, 0, s>>>(P); break;
    case 8: assemble_kernel<4, false><<<nb, 256, 0, s>>>(P); break;
    default: return cudaErrorInvalidValue;          // Ab with d = 512 would need sum_d_model = 1536 > 1024
  }
  return cudaGetLastError();
}
cudaError_t head_launch(const HeadParams& P, cudaStream_t s) {
  const long long n = (P.mode == 0) ? P.rows : P.B;
  head_kernel<<<row_blocks(n), 256, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t row_stats_launch(const __half* hi, const __half* lo, int ld, int C, long long rows, float2* out, cudaStream_t s) {
  const int nb = row_blocks((rows + STAT_ROWS - 1) / STAT_ROWS);
  if (C % 256 != 0) return cudaErrorInvalidValue;
  switch (C >> 8) {
    case 1: row_stats_kernel<1><<<nb, 256, 0, s>>>(hi, lo, ld, rows, out); break;
    case 2: row_stats_kernel<2><<<nb, 256, 0, s>>>(hi, lo, ld, rows, out); break;
    case 3: row_stats_kernel<3><<<nb, 256, 0, s>>>(hi, lo, ld, rows, out); break;
    case 4: row_stats_kernel<4><<<nb, 256, 0, s>>>(hi, lo, ld, rows, out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t weight_colsum_launch(const __half* hi, const __half* lo, int N, int ldw, float* out, cudaStream_t s) {
  weight_colsum_kernel<<<(N + 7) / 8, 256, 0, s>>>(hi, lo, N, ldw, out);
  return cudaGetLastError();
}
cudaError_t tail_gather_launch(const TailGatherParams& P, cudaStream_t s) {
  tail_gather_kernel<<<P.B, 128, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t attn_row_launch(const AttnRowParams& P, cudaStream_t s) {
  if (P.L > ATTN_ROW_MAXL || P.nhead > 8 || P.att_dim != P.nhead * 64) return cudaErrorInvalidValue;
  attn_row_kernel<<<P.B, 256, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t advance_step_launch(int* step, cudaStream_t s) {
  advance_step_kernel<<<1, 1, 0, s>>>(step);
  return cudaGetLastError();
}
cudaError_t set_rng_launch(unsigned long long* rng, unsigned long long seed, unsigned long long seq_offset, cudaStream_t s) {
  set_rng_kernel<<<1, 1, 0, s>>>(rng, seed, seq_offset);
  return cudaGetLastError();
}
cudaError_t validate_launch(const int* tokens, const int* region, const int* chain, const int* order, long long M, int B,
                            int T, int L, int n_tokens, int n_region, int n_side, int* err_flag, cudaStream_t s) {
  const int blocks = static_cast<int>(std::min<long long>((M + 255) / 256, 592));
  validate_kernel<<<blocks, 256, 0, s>>>(tokens, region, chain, order, M, B, T, L, n_tokens, n_region, n_side, err_flag);
  return cudaGetLastError();
}
cudaError_t pack_weight_launch(const float* w, int N, int K, int taps, int Kp, float scale, __half* hi, __half* lo,
                               cudaStream_t s) {
  pack_weight_kernel<<<296, 256, 0, s>>>(w, N, K, taps, Kp, scale, hi, lo);
  return cudaGetLastError();
}
cudaError_t f32_linear_launch(const float* x, const float* w, const float* b, float* y, int rows, int K, int N, int act,
                              cudaStream_t s) {
  f32_linear_kernel<<<(rows * N + 255) / 256, 256, 0, s>>>(x, w, b, y, rows, K, N, act);
  return cudaGetLastError();
}
cudaError_t f32_layernorm_launch(const float* x, const float* g, const float* b, float* y, int rows, int C, int act,
                                 cudaStream_t s) {
  f32_layernorm_kernel<<<(rows + 63) / 64, 64, 0, s>>>(x, g, b, y, rows, C, act);
  return cudaGetLastError();
}
cudaError_t f32_outer_add_launch(const float* xa, const float* xb, float* y, int na, int nb, int C, cudaStream_t s) {
  f32_outer_add_kernel<<<(na * nb * C + 255) / 256, 256, 0, s>>>(xa, xb, y, na, nb, C);
  return cudaGetLastError();
}
cudaError_t f32_add_launch(const float* a, const float* b, float* y, int n, cudaStream_t s) {
  f32_add_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, y, n);
  return cudaGetLastError();
}

}  // namespace hd
