// hudiff_b200 -- shared device/host definitions for the sm_100a kernels.
//
// Data layout:
//   activations           fp32 row-major [B*L, C]   (residual streams)
//   tensor-core operands  fp16 row-major [B*L, Kp]  ("hi" plane, plus a "lo" plane in fp16x3 mode)
//   weights               fp16 [N, taps*Kp] K-major, pre-scaled by a power of two, hi (+lo) planes
// A "row space" maps 128-row MMA tiles onto chains of sequences so that the dilated
// convolutions see zero padding at chain boundaries (TMA out-of-bounds fill) and the ragged
// 139-slot light chain needs no padding in memory.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hd {

// ----------------------------------------------------------------------------------------
// problem descriptors (plain structs, passed to kernels as __grid_constant__ parameters)
// ----------------------------------------------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum Epi {
  EPI_LN_ACT_F16 = 0,   // full-row CTA: bias -> LayerNorm(N) -> act -> fp16 operand
  EPI_RESID_F32 = 1,    // bias (+residual) (*dropout) -> fp32 stream (+ optional fp16 copy)
  EPI_ACT_F16 = 2,      // bias -> act -> fp16 operand
  EPI_QKV_ROPE = 3,     // bias -> RoPE on q,k (q pre-scaled by 1/sqrt(dh)) -> fp16 operand
};
enum DropMode { DROP_OFF = 0, DROP_BITS = 1, DROP_PHILOX = 2 };

struct RowSpace {
  int nseq;         // sequences
  int Lc;           // valid rows per sequence in this chain
  int Lcp;          // tile-space period (multiple of 8, >= Lc)
  int seq_stride;   // rows between consecutive sequences in memory
  int row_off;      // first row of the chain inside a sequence
};

constexpr int LN_SUM_PARTS = 8;   // partial row sums per row: <= 4 N tiles x 2 epilogue warps per lane quadrant

struct EpiParams {
  int epi, act;
  float wscale;                 // 1 / (2^k weight pre-scale)
  const float* bias;            // [N]
  const float* gamma;           // [N]  (EPI_LN_ACT_F16)
  const float* beta;            // [N]
  float* out32; int ld32;       // fp32 output (EPI_RESID_F32)
  __half* out16_hi; __half* out16_lo; int ld16;   // fp16 operand output (lo may be null)
  const float* resid; int ldr;  // residual input (may alias out32)
  // dropout applied after the residual add (reference: F.dropout(x + f(x)))
  int drop_mode; float drop_scale; uint32_t drop_thresh16;  // keep if u16 < thresh (philox)
  int drop_onebit;                                          // p == 0.5 -> 1 random bit / element
  const uint32_t* drop_bits;    // DROP_BITS: keep-bit planes, word index = (seq*Lc + l)*N/32 + c/32
  long long drop_step_words;    // words per sampling step (all calls)
  uint32_t call_id;             // philox stream id of this dropout call
  // RoPE (EPI_QKV_ROPE)
  const float2* rope;           // [seq_len][dh/2] (cos, sin)
  int seq_len, att_dim, head_dim; float qscale;
  // Folded LayerNorm (EPI_ACT_F16 / EPI_QKV_ROPE): the A operand is the RAW row x, the weights carry
  // gamma, and  y = rstd * (acc - mean * colsum) * wscale + bias'   (bias' = b + W beta)
  const float2* ln_stats;       // [rows] (mean, rstd) of the operand rows (row_stats_kernel), or null
  const float* colsum;          // [N] sum over k of the packed (scaled, fp16-rounded) weights W'[n, k]
  int ln_sums_k;                // > 0: ln_stats holds LN_SUM_PARTS partial (sum x, sum x^2) per row over ln_sums_k columns in all,
                                //      written by the PRODUCING residual epilogue (row_sums below) instead of a statistics pass
  float* row_sums;              // EPI_RESID_F32: [rows][LN_SUM_PARTS][2].  Every (N tile, epilogue warp of the lane quadrant)
                                //      that finishes columns of a row owns ONE slot and stores the (sum, sum of squares) of its
                                //      columns there: no atomics, so the statistics -- and with them a whole sampling run -- are
                                //      bit-reproducible; unused slots stay zero (allocation).  The CUDA-core reference GEMMs
                                //      (validation) add into slot 0 instead, which the host zeroes per step for them.

  int sec0;                     // q|k|v section of output column 0 (0: weights start at q, 1: k|v-only GEMM)
  const int* row_pos;           // optional [rows]: sequence position of each (gathered) row; null -> row % seq_len
  // Fused LayerNorm tail (EPI_RESID_F32): once every N tile of a 128-row M tile is stored, the same epilogue
  // warps re-read the finished fp32 rows (L2-resident, written a moment ago by this SM) and emit the operand
  // the NEXT layer consumes, act(LayerNorm(row)) as fp16 planes -- the former stand-alone LN-cast pass.
  const float* ln_gamma;        // [N]; null -> no tail
  const float* ln_beta;         // [N]
  int ln_act;
  __half* ln_out_hi; __half* ln_out_lo; int ln_ld;
};

struct GemmProblem {
  RowSpace rs;
  const CUtensorMap* a_maps_hi;   // 16 maps, box heights 8,16,..,128 rows
  const CUtensorMap* a_maps_lo;   // fp16x3 only
  const CUtensorMap* b_map_hi;    // weight map whose box holds BN / cluster_size rows
  const CUtensorMap* b_map_lo;
  int m_tiles;                    // ceil(nseq*Lcp / 128)
  int N, BN;                      // output width, width per CTA (BN == N for full-row epilogues)
  int kchunks;                    // Kp / 64 per tap
  int k_logical;                  // un-padded input channels (FLOP accounting)
  int taps, dil;                  // taps == 1 for plain linear layers
  int b_box_n;                    // rows per weight TMA box
  // raw pointers for the SIMT reference path
  const __half* a_hi; const __half* a_lo; int lda;
  const __half* w_hi; const __half* w_lo; int ldw;
  EpiParams ep;
};

struct GemmParams {
  GemmProblem p[2];
  int nprob;
  int cs;                         // CTAs per cluster (weight tile multicast width): 1, 2 or 4
  int n_tiles;                    // N tiles per M block (same for every problem of a launch)
  int mblocks[2];                 // M blocks (cs tiles each) per problem
  int total_tiles;                // (mblocks[0] + mblocks[1]) * n_tiles, per cluster work items
  int x3;                         // 1: three-pass split-fp16 accumulation
  const int* step_ptr;            // device step counter (dropout stream position)
  const unsigned long long* rng_ptr;   // device: [0] philox key, [1] global index of this batch's first sequence (both
                                  //         live in device memory so that the captured step graph does not depend on them)
  int pdl;                        // 1: launched with programmatic stream serialization (see pdl_wait)
  int l2hint;                     // 1: activation operand tiles and residual rows are fetched with an L2 evict-first policy
  int reverse;                    // 1: walk the tiles from the last row block to the first.  Consecutive kernels of a step alternate
                                  //    direction, so each one starts on the rows its producer wrote LAST -- the part of the
                                  //    producer's output that is still in the 126 MB L2 (HUDIFF_L2FLIP)
  int ln_tail;                    // 1: M-major tile order (all N tiles of an M block on the same cluster, back to back), required by
                                  //    the fused LayerNorm tail
  int* err_flag;                  // set by watchdogs
  int nsplit;                     // 1: full-row (LayerNorm) tiles are split over the CTA pair by columns: both CTAs work on the
                                  //    same 128 rows, each on N/2 columns with a double-buffered accumulator; the row
                                  //    statistics are exchanged through distributed shared memory
  int two_cta;                    // 1: cta_group::2 UMMA (M = 256 per CTA pair, each CTA holds half of the weight tile)
  int max_sms;                    // > 0: launch at most this many CTAs (two half-batches share the SMs on two streams)
  int dbg;                        // timing experiments, HD_DEBUG_HOOKS builds only (HUDIFF_DBG): 1 = no epilogue work, 2 = no MMA, 4 = no TMA
};

// ----------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one lane of a converged warp (the compiler keeps the enclosing control flow uniform)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- programmatic dependent launch ----------------------------------------------------------
// Kernels of a step are launched with cudaLaunchAttributeProgrammaticStreamSerialization (HUDIFF_PDL): a CTA of the
// NEXT kernel may become resident on an SM as soon as this kernel's CTA there has exited, run its prologue (barrier
// init, TMEM allocation, tensor-map prefetch) under the tail of this kernel, and then blocks in pdl_wait() until the
// whole predecessor grid has completed and its memory is visible.  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (and be reported) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
#ifdef HD_SPIN_SLEEP
    __nanosleep(HD_SPIN_SLEEP);
#endif
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

// ---- TMA -------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- L2 eviction hints ------------------------------------------------------------------------------
// Streams that are read exactly once (activation operand tiles, residual rows, q|k|v) are fetched with an
// evict-first policy so that they do not push the lines the NEXT kernel will want -- this kernel's own output,
// written with the default policy -- out of the 126 MB L2.  policy == 0 selects the plain instructions.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_global_f4_hint(const void* ptr, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(policy) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// weight-tile slice multicast to every CTA of the cluster in `mask` (same smem offset / barrier offset)
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// ---- 2-CTA (cta_group::2) variants: the CTA pair of a cluster shares one UMMA ---------------------
// The mbarrier operand always designates the LEADER CTA's barrier (peer bit cleared), so the
// transaction bytes of both CTAs' loads complete on the barrier the MMA issuer waits on.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                     uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]^T, issued by the leader CTA
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_f16_m256(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// Remote arrive with the default (CTA-scope release) semantics: enough when the arrival only orders
// tcgen05 accesses (tcgen05.fence::before_thread_sync precedes it) and no generic-proxy data is
// published to the other CTA.  The .release.cluster form above costs a cluster-scope memory barrier.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// cluster-scope acquire wait (pairs with mbar_arrive_cluster from the peer CTA)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
// 8-byte store into the shared memory of CTA `rank` of the cluster (same offset as `p` in this CTA)
__device__ __forceinline__ void st_shared_cluster_f2(const void* p, uint32_t rank, float a, float b) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// ---- clusters ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// K-major operand tile, rows of 64 fp16 (128 B), 128-byte swizzle, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset (unused, K fits one atom)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset [32,46)
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: A,B = F16 (K-major), D = F32, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}
// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- numerics ----------------------------------------------------------------------------
// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): 2 MUFU + ~12 FMA-pipe instructions instead
// of libdevice's ~30-instruction erff; the exact-GELU epilogues are instruction-bound on it.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-ax * ax);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  if (act == ACT_GELU) return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));   // nn.GELU() (erf form)
  return x;
}
// fp16 split: x ~= hi + lo, hi = rn(x), lo = rn(x - hi).  Saturating so that an out-of-range
// activation degrades instead of producing inf.
__device__ __forceinline__ __half f2h_sat(float x) {
  x = fminf(fmaxf(x, -65504.0f), 65504.0f);
  return __float2half_rn(x);
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = f2h_sat(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// Philox4x32-10 (Salmon et al. 2011): counter-based RNG for in-kernel dropout / sampling.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// (mean, rstd) of operand row g for the folded-LayerNorm epilogues
__device__ __forceinline__ float2 ln_row_stats(const EpiParams& E, long long g) {
  if (E.ln_sums_k <= 0) return __ldg(E.ln_stats + g);
  float2 s = make_float2(0.f, 0.f);
  {                                                   // partial (sum, sum of squares), added in slot order
    const float4* p4 = reinterpret_cast<const float4*>(E.ln_stats + g * LN_SUM_PARTS);
#pragma unroll
    for (int q = 0; q < LN_SUM_PARTS / 2; ++q) {
      const float4 v = __ldg(p4 + q);
      s.x += v.x; s.y += v.y;
      s.x += v.z; s.y += v.w;
    }
  }
  {                                                   // -> (mean, rstd), biased variance, eps 1e-5
    const float inv = 1.0f / static_cast<float>(E.ln_sums_k);
    const float mean = s.x * inv;
    s.y = 1.0f / sqrtf(fmaxf(s.y * inv - mean * mean, 0.f) + 1e-5f);
    s.x = mean;
  }
  return s;
}

// ---- row-space helpers ---------------------------------------------------------------------
struct RowRef {
  int seq, l;      // sequence index, row inside the chain
  bool valid;
  long long g;     // global row index in memory
};
__device__ __forceinline__ RowRef row_lookup(const RowSpace& rs, int tile_row) {
  RowRef r;
  r.seq = tile_row / rs.Lcp;
  r.l = tile_row - r.seq * rs.Lcp;
  r.valid = (r.seq < rs.nseq) && (r.l < rs.Lc);
  r.g = static_cast<long long>(r.seq) * rs.seq_stride + rs.row_off + r.l;
  return r;
}

// ----------------------------------------------------------------------------------------
// Dropout keep-bits of 32 consecutive columns of one row (bit j = column 32*word + j).
//   DROP_BITS   : replayed planes, word index = didx * (N/32) + word     (didx = seq*Lc + l)
//   DROP_PHILOX : counter-based; p == 0.5 uses one random bit per element, otherwise a 16-bit
//                 uniform per element compared with thresh = round((1-p) * 65536)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dropout_keep_word(const EpiParams& E, unsigned long long seed, int step,
                                                      int didx, int words_per_row, int word) {
  if (E.drop_mode == DROP_BITS) {
    const long long w = static_cast<long long>(didx) * words_per_row + word;
    return __ldg(E.drop_bits + static_cast<long long>(step) * E.drop_step_words + w);
  }
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  if (E.drop_onebit) {
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(didx), E.call_id, static_cast<uint32_t>(step),
                                          0x5EED0000u + static_cast<uint32_t>(word >> 2)), key);
    const int q = word & 3;
    return q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w;
  }
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(didx), E.call_id, static_cast<uint32_t>(step),
                                          0xD0000000u + static_cast<uint32_t>(word * 4 + q)), key);
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bits |= ((w4[u] & 0xFFFFu) < E.drop_thresh16 ? 1u : 0u) << (8 * q + 2 * u);
      bits |= ((w4[u] >> 16) < E.drop_thresh16 ? 1u : 0u) << (8 * q + 2 * u + 1);
    }
  }
  return bits;
}

// Two consecutive keep-words (64 columns, word0 even): one Philox call in the 1-bit mode.
__device__ __forceinline__ void dropout_keep_words2(const EpiParams& E, unsigned long long seed, int step, int didx,
                                                    int words_per_row, int word0, uint32_t& k0, uint32_t& k1) {
  if (E.drop_mode == DROP_PHILOX && E.drop_onebit) {
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(didx), E.call_id, static_cast<uint32_t>(step),
                                          0x5EED0000u + static_cast<uint32_t>(word0 >> 2)), key);
    if (word0 & 2) { k0 = r.z; k1 = r.w; } else { k0 = r.x; k1 = r.y; }
    return;
  }
  k0 = dropout_keep_word(E, seed, step, didx, words_per_row, word0);
  k1 = dropout_keep_word(E, seed, step, didx, words_per_row, word0 + 1);
}

// ----------------------------------------------------------------------------------------
// Reference epilogue (SIMT validation path, gemm_ref.cu): one thread owns one output row and
// writes it directly.  The tcgen05 kernel implements the same arithmetic with a warp-local
// transpose through shared memory so that all global accesses are coalesced (gemm_tc.cu).
// Loader::load32(col0, v) yields 32 consecutive accumulator columns of the row.
// ----------------------------------------------------------------------------------------
template <class Loader>
__device__ __forceinline__ void run_epilogue(const GemmProblem& P, const GemmParams& G, int tile_row, int n0,
                                             Loader& ld) {
  const EpiParams& E = P.ep;
  const RowRef rr = row_lookup(P.rs, tile_row);
  const int nchunks = P.BN / 32;
  float v[32];

  float mean = 0.f, rstd = 1.f;
  if (E.epi == EPI_LN_ACT_F16) {
    // shifted single-pass moments over the full row (BN == N)
    float shift = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      ld.load32(c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = fmaf(v[j], E.wscale, __ldg(E.bias + c * 32 + j));
        if (c == 0 && j == 0) shift = x;
        float d = x - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
    }
    const float inv_n = 1.0f / static_cast<float>(P.N);
    const float md = s1 * inv_n;
    mean = shift + md;
    rstd = 1.0f / sqrtf(fmaxf(s2 * inv_n - md * md, 0.f) + 1e-5f);
  }

  const int step = G.step_ptr ? *G.step_ptr : 0;
  const unsigned long long seed = G.rng_ptr ? G.rng_ptr[0] : 0ull;
  const int seq_off = (G.rng_ptr && E.drop_mode == DROP_PHILOX) ? static_cast<int>(G.rng_ptr[1]) : 0;
  const int didx = (rr.seq + seq_off) * P.rs.Lc + rr.l;
  for (int c = 0; c < nchunks; ++c) {
    const int col0 = n0 + c * 32;
    ld.load32(c * 32, v);
    if (E.ln_stats) {
      const float2 ms = rr.valid ? ln_row_stats(E, rr.g) : make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = fmaf(ms.y * E.wscale, fmaf(-ms.x, __ldg(E.colsum + col0 + j), v[j]), __ldg(E.bias + col0 + j));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], E.wscale, __ldg(E.bias + col0 + j));
    }
    if (!rr.valid) continue;

    if (E.epi == EPI_LN_ACT_F16) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = act_apply((v[j] - mean) * rstd * __ldg(E.gamma + col0 + j) + __ldg(E.beta + col0 + j), E.act);
    } else if (E.epi == EPI_RESID_F32) {
      if (E.resid) {
        const float4* rp = reinterpret_cast<const float4*>(E.resid + rr.g * E.ldr + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 r4 = rp[q];
          v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
        }
      }
      if (E.drop_mode != DROP_OFF) {
        const uint32_t bits = dropout_keep_word(E, seed, step, didx, P.N / 32, col0 >> 5);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] * E.drop_scale : 0.f;
      }
      float4* op = reinterpret_cast<float4*>(E.out32 + rr.g * E.ld32 + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (E.row_sums) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
        atomicAdd(E.row_sums + 2 * LN_SUM_PARTS * rr.g, s1);          // slot 0 (zeroed per step for this validation path)
        atomicAdd(E.row_sums + 2 * LN_SUM_PARTS * rr.g + 1, s2);
      }

    } else if (E.epi == EPI_ACT_F16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], E.act);
    } else if (E.epi == EPI_QKV_ROPE) {
      const int sec = col0 / E.att_dim + E.sec0;   // 0 = q, 1 = k, 2 = v (32 | att_dim)
      if (sec < 2) {
        const int pos = E.row_pos ? __ldg(E.row_pos + rr.g) : static_cast<int>(rr.g % E.seq_len);
        const int j0 = (col0 % E.head_dim) >> 1;   // first complex pair of this chunk
        const float2* rp = E.rope + static_cast<long long>(pos) * (E.head_dim / 2) + j0;
        const float qs = (sec == 0) ? E.qscale : 1.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float2 cs = __ldg(rp + j);
          float xr = v[2 * j], xi = v[2 * j + 1];
          v[2 * j] = (xr * cs.x - xi * cs.y) * qs;
          v[2 * j + 1] = (xr * cs.y + xi * cs.x) * qs;
        }
      }
    }

    if (E.out16_hi) {
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        __half h0, l0, h1, l1;
        split_f16(v[j], h0, l0);
        split_f16(v[j + 1], h1, l1);
        ph[j >> 1] = pack_h2(h0, h1);
        pl[j >> 1] = pack_h2(l0, l1);
      }
      uint4* dh = reinterpret_cast<uint4*>(E.out16_hi + rr.g * E.ld16 + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dh[q] = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
      if (E.out16_lo) {
        uint4* dl = reinterpret_cast<uint4*>(E.out16_lo + rr.g * E.ld16 + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dl[q] = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
      }
    }
  }
}

#endif  // __CUDACC__
}  // namespace hd
