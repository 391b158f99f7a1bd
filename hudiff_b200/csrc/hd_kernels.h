// Launcher declarations and parameter blocks of the non-GEMM kernels (see elementwise.cu,
// attention.cu, gemm_tc.cu, gemm_ref.cu).
#pragma once
#include "hd_common.cuh"

namespace hd {

struct LnCastParams {
  const float* x; int ldx; int C; long long rows;
  const float* gamma[2]; const float* beta[2];   // set 1 is used for rows with (r % seq_len) >= split
  int seq_len, split;                            // seq_len == 0 -> always set 0
  int act;
  __half* out_hi; __half* out_lo; int ldo;
  int reverse;                                   // 1: blocks walk the rows last-to-first (L2 reuse, see GemmParams::reverse)
  int l2hint;                                    // 1: the fp32 rows are read with an L2 evict-first policy (read once)
  int pdl;                                       // 1: programmatic dependent launch (hd_common.cuh pdl_wait)
};

// Optional LayerNorm -> activation -> fp16 operand of the rows a row-wise kernel has just produced (the first
// ByteNet block of each stack consumes act(LN(x)) of the embedding / assembled features): saves the stand-alone
// LN-cast pass and its re-read of the stream.  gamma[1] / beta[1] serve rows with (r % seq_len) >= split (light chain).
struct RowLnOut {
  const float* gamma[2]; const float* beta[2];
  int seq_len, split, act;
  __half* out_hi; __half* out_lo; int ldo;       // out_hi == null -> no LayerNorm output
};

struct AssembleParams {
  const float* e;            // [rows, d] token-branch features
  const int* region;         // [rows]
  const int* chain;          // [2B] (Ab) or null
  const float* pos_table;    // [n_region, L, d]  folded region+position branch
  const float* side_table;   // [n_side, d] folded chain-type branch, or null (Nb)
  float* x; int ldx;         // [rows, 3d] (Ab) / [rows, 2d] (Nb)
  long long rows; int L, B, d, split, n_region, n_side;
  int reverse;
  RowLnOut ln;
};

struct HeadParams {
  int mode;                  // 0: logits for every row, 1: sample one slot per sequence
  const float* x; int ldx; int C; long long rows;
  const float* gamma; const float* beta; const float* w; const float* bias; int n_out;
  float* logits;             // mode 0: [rows, n_out]
  // mode 1
  int B, L, T, n_sample;
  int compact;               // mode 1: x holds one (gathered) row per sequence instead of [B*L] rows
  const int* order;          // [B, T]
  const int* step_ptr;
  int* tokens;               // [B, L], updated in place
  float* sel_logits;         // optional [T, B, n_out]
  const float* exp_draws;    // replayed Exp(1) draws [T, B, n_sample] or null -> philox
  const unsigned long long* rng_ptr;   // device: [0] philox key, [1] global index of the first sequence of this batch
  int* err_flag;
};

// Last-block shortcut of the sampling step: only the slot sampled this step is needed downstream of
// the second attention's keys/values, so its row is gathered per sequence into compact [B, .] buffers.
struct TailGatherParams {
  const int* order; const int* step_ptr; int B, L, T;
  int* row_pos;                                   // [B] out: slot sampled this step
  const __half* op_hi; const __half* op_lo; int ld_op;     // LN1(a) operand rows [B*L, D]
  __half* g_hi; __half* g_lo; int ld_g;                    // compact copy [B, D]
  const float* a32; const float* x32; int ld32;            // residual streams [B*L, D]
  float* ga32; float* gx32;                                // compact copies [B, D]
  const float2* stats; float2* gstats; int stat_slots;     // optional row statistics partials [B*L][slots] -> compact [B][slots]
  int D;
};
// One query row per (sequence, head) against all keys of the sequence (fp32 math on the fp16 planes).
struct AttnRowParams {
  const __half* q_hi; const __half* q_lo; int ld_q;        // compact [B, att_dim]; q carries 1/sqrt(dh) and RoPE
  const __half* kv_hi; const __half* kv_lo; int ld_kv;     // [B*L, 3*att_dim] q|k|v planes (k, v columns are read)
  __half* o_hi; __half* o_lo; int ld_o;                    // compact [B, att_dim]
  int B, L, att_dim, nhead;
};
// (mean, rstd) of every row of an fp16 operand (hi + lo planes), LayerNorm arithmetic (eps 1e-5, biased variance)
cudaError_t row_stats_launch(const __half* hi, const __half* lo, int ld, int C, long long rows, float2* out, cudaStream_t s);
// colsum[n] = sum over k of (hi + lo)[n, k] for packed weights [N, ldw]
cudaError_t weight_colsum_launch(const __half* hi, const __half* lo, int N, int ldw, float* out, cudaStream_t s);
cudaError_t tail_gather_launch(const TailGatherParams& P, cudaStream_t s);
cudaError_t attn_row_launch(const AttnRowParams& P, cudaStream_t s);

struct AttnParams {
  const __half* qkv_hi; const __half* qkv_lo; int ld_qkv;   // [B*L, 3*att_dim], q|k|v column blocks
  __half* o_hi; __half* o_lo; int ld_o;                      // [B*L, att_dim]
  int L, att_dim, nhead, B;
};
cudaError_t attention_launch(const AttnParams& A, bool x3, cudaStream_t stream);
cudaError_t attention_init();
cudaError_t attention_tc2_init();
cudaError_t attention_tc2_clocks(long long* out16);
cudaError_t attention_tc2_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms = 0,
                                 int reverse = 0, int l2hint = 0, int pdl = 0);
cudaError_t attention_tc3_init();
cudaError_t attention_tc3_clocks(long long* out16);
cudaError_t attention_tc3_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms = 0,
                                 int reverse = 0, int l2hint = 0, int pdl = 0);
cudaError_t attention_tc4_init();
cudaError_t attention_tc4_launch(const CUtensorMap* map_q, const CUtensorMap* map_kv, __half* o_hi, int ld_o, int L,
                                 int att_dim, int nhead, int B, int* err_flag, cudaStream_t stream, int max_sms = 0,
                                 int reverse = 0, int l2hint = 0, int pdl = 0);
cudaError_t attention_x3_init();
cudaError_t attention_x3_launch(const CUtensorMap* maps, __half* o_hi, __half* o_lo, int ld_o, int L, int att_dim, int nhead,
                                int B, int* err_flag, cudaStream_t stream, int max_sms = 0, int reverse = 0, int pdl = 0);
cudaError_t gemm_tc_init();
cudaError_t gemm_tc_clocks(unsigned long long* out32, int reset);

cudaError_t kernels_init();   // opt-in shared-memory sizes (must run outside stream capture)
cudaError_t gemm_tc_launch(const GemmParams& G, cudaStream_t stream);
cudaError_t gemm_ref_launch(const GemmParams& G, float* scratch, size_t scratch_floats, cudaStream_t stream);

cudaError_t embed_launch(const int* tok, const float* table, float* x, long long rows, int C, int n_tokens,
                         const RowLnOut& ln, cudaStream_t s);
cudaError_t ln_cast_launch(const LnCastParams& P, cudaStream_t s);
cudaError_t assemble_launch(const AssembleParams& P, cudaStream_t s);
cudaError_t head_launch(const HeadParams& P, cudaStream_t s);
cudaError_t advance_step_launch(int* step, cudaStream_t s);
// rng[0] = seed, rng[1] = seq_offset (device state the step graph reads, so neither is baked into the graph)
cudaError_t set_rng_launch(unsigned long long* rng, unsigned long long seed, unsigned long long seq_offset, cudaStream_t s);
// Range checks of the integer inputs (the reference raises IndexError from its embedding lookups / indexing):
// first violation -> err_flag = 401 (order), 402 (tokens), 403 (region), 404 (chain).  Null arrays are skipped.
cudaError_t validate_launch(const int* tokens, const int* region, const int* chain, const int* order, long long M, int B,
                            int T, int L, int n_tokens, int n_region, int n_side, int* err_flag, cudaStream_t s);
cudaError_t pack_weight_launch(const float* w, int N, int K, int taps, int Kp, float scale, __half* hi, __half* lo,
                               cudaStream_t s);
cudaError_t f32_linear_launch(const float* x, const float* w, const float* b, float* y, int rows, int K, int N, int act,
                              cudaStream_t s);
cudaError_t f32_layernorm_launch(const float* x, const float* g, const float* b, float* y, int rows, int C, int act,
                                 cudaStream_t s);
cudaError_t f32_outer_add_launch(const float* xa, const float* xb, float* y, int na, int nb, int C, cudaStream_t s);
cudaError_t f32_add_launch(const float* a, const float* b, float* y, int n, cudaStream_t s);

}  // namespace hd
