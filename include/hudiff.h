/* hudiff.h -- C-ABI of libhudiff_b200.so: the HuDiff denoiser forward and sampling loop on B200.
 *
 * The reference (TencentAI4S/HuDiff) has no native/plugin interface; its de-facto operator
 * boundary for this path is Python (citations relative to the reference tree):
 *   model construction + weights   utils/train_utils.py:43-55 (model_selected),
 *                                  antibody_scripts/sample.py:456-458, nanosample.py:255-257,269-271
 *   one denoiser forward           model/encoder/model.py:366-384 (AntiTFNet.forward),
 *                                  model/nanoencoder/model.py:325-343 (NanoAntiTFNet.forward)
 *   the sampling loop              antibody_scripts/sample.py:499-513, nanobody_scripts/nanosample.py:316-329
 * Each entry point below names the reference call it replaces.  A reference maintainer binds
 * them with ctypes (INTEGRATION.md shows the stub); `hudiff_b200/capi.py` is that binding.
 *
 * Conventions: every function returns 0 on success or a negative hudiff_status; the text of
 * the last error on the calling thread is `hudiff_last_error()`.  A handle belongs to one CUDA
 * device and is not thread-safe.  The caller owns every pointer it passes; the library owns its
 * packed weights, workspaces and CUDA graphs.  Device-pointer entry points are asynchronous on
 * the given stream; `_host` entry points copy, run and synchronise.
 */
#ifndef HUDIFF_H_
#define HUDIFF_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define HUDIFF_API __attribute__((visibility("default")))
#else
#define HUDIFF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hudiff_ctx* hudiff_handle;
typedef void* hudiff_stream;              /* cudaStream_t */

enum hudiff_status {
  HUDIFF_OK = 0,
  HUDIFF_ERR_INVALID = -1,                /* bad argument / unsupported configuration */
  HUDIFF_ERR_CUDA = -2,                   /* a CUDA runtime or driver call failed */
  HUDIFF_ERR_MISSING_WEIGHT = -3,         /* a state-dict tensor the model needs was not supplied */
  HUDIFF_ERR_STATE = -4,                  /* call order (e.g. forward before load_weights) */
  HUDIFF_ERR_INDEX = -5                   /* a token / region / chain id or a sampling slot is out of range (the reference
                                             raises IndexError from its embedding lookups and `tokens[:, i]`) */
};

enum hudiff_kind { HUDIFF_AB = 0, HUDIFF_NB = 1 };          /* AntiTFNet / NanoAntiTFNet */
enum hudiff_precision { HUDIFF_FP16 = 0, HUDIFF_FP16X3 = 1 };/* tensor-core operand policy */
enum hudiff_act { HUDIFF_ACT_RELU = 1, HUDIFF_ACT_GELU = 2 };

/* The `config.model` section the reference stores inside its checkpoints
 * (configs/antibody_train.yml:3-24, configs/heavy_train.yml:3-21). */
typedef struct hudiff_cfg {
  int32_t kind;                 /* hudiff_kind */
  int32_t n_tokens;             /* 23 */
  int32_t d_model;              /* 256 (== d_embedding == r_model == n_pos_model [== s_model]) */
  int32_t n_encoder_layers;     /* 6 */
  int32_t aa_kernel_size;       /* 7 */
  int32_t r;                    /* 128 -> dilations 2^(n % (log2(r)+1)) */
  int32_t n_side, s_embedding;  /* Ab only: 3, 4 */
  int32_t n_region, r_embedding;/* 7, 4 */
  int32_t max_len;              /* 291 (Ab = 152 + 139) or 152 (Nb) */
  int32_t sum_d_model;          /* 768 (Ab) / 512 (Nb) */
  int32_t dual_layers;          /* 6 */
  int32_t att_model, dim_feedforward, nhead, cs_layers;   /* 512, 256, 8, 5 */
  int32_t activation;           /* hudiff_act of the aa_encoder ('gelu' in the released configs) */
  int32_t precision;            /* hudiff_precision */
  double dropout;               /* config.model.dropout; > 0 enables the reference's live eval dropout */
} hudiff_cfg;

/* One state-dict entry, named with the reference's key (e.g. "self_at.layers.0.attn_hl.rope"). */
typedef struct hudiff_tensor {
  const char* name;
  const void* data;             /* contiguous, host or device memory */
  int32_t dtype;                /* 0 = float32, 1 = complex64 (interleaved re, im) */
  int32_t ndim;
  int64_t shape[4];
  int32_t on_device;            /* 0 = host pointer, 1 = device pointer */
} hudiff_tensor;

/* Random streams of one call.  mode 0: counter-based Philox inside the kernels (throughput mode,
 * statistically equivalent).  mode 1: replay of host-generated streams -- the keep-bits of every
 * functional dropout call (model/encoder/model.py:176-178,295-303) and the Exp(1) draws behind
 * torch.multinomial (sample.py:512) -- which makes the loop token-for-token equal to the
 * reference under the same torch seed. */
typedef struct hudiff_rng {
  int32_t mode;                 /* 0 = philox, 1 = replay */
  int32_t dropout;              /* 0 = off, 1 = live (ignored when cfg.dropout == 0) */
  uint64_t seed;                /* philox key */
  const uint32_t* dropout_bits; /* replay: [T][hudiff_dropout_words(B)] keep-bit planes, reference call order */
  const float* exp_draws;       /* replay: [T][B][22] */
  uint64_t seq_offset;          /* philox: global index of this batch's first sequence; sequence b draws the stream of
                                   global sequence seq_offset + b, so a batch sharded over GPUs (or calls) with one seed
                                   reproduces the un-sharded run and shards never repeat each other's draws */
} hudiff_rng;

/* replaces: model_selected(config)  (utils/train_utils.py:43-55) */
HUDIFF_API int hudiff_create(hudiff_handle* out, const hudiff_cfg* cfg, int device);

/* replaces: model.load_state_dict(ckpt['model'])  (sample.py:457).  Tensors are looked up by the
 * reference's key names; the library repacks them (fp16 planes, folded constant branches). */
HUDIFF_API int hudiff_load_weights(hudiff_handle h, const hudiff_tensor* tensors, int n);

/* Device bytes the library will hold for batches of B sequences (weights excluded). */
HUDIFF_API size_t hudiff_workspace_bytes(hudiff_handle h, int B);

/* 32-bit words of dropout keep-bits one forward consumes for batch B (0 if cfg.dropout == 0). */
HUDIFF_API size_t hudiff_dropout_words(hudiff_handle h, int B);

/* replaces: logits = model(tokens, region, chain)  (model.py:366-384 / nanoencoder/model.py:325-343)
 * tokens, region: int32 [B, L]; chain: int32 [2B] = [h]*B + [l]*B (Ab) or NULL (Nb);
 * logits_out: float32 [B, L, n_tokens].  All device pointers. */
HUDIFF_API int hudiff_forward(hudiff_handle h, const int32_t* tokens, const int32_t* region, const int32_t* chain, int B,
                   const hudiff_rng* rng, float* logits_out, hudiff_stream stream);

/* replaces: the `for i in loc:` loop  (sample.py:499-513 / nanosample.py:316-329), generalised to a
 * per-sequence order.  tokens_inout int32 [B, L] is updated in place: at step t the slot
 * order[b*T + t] of sequence b is resampled from softmax(logits[b, slot, :22]).
 * sel_logits_out: optional float32 [T, B, n_tokens] (the logits row each step sampled from). */
HUDIFF_API int hudiff_sample(hudiff_handle h, int32_t* tokens_inout, const int32_t* region, const int32_t* chain,
                  const int32_t* order, int B, int T, const hudiff_rng* rng, float* sel_logits_out,
                  hudiff_stream stream);

/* Same as hudiff_sample with HOST buffers (and host rng streams): copies in, runs the loop,
 * copies the tokens back and synchronises.  This is the end-to-end call bench.py times. */
HUDIFF_API int hudiff_sample_host(hudiff_handle h, int32_t* tokens_inout, const int32_t* region, const int32_t* chain,
                       const int32_t* order, int B, int T, const hudiff_rng* rng, float* sel_logits_out);

/* The device entry points are asynchronous: this call synchronises the library's stream and returns what the
 * device flagged during the last hudiff_forward / hudiff_sample -- HUDIFF_ERR_INDEX for an out-of-range
 * token / region / chain id or sampling slot (such a slot is never written), HUDIFF_ERR_CUDA if a pipeline
 * watchdog of the tcgen05 kernels fired.  hudiff_sample_host checks by itself. */
HUDIFF_API int hudiff_check_error(hudiff_handle h);

/* Kernel launches issued by the last hudiff_forward / hudiff_sample call on this handle. */
HUDIFF_API long long hudiff_last_launch_count(hudiff_handle h);

/* Debug: 0 = tcgen05 GEMMs (default), 1 = SIMT reference GEMMs (validation only). */
HUDIFF_API int hudiff_set_gemm_impl(hudiff_handle h, int impl);

/* Measurement hook used by bench.py: one denoiser forward run eagerly with a CUDA-event pair
 * around every launch (on the library's launch stream).  Outputs are indexed by kernel class:
 * 0 = tcgen05 GEMM / dilated conv, 1 = attention core, 2 = row-wise (embed, LN-cast, assemble),
 * 3 = head.  flops_by_kind holds the algorithmic FLOPs (2 per MAC, valid rows only). */
HUDIFF_API int hudiff_profile_step(hudiff_handle h, const int32_t* tokens, const int32_t* region, const int32_t* chain,
                        int B, const hudiff_rng* rng, double* ms_by_kind, long long* launches_by_kind,
                        double* flops_by_kind);

/* Test / measurement hook: the attention core alone (reference model/encoder/cross_attention.py:131-173 after the
 * q|k|v projection and RoPE), any sequence length 1..320.  qkv: DEVICE fp16 [B*L, 3*att_dim] (q|k|v column blocks, q
 * already scaled by 1/sqrt(64)); out: DEVICE fp16 [B*L, att_dim]; head dim 64.  impl: 1 = mma.sync validation kernel,
 * 2 / 3 / 4 = the tcgen05 kernels attention_tc2 / tc3 / tc4.  iters > 0: also time `iters` launches (mean ms). */
HUDIFF_API int hudiff_debug_attention(const void* qkv, void* out, int B, int L, int att_dim, int nhead, int impl, int iters,
                           float* ms_out);
/* Diagnostic: per-phase clock totals of CTA 0 of the last persistent-attention launch
 * [0] tiles, [1] wait-for-S, [2] TMEM load, [3] max exchange, [4] exp + P store, [5] O read-back. */
HUDIFF_API int hudiff_debug_attn_clocks(long long* out16);
/* Diagnostic (HUDIFF_DBG bit 8): epilogue phase clocks of the GEMM kernel summed per epilogue kind, [4][8]:
 * [0] tiles, [1] wait for accumulator, [2] LayerNorm statistics, [3] TMEM -> smem, [4] math + global I/O. */
HUDIFF_API int hudiff_debug_gemm_clocks(unsigned long long* out32, int reset);
/* Debug/test hook: device pointer, leading dimension and element type of a named workspace buffer
 * ("E32", "X32", "A32" fp32; "opE", "hE1", "hE2", "opD", "hD1", "hD2", "qkv", "o16", "f16" fp16) as
 * left by the last forward of batch size B. */
HUDIFF_API int hudiff_debug_buffer(hudiff_handle h, int B, const char* name, void** ptr, int* ld, int* is_half);

HUDIFF_API int hudiff_destroy(hudiff_handle h);
HUDIFF_API const char* hudiff_last_error(void);
HUDIFF_API const char* hudiff_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HUDIFF_H_ */
