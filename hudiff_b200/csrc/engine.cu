// Host side of libhudiff_b200.so: weight repacking, workspaces, tensor maps, the per-step kernel
// schedule (captured into a CUDA graph) and the C-ABI of include/hudiff.h.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/hudiff.h"
#include "hd_kernels.h"

namespace {

using namespace hd;

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(HUDIFF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + " @" + \
                                       __FILE__ + ":" + std::to_string(__LINE__));            \
  } while (0)
#define CKI(expr)                          \
  do {                                     \
    int r_ = (expr);                       \
    if (r_ != HUDIFF_OK) return r_;        \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- driver entry point for tensor-map encoding (no link-time dependency on libcuda) ---------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16, 128-byte swizzle, inner box = 64 elements; zero fill outside the tensor
int encode_map(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(HUDIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string m = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r) + " rank=" + std::to_string(rank) +
                    " base=" + std::to_string(reinterpret_cast<uintptr_t>(base)) + " dims=";
    for (int i = 0; i < rank; ++i) m += std::to_string(dims[i]) + ",";
    m += " strides=";
    for (int i = 0; i + 1 < rank; ++i) m += std::to_string(strides[i]) + ",";
    m += " box=";
    for (int i = 0; i < rank; ++i) m += std::to_string(box[i]) + ",";
    return fail(HUDIFF_ERR_CUDA, m);
  }
  return HUDIFF_OK;
}

// ---- packed model ------------------------------------------------------------------------------
struct PackedLinear {
  int N = 0, K = 0, Kp = 0, taps = 1;
  __half* w_hi = nullptr; __half* w_lo = nullptr;
  float* bias = nullptr;
  float* colsum = nullptr;        // folded-LayerNorm packs only: sum over k of the packed weights of each output column
  float wscale_inv = 1.f;
  CUtensorMap* map_hi = nullptr; CUtensorMap* map_lo = nullptr;   // device copies
  int BN = 0, box_n = 0;
};
struct LNW { float* g = nullptr; float* b = nullptr; int C = 0; };
struct BlockW { LNW ln1, ln2, ln3; PackedLinear pff1, conv, pff2; int dil = 1; };
struct AttW { PackedLinear qkv, out; float2* rope = nullptr; };
struct AttBlockW {
  AttW a1, a2; LNW n1, n2; PackedLinear ff1, ff2;
  PackedLinear a2_q, a2_kv;      // last block only: query-only and key|value-only projections of the second attention
  bool has_split = false;
  // folded-LayerNorm packs (consume the raw residual row + its statistics): norm_hl1 -> second attention, norm_hl2 -> FFN
  PackedLinear a2_qkv_f, ff1_f, a2_q_f, a2_kv_f;
  bool has_fold = false;
};

enum View { VIEW_A = 0, VIEW_B = 1, VIEW_FLAT = 2 };

struct OperandBuf {
  __half* hi = nullptr; __half* lo = nullptr; int ld = 0;
  CUtensorMap* maps_hi[3] = {nullptr, nullptr, nullptr};
  CUtensorMap* maps_lo[3] = {nullptr, nullptr, nullptr};
};

struct Workspace {
  int B = 0; long long M = 0;
  RowSpace rs[3];
  bool has_view[3] = {false, false, false};
  int *tokens = nullptr, *region = nullptr, *chain = nullptr, *order = nullptr, *step = nullptr, *err_flag = nullptr;
  unsigned long long* rng = nullptr;      // [0] philox key, [1] global index of the first sequence (read by the kernels)
  int order_cap = 0, order_T = 0;
  int *s_tok = nullptr, *s_reg = nullptr, *s_chain = nullptr, *s_order = nullptr;   // device staging of hudiff_sample_host inputs
  int s_order_cap = 0;
  float *E32 = nullptr, *X32 = nullptr, *A32 = nullptr, *logits = nullptr;
  OperandBuf opE, hE1, hE2, opD, hD1, hD2, qkv, o16, f16;
  // compact [B, .] buffers of the sampling shortcut (one gathered row per sequence)
  OperandBuf tg16, tq16, to16, tf16;
  float *tA32 = nullptr, *tX32 = nullptr;
  int* t_pos = nullptr;
  float2 *lnstat = nullptr, *t_lnstat = nullptr;     // (mean, rstd) of the raw operand rows (folded LayerNorm), [M] and compact [B]
  float* lnsums = nullptr;                           // HUDIFF_LNFOLD=2: per LayerNorm instance (2 per attention block) [M][LN_SUM_PARTS][2]
                                                     // (sum, sum of squares), accumulated by the producing out-projection
  size_t lnsums_floats = 0;
  RowSpace rs_compact;
  float* scratch = nullptr; size_t scratch_floats = 0;
  CUtensorMap* attn_maps = nullptr;     // [0] query tile box, [1] key/value half-sequence box (tcgen05 attention); [2], [3]: lo planes
  std::vector<void*> allocs;
  size_t bytes = 0;
};

}  // namespace

struct hudiff_ctx {
  hudiff_cfg cfg;
  int device = 0;
  bool x3 = false;
  bool loaded = false;
  int gemm_impl = 0;
  int attn_impl = 0;              // 0 = persistent tcgen05 kernels (attention_tc3: fp16, attention_x3: fp16x3),
                                  // 1 = mma.sync kernel (HUDIFF_ATTN=mma, validation),
                                  // 2 = second-generation fp16 kernel (HUDIFF_ATTN=tc2: four warps per row, P in smem),
                                  // 4 = attention_tc4 (HUDIFF_ATTN=tc4: one query tile in flight, its key halves in parallel; measured slower)
  int two_cta = 1;                // cta_group::2 UMMA pairs (M = 256 per CTA pair, each CTA stages half of the weight tile;
                                  // needs cs == 2).  HUDIFF_2CTA=0 selects the multicast variant.
  int cs = 2;                     // CTAs per cluster sharing a weight tile by TMA multicast (HUDIFF_CLUSTER)
  int max_sms = 0;                // HUDIFF_MAX_SMS: cap of the persistent grids (0 = all SMs)
  int nsplit = 0;                 // HUDIFF_NSPLIT=1: LayerNorm-fused GEMMs split the tile over the CTA pair by columns (row
                                  // statistics through DSMEM, double-buffered accumulators).  Measured neutral for N = 384 and
                                  // slower for N = 128 (each CTA streams its own weight half: twice the L2 bytes per FLOP) -> off.
  int lnfold = 2;                 // default.  HUDIFF_LNFOLD=2: like 1, and the row statistics come from the out-projection epilogues themselves
                                  // (sum / sum of squares accumulated while the rows are stored: no statistics pass at all).
                                  // HUDIFF_LNFOLD=1: fold norm_hl1 / norm_hl2 into the projections that consume them (raw fp16
                                  // rows + a row-statistics pass instead of the LayerNorm-cast pass).  Parity-tested; measured
                                  // gain on B200 is below 1 % of a step, so the plain pass stays the default.
  int tail = 1;                   // sampling shortcut for the last attention block (HUDIFF_TAIL=0 disables)
  int l2flip = 1;                 // consecutive kernels of a step sweep the rows in alternating directions (HUDIFF_L2FLIP=0: all forward)
  int pdl = 0;                    // HUDIFF_PDL=1: GEMM / attention / LN-cast kernels use programmatic dependent launch
  int l2hint = 0;                 // HUDIFF_L2HINT=1: once-read streams (operand tiles, LN-cast input, q|k|v) use an L2 evict-first policy
  int fuse0 = 1;                  // embed / assemble also emit the LN -> act -> fp16 operand of their stack's first block
                                  // (HUDIFF_FUSE0=0: stand-alone LN-cast passes)
  int lntail = 0;                 // HUDIFF_LNTAIL: LayerNorm-cast passes fused into the producing residual epilogues (two extra
                                  // LN warps per GEMM CTA re-read the finished rows); bit mask 1 = attention blocks, 2 = wide
                                  // ByteNet stack, 4 = narrow ByteNet stack ("1" alone = all).  Parity-tested, measured SLOWER on
                                  // B200 for every subset (profiles/r02_lntail_experiment.md: the rows have left L2 by the time
                                  // they are re-read and two warps cannot issue the LayerNorm arithmetic of 128 rows per tile;
                                  // the narrow stack alone, whose stream fits L2: 1471 -> 1436 seq/s at T = 20) -> off.
  int L = 0, split = 0;           // sequence length; heavy/light boundary (Ab) or L (Nb)
  int d = 0, dh = 0, D = 0, Dh = 0, A = 0, F = 0;
  std::vector<int> dil_enc, dil_dual;
  // weights
  std::vector<void*> wallocs;
  float* emb = nullptr;           // [n_tokens, d]
  std::vector<BlockW> enc[2];     // [chain][layer]  (Nb uses chain 0 only)
  std::vector<BlockW> dual[2];
  std::vector<AttBlockW> att;
  LNW last_norm; float* dec_w = nullptr; float* dec_b = nullptr;
  float* pos_table = nullptr;     // [n_region, L, d]
  float* side_table = nullptr;    // [n_side, d]
  std::map<int, std::unique_ptr<Workspace>> ws;
  long long last_launches = 0;
  void* last_ws = nullptr;        // Workspace of the last forward / sample call (hudiff_check_error)
  // graph cache for the sampling step
  cudaGraphExec_t graph_exec = nullptr;
  std::string graph_key;
  long long graph_per_step = 0;
  // all device work runs on this non-blocking stream, ordered after / before the caller's stream
  // with events (stream capture is not allowed on the legacy default stream PyTorch uses)
  cudaStream_t work = nullptr;
  cudaStream_t io = nullptr;      // host <-> device copies of hudiff_sample_host
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

namespace {

// ---- allocation helpers ---------------------------------------------------------------------------
template <class T>
int dev_alloc(std::vector<void*>& track, T** out, size_t count, size_t* total = nullptr, bool zero = true) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 256);
  CK(cudaMalloc(&p, bytes));
  if (zero) CK(cudaMemset(p, 0, bytes));
  track.push_back(p);
  if (total) *total += bytes;
  *out = static_cast<T*>(p);
  return HUDIFF_OK;
}

struct TensorTable {
  std::map<std::string, hudiff_tensor> t;
  const hudiff_tensor* find(const std::string& k) const {
    auto it = t.find(k);
    return it == t.end() ? nullptr : &it->second;
  }
};

long long numel(const hudiff_tensor& t) {
  long long n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  return n * (t.dtype == 1 ? 2 : 1);
}

int fetch_host(const TensorTable& tt, const std::string& name, std::vector<float>& out, long long expect = -1) {
  const hudiff_tensor* t = tt.find(name);
  if (!t) return fail(HUDIFF_ERR_MISSING_WEIGHT, "missing state-dict tensor: " + name);
  const long long n = numel(*t);
  if (expect >= 0 && n != expect)
    return fail(HUDIFF_ERR_INVALID, "tensor " + name + " has " + std::to_string(n) + " elements, expected " +
                                        std::to_string(expect));
  out.resize(n);
  if (t->on_device) CK(cudaMemcpy(out.data(), t->data, n * sizeof(float), cudaMemcpyDeviceToHost));
  else std::memcpy(out.data(), t->data, n * sizeof(float));
  return HUDIFF_OK;
}

int upload_f32(hudiff_ctx* c, const std::vector<float>& h, float** out) {
  CKI(dev_alloc(c->wallocs, out, h.size(), nullptr, false));
  CK(cudaMemcpy(*out, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return HUDIFF_OK;
}

int load_f32(hudiff_ctx* c, const TensorTable& tt, const std::string& name, float** out, long long expect) {
  std::vector<float> h;
  CKI(fetch_host(tt, name, h, expect));
  return upload_f32(c, h, out);
}

int load_ln(hudiff_ctx* c, const TensorTable& tt, const std::string& pfx, int C, LNW* ln) {
  ln->C = C;
  CKI(load_f32(c, tt, pfx + ".weight", &ln->g, C));
  CKI(load_f32(c, tt, pfx + ".bias", &ln->b, C));
  return HUDIFF_OK;
}

// Pack fp32 weights (rows = concatenation of `names`, each [N_i, K(, taps)]) into scaled fp16 planes.
// `fold_ln` (LayerNorm weight prefix, e.g. "self_at.layers.0.norm_hl1"): the layer consumes LN(x) in the
// reference; here it consumes the raw x and the LayerNorm is folded into the weights and the epilogue:
//   W'[n, k] = W[n, k] * gamma[k]     bias'[n] = b[n] + sum_k W[n, k] * beta[k]     (taps == 1 only)
int pack_linear(hudiff_ctx* c, const TensorTable& tt, const std::vector<std::string>& wnames,
                const std::vector<std::string>& bnames, int N_each, int K, int taps, bool full_row, PackedLinear* out,
                const std::string& fold_ln = std::string()) {
  std::vector<float> w, b;
  for (size_t i = 0; i < wnames.size(); ++i) {
    std::vector<float> wi, bi;
    CKI(fetch_host(tt, wnames[i], wi, static_cast<long long>(N_each) * K * taps));
    CKI(fetch_host(tt, bnames[i], bi, N_each));
    w.insert(w.end(), wi.begin(), wi.end());
    b.insert(b.end(), bi.begin(), bi.end());
  }
  const int N = N_each * static_cast<int>(wnames.size());
  if (!fold_ln.empty()) {
    if (taps != 1) return fail(HUDIFF_ERR_INVALID, "LayerNorm folding needs a 1-tap layer");
    std::vector<float> g, be;
    CKI(fetch_host(tt, fold_ln + ".weight", g, K));
    CKI(fetch_host(tt, fold_ln + ".bias", be, K));
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        float& wv = w[static_cast<size_t>(n) * K + k];
        acc += static_cast<double>(wv) * be[k];
        wv *= g[k];
      }
      b[n] = static_cast<float>(b[n] + acc);
    }
  }
  float mx = 0.f;
  for (float v : w) mx = std::max(mx, std::fabs(v));
  // power-of-two pre-scale: largest |w| lands in (64, 128] so small weights stay in fp16's normal range
  int e = 0;
  if (mx > 0.f && std::isfinite(mx)) e = 7 - static_cast<int>(std::ceil(std::log2(mx)));
  e = std::max(-14, std::min(e, 24));
  const float scale = std::ldexp(1.0f, e);
  out->N = N; out->K = K; out->taps = taps; out->Kp = round_up(K, 64);
  out->wscale_inv = 1.0f / scale;
  if (N % 64 != 0) return fail(HUDIFF_ERR_INVALID, "layer width " + std::to_string(N) + " must be a multiple of 64");
  if (full_row) {
    if (N > 512) return fail(HUDIFF_ERR_INVALID, "LayerNorm-fused layer width " + std::to_string(N) + " exceeds 512 TMEM columns");
    out->BN = N;
  } else {
    out->BN = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : (N <= 256 ? N : 64);
  }
  if (out->BN % (64 * c->cs) != 0)
    return fail(HUDIFF_ERR_INVALID, "tile width " + std::to_string(out->BN) + " must be a multiple of 64 x cluster size");
  out->box_n = out->BN / c->cs;                       // the slice one CTA fetches and multicasts
  if (out->box_n > 256) out->box_n /= 2;
  if (c->two_cta) out->box_n = (out->BN % 256 == 0) ? 128 : 64;   // per-CTA half of each (<= 256)-column chunk
  float* wdev = nullptr;
  std::vector<void*> tmp;
  CKI(dev_alloc(tmp, &wdev, w.size(), nullptr, false));
  CK(cudaMemcpy(wdev, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  const size_t packed = static_cast<size_t>(N) * taps * out->Kp;
  CKI(dev_alloc(c->wallocs, &out->w_hi, packed));
  if (c->x3) CKI(dev_alloc(c->wallocs, &out->w_lo, packed));
  CK(pack_weight_launch(wdev, N, K, taps, out->Kp, scale, out->w_hi, out->w_lo, 0));
  CK(cudaDeviceSynchronize());
  cudaFree(wdev);
  CKI(upload_f32(c, b, &out->bias));
  if (!fold_ln.empty()) {
    CKI(dev_alloc(c->wallocs, &out->colsum, N, nullptr, false));
    CK(weight_colsum_launch(out->w_hi, out->w_lo, N, taps * out->Kp, out->colsum, 0));
    CK(cudaDeviceSynchronize());
  }
  // weight tensor maps: [N, taps*Kp] fp16, box = 64 k x box_n rows
  CUtensorMap hm;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(taps) * out->Kp, static_cast<cuuint64_t>(N)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(taps) * out->Kp * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(out->box_n)};
  CKI(encode_map(&hm, out->w_hi, 2, dims, strides, box));
  CKI(dev_alloc(c->wallocs, &out->map_hi, 1, nullptr, false));
  CK(cudaMemcpy(out->map_hi, &hm, sizeof(hm), cudaMemcpyHostToDevice));
  if (c->x3) {
    CKI(encode_map(&hm, out->w_lo, 2, dims, strides, box));
    CKI(dev_alloc(c->wallocs, &out->map_lo, 1, nullptr, false));
    CK(cudaMemcpy(out->map_lo, &hm, sizeof(hm), cudaMemcpyHostToDevice));
  }
  return HUDIFF_OK;
}

int load_block(hudiff_ctx* c, const TensorTable& tt, const std::string& pfx, int d_in, int d_h, int k, int dil,
               BlockW* bw) {
  bw->dil = dil;
  CKI(load_ln(c, tt, pfx + ".sequence1.0", d_in, &bw->ln1));
  CKI(pack_linear(c, tt, {pfx + ".sequence1.2.conv.weight"}, {pfx + ".sequence1.2.conv.bias"}, d_h, d_in, 1, true, &bw->pff1));
  CKI(load_ln(c, tt, pfx + ".sequence1.3", d_h, &bw->ln2));
  CKI(pack_linear(c, tt, {pfx + ".conv.weight"}, {pfx + ".conv.bias"}, d_h, d_h, k, true, &bw->conv));
  CKI(load_ln(c, tt, pfx + ".sequence2.0", d_h, &bw->ln3));
  CKI(pack_linear(c, tt, {pfx + ".sequence2.2.conv.weight"}, {pfx + ".sequence2.2.conv.bias"}, d_in, d_h, 1, false, &bw->pff2));
  return HUDIFF_OK;
}

int load_att(hudiff_ctx* c, const TensorTable& tt, const std::string& pfx, AttW* aw) {
  CKI(pack_linear(c, tt, {pfx + ".query.weight", pfx + ".key.weight", pfx + ".value.weight"},
                  {pfx + ".query.bias", pfx + ".key.bias", pfx + ".value.bias"}, c->A, c->D, 1, false, &aw->qkv));
  CKI(pack_linear(c, tt, {pfx + ".out_put.weight"}, {pfx + ".out_put.bias"}, c->D, c->A, 1, false, &aw->out));
  float* rope = nullptr;
  const int hd2 = (c->A / c->cfg.nhead) / 2;
  CKI(load_f32(c, tt, pfx + ".rope", &rope, static_cast<long long>(c->L) * hd2 * 2));
  aw->rope = reinterpret_cast<float2*>(rope);
  return HUDIFF_OK;
}

// Constant folding: the region/position and chain-type branches do not depend on the tokens.
int fold_constants(hudiff_ctx* c, const TensorTable& tt) {
  const int d = c->d, L = c->L, nr = c->cfg.n_region, re = c->cfg.r_embedding;
  std::vector<void*> tmp;
  auto cleanup = [&]() { for (void* p : tmp) cudaFree(p); };
  auto up = [&](const std::string& name, long long expect, float** out) -> int {
    std::vector<float> h;
    CKI(fetch_host(tt, name, h, expect));
    CKI(dev_alloc(tmp, out, h.size(), nullptr, false));
    CK(cudaMemcpy(*out, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return HUDIFF_OK;
  };
  float *remb, *g0, *b0, *pw, *pb, *g1, *b1, *pe, *w1, *bb1, *w2, *bb2;
  int rc = HUDIFF_OK;
  do {
    if ((rc = up("region_encoder.region_embedding.weight", static_cast<long long>(nr) * re, &remb))) break;
    if ((rc = up("region_encoder.region_layer1.0.weight", re, &g0))) break;
    if ((rc = up("region_encoder.region_layer1.0.bias", re, &b0))) break;
    if ((rc = up("region_encoder.region_layer1.2.conv.weight", static_cast<long long>(d) * re, &pw))) break;
    if ((rc = up("region_encoder.region_layer1.2.conv.bias", d, &pb))) break;
    if ((rc = up("region_encoder.region_layer1.3.weight", d, &g1))) break;
    if ((rc = up("region_encoder.region_layer1.3.bias", d, &b1))) break;
    {
      // the buffer may be longer than L in principle; the reference slices pe[:L]
      const hudiff_tensor* t = tt.find("pos_encoder.pos_embedding.pe");
      if (!t) { rc = fail(HUDIFF_ERR_MISSING_WEIGHT, "missing state-dict tensor: pos_encoder.pos_embedding.pe"); break; }
      if (numel(*t) < static_cast<long long>(L) * d) { rc = fail(HUDIFF_ERR_INVALID, "pos_embedding.pe shorter than max_len"); break; }
      if ((rc = up("pos_encoder.pos_embedding.pe", -1, &pe))) break;
    }
    if ((rc = up("pos_encoder.pos_lin.ln1.weight", 2LL * d * d, &w1))) break;
    if ((rc = up("pos_encoder.pos_lin.ln1.bias", 2 * d, &bb1))) break;
    if ((rc = up("pos_encoder.pos_lin.ln2.weight", 2LL * d * d, &w2))) break;
    if ((rc = up("pos_encoder.pos_lin.ln2.bias", d, &bb2))) break;
    float *t0, *t1, *R, *x, *hbuf, *m;
    if ((rc = dev_alloc(tmp, &t0, nr * re))) break;
    if ((rc = dev_alloc(tmp, &t1, nr * d))) break;
    if ((rc = dev_alloc(tmp, &R, nr * d))) break;
    if ((rc = dev_alloc(tmp, &x, static_cast<size_t>(nr) * L * d))) break;
    if ((rc = dev_alloc(tmp, &hbuf, static_cast<size_t>(nr) * L * 2 * d))) break;
    if ((rc = dev_alloc(tmp, &m, static_cast<size_t>(nr) * L * d))) break;
    if ((rc = dev_alloc(c->wallocs, &c->pos_table, static_cast<size_t>(nr) * L * d))) break;
    cudaError_t e;
    e = f32_layernorm_launch(remb, g0, b0, t0, nr, re, ACT_RELU, 0);                       // LN(4) + ReLU
    if (e == cudaSuccess) e = f32_linear_launch(t0, pw, pb, t1, nr, re, d, ACT_NONE, 0);     // PFF 4 -> d
    if (e == cudaSuccess) e = f32_layernorm_launch(t1, g1, b1, R, nr, d, ACT_RELU, 0);       // LN(d) + ReLU
    if (e == cudaSuccess) e = f32_outer_add_launch(R, pe, x, nr, L, d, 0);                   // + pe[l]
    if (e == cudaSuccess) e = f32_linear_launch(x, w1, bb1, hbuf, nr * L, d, 2 * d, ACT_GELU, 0);
    if (e == cudaSuccess) e = f32_linear_launch(hbuf, w2, bb2, m, nr * L, 2 * d, d, ACT_NONE, 0);
    if (e == cudaSuccess) e = f32_add_launch(x, m, c->pos_table, nr * L * d, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { rc = fail(HUDIFF_ERR_CUDA, std::string("constant folding (pos): ") + cudaGetErrorString(e)); break; }
    if (c->cfg.kind == HUDIFF_AB) {
      const int ns = c->cfg.n_side, se = c->cfg.s_embedding;
      float *semb, *sw0, *sb0, *sg, *sb, *sw1, *sb1, *u0, *u1;
      if ((rc = up("side_encoder.side_embeddinng.weight", static_cast<long long>(ns) * se, &semb))) break;
      if ((rc = up("side_encoder.side_mlp.0.weight", static_cast<long long>(d) * se, &sw0))) break;
      if ((rc = up("side_encoder.side_mlp.0.bias", d, &sb0))) break;
      if ((rc = up("side_encoder.side_mlp.1.weight", d, &sg))) break;
      if ((rc = up("side_encoder.side_mlp.1.bias", d, &sb))) break;
      if ((rc = up("side_encoder.side_mlp.3.weight", static_cast<long long>(d) * d, &sw1))) break;
      if ((rc = up("side_encoder.side_mlp.3.bias", d, &sb1))) break;
      if ((rc = dev_alloc(tmp, &u0, ns * d))) break;
      if ((rc = dev_alloc(tmp, &u1, ns * d))) break;
      if ((rc = dev_alloc(c->wallocs, &c->side_table, ns * d))) break;
      e = f32_linear_launch(semb, sw0, sb0, u0, ns, se, d, ACT_NONE, 0);
      if (e == cudaSuccess) e = f32_layernorm_launch(u0, sg, sb, u1, ns, d, ACT_RELU, 0);
      if (e == cudaSuccess) e = f32_linear_launch(u1, sw1, sb1, c->side_table, ns, d, d, ACT_NONE, 0);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { rc = fail(HUDIFF_ERR_CUDA, std::string("constant folding (side): ") + cudaGetErrorString(e)); break; }
    }
  } while (0);
  cleanup();
  return rc;
}

// ---- workspace --------------------------------------------------------------------------------------
int make_operand(hudiff_ctx* c, Workspace* w, OperandBuf* ob, int ld, const std::vector<int>& views, long long rows = -1,
                 const RowSpace* rs_override = nullptr) {
  ob->ld = ld;
  const size_t n = static_cast<size_t>(rows >= 0 ? rows : w->M) * ld;
  CKI(dev_alloc(w->allocs, &ob->hi, n, &w->bytes));
  if (c->x3) CKI(dev_alloc(w->allocs, &ob->lo, n, &w->bytes));
  for (int v : views) {
    const RowSpace& rs = rs_override ? *rs_override : w->rs[v];
    for (int plane = 0; plane < (c->x3 ? 2 : 1); ++plane) {
      CUtensorMap host[16];
      const __half* base = (plane ? ob->lo : ob->hi) + static_cast<size_t>(rs.row_off) * ld;
      for (int hi = 0; hi < 16; ++hi) {
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rs.Lc), static_cast<cuuint64_t>(rs.nseq)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2,
                                 static_cast<cuuint64_t>(std::max(rs.seq_stride, 1)) * ld * 2};
        cuuint32_t box[3] = {64, static_cast<cuuint32_t>((hi + 1) * 8), 1};
        CKI(encode_map(&host[hi], base, 3, dims, strides, box));
      }
      CUtensorMap* dev = nullptr;
      CKI(dev_alloc(w->allocs, &dev, 16, &w->bytes, false));
      CK(cudaMemcpy(dev, host, sizeof(host), cudaMemcpyHostToDevice));
      (plane ? ob->maps_lo : ob->maps_hi)[v] = dev;
    }
  }
  return HUDIFF_OK;
}

size_t workspace_estimate(const hudiff_ctx* c, int B) {
  const size_t M = static_cast<size_t>(B) * c->L;
  const size_t planes = c->x3 ? 2 : 1;
  size_t f32 = M * (c->d + 2 * c->D + c->cfg.n_tokens) * 4;
  size_t f16 = M * (round_up(c->d, 64) + 2 * round_up(c->dh, 64) + round_up(c->D, 64) + 2 * round_up(c->Dh, 64) +
                    3 * c->A + c->A + round_up(c->F, 64)) * 2 * planes;
  return f32 + f16 + M * 8 + (1 << 20);
}

int get_workspace(hudiff_ctx* c, int B, Workspace** out) {
  auto it = c->ws.find(B);
  if (it != c->ws.end()) { *out = it->second.get(); return HUDIFF_OK; }
  // keep at most one cached batch size besides the new one (180 GB is plenty, but be tidy)
  if (c->ws.size() >= 2) {
    for (auto& kv : c->ws) for (void* p : kv.second->allocs) cudaFree(p);
    c->ws.clear();
    c->last_ws = nullptr;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear(); }
  }
  std::unique_ptr<Workspace> w(new Workspace());
  w->B = B; w->M = static_cast<long long>(B) * c->L;
  const int L = c->L;
  if (c->cfg.kind == HUDIFF_AB) {
    w->rs[VIEW_A] = RowSpace{B, c->split, round_up(c->split, 8), L, 0};
    w->rs[VIEW_B] = RowSpace{B, L - c->split, round_up(L - c->split, 8), L, c->split};
    w->has_view[VIEW_A] = w->has_view[VIEW_B] = true;
  } else {
    w->rs[VIEW_A] = RowSpace{B, L, round_up(L, 8), L, 0};
    w->has_view[VIEW_A] = true;
  }
  w->rs[VIEW_FLAT] = RowSpace{1, static_cast<int>(w->M), round_up(static_cast<int>(w->M), 128), static_cast<int>(w->M), 0};
  w->has_view[VIEW_FLAT] = true;
  CKI(dev_alloc(w->allocs, &w->tokens, w->M, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->region, w->M, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->chain, 2 * B, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->step, 4, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->err_flag, 4, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->rng, 4, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->E32, w->M * c->d, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->X32, w->M * c->D, &w->bytes));
  CKI(dev_alloc(w->allocs, &w->A32, w->M * c->D, &w->bytes));
  const bool ab = c->cfg.kind == HUDIFF_AB;
  const std::vector<int> conv_views = ab ? std::vector<int>{VIEW_A, VIEW_B} : std::vector<int>{VIEW_A};
  CKI(make_operand(c, w.get(), &w->opE, round_up(c->d, 64), conv_views));
  CKI(make_operand(c, w.get(), &w->hE1, round_up(c->dh, 64), conv_views));
  CKI(make_operand(c, w.get(), &w->hE2, round_up(c->dh, 64), conv_views));
  const std::vector<int> d_views = ab ? std::vector<int>{VIEW_A, VIEW_B, VIEW_FLAT} : std::vector<int>{VIEW_A, VIEW_FLAT};
  CKI(make_operand(c, w.get(), &w->opD, round_up(c->D, 64), d_views));
  CKI(make_operand(c, w.get(), &w->hD1, round_up(c->Dh, 64), conv_views));
  CKI(make_operand(c, w.get(), &w->hD2, round_up(c->Dh, 64), conv_views));
  CKI(make_operand(c, w.get(), &w->qkv, 3 * c->A, std::vector<int>{}));
  {
    // tcgen05 attention: 3-D maps (columns, rows of one sequence, sequence) over the q|k|v operand plane;
    // rows past the sequence end are zero-filled by the TMA unit
    CUtensorMap host[4];
    const int LK = round_up(c->L, 64);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(3 * c->A), static_cast<cuuint64_t>(c->L), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(3 * c->A) * 2, static_cast<cuuint64_t>(c->L) * 3 * c->A * 2};
    cuuint32_t box_q[3] = {64, 128, 1};
    cuuint32_t box_kv[3] = {64, static_cast<cuuint32_t>(LK / 2), 1};
    CKI(encode_map(&host[0], w->qkv.hi, 3, dims, strides, box_q));
    CKI(encode_map(&host[1], w->qkv.hi, 3, dims, strides, box_kv));
    if (c->x3) {                                         // lo planes of the split-fp16 policy (attention_x3.cu)
      CKI(encode_map(&host[2], w->qkv.lo, 3, dims, strides, box_q));
      CKI(encode_map(&host[3], w->qkv.lo, 3, dims, strides, box_kv));
    } else {
      host[2] = host[0]; host[3] = host[1];
    }
    CKI(dev_alloc(w->allocs, &w->attn_maps, 4, &w->bytes, false));
    CK(cudaMemcpy(w->attn_maps, host, sizeof(host), cudaMemcpyHostToDevice));
  }
  CKI(make_operand(c, w.get(), &w->o16, c->A, std::vector<int>{VIEW_FLAT}));
  CKI(make_operand(c, w.get(), &w->f16, round_up(c->F, 64), std::vector<int>{VIEW_FLAT}));
  {
    // compact row space: one row per sequence (tile rows past B are out of bounds -> zero-filled / not stored)
    w->rs_compact = RowSpace{1, B, round_up(B, 128), B, 0};
    // the TMA maps span at least one full 128-row box of (zero-initialised) memory
    const int Brows = std::max(B, 128);
    const RowSpace rs_map{1, Brows, round_up(Brows, 128), Brows, 0};
    const std::vector<int> cv{VIEW_FLAT};
    CKI(make_operand(c, w.get(), &w->tg16, round_up(c->D, 64), cv, Brows, &rs_map));
    CKI(make_operand(c, w.get(), &w->tq16, c->A, std::vector<int>{}, Brows));
    CKI(make_operand(c, w.get(), &w->to16, c->A, cv, Brows, &rs_map));
    CKI(make_operand(c, w.get(), &w->tf16, round_up(c->F, 64), cv, Brows, &rs_map));
    CKI(dev_alloc(w->allocs, &w->tA32, static_cast<size_t>(B) * c->D, &w->bytes));
    CKI(dev_alloc(w->allocs, &w->tX32, static_cast<size_t>(B) * c->D, &w->bytes));
    CKI(dev_alloc(w->allocs, &w->t_pos, B, &w->bytes));
    CKI(dev_alloc(w->allocs, &w->lnstat, w->M, &w->bytes));
    CKI(dev_alloc(w->allocs, &w->t_lnstat, static_cast<size_t>(std::max(B, 128)) * LN_SUM_PARTS, &w->bytes));
    if (c->lnfold == 2) {
      w->lnsums_floats = static_cast<size_t>(2) * std::max(1, c->cfg.cs_layers) * w->M * LN_SUM_PARTS * 2;
      CKI(dev_alloc(w->allocs, &w->lnsums, w->lnsums_floats, &w->bytes));
    }
  }
  *out = w.get();
  c->ws[B] = std::move(w);
  return HUDIFF_OK;
}

// ---- the per-forward kernel schedule --------------------------------------------------------------------
struct DropCfg {
  int mode = DROP_OFF;             // DROP_OFF / DROP_BITS / DROP_PHILOX
  const uint32_t* bits = nullptr;
  long long step_words = 0;
};

enum OpKind { OP_GEMM = 0, OP_ATTN = 1, OP_ROWWISE = 2, OP_HEAD = 3, OP_KINDS = 4 };
struct Plan {
  std::vector<std::function<cudaError_t(cudaStream_t)>> ops;
  std::vector<int> kinds;          // OpKind of each op (profiling)
  std::vector<double> flops;       // algorithmic FLOPs (2*MAC) of each op, 0 for non-contractions
  std::vector<std::string> labels; // per-op names for HUDIFF_PROFILE_DUMP
  long long launches_per_run = 0;
  int flip = 0, dir = 0;           // flip: consecutive row-parallel kernels alternate their sweep direction (L2 reuse)
  int next_dir() { if (flip) dir ^= 1; return dir; }
  void add(std::function<cudaError_t(cudaStream_t)> f, int kind, double fl = 0.0, int launches = 1,
           std::string label = std::string()) {
    ops.push_back(std::move(f)); kinds.push_back(kind); flops.push_back(fl); labels.push_back(std::move(label));
    launches_per_run += launches;
  }
};

GemmProblem make_problem(const hudiff_ctx* c, const Workspace* w, const PackedLinear& W, const OperandBuf& A, int view,
                         int dil, bool full_row, const EpiParams& ep, const RowSpace* rs_override = nullptr) {
  GemmProblem P;
  std::memset(&P, 0, sizeof(P));
  P.rs = rs_override ? *rs_override : w->rs[view];
  P.a_maps_hi = A.maps_hi[view];
  P.a_maps_lo = A.maps_lo[view];
  P.b_map_hi = W.map_hi;
  P.b_map_lo = W.map_lo;
  P.m_tiles = (P.rs.nseq * P.rs.Lcp + 127) / 128;
  P.N = W.N;
  P.BN = full_row ? W.N : W.BN;
  P.kchunks = W.Kp / 64;
  P.k_logical = W.K;
  P.taps = W.taps;
  P.dil = dil;
  P.b_box_n = W.box_n;
  P.a_hi = A.hi; P.a_lo = A.lo; P.lda = A.ld;
  P.w_hi = W.w_hi; P.w_lo = W.w_lo; P.ldw = W.taps * W.Kp;
  P.ep = ep;
  P.ep.wscale = W.wscale_inv;
  P.ep.bias = W.bias;
  return P;
}

EpiParams epi_ln(const LNW& ln, int act, OperandBuf& out) {
  EpiParams e;
  std::memset(&e, 0, sizeof(e));
  e.epi = EPI_LN_ACT_F16; e.act = act; e.gamma = ln.g; e.beta = ln.b;
  e.out16_hi = out.hi; e.out16_lo = out.lo; e.ld16 = out.ld;
  return e;
}
EpiParams epi_resid(float* out32, int ld32, const float* resid, int ldr, OperandBuf* copy16) {
  EpiParams e;
  std::memset(&e, 0, sizeof(e));
  e.epi = EPI_RESID_F32; e.out32 = out32; e.ld32 = ld32; e.resid = resid; e.ldr = ldr;
  if (copy16) { e.out16_hi = copy16->hi; e.out16_lo = copy16->lo; e.ld16 = copy16->ld; }
  return e;
}

// Fused LayerNorm tail of a residual epilogue: the GEMM also emits act(LN(row)) of its finished fp32 rows as
// the fp16 operand of the next layer (replaces a stand-alone LN-cast pass over the stream).
void set_ln_tail(EpiParams& e, const LNW& ln, int act, OperandBuf& out) {
  e.ln_gamma = ln.g; e.ln_beta = ln.b; e.ln_act = act;
  e.ln_out_hi = out.hi; e.ln_out_lo = out.lo; e.ln_ld = out.ld;
}

void set_dropout(EpiParams& e, const DropCfg& dc, double p, long long* word_off, long long words, uint32_t call_id) {
  if (dc.mode == DROP_OFF) return;
  e.drop_mode = dc.mode;
  const float keep = static_cast<float>(1.0 - p);
  e.drop_scale = 1.0f / keep;
  e.drop_thresh16 = static_cast<uint32_t>(std::lround((1.0 - p) * 65536.0));
  e.drop_onebit = (p == 0.5) ? 1 : 0;
  e.call_id = call_id;
  if (dc.mode == DROP_BITS) {
    e.drop_bits = dc.bits + *word_off;
    e.drop_step_words = dc.step_words;
  }
  *word_off += words;
}

long long dropout_words(const hudiff_ctx* c, int B) {
  if (!(c->cfg.dropout > 0.0)) return 0;
  const long long rows = static_cast<long long>(B) * c->L;
  return rows * (c->d / 32) * c->cfg.n_encoder_layers + rows * (c->D / 32) * c->cfg.dual_layers;
}

int launch_gemm(hudiff_ctx* c, Workspace* w, Plan& plan, GemmParams G) {
  G.x3 = c->x3 ? 1 : 0;
  G.cs = c->cs;
  G.two_cta = c->two_cta;
  G.max_sms = c->max_sms;
  G.nsplit = (c->nsplit && c->cs == 2 && !c->two_cta && G.p[0].ep.epi == EPI_LN_ACT_F16) ? 1 : 0;
#ifdef HD_DEBUG_HOOKS
  if (const char* d = std::getenv("HUDIFF_DBG")) G.dbg = std::atoi(d);      // diagnostic builds only (build.sh HD_DEBUG_HOOKS=1)
#endif
  G.step_ptr = w->step;
  G.reverse = plan.next_dir();
  G.l2hint = c->l2hint;
  G.pdl = c->pdl;
  G.rng_ptr = w->rng;
  G.err_flag = w->err_flag;
  double fl = 0.0;                 // algorithmic FLOPs: valid rows x N x logical K x taps, 2 per MAC
  for (int i = 0; i < G.nprob; ++i)
    fl += 2.0 * G.p[i].rs.nseq * G.p[i].rs.Lc * static_cast<double>(G.p[i].N) * G.p[i].k_logical * G.p[i].taps;
  if (c->gemm_impl == 1 && G.p[0].ep.ln_gamma) {
    // SIMT validation path: the LayerNorm tail runs as the stand-alone LN-cast pass it replaces
    LnCastParams lp;
    std::memset(&lp, 0, sizeof(lp));
    const EpiParams& e0 = G.p[0].ep;
    lp.x = e0.out32; lp.ldx = e0.ld32; lp.C = G.p[0].N; lp.rows = w->M;
    lp.gamma[0] = e0.ln_gamma; lp.beta[0] = e0.ln_beta;
    if (G.nprob == 2) { lp.gamma[1] = G.p[1].ep.ln_gamma; lp.beta[1] = G.p[1].ep.ln_beta; lp.seq_len = c->L; lp.split = c->split; }
    lp.act = e0.ln_act; lp.out_hi = e0.ln_out_hi; lp.out_lo = e0.ln_out_lo; lp.ldo = e0.ln_ld;
    for (int i = 0; i < G.nprob; ++i) { G.p[i].ep.ln_gamma = nullptr; G.p[i].ep.ln_beta = nullptr; G.p[i].ep.ln_out_hi = nullptr; G.p[i].ep.ln_out_lo = nullptr; }
    CKI(launch_gemm(c, w, plan, G));
    plan.add([lp](cudaStream_t s) { return ln_cast_launch(lp, s); }, OP_ROWWISE, 0.0, 1, "ln_cast (simt tail)");
    return HUDIFF_OK;
  }
  if (c->gemm_impl == 1) {
    size_t need = 0;
    int m_tiles = 0, ldn = 0;
    for (int i = 0; i < G.nprob; ++i) { m_tiles += G.p[i].m_tiles; ldn = std::max(ldn, G.p[i].N); }
    need = static_cast<size_t>(m_tiles) * 128 * ldn;
    if (need > w->scratch_floats) {
      size_t want = static_cast<size_t>((w->M + 512)) * std::max(3 * c->A, c->D);
      want = std::max(want, need);
      CKI(dev_alloc(w->allocs, &w->scratch, want, &w->bytes, false));
      w->scratch_floats = want;
    }
    float* scratch = w->scratch; size_t nf = w->scratch_floats;
    plan.add([G, scratch, nf](cudaStream_t s) { return gemm_ref_launch(G, scratch, nf, s); }, OP_GEMM, fl, 2);
  } else {
    char lab[96];
    std::snprintf(lab, sizeof(lab), "gemm epi%d N%d K%d taps%d np%d", G.p[0].ep.epi, G.p[0].N, G.p[0].k_logical, G.p[0].taps, G.nprob);
    plan.add([G](cudaStream_t s) { return gemm_tc_launch(G, s); }, OP_GEMM, fl, 1, lab);
  }
  return HUDIFF_OK;
}

// One ByteNet stack (6 blocks) over the heavy/light chains (Ab) or the single chain (Nb).
// LayerNorm output block of the kernel that PRODUCES the input of a ByteNet stack (embed / assemble): the operand of
// the stack's first block, act(LN1(x)), with the per-chain parameter sets of the Ab model.
RowLnOut first_block_ln(const hudiff_ctx* c, const std::vector<BlockW>* blocks, int act, OperandBuf& op) {
  RowLnOut ln;
  std::memset(&ln, 0, sizeof(ln));
  if (!c->fuse0 || blocks[0].empty()) return ln;
  ln.gamma[0] = blocks[0][0].ln1.g; ln.beta[0] = blocks[0][0].ln1.b;
  if (c->cfg.kind == HUDIFF_AB) { ln.gamma[1] = blocks[1][0].ln1.g; ln.beta[1] = blocks[1][0].ln1.b; ln.seq_len = c->L; ln.split = c->split; }
  ln.act = act; ln.out_hi = op.hi; ln.out_lo = op.lo; ln.ldo = op.ld;
  return ln;
}

int plan_stack(hudiff_ctx* c, Workspace* w, Plan& plan, std::vector<BlockW>* blocks /*[2]*/, float* x32, int width,
               OperandBuf& op, OperandBuf& h1, OperandBuf& h2, int act, double drop_p, const DropCfg& dc,
               long long* word_off, bool interleaved, uint32_t call_base, OperandBuf* final_copy16) {
  const bool ab = c->cfg.kind == HUDIFF_AB;
  const int nchain = ab ? 2 : 1;
  const int nl = static_cast<int>(blocks[0].size());
  const bool live = dc.mode != DROP_OFF && c->cfg.dropout > 0.0;
  // dropout-bit offsets follow the reference's call order: the aa_encoder interleaves H and L per
  // layer (model.py:173-178) while DualConv runs all heavy layers, then all light layers (:292-304).
  std::vector<long long> off(static_cast<size_t>(nl) * nchain, 0);
  if (live) {
    long long o = *word_off;
    auto words = [&](int ch) { return static_cast<long long>(w->B) * w->rs[ch].Lc * (width / 32); };
    if (interleaved) {
      for (int n = 0; n < nl; ++n) for (int ch = 0; ch < nchain; ++ch) { off[n * nchain + ch] = o; o += words(ch); }
    } else {
      for (int ch = 0; ch < nchain; ++ch) for (int n = 0; n < nl; ++n) { off[n * nchain + ch] = o; o += words(ch); }
    }
    *word_off = o;
  }
  for (int n = 0; n < nl; ++n) {
    // LN(d_in) -> act -> fp16 operand: a stand-alone pass for the first block only; every later block gets its
    // operand from the LayerNorm tail of the previous block's residual epilogue (c->lntail)
    const bool lt = (c->lntail & (width <= 256 ? 4 : 2)) != 0;
    if ((n == 0 && !c->fuse0) || (n > 0 && !lt)) {             // block 0: emitted by embed / assemble (fuse0)
      LnCastParams lp;
      std::memset(&lp, 0, sizeof(lp));
      lp.x = x32; lp.ldx = width; lp.C = width; lp.rows = w->M;
      lp.gamma[0] = blocks[0][n].ln1.g; lp.beta[0] = blocks[0][n].ln1.b;
      if (ab) { lp.gamma[1] = blocks[1][n].ln1.g; lp.beta[1] = blocks[1][n].ln1.b; lp.seq_len = c->L; lp.split = c->split; }
      lp.act = act; lp.out_hi = op.hi; lp.out_lo = op.lo; lp.ldo = op.ld;
      lp.reverse = plan.next_dir();
      lp.l2hint = c->l2hint;
      lp.pdl = c->pdl;
      plan.add([lp](cudaStream_t s) { return ln_cast_launch(lp, s); }, OP_ROWWISE, 0.0, 1, "ln_cast C" + std::to_string(width));
    }

    GemmParams g1, g2, g3;
    std::memset(&g1, 0, sizeof(g1)); std::memset(&g2, 0, sizeof(g2)); std::memset(&g3, 0, sizeof(g3));
    g1.nprob = g2.nprob = g3.nprob = nchain;
      for (int ch = 0; ch < nchain; ++ch) {
      const BlockW& bw = blocks[ch][n];
      g1.p[ch] = make_problem(c, w, bw.pff1, op, ch, 1, true, epi_ln(bw.ln2, act, h1));
      g2.p[ch] = make_problem(c, w, bw.conv, h1, ch, bw.dil, true, epi_ln(bw.ln3, act, h2));
      EpiParams e3 = epi_resid(x32, width, x32, width, (final_copy16 && n == nl - 1) ? final_copy16 : nullptr);
      if (lt && n + 1 < nl) set_ln_tail(e3, blocks[ch][n + 1].ln1, act, op);
      if (live) {
        long long o = off[n * nchain + ch];
        set_dropout(e3, dc, drop_p, &o, 0, call_base + static_cast<uint32_t>(n * 2 + ch));
      }
      g3.p[ch] = make_problem(c, w, bw.pff2, h2, ch, 1, false, e3);
    }
    CKI(launch_gemm(c, w, plan, g1));
    CKI(launch_gemm(c, w, plan, g2));
    CKI(launch_gemm(c, w, plan, g3));
  }
  return HUDIFF_OK;
}

EpiParams epi_qkv(const hudiff_ctx* c, const AttW& aw, __half* out_hi, __half* out_lo, int ld16, int sec0, const int* row_pos,
                  const float2* ln_stats = nullptr, const float* colsum = nullptr, int sums_k = 0) {
  EpiParams eq;
  std::memset(&eq, 0, sizeof(eq));
  eq.epi = EPI_QKV_ROPE;
  eq.out16_hi = out_hi; eq.out16_lo = out_lo; eq.ld16 = ld16;
  eq.rope = aw.rope; eq.seq_len = c->L; eq.att_dim = c->A; eq.head_dim = c->A / c->cfg.nhead;
  eq.qscale = 1.0f / std::sqrt(static_cast<float>(c->A) / static_cast<float>(c->cfg.nhead));
  eq.sec0 = sec0; eq.row_pos = row_pos;
  eq.ln_stats = ln_stats; eq.colsum = ln_stats ? colsum : nullptr; eq.ln_sums_k = ln_stats ? sums_k : 0;
  return eq;
}

// fold_in: statistics of the rows this attention's q|k|v projection consumes (folded LayerNorm); sums_out: the
// out-projection accumulates (sum, sum of squares) of the rows it writes for the NEXT folded consumer (HUDIFF_LNFOLD=2)
int plan_attention(hudiff_ctx* c, Workspace* w, Plan& plan, const AttW& aw, const float* resid, float* out32,
                   OperandBuf* copy16, const PackedLinear* qkv_folded = nullptr, const LNW* ln_after = nullptr,
                   const float2* fold_in = nullptr, float* sums_out = nullptr) {
  GemmParams g;
  std::memset(&g, 0, sizeof(g));
  g.nprob = 1;
  // qkv_folded: opD holds the RAW rows and w->lnstat their statistics (LayerNorm folded into the projection)
  const PackedLinear& Wqkv = qkv_folded ? *qkv_folded : aw.qkv;
  g.p[0] = make_problem(c, w, Wqkv, w->opD, VIEW_FLAT, 1, false,
                        epi_qkv(c, aw, w->qkv.hi, w->qkv.lo, w->qkv.ld, 0, nullptr,
                                qkv_folded ? (fold_in ? fold_in : w->lnstat) : nullptr, Wqkv.colsum, fold_in ? c->D : 0));
  CKI(launch_gemm(c, w, plan, g));
  AttnParams ap;
  ap.qkv_hi = w->qkv.hi; ap.qkv_lo = w->qkv.lo; ap.ld_qkv = w->qkv.ld;
  ap.o_hi = w->o16.hi; ap.o_lo = w->o16.lo; ap.ld_o = w->o16.ld;
  ap.L = c->L; ap.att_dim = c->A; ap.nhead = c->cfg.nhead; ap.B = w->B;
  const bool x3 = c->x3;
  // QK^T and PV: 2 * (2 * L * L * dh) FLOPs per (sequence, head)
  const double afl = 4.0 * static_cast<double>(w->B) * c->cfg.nhead * c->L * c->L * (c->A / c->cfg.nhead);
  if (!x3 && c->attn_impl != 1) {
    const CUtensorMap* mq = w->attn_maps; const CUtensorMap* mkv = w->attn_maps + 1;
    __half* o_hi = w->o16.hi; const int ld_o = w->o16.ld, L = c->L, A = c->A, nh = c->cfg.nhead, Bq = w->B;
    int* ef = w->err_flag;
    const int msm = c->max_sms, rev = plan.next_dir(), hint = c->l2hint, pdl = c->pdl;
    if (c->attn_impl == 2)
      plan.add([=](cudaStream_t s) { return attention_tc2_launch(mq, mkv, o_hi, ld_o, L, A, nh, Bq, ef, s, msm, rev, hint, pdl); }, OP_ATTN, afl, 1, "attention_tc2");
    else if (c->attn_impl != 4 || L < 17)
      plan.add([=](cudaStream_t s) { return attention_tc3_launch(mq, mkv, o_hi, ld_o, L, A, nh, Bq, ef, s, msm, rev, hint, pdl); }, OP_ATTN, afl, 1, "attention_tc3");
    else
      plan.add([=](cudaStream_t s) { return attention_tc4_launch(mq, mkv, o_hi, ld_o, L, A, nh, Bq, ef, s, msm, rev, hint, pdl); }, OP_ATTN, afl, 1, "attention_tc4");
  } else if (x3 && c->attn_impl != 1) {
    const CUtensorMap* maps = w->attn_maps;
    __half* o_hi = w->o16.hi; __half* o_lo = w->o16.lo;
    const int ld_o = w->o16.ld, L = c->L, A = c->A, nh = c->cfg.nhead, Bq = w->B, msm = c->max_sms, rev = plan.next_dir(), pdl = c->pdl;
    int* ef = w->err_flag;
    plan.add([=](cudaStream_t s) { return attention_x3_launch(maps, o_hi, o_lo, ld_o, L, A, nh, Bq, ef, s, msm, rev, pdl); }, OP_ATTN, afl, 1, "attention_x3");
  } else {
    plan.add([ap, x3](cudaStream_t s) { return attention_launch(ap, x3, s); }, OP_ATTN, afl, 1, "attention_mma");
  }
  GemmParams go;
  std::memset(&go, 0, sizeof(go));
  go.nprob = 1;
  EpiParams eo = epi_resid(out32, c->D, resid, c->D, copy16);
  eo.row_sums = sums_out;
  if (ln_after) set_ln_tail(eo, *ln_after, ACT_NONE, w->opD);      // LN(out32 row) -> opD for the next projection
  go.p[0] = make_problem(c, w, aw.out, w->o16, VIEW_FLAT, 1, false, eo);
  CKI(launch_gemm(c, w, plan, go));
  return HUDIFF_OK;
}

void add_ln_cast(Plan& plan, const float* x, int ld, int C, long long rows, const LNW& ln, OperandBuf& out, int l2hint = 0, int pdl = 0) {
  LnCastParams lp;
  std::memset(&lp, 0, sizeof(lp));
  lp.x = x; lp.ldx = ld; lp.C = C; lp.rows = rows;
  lp.gamma[0] = ln.g; lp.beta[0] = ln.b; lp.act = ACT_NONE;
  lp.out_hi = out.hi; lp.out_lo = out.lo; lp.ldo = out.ld;
  lp.reverse = plan.next_dir();
  lp.l2hint = l2hint;
  lp.pdl = pdl;
  plan.add([lp](cudaStream_t s) { return ln_cast_launch(lp, s); }, OP_ROWWISE, 0.0, 1, "ln_cast C" + std::to_string(C) + (rows < 4096 ? " tail" : ""));
}

void add_row_stats(Plan& plan, const OperandBuf& op, int C, long long rows, float2* out) {
  const __half* hi = op.hi; const __half* lo = op.lo; const int ld = op.ld;
  plan.add([=](cudaStream_t s) { return row_stats_launch(hi, lo, ld, C, rows, out, s); }, OP_ROWWISE, 0.0, 1, "row_stats C" + std::to_string(C));
}

// Second half of the LAST attention block when only one slot per sequence is consumed downstream
// (sampling step): keys/values for all rows, everything after them for the gathered row only.
//   a' = a + Att2(LN1(a))   x_out = FFN(LN2(a')) + x      (cross_attention.py:284-286), evaluated at row (b, order[b, step])
// Reference-equivalent FLOPs are unchanged in the accounting (skipped work counts).
int plan_tail(hudiff_ctx* c, Workspace* w, Plan& plan, const AttBlockW& bw, bool fold, const float2* sums_in = nullptr) {
  const float2* stats = sums_in ? sums_in : w->lnstat;          // (sum, sum of squares) when sums_in, else (mean, rstd)
  const int sums_k = sums_in ? c->D : 0;
  const RowSpace* rc = &w->rs_compact;
  // opD = fp16(LN1(a)) for all rows was produced by the caller (fold: the raw rows fp16(a) and w->lnstat)
  {
    TailGatherParams tp;
    std::memset(&tp, 0, sizeof(tp));
    tp.order = w->order; tp.step_ptr = w->step; tp.B = w->B; tp.L = c->L; tp.T = w->order_T;
    tp.row_pos = w->t_pos;
    tp.op_hi = w->opD.hi; tp.op_lo = w->opD.lo; tp.ld_op = w->opD.ld;
    tp.g_hi = w->tg16.hi; tp.g_lo = w->tg16.lo; tp.ld_g = w->tg16.ld;
    tp.a32 = w->A32; tp.x32 = w->X32; tp.ld32 = c->D; tp.ga32 = w->tA32; tp.gx32 = w->tX32; tp.D = c->D;
    if (fold) { tp.stats = stats; tp.gstats = w->t_lnstat; tp.stat_slots = sums_in ? LN_SUM_PARTS : 1; }
    plan.add([tp](cudaStream_t s) { return tail_gather_launch(tp, s); }, OP_ROWWISE, 0.0, 1, "tail_gather");
  }
  GemmParams gkv, gq;
  std::memset(&gkv, 0, sizeof(gkv)); std::memset(&gq, 0, sizeof(gq));
  gkv.nprob = gq.nprob = 1;
  // k | v for every row into columns [A, 3A) of the q|k|v planes
  const PackedLinear& Wkv = fold ? bw.a2_kv_f : bw.a2_kv;
  const PackedLinear& Wq = fold ? bw.a2_q_f : bw.a2_q;
  gkv.p[0] = make_problem(c, w, Wkv, w->opD, VIEW_FLAT, 1, false,
                          epi_qkv(c, bw.a2, w->qkv.hi + c->A, w->qkv.lo ? w->qkv.lo + c->A : nullptr, w->qkv.ld, 1, nullptr,
                                  fold ? stats : nullptr, Wkv.colsum, sums_k));
  CKI(launch_gemm(c, w, plan, gkv));
  // q for the gathered rows (rotary phase of the sampled slot)
  gq.p[0] = make_problem(c, w, Wq, w->tg16, VIEW_FLAT, 1, false,
                         epi_qkv(c, bw.a2, w->tq16.hi, w->tq16.lo, w->tq16.ld, 0, w->t_pos, fold ? w->t_lnstat : nullptr, Wq.colsum, sums_k), rc);
  CKI(launch_gemm(c, w, plan, gq));
  {
    AttnRowParams ar;
    std::memset(&ar, 0, sizeof(ar));
    ar.q_hi = w->tq16.hi; ar.q_lo = w->tq16.lo; ar.ld_q = w->tq16.ld;
    ar.kv_hi = w->qkv.hi; ar.kv_lo = w->qkv.lo; ar.ld_kv = w->qkv.ld;
    ar.o_hi = w->to16.hi; ar.o_lo = w->to16.lo; ar.ld_o = w->to16.ld;
    ar.B = w->B; ar.L = c->L; ar.att_dim = c->A; ar.nhead = c->cfg.nhead;
    plan.add([ar](cudaStream_t s) { return attn_row_launch(ar, s); }, OP_ATTN, 0.0, 1, "attention_row");
  }
  GemmParams go, g1, g2;
  std::memset(&go, 0, sizeof(go)); std::memset(&g1, 0, sizeof(g1)); std::memset(&g2, 0, sizeof(g2));
  go.nprob = g1.nprob = g2.nprob = 1;
  EpiParams eo = epi_resid(w->tA32, c->D, w->tA32, c->D, nullptr);
  if (c->lntail & 1) set_ln_tail(eo, bw.n2, ACT_NONE, w->tg16);
  go.p[0] = make_problem(c, w, bw.a2.out, w->to16, VIEW_FLAT, 1, false, eo, rc);
  CKI(launch_gemm(c, w, plan, go));
  if (!(c->lntail & 1)) add_ln_cast(plan, w->tA32, c->D, c->D, w->B, bw.n2, w->tg16);
  EpiParams e1;
  std::memset(&e1, 0, sizeof(e1));
  e1.epi = EPI_ACT_F16; e1.act = ACT_RELU;
  e1.out16_hi = w->tf16.hi; e1.out16_lo = w->tf16.lo; e1.ld16 = w->tf16.ld;
  g1.p[0] = make_problem(c, w, bw.ff1, w->tg16, VIEW_FLAT, 1, false, e1, rc);
  CKI(launch_gemm(c, w, plan, g1));
  g2.p[0] = make_problem(c, w, bw.ff2, w->tf16, VIEW_FLAT, 1, false, epi_resid(w->tX32, c->D, w->tX32, c->D, nullptr), rc);
  CKI(launch_gemm(c, w, plan, g2));
  return HUDIFF_OK;
}

void plan_ln(hudiff_ctx* c, Workspace* w, Plan& plan, const float* x, const LNW& ln, OperandBuf& out) {
  add_ln_cast(plan, x, c->D, c->D, w->M, ln, out, c->l2hint, c->pdl);
}

// Everything from tokens to the final residual stream X32 (fp32 [M, D]).
int plan_trunk(hudiff_ctx* c, Workspace* w, Plan& plan, const DropCfg& dc, bool tail = false) {
  const bool ab = c->cfg.kind == HUDIFF_AB;
  long long word_off = 0;
  {
    const int* tok = w->tokens; const float* emb = c->emb; float* e32 = w->E32;
    const long long M = w->M; const int d = c->d, nt = c->cfg.n_tokens;
    const RowLnOut ln0 = first_block_ln(c, c->enc, c->cfg.activation, w->opE);
    plan.add([=](cudaStream_t s) { return embed_launch(tok, emb, e32, M, d, nt, ln0, s); }, OP_ROWWISE, 0.0, 1, "embed");
  }
  CKI(plan_stack(c, w, plan, c->enc, w->E32, c->d, w->opE, w->hE1, w->hE2, c->cfg.activation, c->cfg.dropout, dc,
                 &word_off, true, 0u, nullptr));
  {
    AssembleParams ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.e = w->E32; ap.region = w->region; ap.chain = ab ? w->chain : nullptr;
    ap.pos_table = c->pos_table; ap.side_table = ab ? c->side_table : nullptr;
    ap.x = w->X32; ap.ldx = c->D; ap.rows = w->M; ap.L = c->L; ap.B = w->B; ap.d = c->d; ap.split = c->split;
    ap.n_region = c->cfg.n_region; ap.n_side = c->cfg.n_side;
    ap.reverse = plan.next_dir();
    ap.ln = first_block_ln(c, c->dual, ab ? ACT_RELU : ACT_GELU, w->opD);
    plan.add([ap](cudaStream_t s) { return assemble_launch(ap, s); }, OP_ROWWISE, 0.0, 1, "assemble");
  }
  // DualConv keeps ByteNetBlock's default ReLU (model.py:345,252); NanoConv defaults to GELU
  // (nanoencoder/model.py:242).  Both use F.dropout's default p = 0.5.
  CKI(plan_stack(c, w, plan, c->dual, w->X32, c->D, w->opD, w->hD1, w->hD2, ab ? ACT_RELU : ACT_GELU, 0.5, dc,
                 &word_off, false, 1000u, &w->opD));
  const bool sums = c->lnfold == 2 && w->lnsums != nullptr;
  if (sums && c->gemm_impl == 1) {
    // the tcgen05 epilogues STORE their partial sums (one slot per N tile and epilogue warp); only the CUDA-core
    // validation GEMMs accumulate atomically and need the buffer cleared
    float* p = w->lnsums; const size_t nb = w->lnsums_floats * sizeof(float);
    plan.add([p, nb](cudaStream_t s) { return cudaMemsetAsync(p, 0, nb, s); }, OP_ROWWISE, 0.0, 1, "zero row sums");
  }
  for (size_t n = 0; n < c->att.size(); ++n) {
    const AttBlockW& ab_w = c->att[n];
    float* sums1 = sums ? w->lnsums + (2 * n) * w->M * LN_SUM_PARTS * 2 : nullptr;        // rows of `a` after the first attention  -> norm_hl1
    float* sums2 = sums ? w->lnsums + (2 * n + 1) * w->M * LN_SUM_PARTS * 2 : nullptr;    // rows of `a` after the second attention -> norm_hl2
    const float2* st1 = sums ? reinterpret_cast<const float2*>(sums1) : nullptr;
    const float2* st2 = sums ? reinterpret_cast<const float2*>(sums2) : nullptr;
    // Folded LayerNorm: the out-projections also emit the raw fp16 rows of `a`, a light pass reduces
    // them to (mean, rstd) per row, and norm_hl1 / norm_hl2 are applied inside the epilogues of the
    // projections that consume them (no fp32 re-read of the stream, no normalised copy in HBM).
    const bool fold = c->lnfold && ab_w.has_fold;
    // a = x + Att1(x)                         (cross_attention.py:282)
    const bool lt = (c->lntail & 1) && !fold;    // LN1 / LN2 as tails of the out-projections
    CKI(plan_attention(c, w, plan, ab_w.a1, w->X32, w->A32, fold ? &w->opD : nullptr, nullptr, lt ? &ab_w.n1 : nullptr,
                       nullptr, fold ? sums1 : nullptr));
    // a = a + Att2(LN1(a))                    (:284)
    if (fold && !sums) add_row_stats(plan, w->opD, c->D, w->M, w->lnstat);
    else if (!fold && !lt) plan_ln(c, w, plan, w->A32, ab_w.n1, w->opD);
    if (tail && n + 1 == c->att.size() && ab_w.has_split) {
      CKI(plan_tail(c, w, plan, ab_w, fold, fold ? st1 : nullptr));    // rest of the block for the sampled slot of every sequence only
      break;
    }
    CKI(plan_attention(c, w, plan, ab_w.a2, w->A32, w->A32, fold ? &w->opD : nullptr, fold ? &ab_w.a2_qkv_f : nullptr,
                       lt ? &ab_w.n2 : nullptr, fold ? st1 : nullptr, fold ? sums2 : nullptr));
    // x = FFN(LN2(a)) + x                     (:286)  residual is the block input
    if (fold && !sums) add_row_stats(plan, w->opD, c->D, w->M, w->lnstat);
    else if (!fold && !lt) plan_ln(c, w, plan, w->A32, ab_w.n2, w->opD);
    GemmParams g1, g2;
    std::memset(&g1, 0, sizeof(g1)); std::memset(&g2, 0, sizeof(g2));
    g1.nprob = g2.nprob = 1;
    EpiParams e1;
    std::memset(&e1, 0, sizeof(e1));
    e1.epi = EPI_ACT_F16; e1.act = ACT_RELU;
    e1.out16_hi = w->f16.hi; e1.out16_lo = w->f16.lo; e1.ld16 = w->f16.ld;
    if (fold) { e1.ln_stats = sums ? st2 : w->lnstat; e1.colsum = ab_w.ff1_f.colsum; e1.ln_sums_k = sums ? c->D : 0; }
    g1.p[0] = make_problem(c, w, fold ? ab_w.ff1_f : ab_w.ff1, w->opD, VIEW_FLAT, 1, false, e1);
    CKI(launch_gemm(c, w, plan, g1));
    const bool last = (n + 1 == c->att.size());
    g2.p[0] = make_problem(c, w, ab_w.ff2, w->f16, VIEW_FLAT, 1, false,
                           epi_resid(w->X32, c->D, w->X32, c->D, last ? nullptr : &w->opD));
    CKI(launch_gemm(c, w, plan, g2));
  }
  return HUDIFF_OK;
}

HeadParams head_common(hudiff_ctx* c, Workspace* w) {
  HeadParams hp;
  std::memset(&hp, 0, sizeof(hp));
  hp.x = w->X32; hp.ldx = c->D; hp.C = c->D; hp.rows = w->M;
  hp.gamma = c->last_norm.g; hp.beta = c->last_norm.b; hp.w = c->dec_w; hp.bias = c->dec_b;
  hp.n_out = c->cfg.n_tokens; hp.B = w->B; hp.L = c->L;
  hp.rng_ptr = w->rng; hp.err_flag = w->err_flag;
  return hp;
}

int run_plan(const Plan& plan, cudaStream_t s) {
  for (const auto& op : plan.ops) CK(op(s));
  return HUDIFF_OK;
}

int check_err_flag(Workspace* w, cudaStream_t s) {
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, w->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (flag >= 401 && flag <= 404) {
    static const char* what[] = {"order (sampling slot)", "tokens", "region", "chain"};
    return fail(HUDIFF_ERR_INDEX, std::string("index out of range in ") + what[flag - 401] + " (code " + std::to_string(flag) + ")");
  }
  if (flag != 0) return fail(HUDIFF_ERR_CUDA, "device pipeline watchdog fired, code " + std::to_string(flag));
  return HUDIFF_OK;
}

DropCfg drop_from_rng(const hudiff_ctx* c, const hudiff_rng* rng, int B) {
  DropCfg dc;
  if (!rng || !rng->dropout || !(c->cfg.dropout > 0.0)) return dc;
  if (rng->mode == 1) {
    dc.mode = DROP_BITS; dc.bits = rng->dropout_bits; dc.step_words = dropout_words(c, B);
  } else {
    dc.mode = DROP_PHILOX;
  }
  return dc;
}

}  // namespace

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" {

const char* hudiff_last_error(void) { return g_err.c_str(); }
const char* hudiff_version(void) { return "hudiff_b200 0.1.0 (sm_100a)"; }

int hudiff_create(hudiff_handle* out, const hudiff_cfg* cfg, int device) {
  if (!out || !cfg) return fail(HUDIFF_ERR_INVALID, "null argument");
  const hudiff_cfg& k = *cfg;
  if (k.kind != HUDIFF_AB && k.kind != HUDIFF_NB) return fail(HUDIFF_ERR_INVALID, "cfg.kind must be 0 (Ab) or 1 (Nb)");
  const int nparts = (k.kind == HUDIFF_AB) ? 3 : 2;
  if (k.sum_d_model != nparts * k.d_model) return fail(HUDIFF_ERR_INVALID, "sum_d_model must be d_model * (3 for Ab, 2 for Nb)");
  if (k.d_model % 128 != 0 || k.sum_d_model % 128 != 0 || k.sum_d_model > 1024)
    return fail(HUDIFF_ERR_INVALID, "d_model and sum_d_model must be multiples of 128 (sum_d_model <= 1024)");
  if (k.nhead <= 0 || k.att_model % k.nhead != 0 || k.att_model / k.nhead != 64)
    return fail(HUDIFF_ERR_INVALID, "attention head dim must be 64 (att_model / nhead)");
  if (k.att_model % 256 != 0) return fail(HUDIFF_ERR_INVALID, "att_model must be a multiple of 256 (q|k|v column blocks hold whole N tiles)");
  if (k.aa_kernel_size % 2 != 1 || k.aa_kernel_size < 1) return fail(HUDIFF_ERR_INVALID, "aa_kernel_size must be odd");
  if (k.n_tokens < 2 || k.n_tokens > 32) return fail(HUDIFF_ERR_INVALID, "n_tokens must be in [2, 32]");
  if (k.kind == HUDIFF_AB && k.max_len != 291) return fail(HUDIFF_ERR_INVALID, "HuDiff-Ab is hard-wired to max_len 291 = 152 + 139 (model/encoder/model.py:185)");
  if (k.max_len < 128 || k.max_len > 320) return fail(HUDIFF_ERR_INVALID, "max_len must be in [128, 320]");
  if (k.dim_feedforward % 64 != 0 || (k.d_model / 2) % 64 != 0 || (k.sum_d_model / 2) % 64 != 0)
    return fail(HUDIFF_ERR_INVALID, "hidden widths must be multiples of 64");
  if (k.activation != HUDIFF_ACT_RELU && k.activation != HUDIFF_ACT_GELU) return fail(HUDIFF_ERR_INVALID, "activation must be relu or gelu");
  if (!(k.dropout >= 0.0 && k.dropout < 1.0)) return fail(HUDIFF_ERR_INVALID, "dropout must be in [0, 1)");
  CK(cudaSetDevice(device));
  int major = 0, minor = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) return fail(HUDIFF_ERR_INVALID, "libhudiff_b200 needs a Blackwell sm_100 device, found sm_" +
                                                       std::to_string(major) + std::to_string(minor));
  CK(kernels_init());
  hudiff_ctx* c = new hudiff_ctx();
  c->cfg = k; c->device = device; c->x3 = (k.precision == HUDIFF_FP16X3);
  c->L = k.max_len; c->split = (k.kind == HUDIFF_AB) ? 152 : k.max_len;
  c->d = k.d_model; c->dh = k.d_model / 2; c->D = k.sum_d_model; c->Dh = k.sum_d_model / 2;
  c->A = k.att_model; c->F = k.dim_feedforward;
  const int lg = static_cast<int>(std::log2(static_cast<double>(k.r))) + 1;
  for (int n = 0; n < k.n_encoder_layers; ++n) c->dil_enc.push_back(1 << (n % lg));
  for (int n = 0; n < k.dual_layers; ++n) c->dil_dual.push_back(1 << (n % lg));
  CK(cudaStreamCreateWithFlags(&c->work, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->io, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
  if (const char* e2 = std::getenv("HUDIFF_CLUSTER")) {
    const int v = std::atoi(e2);
    if (v == 1 || v == 2 || v == 4) c->cs = v;
  }
  if (const char* e4 = std::getenv("HUDIFF_2CTA")) c->two_cta = std::atoi(e4) != 0 ? 1 : 0;
  if (c->cs != 2) c->two_cta = 0;
  if (const char* e8 = std::getenv("HUDIFF_MAX_SMS")) c->max_sms = std::max(0, std::atoi(e8));
  if (const char* e6 = std::getenv("HUDIFF_NSPLIT")) c->nsplit = std::atoi(e6) != 0 ? 1 : 0;
  if (const char* e7 = std::getenv("HUDIFF_LNFOLD")) { const int v = std::atoi(e7); c->lnfold = (v == 1 || v == 2) ? v : 0; }
  if (const char* e5 = std::getenv("HUDIFF_TAIL")) c->tail = std::atoi(e5) != 0 ? 1 : 0;
  // bit mask: 1 = attention blocks, 2 = wide ByteNet stack, 4 = narrow ByteNet stack; "1" keeps its old meaning (all)
  if (const char* e9 = std::getenv("HUDIFF_LNTAIL")) { const int v = std::atoi(e9); c->lntail = v == 1 ? 7 : (v & 7); }
  if (const char* e10 = std::getenv("HUDIFF_L2FLIP")) c->l2flip = std::atoi(e10) != 0 ? 1 : 0;
  if (const char* e11 = std::getenv("HUDIFF_FUSE0")) c->fuse0 = std::atoi(e11) != 0 ? 1 : 0;
  if (const char* e12 = std::getenv("HUDIFF_L2HINT")) c->l2hint = std::atoi(e12) != 0 ? 1 : 0;
  if (const char* e13 = std::getenv("HUDIFF_PDL")) { const int v = std::atoi(e13); c->pdl = (v == 1 || v == 2) ? v : 0; }
  if (const char* e3 = std::getenv("HUDIFF_ATTN")) c->attn_impl = (std::string(e3) == "mma") ? 1 : (std::string(e3) == "tc2") ? 2 : (std::string(e3) == "tc4") ? 4 : 0;
  const char* env = std::getenv("HUDIFF_GEMM");
  if (env && std::string(env) == "simt") c->gemm_impl = 1;
  *out = c;
  return HUDIFF_OK;
}

int hudiff_set_gemm_impl(hudiff_handle c, int impl) {
  if (!c || (impl != 0 && impl != 1)) return fail(HUDIFF_ERR_INVALID, "bad argument");
  c->gemm_impl = impl;
  if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear(); }
  return HUDIFF_OK;
}

int hudiff_load_weights(hudiff_handle c, const hudiff_tensor* tensors, int n) {
  if (!c || !tensors || n <= 0) return fail(HUDIFF_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  if (c->loaded) return fail(HUDIFF_ERR_STATE, "weights already loaded; create a new handle");
  TensorTable tt;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].name || !tensors[i].data) return fail(HUDIFF_ERR_INVALID, "tensor entry with null name/data");
    std::string name = tensors[i].name;
    if (name.rfind("module.", 0) == 0) name = name.substr(7);      // nn.DataParallel prefix (antibody_train.py:23-30)
    tt.t[name] = tensors[i];
  }
  const hudiff_cfg& k = c->cfg;
  const bool ab = k.kind == HUDIFF_AB;
  CKI(load_f32(c, tt, "aa_encoder.embedder.weight", &c->emb, static_cast<long long>(k.n_tokens) * c->d));
  const char* enc_names[2] = {ab ? "aa_encoder.h_layers." : "aa_encoder.layers.", "aa_encoder.l_layers."};
  const char* dual_names[2] = {ab ? "dual_conv_block.h_layers." : "nano_conv_block.layers.", "dual_conv_block.l_layers."};
  for (int ch = 0; ch < (ab ? 2 : 1); ++ch) {
    c->enc[ch].resize(k.n_encoder_layers);
    for (int i = 0; i < k.n_encoder_layers; ++i)
      CKI(load_block(c, tt, enc_names[ch] + std::to_string(i), c->d, c->dh, k.aa_kernel_size, c->dil_enc[i], &c->enc[ch][i]));
    c->dual[ch].resize(k.dual_layers);
    for (int i = 0; i < k.dual_layers; ++i)
      CKI(load_block(c, tt, dual_names[ch] + std::to_string(i), c->D, c->Dh, k.aa_kernel_size, c->dil_dual[i], &c->dual[ch][i]));
  }
  c->att.resize(k.cs_layers);
  for (int i = 0; i < k.cs_layers; ++i) {
    const std::string p = "self_at.layers." + std::to_string(i);
    CKI(load_att(c, tt, p + ".attn_hl", &c->att[i].a1));
    CKI(load_att(c, tt, p + ".attn_hl_c", &c->att[i].a2));
    CKI(load_ln(c, tt, p + ".norm_hl1", c->D, &c->att[i].n1));
    CKI(load_ln(c, tt, p + ".norm_hl2", c->D, &c->att[i].n2));
    CKI(pack_linear(c, tt, {p + ".ff_hl.0.weight"}, {p + ".ff_hl.0.bias"}, c->F, c->D, 1, false, &c->att[i].ff1));
    CKI(pack_linear(c, tt, {p + ".ff_hl.2.weight"}, {p + ".ff_hl.2.bias"}, c->D, c->F, 1, false, &c->att[i].ff2));
    if (c->lnfold && c->D % 256 == 0) {
      const std::string a = p + ".attn_hl_c";
      CKI(pack_linear(c, tt, {a + ".query.weight", a + ".key.weight", a + ".value.weight"},
                      {a + ".query.bias", a + ".key.bias", a + ".value.bias"}, c->A, c->D, 1, false, &c->att[i].a2_qkv_f, p + ".norm_hl1"));
      CKI(pack_linear(c, tt, {p + ".ff_hl.0.weight"}, {p + ".ff_hl.0.bias"}, c->F, c->D, 1, false, &c->att[i].ff1_f, p + ".norm_hl2"));
      c->att[i].has_fold = true;
    }
  }
  if (k.cs_layers > 0) {
    // sampling shortcut (plan_tail): the second attention of the last block needs q for one slot per sequence only
    AttBlockW& lb = c->att[k.cs_layers - 1];
    const std::string p = "self_at.layers." + std::to_string(k.cs_layers - 1) + ".attn_hl_c";
    CKI(pack_linear(c, tt, {p + ".query.weight"}, {p + ".query.bias"}, c->A, c->D, 1, false, &lb.a2_q));
    CKI(pack_linear(c, tt, {p + ".key.weight", p + ".value.weight"}, {p + ".key.bias", p + ".value.bias"}, c->A, c->D, 1, false, &lb.a2_kv));
    lb.has_split = true;
    if (lb.has_fold) {
      const std::string ln = "self_at.layers." + std::to_string(k.cs_layers - 1) + ".norm_hl1";
      CKI(pack_linear(c, tt, {p + ".query.weight"}, {p + ".query.bias"}, c->A, c->D, 1, false, &lb.a2_q_f, ln));
      CKI(pack_linear(c, tt, {p + ".key.weight", p + ".value.weight"}, {p + ".key.bias", p + ".value.bias"}, c->A, c->D, 1, false, &lb.a2_kv_f, ln));
    }
  }
  CKI(load_ln(c, tt, "last_norm", c->D, &c->last_norm));
  CKI(load_f32(c, tt, "decoder.weight", &c->dec_w, static_cast<long long>(k.n_tokens) * c->D));
  CKI(load_f32(c, tt, "decoder.bias", &c->dec_b, k.n_tokens));
  CKI(fold_constants(c, tt));
  c->loaded = true;
  return HUDIFF_OK;
}

size_t hudiff_workspace_bytes(hudiff_handle c, int B) {
  if (!c || B <= 0) return 0;
  auto it = c->ws.find(B);
  if (it != c->ws.end()) return it->second->bytes;
  return workspace_estimate(c, B);
}

size_t hudiff_dropout_words(hudiff_handle c, int B) {
  if (!c || B <= 0) return 0;
  return static_cast<size_t>(dropout_words(c, B));
}

long long hudiff_last_launch_count(hudiff_handle c) { return c ? c->last_launches : 0; }

// Synchronises the library's launch stream and reports what the device flagged during the last
// hudiff_forward / hudiff_sample on this handle: an out-of-range integer input (HUDIFF_ERR_INDEX) or a
// pipeline watchdog (HUDIFF_ERR_CUDA).  hudiff_sample_host performs this check itself.
int hudiff_check_error(hudiff_handle c) {
  if (!c) return fail(HUDIFF_ERR_INVALID, "null handle");
  CK(cudaSetDevice(c->device));
  if (!c->last_ws) return HUDIFF_OK;
  return check_err_flag(static_cast<Workspace*>(c->last_ws), c->work);
}

int hudiff_forward(hudiff_handle c, const int32_t* tokens, const int32_t* region, const int32_t* chain, int B,
                   const hudiff_rng* rng, float* logits_out, hudiff_stream stream) {
  if (!c || !tokens || !region || !logits_out || B <= 0) return fail(HUDIFF_ERR_INVALID, "bad argument");
  if (!c->loaded) return fail(HUDIFF_ERR_STATE, "hudiff_load_weights has not been called");
  if (c->cfg.kind == HUDIFF_AB && !chain) return fail(HUDIFF_ERR_INVALID, "HuDiff-Ab needs the chain-type vector [2B]");
  if (rng && rng->mode == 1 && rng->dropout && c->cfg.dropout > 0.0 && !rng->dropout_bits)
    return fail(HUDIFF_ERR_INVALID, "replay mode needs dropout_bits");
  CK(cudaSetDevice(c->device));
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  cudaStream_t s = c->work;
  CK(cudaEventRecord(c->ev_in, user));
  CK(cudaStreamWaitEvent(s, c->ev_in, 0));
  Workspace* w = nullptr;
  CKI(get_workspace(c, B, &w));
  CK(cudaMemcpyAsync(w->tokens, tokens, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(w->region, region, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (chain) CK(cudaMemcpyAsync(w->chain, chain, 2 * B * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(w->step, 0, sizeof(int), s));
  CK(cudaMemsetAsync(w->err_flag, 0, sizeof(int), s));
  CK(set_rng_launch(w->rng, rng ? rng->seed : 0, rng ? rng->seq_offset : 0, s));
  CK(validate_launch(w->tokens, w->region, chain ? w->chain : nullptr, nullptr, w->M, B, 0, c->L, c->cfg.n_tokens,
                     c->cfg.n_region, c->cfg.n_side, w->err_flag, s));
  c->last_ws = w;
  Plan plan;
  plan.flip = c->l2flip;
  CKI(plan_trunk(c, w, plan, drop_from_rng(c, rng, B)));
  HeadParams hp = head_common(c, w);
  hp.mode = 0; hp.logits = logits_out;
  plan.add([hp](cudaStream_t st) { return head_launch(hp, st); }, OP_HEAD);
  CKI(run_plan(plan, s));
  c->last_launches = plan.launches_per_run;
  CK(cudaEventRecord(c->ev_out, s));
  CK(cudaStreamWaitEvent(user, c->ev_out, 0));
  return HUDIFF_OK;
}

int hudiff_sample(hudiff_handle c, int32_t* tokens_inout, const int32_t* region, const int32_t* chain,
                  const int32_t* order, int B, int T, const hudiff_rng* rng, float* sel_logits_out,
                  hudiff_stream stream) {
  if (!c || !tokens_inout || !region || !order || B <= 0 || T <= 0) return fail(HUDIFF_ERR_INVALID, "bad argument");
  if (!c->loaded) return fail(HUDIFF_ERR_STATE, "hudiff_load_weights has not been called");
  if (c->cfg.kind == HUDIFF_AB && !chain) return fail(HUDIFF_ERR_INVALID, "HuDiff-Ab needs the chain-type vector [2B]");
  if (rng && rng->mode == 1) {
    if (!rng->exp_draws) return fail(HUDIFF_ERR_INVALID, "replay mode needs exp_draws");
    if (rng->dropout && c->cfg.dropout > 0.0 && !rng->dropout_bits) return fail(HUDIFF_ERR_INVALID, "replay mode needs dropout_bits");
  }
  CK(cudaSetDevice(c->device));
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  cudaStream_t s = c->work;
  CK(cudaEventRecord(c->ev_in, user));
  CK(cudaStreamWaitEvent(s, c->ev_in, 0));
  Workspace* w = nullptr;
  CKI(get_workspace(c, B, &w));
  if (w->order_cap < B * T) {
    CKI(dev_alloc(w->allocs, &w->order, static_cast<size_t>(B) * T, &w->bytes, false));
    w->order_cap = B * T;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear(); }
  }
  CK(cudaMemcpyAsync(w->tokens, tokens_inout, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(w->region, region, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (chain) CK(cudaMemcpyAsync(w->chain, chain, 2 * B * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(w->order, order, static_cast<size_t>(B) * T * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(w->step, 0, sizeof(int), s));
  CK(cudaMemsetAsync(w->err_flag, 0, sizeof(int), s));
  // the philox key and the shard's sequence offset live in device memory: the cached step graph serves every seed
  CK(set_rng_launch(w->rng, rng ? rng->seed : 0, rng ? rng->seq_offset : 0, s));
  CK(validate_launch(w->tokens, w->region, chain ? w->chain : nullptr, w->order, w->M, B, T, c->L, c->cfg.n_tokens,
                     c->cfg.n_region, c->cfg.n_side, w->err_flag, s));
  c->last_ws = w;

  const DropCfg dc = drop_from_rng(c, rng, B);
  w->order_T = T;
  const bool use_tail = c->tail && !c->att.empty() && c->att.back().has_split && c->cfg.nhead <= 8 && c->L <= 320;
  HeadParams hp = head_common(c, w);
  if (use_tail) { hp.x = w->tX32; hp.compact = 1; }               // the last block leaves one row per sequence
  hp.mode = 1; hp.T = T; hp.n_sample = c->cfg.n_tokens - 1;     // sample.py:510  [:len(toks)-1]
  hp.order = w->order; hp.step_ptr = w->step; hp.tokens = w->tokens; hp.sel_logits = sel_logits_out;
  hp.exp_draws = (rng && rng->mode == 1) ? rng->exp_draws : nullptr;

  // The step schedule only depends on these values -> reuse the instantiated graph when they repeat.
  char key[256];
  std::snprintf(key, sizeof(key), "B%d T%d g%d t%d dm%d bits%p exp%p sel%p x3%d", B, T, c->gemm_impl, use_tail ? 1 : 0, dc.mode,
                static_cast<const void*>(dc.bits), static_cast<const void*>(hp.exp_draws), static_cast<void*>(sel_logits_out),
                c->x3 ? 1 : 0);
  long long per_step = 0;
  const char* nograph = std::getenv("HUDIFF_NO_GRAPH");
  const bool use_graph = !(nograph && nograph[0] == '1');
  if (!use_graph || !c->graph_exec || c->graph_key != key) {
    Plan plan;
    plan.flip = c->l2flip;
    CKI(plan_trunk(c, w, plan, dc, use_tail));
    plan.add([hp](cudaStream_t st) { return head_launch(hp, st); }, OP_HEAD);
    int* stp = w->step;
    plan.add([stp](cudaStream_t st) { return advance_step_launch(stp, st); }, OP_HEAD);
    per_step = plan.launches_per_run;
    if (!use_graph) {
      for (int t = 0; t < T; ++t) CKI(run_plan(plan, s));
    } else {
      if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear(); }
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rc = run_plan(plan, s);
      cudaError_t e = cudaStreamEndCapture(s, &graph);
      if (rc != HUDIFF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return fail(HUDIFF_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
      e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { c->graph_exec = nullptr; return fail(HUDIFF_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
      c->graph_key = key;
      c->graph_per_step = per_step;
    }
  } else {
    per_step = c->graph_per_step;
  }
  if (use_graph) {
    for (int t = 0; t < T; ++t) CK(cudaGraphLaunch(c->graph_exec, s));
  }
  c->last_launches = per_step * T;
  CK(cudaMemcpyAsync(tokens_inout, w->tokens, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaEventRecord(c->ev_out, s));
  CK(cudaStreamWaitEvent(user, c->ev_out, 0));
  return HUDIFF_OK;
}

int hudiff_sample_host(hudiff_handle c, int32_t* tokens_inout, const int32_t* region, const int32_t* chain,
                       const int32_t* order, int B, int T, const hudiff_rng* rng, float* sel_logits_out) {
  if (!c || !tokens_inout || !region || !order || B <= 0 || T <= 0) return fail(HUDIFF_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  const size_t M = static_cast<size_t>(B) * c->L;
  Workspace* w = nullptr;
  CKI(get_workspace(c, B, &w));
  // device staging of the host inputs lives in the workspace (allocated once per batch size / T)
  if (!w->s_tok) {
    CKI(dev_alloc(w->allocs, &w->s_tok, M, &w->bytes, false));
    CKI(dev_alloc(w->allocs, &w->s_reg, M, &w->bytes, false));
    CKI(dev_alloc(w->allocs, &w->s_chain, 2 * static_cast<size_t>(B), &w->bytes, false));
  }
  if (w->s_order_cap < B * T) {
    CKI(dev_alloc(w->allocs, &w->s_order, static_cast<size_t>(B) * T, &w->bytes, false));
    w->s_order_cap = B * T;
  }
  std::vector<void*> tmp;                     // replay-mode streams and logits are per call (sizes depend on T)
  auto cleanup = [&]() { for (void* p : tmp) cudaFree(p); };
  int *d_tok = w->s_tok, *d_reg = w->s_reg, *d_chain = chain ? w->s_chain : nullptr, *d_order = w->s_order;
  uint32_t* d_bits = nullptr; float* d_exp = nullptr; float* d_sel = nullptr;
  int rc = HUDIFF_OK;
  cudaStream_t s = c->io;
  do {
    cudaError_t e = cudaMemcpyAsync(d_tok, tokens_inout, M * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_reg, region, M * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_order, order, static_cast<size_t>(B) * T * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && chain) e = cudaMemcpyAsync(d_chain, chain, 2 * static_cast<size_t>(B) * sizeof(int), cudaMemcpyHostToDevice, s);
    hudiff_rng r2;
    std::memset(&r2, 0, sizeof(r2));
    if (rng) r2 = *rng;
    if (e == cudaSuccess && rng && rng->mode == 1) {
      const size_t words = static_cast<size_t>(dropout_words(c, B)) * T;
      if (rng->dropout && words > 0 && rng->dropout_bits) {
        if ((rc = dev_alloc(tmp, &d_bits, words, nullptr, false))) break;
        e = cudaMemcpyAsync(d_bits, rng->dropout_bits, words * 4, cudaMemcpyHostToDevice, s);
        r2.dropout_bits = d_bits;
      }
      if (e == cudaSuccess && rng->exp_draws) {
        const size_t ne = static_cast<size_t>(T) * B * (c->cfg.n_tokens - 1);
        if ((rc = dev_alloc(tmp, &d_exp, ne, nullptr, false))) break;
        e = cudaMemcpyAsync(d_exp, rng->exp_draws, ne * 4, cudaMemcpyHostToDevice, s);
        r2.exp_draws = d_exp;
      }
    }
    if (e != cudaSuccess) { rc = fail(HUDIFF_ERR_CUDA, std::string("host->device copy: ") + cudaGetErrorString(e)); break; }
    const size_t nsel = static_cast<size_t>(T) * B * c->cfg.n_tokens;
    if (sel_logits_out && (rc = dev_alloc(tmp, &d_sel, nsel, nullptr, false))) break;
    if ((rc = hudiff_sample(c, d_tok, d_reg, d_chain, d_order, B, T, &r2, d_sel, s))) break;
    e = cudaMemcpyAsync(tokens_inout, d_tok, M * sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && sel_logits_out) e = cudaMemcpyAsync(sel_logits_out, d_sel, nsel * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { rc = fail(HUDIFF_ERR_CUDA, std::string("sample/device->host: ") + cudaGetErrorString(e)); break; }
    rc = check_err_flag(w, s);
  } while (0);
  // a graph that baked in this call's temporary stream/logit buffers must go with them
  if ((d_bits || d_exp || d_sel) && c->graph_exec) {
    cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key.clear();
  }
  cleanup();
  return rc;
}

// One forward of the sampling step, run eagerly with a CUDA-event pair around every launch on the
// stream the kernels run on.  Used by bench.py for the roofline of the dominant kernel.
int hudiff_profile_step(hudiff_handle c, const int32_t* tokens, const int32_t* region, const int32_t* chain, int B,
                        const hudiff_rng* rng, double* ms_by_kind, long long* launches_by_kind, double* flops_by_kind) {
  if (!c || !tokens || !region || B <= 0 || !ms_by_kind || !launches_by_kind || !flops_by_kind)
    return fail(HUDIFF_ERR_INVALID, "bad argument");
  if (!c->loaded) return fail(HUDIFF_ERR_STATE, "hudiff_load_weights has not been called");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  cudaStream_t s = c->work;
  Workspace* w = nullptr;
  CKI(get_workspace(c, B, &w));
  CK(cudaMemcpyAsync(w->tokens, tokens, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(w->region, region, w->M * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (chain) CK(cudaMemcpyAsync(w->chain, chain, 2 * B * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(w->step, 0, sizeof(int), s));
  CK(cudaMemsetAsync(w->err_flag, 0, sizeof(int), s));
  CK(set_rng_launch(w->rng, rng ? rng->seed : 0, rng ? rng->seq_offset : 0, s));
  // profile the schedule of the SAMPLING step (last-block shortcut included); without a sampling call
  // before, slot 0 of every sequence stands in for the sampled slot
  const bool use_tail = c->tail && !c->att.empty() && c->att.back().has_split && c->cfg.nhead <= 8 && c->L <= 320;
  if (use_tail && !w->order) {
    CKI(dev_alloc(w->allocs, &w->order, static_cast<size_t>(B), &w->bytes, true));
    w->order_cap = B; w->order_T = 1;
  }
  Plan plan;
  plan.flip = c->l2flip;
  CKI(plan_trunk(c, w, plan, drop_from_rng(c, rng, B), use_tail));
  const size_t n = plan.ops.size();
  std::vector<cudaEvent_t> ev(2 * n);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  const char* eager = std::getenv("HUDIFF_PROFILE_EAGER");
  if (eager && eager[0] == '1') {
    for (size_t i = 0; i < n; ++i) {
      CK(cudaEventRecord(ev[2 * i], s));
      CK(plan.ops[i](s));
      CK(cudaEventRecord(ev[2 * i + 1], s));
    }
    CK(cudaStreamSynchronize(s));
  } else {
    // The forward is captured into a CUDA graph exactly like the timed sampling step, with external event-record
    // nodes between the kernels: the per-kernel times are taken INSIDE a graph launch (same launch gaps and cache
    // state as the benchmarked loop), so their sum reproduces ms_per_step / T.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = HUDIFF_OK;
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
      e = cudaEventRecordWithFlags(ev[2 * i], s, cudaEventRecordExternal);
      if (e == cudaSuccess) e = plan.ops[i](s);
      if (e == cudaSuccess) e = cudaEventRecordWithFlags(ev[2 * i + 1], s, cudaEventRecordExternal);
    }
    cudaError_t e2 = cudaStreamEndCapture(s, &graph);
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      rc = fail(HUDIFF_ERR_CUDA, std::string("profile capture: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    if (rc == HUDIFF_OK) {
      e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) rc = fail(HUDIFF_ERR_CUDA, std::string("profile instantiate: ") + cudaGetErrorString(e));
    }
    if (rc == HUDIFF_OK) {
      for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) e = cudaGraphLaunch(exec, s);      // the last launch is the one read
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaGraphExecDestroy(exec);
      if (e != cudaSuccess) rc = fail(HUDIFF_ERR_CUDA, std::string("profile launch: ") + cudaGetErrorString(e));
    }
    if (rc != HUDIFF_OK) { for (auto& x : ev) cudaEventDestroy(x); return rc; }
  }
  for (int k = 0; k < OP_KINDS; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; flops_by_kind[k] = 0.0; }
  for (size_t i = 0; i < n; ++i) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
    ms_by_kind[plan.kinds[i]] += ms;
    launches_by_kind[plan.kinds[i]] += 1;
    flops_by_kind[plan.kinds[i]] += plan.flops[i];
  }
  if (const char* dump = std::getenv("HUDIFF_PROFILE_DUMP")) {
    if (FILE* f = std::fopen(dump, "a")) {
      const char* dbg = std::getenv("HUDIFF_DBG");
      for (size_t i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
        std::fprintf(f, "%s\t%zu\t%s\t%.4f\t%.3f\n", dbg ? dbg : "0", i, plan.labels[i].c_str(), ms, plan.flops[i] * 1e-9);
      }
      std::fclose(f);
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return check_err_flag(w, s);
}

// Test / measurement hook: the attention core alone, any sequence length.  qkv: device fp16 [B*L, 3*att_dim] (q|k|v column
// blocks; q already scaled and rotated), out: device fp16 [B*L, att_dim].  impl: 1 = mma.sync validation kernel, 2 = tc2,
// 3 = tc3, 4 = tc4.  iters > 0 repeats the launch and returns the mean device time per launch in *ms_out.
int hudiff_debug_attention(const void* qkv, void* out, int B, int L, int att_dim, int nhead, int impl, int iters, float* ms_out) {
  if (!qkv || !out || B < 1 || L < 1 || L > 320 || nhead < 1 || att_dim != nhead * 64) return fail(HUDIFF_ERR_INVALID, "bad attention problem");
  CK(kernels_init());
  __half* q = static_cast<__half*>(const_cast<void*>(qkv));
  __half* o = static_cast<__half*>(out);
  CUtensorMap host[2];
  CUtensorMap* dmaps = nullptr;
  int* err = nullptr;
  if (impl != 1) {
    const int LK = round_up(L, 64);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(3 * att_dim), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(3 * att_dim) * 2, static_cast<cuuint64_t>(L) * 3 * att_dim * 2};
    cuuint32_t box_q[3] = {64, 128, 1};
    cuuint32_t box_kv[3] = {64, static_cast<cuuint32_t>(LK / 2), 1};
    CKI(encode_map(&host[0], q, 3, dims, strides, box_q));
    CKI(encode_map(&host[1], q, 3, dims, strides, box_kv));
    CK(cudaMalloc(&dmaps, sizeof(host)));
    CK(cudaMemcpy(dmaps, host, sizeof(host), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&err, sizeof(int)));
  CK(cudaMemset(err, 0, sizeof(int)));
  auto launch = [&]() -> cudaError_t {
    if (impl == 1) {
      AttnParams ap;
      ap.qkv_hi = q; ap.qkv_lo = nullptr; ap.ld_qkv = 3 * att_dim; ap.o_hi = o; ap.o_lo = nullptr; ap.ld_o = att_dim;
      ap.L = L; ap.att_dim = att_dim; ap.nhead = nhead; ap.B = B;
      return attention_launch(ap, false, 0);
    }
    if (impl == 2) return attention_tc2_launch(dmaps, dmaps + 1, o, att_dim, L, att_dim, nhead, B, err, 0);
    if (impl == 4) return attention_tc4_launch(dmaps, dmaps + 1, o, att_dim, L, att_dim, nhead, B, err, 0);
    return attention_tc3_launch(dmaps, dmaps + 1, o, att_dim, L, att_dim, nhead, B, err, 0);
  };
  cudaError_t e = launch();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && iters > 0 && ms_out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    for (int i = 0; i < iters && e == cudaSuccess; ++i) e = launch();
    cudaEventRecord(b, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / static_cast<float>(iters);
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
  int flag = 0;
  if (e == cudaSuccess) cudaMemcpy(&flag, err, sizeof(int), cudaMemcpyDeviceToHost);
  if (dmaps) cudaFree(dmaps);
  cudaFree(err);
  if (e != cudaSuccess) return fail(HUDIFF_ERR_CUDA, std::string("attention launch: ") + cudaGetErrorString(e));
  if (flag != 0) return fail(HUDIFF_ERR_CUDA, "attention pipeline watchdog " + std::to_string(flag));
  return HUDIFF_OK;
}

// Diagnostic: phase clocks of the persistent attention kernel's CTA 0 (last launch).
int hudiff_debug_attn_clocks(long long* out16) {
  if (!out16) return fail(HUDIFF_ERR_INVALID, "null argument");
  CK(cudaDeviceSynchronize());
  CK(attention_tc2_clocks(out16));                 // [0..7]: second-generation kernel (HUDIFF_ATTN=tc2)
  CK(attention_tc3_clocks(out16 + 8));             // [8..11]: tiles, total / wait-S / wait-O cycles of CTA 0, slot 0
  return HUDIFF_OK;
}

// Diagnostic: epilogue phase clocks of the GEMM kernel (HUDIFF_DBG bit 8), [4 epilogue kinds][8]; optionally reset.
int hudiff_debug_gemm_clocks(unsigned long long* out32, int reset) {
  if (!out32) return fail(HUDIFF_ERR_INVALID, "null argument");
  CK(cudaDeviceSynchronize());
  CK(gemm_tc_clocks(out32, reset));
  return HUDIFF_OK;
}

// Debug/test hook: device pointer of a workspace buffer after the last forward of batch B.
int hudiff_debug_buffer(hudiff_handle c, int B, const char* name, void** ptr, int* ld, int* is_half) {
  if (!c || !name || !ptr || !ld || !is_half) return fail(HUDIFF_ERR_INVALID, "bad argument");
  auto it = c->ws.find(B);
  if (it == c->ws.end()) return fail(HUDIFF_ERR_STATE, "no workspace for this batch size yet");
  Workspace* w = it->second.get();
  const std::string n = name;
  *is_half = 1;
  if (n == "E32") { *ptr = w->E32; *ld = c->d; *is_half = 0; }
  else if (n == "X32") { *ptr = w->X32; *ld = c->D; *is_half = 0; }
  else if (n == "A32") { *ptr = w->A32; *ld = c->D; *is_half = 0; }
  else if (n == "opE") { *ptr = w->opE.hi; *ld = w->opE.ld; }
  else if (n == "hE1") { *ptr = w->hE1.hi; *ld = w->hE1.ld; }
  else if (n == "hE2") { *ptr = w->hE2.hi; *ld = w->hE2.ld; }
  else if (n == "opD") { *ptr = w->opD.hi; *ld = w->opD.ld; }
  else if (n == "hD1") { *ptr = w->hD1.hi; *ld = w->hD1.ld; }
  else if (n == "hD2") { *ptr = w->hD2.hi; *ld = w->hD2.ld; }
  else if (n == "qkv") { *ptr = w->qkv.hi; *ld = w->qkv.ld; }
  else if (n == "o16") { *ptr = w->o16.hi; *ld = w->o16.ld; }
  else if (n == "f16") { *ptr = w->f16.hi; *ld = w->f16.ld; }
  else return fail(HUDIFF_ERR_INVALID, "unknown buffer name");
  return HUDIFF_OK;
}

int hudiff_destroy(hudiff_handle c) {
  if (!c) return HUDIFF_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->work) cudaStreamDestroy(c->work);
  if (c->io) cudaStreamDestroy(c->io);
  if (c->ev_in) cudaEventDestroy(c->ev_in);
  if (c->ev_out) cudaEventDestroy(c->ev_out);
  for (auto& kv : c->ws) for (void* p : kv.second->allocs) cudaFree(p);
  for (void* p : c->wallocs) cudaFree(p);
  delete c;
  return HUDIFF_OK;
}

}  // extern "C"
