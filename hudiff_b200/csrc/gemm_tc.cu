// Persistent tcgen05 GEMM / implicit-GEMM dilated convolution for sm_100a.
//
//   D[128 x BN] (fp32, TMEM) = sum over K-steps  A_tile[128 x 64] (fp16, smem)  *  W_tile[BN x 64]^T
//
// Grid: one CTA per SM, grouped in clusters of `cs` CTAs.  A cluster walks the (M block, N tile)
// work list; its CTAs own consecutive 128-row M tiles and share the weight tile: every CTA
// fetches 1/cs of it and TMA-multicasts the slice to all peers, which divides the L2->SM operand
// traffic (the measured limiter of the first version) by up to cs.
// Warp roles per CTA: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one
// lane), warps 2..5 = epilogue.  The accumulator is double-buffered in TMEM (2 x 256 columns) so
// the epilogue of tile i overlaps the K loop of tile i+1.
// A K-step is one (tap, 64-channel chunk): the A tile of tap t is the activation tile shifted by
// (t - taps/2) * dilation rows, fetched through a 3-D tensor map (channels, rows-in-chain,
// sequence) whose out-of-bounds zero fill supplies the convolution padding and the ragged chain
// tails.  fp16x3 mode runs three passes (hi*hi, hi*lo, lo*hi) into the same accumulator.
// Epilogue: each warp dumps its 32 accumulator rows (TMEM lane quadrant) 64 columns at a time
// into a private XOR-swizzled smem buffer (thread = row), then re-reads it with thread = column
// pair so that bias / LayerNorm / activation / residual / dropout / RoPE and every global load
// and store are coalesced 128-256 B row segments.
#include "hd_common.cuh"
#include "hd_kernels.h"

#include <cstring>

namespace hd {

constexpr int BM = 128;
constexpr int BK = 64;               // fp16 elements per K-step = one 128-byte swizzle atom
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;         // two warps per TMEM lane quadrant, alternating column steps
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_STEP = 64;         // columns per epilogue step
constexpr int STAGING_BYTES_PER_WARP = 32 * EPI_STEP * 4;        // 32 rows x 64 fp32, XOR-swizzled
constexpr int ACC_COLS = 256;        // TMEM columns per accumulator stage when double-buffered

struct RowInfo {                     // per accumulator row, written once per tile by its owner lane
  long long g;                       // global row index, -1 if the row does not exist
  int didx;                          // seq * Lc + l   (dropout plane row)
  int pos;                           // position inside the sequence (RoPE phase)
  float mean, rstd;                  // LayerNorm statistics (EPI_LN_ACT_F16)
  int pad[2];
};
static_assert(sizeof(RowInfo) == 32, "RowInfo must be 32 bytes");

struct TileCoord { int pi, tile_m, n0; };

__device__ __forceinline__ TileCoord tile_lookup(const GemmParams& G, int t, int rank) {
  TileCoord tc;
  int mblk = t / G.n_tiles;
  const int nt = t - mblk * G.n_tiles;
  tc.pi = 0;
  if (mblk >= G.mblocks[0]) { tc.pi = 1; mblk -= G.mblocks[0]; }
  tc.tile_m = mblk * G.cs + rank;
  tc.n0 = nt * G.p[tc.pi].BN;
  return tc;
}

// Phase 2 works on row pairs: lanes 0-15 take row 2i, lanes 16-31 row 2i+1, each lane 4 columns.
// Global loads (residual rows / RoPE phases) of 8 iterations are issued ahead of their use.
template <int EPI>
__device__ __forceinline__ void epi_issue_loads(const EpiParams& E, const RowInfo* info, int it0, int sub, int c0,
                                                int rope_j, int sec, float4* out) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    out[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r = 2 * (it0 + u) + sub;
    if (EPI == EPI_RESID_F32) {
      const long long g = info[r].g;
      if (E.resid && g >= 0) out[u] = *reinterpret_cast<const float4*>(E.resid + g * E.ldr + c0);
    } else if (EPI == EPI_QKV_ROPE) {
      if (sec < 2)
        out[u] = __ldg(reinterpret_cast<const float4*>(E.rope + static_cast<long long>(info[r].pos) * (E.head_dim / 2) + rope_j));
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epi_process_rows(const EpiParams& E, const RowInfo* info, int it0, int sub, int cg,
                                                 const uint8_t* sbuf, const uint32_t* kbuf, int c0, float4 bias4,
                                                 float4 gam4, float4 bet4, int sec, float qs, const float4* ld) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int r = 2 * (it0 + u) + sub;
    const float4 a = *reinterpret_cast<const float4*>(sbuf + r * 256 + ((cg ^ (r & 15)) << 4));
    float x[4] = {fmaf(a.x, E.wscale, bias4.x), fmaf(a.y, E.wscale, bias4.y), fmaf(a.z, E.wscale, bias4.z),
                  fmaf(a.w, E.wscale, bias4.w)};
    const long long g = info[r].g;
    if (EPI == EPI_LN_ACT_F16) {
      const float mean = info[r].mean, rstd = info[r].rstd;
      x[0] = act_apply((x[0] - mean) * rstd * gam4.x + bet4.x, E.act);
      x[1] = act_apply((x[1] - mean) * rstd * gam4.y + bet4.y, E.act);
      x[2] = act_apply((x[2] - mean) * rstd * gam4.z + bet4.z, E.act);
      x[3] = act_apply((x[3] - mean) * rstd * gam4.w + bet4.w, E.act);
    } else if (EPI == EPI_RESID_F32) {
      x[0] += ld[u].x; x[1] += ld[u].y; x[2] += ld[u].z; x[3] += ld[u].w;
      if (E.drop_mode != DROP_OFF) {
        const uint32_t kw = kbuf[r * 2 + (cg >> 3)] >> ((cg & 7) * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ((kw >> j) & 1u) ? x[j] * E.drop_scale : 0.f;
      }
      if (g >= 0) *reinterpret_cast<float4*>(E.out32 + g * E.ld32 + c0) = make_float4(x[0], x[1], x[2], x[3]);
    } else if (EPI == EPI_ACT_F16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = act_apply(x[j], E.act);
    } else {  // EPI_QKV_ROPE: two (re, im) pairs per lane
      if (sec < 2) {
        const float r0 = x[0], i0 = x[1], r1 = x[2], i1 = x[3];
        x[0] = (r0 * ld[u].x - i0 * ld[u].y) * qs;
        x[1] = (r0 * ld[u].y + i0 * ld[u].x) * qs;
        x[2] = (r1 * ld[u].z - i1 * ld[u].w) * qs;
        x[3] = (r1 * ld[u].w + i1 * ld[u].z) * qs;
      }
    }
    if (E.out16_hi && g >= 0) {
      __half h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_f16(x[j], h[j], l[j]);
      *reinterpret_cast<uint2*>(E.out16_hi + g * E.ld16 + c0) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
      if (E.out16_lo)
        *reinterpret_cast<uint2*>(E.out16_lo + g * E.ld16 + c0) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams G, int stages, int stage_bytes, int acc_stages) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms (identical offset in every CTA)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster = static_cast<int>(cluster_id_x());
  const int nclusters = static_cast<int>(cluster_nctaid_x());
  const uint16_t cmask = static_cast<uint16_t>((1u << G.cs) - 1u);

  uint8_t* tiles = smem;
  uint8_t* staging = smem + stages * stage_bytes;
  RowInfo* rowinfo = reinterpret_cast<RowInfo*>(staging + EPI_WARPS * STAGING_BYTES_PER_WARP);   // [EPI_WARPS][32]
  uint32_t* keepw = reinterpret_cast<uint32_t*>(rowinfo + EPI_WARPS * 32);                       // [EPI_WARPS][32][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(keepw + EPI_WARPS * 32 * 2);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], G.cs); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t it = 0;
      for (int t = cluster; t < G.total_tiles; t += nclusters) {
        const TileCoord tc = tile_lookup(G, t, rank);
        const GemmProblem& P = G.p[tc.pi];
        const RowSpace rs = P.rs;
        const int total_rows = rs.nseq * rs.Lcp;
        const int r_begin = tc.tile_m * BM;
        uint32_t a_bytes = 0;
        {
          int r = r_begin, rem = BM;
          while (rem > 0 && r < total_rows) {
            const int l0 = r % rs.Lcp;
            const int h = min(rem, rs.Lcp - l0);
            a_bytes += static_cast<uint32_t>(h) * BK * 2;
            r += h; rem -= h;
          }
        }
        const uint32_t b_bytes = static_cast<uint32_t>(P.BN) * BK * 2;
        const int slice = P.BN / G.cs;                  // weight rows this CTA fetches for the cluster
        const int half_taps = P.taps / 2;
        const int nk1 = P.taps * P.kchunks;
        const int nk = G.x3 ? 3 * nk1 : nk1;
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1, G.err_flag, 101);
          const int pass = ks / nk1;                    // 0: hi*hi  1: hi*lo  2: lo*hi
          const int k1 = ks - pass * nk1;
          const int tap = k1 / P.kchunks;
          const int kc = k1 - tap * P.kchunks;
          const CUtensorMap* amaps = (pass == 2) ? P.a_maps_lo : P.a_maps_hi;
          const CUtensorMap* bmap = (pass == 1) ? P.b_map_lo : P.b_map_hi;
          uint8_t* a_dst = tiles + s * stage_bytes;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          if (G.dbg & 4) { mbar_arrive(&full_bar[s]); continue; }
          mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes);
          const int shift = (tap - half_taps) * P.dil;
          int r = r_begin, rem = BM;
          while (rem > 0 && r < total_rows) {
            const int seq = r / rs.Lcp;
            const int l0 = r - seq * rs.Lcp;
            const int h = min(rem, rs.Lcp - l0);
            tma_load_3d(a_dst + (BM - rem) * (BK * 2), amaps + (h / 8 - 1), &full_bar[s], kc * BK, l0 + shift, seq);
            r += h; rem -= h;
          }
          const int bk = (tap * P.kchunks + kc) * BK;
          for (int nb = rank * slice; nb < (rank + 1) * slice; nb += P.b_box_n) {
            if (G.cs > 1) tma_load_2d_mc(b_dst + nb * (BK * 2), bmap, &full_bar[s], bk, tc.n0 + nb, cmask);
            else tma_load_2d(b_dst + nb * (BK * 2), bmap, &full_bar[s], bk, tc.n0 + nb);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      uint32_t it = 0, tl = 0;
      for (int t = cluster; t < G.total_tiles; t += nclusters, ++tl) {
        const TileCoord tc = tile_lookup(G, t, rank);
        const GemmProblem& P = G.p[tc.pi];
        const int nk1 = P.taps * P.kchunks;
        const int nk = G.x3 ? 3 * nk1 : nk1;
        const int acc = tl % acc_stages;
        const uint32_t aph = (tl / acc_stages) & 1;
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1, G.err_flag, 104);     // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * ACC_COLS);
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&full_bar[s], ph, G.err_flag, 102);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + s * stage_bytes);
          const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < ((G.dbg & 2) ? 0 : BK / 16); ++kk) {
            const uint64_t adesc = umma_smem_desc_sw128(a_addr + kk * 32);
            for (int nc = 0; nc < P.BN; nc += 256) {
              const int n_instr = min(256, P.BN - nc);
              const uint64_t bdesc = umma_smem_desc_sw128(b_addr + nc * (BK * 2) + kk * 32);
              umma_f16(d_tmem + static_cast<uint32_t>(nc), adesc, bdesc, umma_idesc_f16(n_instr),
                       (ks > 0 || kk > 0) ? 1u : 0u);
            }
          }
          // the stage is free once every CTA of the cluster has consumed it (peers multicast into it)
          if (G.cs > 1) umma_commit_mc(&empty_bar[s], cmask);
          else umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);            // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (8 warps: 2 per TMEM lane quadrant) =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int ew = warp - 2;                         // private smem slice
    const int grp = ew >> 2;                         // column steps with (st & 1) == grp are mine
    uint8_t* sbuf = staging + ew * STAGING_BYTES_PER_WARP;
    RowInfo* my_info = rowinfo + ew * 32;
    uint32_t* kbuf = keepw + ew * (32 * 2);
    const int step = G.step_ptr ? *G.step_ptr : 0;
    const int sub = lane >> 4, cg = lane & 15;       // phase 2: row parity and 4-column group of this lane
    uint32_t tl = 0;
    for (int t = cluster; t < G.total_tiles; t += nclusters, ++tl) {
      const TileCoord tc = tile_lookup(G, t, rank);
      const GemmProblem& P = G.p[tc.pi];
      const EpiParams& E = P.ep;
      const int acc = tl % acc_stages;
      const uint32_t aph = (tl / acc_stages) & 1;
      const RowRef rr = row_lookup(P.rs, tc.tile_m * BM + quad * 32 + lane);
      mbar_wait(&tmem_full_bar[acc], aph, G.err_flag, 103);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * ACC_COLS);
      const int nsteps = P.BN / EPI_STEP;
      float v[32];
      if (G.dbg & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }

      float mean = 0.f, rstd = 1.f;
      if (EPI == EPI_LN_ACT_F16) {
        // shifted single-pass moments of this thread's row over the full width (BN == N); both warps
        // of a quadrant compute them (TMEM reads are cheap, no cross-warp exchange needed)
        float shift = 0.f, s1 = 0.f, s2 = 0.f;
        for (int c = 0; c < P.BN / 32; ++c) {
          tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v);
          const float4* b4 = reinterpret_cast<const float4*>(E.bias + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = __ldg(b4 + q);
            const float xs[4] = {fmaf(v[4 * q], E.wscale, bb.x), fmaf(v[4 * q + 1], E.wscale, bb.y),
                                 fmaf(v[4 * q + 2], E.wscale, bb.z), fmaf(v[4 * q + 3], E.wscale, bb.w)};
            if (c == 0 && q == 0) shift = xs[0];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float d = xs[j] - shift; s1 += d; s2 = fmaf(d, d, s2); }
          }
        }
        const float inv_n = 1.0f / static_cast<float>(P.N);
        const float md = s1 * inv_n;
        mean = shift + md;
        rstd = 1.0f / sqrtf(fmaxf(s2 * inv_n - md * md, 0.f) + 1e-5f);
      }
      __syncwarp();                                  // previous tile's readers are done with my_info / sbuf
      {
        RowInfo ri;
        ri.g = rr.valid ? rr.g : -1;
        ri.didx = rr.seq * P.rs.Lc + rr.l;
        ri.pos = (EPI == EPI_QKV_ROPE && rr.valid) ? static_cast<int>(rr.g % E.seq_len) : 0;
        ri.mean = mean; ri.rstd = rstd; ri.pad[0] = ri.pad[1] = 0;
        my_info[lane] = ri;
      }
      __syncwarp();

      int last_mine = -1;
      for (int st = grp; st < nsteps; st += 2) last_mine = st;
      if (last_mine < 0) {                           // nothing to do for this warp (single-step tiles)
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      for (int st = grp; st < nsteps; st += 2) {
        // per-lane column constants of this step (phase 2: lane = 4 consecutive columns)
        const int c0 = tc.n0 + st * EPI_STEP + 4 * cg;
        const float4 bias4 = *reinterpret_cast<const float4*>(E.bias + c0);
        float4 gam4 = make_float4(1.f, 1.f, 1.f, 1.f), bet4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI == EPI_LN_ACT_F16) {
          gam4 = *reinterpret_cast<const float4*>(E.gamma + c0);
          bet4 = *reinterpret_cast<const float4*>(E.beta + c0);
        }
        const int sec = (EPI == EPI_QKV_ROPE) ? c0 / E.att_dim : 2;
        const int rope_j = (EPI == EPI_QKV_ROPE) ? ((c0 % E.head_dim) >> 1) : 0;
        const float qs = (sec == 0) ? E.qscale : 1.0f;
        float4 ld0[8], ld1[8];
        epi_issue_loads<EPI>(E, my_info, 0, sub, c0, rope_j, sec, ld0);
        // ---- phase 1: thread = row.  TMEM -> registers -> swizzled smem (raw accumulators)
#pragma unroll
        for (int hch = 0; hch < 2; ++hch) {
          tmem_ld32(taddr + static_cast<uint32_t>(st * EPI_STEP + hch * 32), v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = hch * 8 + q;               // 16-byte chunk index inside the 256-byte row
            *reinterpret_cast<float4*>(sbuf + lane * 256 + ((j ^ (lane & 15)) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        if (EPI == EPI_RESID_F32 && E.drop_mode != DROP_OFF && rr.valid) {
          uint32_t k0, k1;
          dropout_keep_words2(E, G.seed, step, rr.seq * P.rs.Lc + rr.l, P.N / 32, (tc.n0 + st * EPI_STEP) >> 5, k0, k1);
          kbuf[lane * 2 + 0] = k0;
          kbuf[lane * 2 + 1] = k1;
        }
        if (st == last_mine) {                       // my share of the accumulator is read: release it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        } else {
          __syncwarp();
        }
        // ---- phase 2: coalesced math + global I/O, two batches of 8 row pairs
        epi_issue_loads<EPI>(E, my_info, 8, sub, c0, rope_j, sec, ld1);
        epi_process_rows<EPI>(E, my_info, 0, sub, cg, sbuf, kbuf, c0, bias4, gam4, bet4, sec, qs, ld0);
        epi_process_rows<EPI>(E, my_info, 8, sub, cg, sbuf, kbuf, c0, bias4, gam4, bet4, sec, qs, ld1);
        __syncwarp();                                // sbuf / kbuf are reused by the next step
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // no CTA leaves while a peer may still multicast into it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------
static int g_num_sms = 0;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int EPI_SMEM = EPI_WARPS * (STAGING_BYTES_PER_WARP + 32 * 32 + 32 * 2 * 4);

static size_t gemm_tc_smem_bytes(int BN, int stages) {
  return static_cast<size_t>(stages) * (A_TILE_BYTES + BN * BK * 2) + EPI_SMEM + (2 * stages + 4) * 8 + 16 + 1024;
}

static int gemm_tc_pick_stages(int BN) {
  int s = 6;
  while (s > 1 && gemm_tc_smem_bytes(BN, s) > static_cast<size_t>(SMEM_LIMIT)) --s;
  return s;
}

typedef void (*GemmKernel)(GemmParams, int, int, int);
static GemmKernel gemm_kernel_for(int epi) {
  switch (epi) {
    case EPI_LN_ACT_F16: return gemm_tc_kernel<EPI_LN_ACT_F16>;
    case EPI_RESID_F32: return gemm_tc_kernel<EPI_RESID_F32>;
    case EPI_ACT_F16: return gemm_tc_kernel<EPI_ACT_F16>;
    default: return gemm_tc_kernel<EPI_QKV_ROPE>;
  }
}

cudaError_t gemm_tc_init() {
  cudaError_t e = cudaSuccess;
  for (int epi = 0; epi < 4 && e == cudaSuccess; ++epi)
    e = cudaFuncSetAttribute(gemm_kernel_for(epi), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}
cudaError_t kernels_init() {
  cudaError_t e = gemm_tc_init();
  if (e != cudaSuccess) return e;
  return attention_init();
}

// Fills the scheduling fields of G (cs must be set by the caller) and launches the persistent grid.
cudaError_t gemm_tc_launch(const GemmParams& Gin, cudaStream_t stream) {
  GemmParams G = Gin;
  const int cs = G.cs < 1 ? 1 : G.cs;
  G.cs = cs;
  int BN = 0;
  G.n_tiles = 1;
  G.mblocks[0] = G.mblocks[1] = 0;
  for (int i = 0; i < G.nprob; ++i) {
    const GemmProblem& P = G.p[i];
    BN = max(BN, P.BN);
    G.n_tiles = max(G.n_tiles, (P.N + P.BN - 1) / P.BN);
    G.mblocks[i] = (P.m_tiles + cs - 1) / cs;
  }
  G.total_tiles = (G.mblocks[0] + G.mblocks[1]) * G.n_tiles;
  const int stages = gemm_tc_pick_stages(BN);
  const int stage_bytes = A_TILE_BYTES + BN * BK * 2;
  const int acc_stages = (BN <= ACC_COLS) ? 2 : 1;
  const size_t smem = gemm_tc_smem_bytes(BN, stages);
  if (g_num_sms <= 0) {
    cudaError_t e = gemm_tc_init();
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((g_num_sms / cs) * cs, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // clusters that can be co-resident (GPC granularity can strand SMs for cs = 4): the static
  // round-robin schedule wants every launched cluster to be running
  static int max_clusters[5][4] = {};
  const int sidx = (BN <= 128) ? 0 : (BN <= 256) ? 1 : 2;
  if (max_clusters[cs][sidx] == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel_for(G.p[0].ep.epi), &cfg) != cudaSuccess || n <= 0) n = g_num_sms / cs;
    max_clusters[cs][sidx] = n;
    (void)cudaGetLastError();
  }
  const int nclusters = min(G.total_tiles, min(max_clusters[cs][sidx], g_num_sms / cs));
  cfg.gridDim = dim3(nclusters * cs, 1, 1);
  return cudaLaunchKernelEx(&cfg, gemm_kernel_for(G.p[0].ep.epi), G, stages, stage_bytes, acc_stages);
}

}  // namespace hd
