// dcpgpu.cu -- the C ABI of include/dcpgpu.h over the sm_100a kernels.
//
// Host side of the boundary described in include/dcpgpu.h: profile residency
// (work_setup/protein_setup_viterbi, c-core/work.c:24-46, protein.c:353-394), read
// packing (batch_encode, c-core/batch.c:60), per-window special transitions
// (xtrans_setup, c-core/xtrans.c:21-68) and the launches of the score and trace kernels.
// No CPU compute path exists here: every entry point needs a CUDA device.
#include "../../include/dcpgpu.h"
#include "generic_kernel.cuh"
#include "layout.cuh"
#include "kernels.h"
#include "match_kernel.cuh"
#include "press_kernel.cuh"
#include "row_kernel.cuh"
#include "trace_argmin.cuh"
#include "trace_walk.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace dcp;

namespace {

struct Slab
{
  char *base;
  size_t size;
  size_t used;
};

struct PoolBlock
{
  float *em;    // [n][1364] log-probs
  float *trans; // [n][7]
  int64_t first;
  int64_t n;
};

struct NodeRef
{
  float const *em;
  float const *trans;
};

} // namespace

struct dcpgpu_ctx
{
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;

  std::vector<Slab> slabs;
  size_t profile_bytes = 0;

  std::vector<PoolBlock> pool;
  int64_t pool_nodes = 0;

  std::vector<ProfileDesc> h_profiles;
  std::vector<char> h_unsafe; // a cost is negative/NaN: only the generic kernel may run it
  ProfileDesc *d_profiles = nullptr;
  size_t d_profiles_cap = 0;
  // per-profile decode tables for the match strings (match_kernel.cuh)
  std::vector<DecoderDesc> h_decoders;
  DecoderDesc *d_decoders = nullptr;
  size_t d_decoders_cap = 0;
  // match pass state
  int *d_m_hit = nullptr, *d_m_start = nullptr, *d_m_stop = nullptr, *d_m_begin = nullptr, *d_m_end = nullptr, *d_m_bad = nullptr;
  long long *d_m_len = nullptr, *d_m_off = nullptr;
  size_t m_cap[7] = {0, 0, 0, 0, 0, 0, 0};
  uint8_t *d_m_codon = nullptr;
  char *d_m_amino = nullptr, *d_m_text = nullptr;
  size_t m_codon_cap = 0, m_amino_cap = 0, m_text_cap = 0, m_bad_cap = 0;
  std::vector<long long> m_text_off;
  bool matched = false;
  bool steps_compact = false;
  // segmented copies of the profiles of more than 256 nodes (strip_kernel.cuh)
  std::vector<ProfileDesc> h_segs;
  std::vector<int> h_seg_first, h_seg_count; // per profile; first = -1: not segmented
  ProfileDesc *d_segs = nullptr;
  int *d_seg_first = nullptr;
  size_t d_segs_cap = 0, d_seg_first_cap = 0;
  static constexpr int MAXSEG = 8;
  cudaEvent_t ev_level[MAXSEG] = {};
  unsigned long long *d_seg_cursor = nullptr; // [MAXSEG * 17] work cursors of the segment launches
  int *d_seg_list = nullptr;                  // profile / pair lists of the segment launches
  long long *d_seg_order = nullptr, *d_seg_colmap = nullptr;
  size_t seg_list_cap = 0, seg_order_cap = 0, seg_colmap_cap = 0;
  bool profiles_dirty = false;

  // upload staging (device)
  char *d_stage = nullptr;
  size_t d_stage_cap = 0;

  // reads
  uint32_t *d_words = nullptr;
  long long nwords = 0;
  uint16_t *d_hist = nullptr; // ten-bit history of every position of the packed stream (row_kernel.cuh)
  long long *d_seq_word = nullptr;
  int *d_seq_len = nullptr;
  std::vector<int> h_seq_len;
  int nseq = 0;
  int maxlen = 0;

  // special transitions per flag combination
  float *d_xt[4] = {nullptr, nullptr, nullptr, nullptr};
  int xt_len[4] = {-1, -1, -1, -1};

  // score pass state
  float2 *d_out = nullptr;
  size_t out_cap = 0;
  int64_t last_n = 0;
  unsigned long long *d_counters = nullptr; // [NSLOTS]: class work cursors, then the SLOT_* counters
  int *d_class_profiles = nullptr;
  size_t class_profiles_cap = 0;
  Pair *d_pairs = nullptr;
  size_t pairs_cap = 0;
  long long *d_order = nullptr;
  size_t order_cap = 0;
  float *d_scratch = nullptr;
  size_t scratch_cap = 0; // floats
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // side streams: the per-class kernels of one pass are independent, so they are spread over these
  // and the tail of one class overlaps the head of the next
#ifndef DCPGPU_NSIDE
#define DCPGPU_NSIDE 8
#endif
  static constexpr int NSIDE = DCPGPU_NSIDE;
  cudaStream_t side[NSIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {};
  bool forked = false;
  int side_next = 0;
  cudaStream_t pinned = nullptr; // overrides the rotation (kernels that depend on each other)
  Mail *d_col = nullptr; // boundary columns of the one-strip-per-launch kernels
  size_t col_cap = 0;
  int stage = 1; // DCPGPU_STAGE=0: never use the profile-stationary (TMA-staged) kernels of row_kernel.cuh
  size_t col_budget = size_t(16) << 30; // DCPGPU_COL_BUDGET_MB: cap of the boundary columns (tests force chunking)
  long long lz_slack = 64;              // DCPGPU_LZ_SLACK: slack of a lazily walked path's slot (tests force the rerun)
  bool subwarp = true;           // DCPGPU_SUBWARP=0: profiles of <= 128 nodes keep a whole warp (A/B switch)
  bool timed = false;
  double last_cells = 0;
  int64_t launches = 0; // cumulative count of kernels this library launched
  int64_t h2d_bytes = 0, d2h_bytes = 0; // cumulative bytes of the copies this library issued
  double total_cells = 0; // cumulative DP cells over the score passes (dcpgpu_counter)

  // trace pass state (buffers grow on demand and are reused across calls)
  std::vector<Pair> t_pairs;
  Pair *d_tpairs = nullptr;
  size_t tpairs_cap = 0;
  uint32_t *d_xnodes = nullptr;
  size_t xnodes_cap = 0;
  uint16_t *d_nodes = nullptr;
  size_t nodes_cap = 0;
  long long *d_xnode_off = nullptr, *d_node_off = nullptr, *d_step_off = nullptr;
  size_t xnode_off_cap = 0, node_off_cap = 0, step_off_cap = 0;
  int *d_nsteps = nullptr;
  size_t nsteps_cap = 0;
  float2 *d_tout = nullptr;
  size_t tout_cap = 0;
  long long *d_redo = nullptr;
  size_t redo_cap = 0;
  Pair *d_redo_pairs = nullptr;
  size_t redo_pairs_cap = 0;
  long long *d_redo_order = nullptr, *d_redo_out = nullptr;
  size_t redo_order_cap = 0, redo_out_cap = 0;
  int64_t last_redo = 0;
  long long *d_hit_idx = nullptr;
  size_t hit_idx_cap = 0;
  float *d_dump = nullptr;
  size_t dump_cap = 0;
  long long *d_dump_off = nullptr, *d_tile_off = nullptr;
  size_t dump_off_cap = 0, tile_off_cap = 0;
  uint16_t *d_step_ids = nullptr;
  size_t step_ids_cap = 0;
  uint8_t *d_step_sz = nullptr;
  size_t step_sz_cap = 0;
  std::vector<long long> t_xnode_off, t_node_off;
  std::vector<int> t_nsteps;
  bool traced = false;
  // paths-only trace route (trace_walk.cuh): steps in the order the walks finished
  uint16_t *d_lz_ids = nullptr;
  size_t lz_ids_cap = 0;
  uint8_t *d_lz_sz = nullptr;
  size_t lz_sz_cap = 0;
  long long *d_lz_off = nullptr; // per pair; -1 = traced through a kept trellis
  size_t lz_off_cap = 0;
};

namespace {

int fail_cuda(dcpgpu_ctx *c, cudaError_t e, char const *what)
{
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  if (c) c->err = buf;
  return e == cudaErrorMemoryAllocation ? DCPGPU_ENOMEM : DCPGPU_ECUDA;
}

// DCPGPU_TIMING=1: host wall time of the phases of a trace pass on stderr (development aid)
struct TracePhases
{
  bool on = std::getenv("DCPGPU_TIMING") != nullptr;
  double t[6] = {0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point mark = std::chrono::steady_clock::now();
  void lap(int i, cudaStream_t st)
  {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto const now = std::chrono::steady_clock::now();
    t[i] += std::chrono::duration<double>(now - mark).count();
    mark = now;
  }
};

int fail(dcpgpu_ctx *c, int code, char const *what)
{
  if (c) c->err = what;
  return code;
}

// every host<->device copy of the library goes through here so that dcpgpu_counter can report them
inline cudaError_t counted_copy(dcpgpu_ctx *ctx, void *dst, void const *src, size_t bytes, cudaMemcpyKind kind,
                                cudaStream_t st)
{
  if (kind == cudaMemcpyHostToDevice) ctx->h2d_bytes += (int64_t)bytes;
  else if (kind == cudaMemcpyDeviceToHost) ctx->d2h_bytes += (int64_t)bytes;
  return cudaMemcpyAsync(dst, src, bytes, kind, st);
}

#define CU(call)                                                                                 \
  do                                                                                             \
  {                                                                                              \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);                                     \
  } while (0)

template <class T>
int ensure(dcpgpu_ctx *ctx, T *&ptr, size_t &cap, size_t need)
{
  if (need <= cap && ptr) return 0;
  if (ptr) CU(cudaFree(ptr));
  ptr = nullptr;
  cap = 0;
  size_t n = std::max<size_t>(need + need / 4, 16); // head-room: sizes vary a little from call to call
  CU(cudaMalloc(reinterpret_cast<void **>(&ptr), n * sizeof(T)));
  cap = n;
  return 0;
}

int arena_alloc(dcpgpu_ctx *ctx, size_t bytes, void **out)
{
  bytes = (bytes + 255) & ~size_t(255);
  for (auto &s : ctx->slabs)
    if (s.size - s.used >= bytes)
    {
      *out = s.base + s.used;
      s.used += bytes;
      ctx->profile_bytes += bytes;
      return 0;
    }
  size_t const slab = std::max<size_t>(bytes, size_t(512) << 20);
  char *p = nullptr;
  CU(cudaMalloc(reinterpret_cast<void **>(&p), slab));
  ctx->slabs.push_back({p, slab, bytes});
  ctx->profile_bytes += bytes;
  *out = p;
  return 0;
}

// ---- pack kernels: .dcp log-probs -> cost-form device layout (protein.c:353-394) -------

// Nodes [k0, k0 + K) of a profile of Ktot nodes (the whole profile, or one segment of it).
__global__ void pack_em_kernel(NodeRef const *nodes, int k0, int K, int Q, int VL, int Kpad, float *em, int *bad)
{
  // x: node index over Kpad, y: code
  int const k = blockIdx.x * blockDim.x + threadIdx.x;
  int const code = blockIdx.y;
  if (k >= Kpad) return;
  float v = CUDART_INF_F;
  if (k < K) v = -nodes[k0 + k].em[code]; // viterbi_set_match(v, -emission[i], k, i), protein.c:390-391
  if (!(v >= 0.0f)) atomicOr(bad, 1);  // negative or NaN cost: the register kernels' unsigned-order tricks do not apply
  em[(size_t)code * Kpad + layout_pos(k, Q, VL)] = v;
}

__global__ void pack_core_kernel(NodeRef const *nodes, float const *BMk, int k0, int K, int Ktot, int Q, int VL,
                                 int Kpad, float *core, int *bad)
{
  int const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  float v[C_ROWS];
#pragma unroll
  for (int i = 0; i < C_ROWS; ++i)
    v[i] = CUDART_INF_F;
  if (k < K)
  {
    int const g = k0 + k; // node of the whole profile
    v[C_BM] = -BMk[g]; // protein.c:361-364
    if (g >= 1)
    { // transitions out of node g-1 land on node g (protein.c:374-381)
      float const *t = nodes[g - 1].trans;
      v[C_MM] = -t[0];
      v[C_MD] = -t[2];
      v[C_IM] = -t[3];
      v[C_DM] = -t[5];
      v[C_DD] = -t[6];
    }
    if (g + 1 < Ktot)
    { // MI, II stay on node g; node K-1 keeps +INF (protein.c:382-383)
      float const *t = nodes[g].trans;
      v[C_MI] = -t[1];
      v[C_II] = -t[4];
    }
  }
  int const pos = layout_pos(k, Q, VL);
#pragma unroll
  for (int i = 0; i < C_ROWS; ++i)
  {
    if (!(v[i] >= 0.0f)) atomicOr(bad, 1);
    core[(size_t)i * Kpad + pos] = v[i];
  }
}

__global__ void pack_nulbg_kernel(float const *nul, float const *bg, float2 *out, int *bad)
{
  int const c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NCODES) return;
  float2 const v = make_float2(-nul[c], -bg[c]); // protein.c:387-388
  if (!(v.x >= 0.0f) || !(v.y >= 0.0f)) atomicOr(bad, 1);
  out[c] = v;
}

__global__ void hits_fill_kernel(float2 const *out, long long n, unsigned long long *cursor,
                                 long long cap, long long *idx)
{
  long long const i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 const v = out[i];
  float const d = v.y - v.x;
  if (d <= 0.0f && d > -CUDART_INF_F)
  {
    unsigned long long const at = atomicAdd(cursor, 1ULL);
    if ((long long)at < cap) idx[at] = i;
  }
}

__global__ void gather_scores_kernel(float2 const *out, long long const *idx, long long n, float2 *dst)
{
  long long const i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = out[idx[i]];
}

// The special transitions of one window (xtrans.c:21-68 with thread.c:112's max(L/3, 1)).
// Same expressions as the reference: float operands, double log(), results stored to float.
void host_xtrans(int window_len, bool multi_hits, bool hmmer3_compat, float *out)
{
  int seq_size = window_len / 3;
  if (seq_size < 1) seq_size = 1;
  float const L = (float)seq_size;
  float q = 0.0f;
  float log_q = -INFINITY;
  if (multi_hits)
  {
    q = 0.5f;
    log_q = (float)::log(0.5);
  }
  float const denom = L + 2 + q / (1 - q);
  float const lp = (float)(::log((double)L) - ::log((double)denom));
  float const l1p = (float)(::log((double)(2 + q / (1 - q))) - ::log((double)denom));
  float const lr = (float)(::log((double)L) - ::log((double)(L + 1)));
  float NN = lp, CC = lp, JJ = lp;
  float const NB = l1p, CT = l1p, JB = l1p, RR = lr, EJ = log_q;
  float const EC = (float)::log((double)(1 - q));
  if (hmmer3_compat) NN = CC = JJ = ::logf(1.0f);
  out[X_RR] = -RR;
  out[X_SN] = -0 - NN;
  out[X_NN] = -NN;
  out[X_SB] = -0 - NB;
  out[X_NB] = -NB;
  out[X_EB] = -EJ - JB;
  out[X_JB] = -JB;
  out[X_EJ] = -EJ - JJ;
  out[X_JJ] = -JJ;
  out[X_EC] = -EC - CC;
  out[X_CC] = -CC;
  out[X_ET] = -EC - CT;
  out[X_CT] = -CT;
}

int ensure_xt(dcpgpu_ctx *ctx, uint32_t flags, int maxlen)
{
  int const f = (int)(flags & 3u);
  if (ctx->xt_len[f] >= maxlen && ctx->d_xt[f]) return 0;
  int const n = std::max(maxlen, 1);
  std::vector<float> h((size_t)(n + 1) * X_STRIDE, 0.0f);
  for (int L = 1; L <= n; ++L)
    host_xtrans(L, flags & DCPGPU_MULTI_HITS, flags & DCPGPU_HMMER3_COMPAT, &h[(size_t)L * X_STRIDE]);
  if (ctx->d_xt[f]) CU(cudaFree(ctx->d_xt[f]));
  ctx->d_xt[f] = nullptr;
  CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_xt[f]), h.size() * sizeof(float)));
  CU(counted_copy(ctx, ctx->d_xt[f], h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->xt_len[f] = n;
  return 0;
}

int sync_profiles(dcpgpu_ctx *ctx)
{
  if (!ctx->profiles_dirty) return 0;
  size_t const n = ctx->h_profiles.size();
  if (n > ctx->d_profiles_cap)
  {
    if (ctx->d_profiles) CU(cudaFree(ctx->d_profiles));
    ctx->d_profiles = nullptr;
    size_t const cap = std::max<size_t>(n * 2, 1024);
    CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_profiles), cap * sizeof(ProfileDesc)));
    ctx->d_profiles_cap = cap;
  }
  CU(counted_copy(ctx, ctx->d_profiles, ctx->h_profiles.data(), n * sizeof(ProfileDesc),
                     cudaMemcpyHostToDevice, ctx->stream));
  int rc;
  ctx->h_decoders.resize(n, DecoderDesc{nullptr, {0}});
  if ((rc = ensure(ctx, ctx->d_decoders, ctx->d_decoders_cap, n))) return rc;
  CU(counted_copy(ctx, ctx->d_decoders, ctx->h_decoders.data(), n * sizeof(DecoderDesc), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = ensure(ctx, ctx->d_seg_first, ctx->d_seg_first_cap, n))) return rc;
  CU(counted_copy(ctx, ctx->d_seg_first, ctx->h_seg_first.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (!ctx->h_segs.empty())
  {
    if ((rc = ensure(ctx, ctx->d_segs, ctx->d_segs_cap, ctx->h_segs.size()))) return rc;
    CU(counted_copy(ctx, ctx->d_segs, ctx->h_segs.data(), ctx->h_segs.size() * sizeof(ProfileDesc),
                       cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->profiles_dirty = false;
  return 0;
}

// Kernel classes: 0 = generic kernel; 1..8 = one warp per pair, Q nodes per lane;
// 9..20 = W = 2/4/8 warps per pair, Q = 5..8; 21..32 = G = 2/4/8 pairs per warp (profiles of at
// most 128/64/32 nodes on 16/8/4 lanes), Q = 5..8.
constexpr int NCLASS = 33;
// d_counters slots: work cursors of the classes first, then
enum
{
  SLOT_OVERFLOW = 40,    // lazy walk: paths that outgrew their slot
  SLOT_NREDO = 41,       // strip kernels: pairs to redo
  SLOT_NHITS = 42,       // score pass: pairs with lrt >= 0
  SLOT_FILL = 43,        // hits_fill cursor
  SLOT_TRACE_CUR = 44,   // 4 cursors of the trellis-keeping trace kernels
  SLOT_TRACE_NHITS = 48,
  SLOT_BAD = 49,         // pack kernels: negative cost seen
  NSLOTS = 64
};

ReadsView reads_view(dcpgpu_ctx const *ctx);

// stream the next class kernel goes to
cudaStream_t launch_stream(dcpgpu_ctx *ctx)
{
  if (ctx->pinned) return ctx->pinned;
  if (!ctx->forked) return ctx->stream;
  cudaStream_t s = ctx->side[ctx->side_next];
  ctx->side_next = (ctx->side_next + 1) % dcpgpu_ctx::NSIDE;
  return s;
}

int fork_streams(dcpgpu_ctx *ctx)
{
  CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
  for (auto s : ctx->side) CU(cudaStreamWaitEvent(s, ctx->ev_fork, 0));
  ctx->forked = true;
  ctx->side_next = 0;
  return 0;
}

int join_streams(dcpgpu_ctx *ctx)
{
  ctx->forked = false;
  for (int i = 0; i < dcpgpu_ctx::NSIDE; ++i)
  {
    CU(cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
  }
  return 0;
}

// classes in the order a pass launches them: generic, W = 8..2, one warp (Q = 8..1), sub-warp
int launch_order(int i)
{
  if (i == 0) return 0;
  if (i <= 20) return 21 - i;
  return i;
}

int kernel_class(dcpgpu_ctx const *ctx, int profile)
{
  ProfileDesc const &p = ctx->h_profiles[(size_t)profile];
  if (ctx->h_unsafe[(size_t)profile] || p.Q > MAXQ_REG) return 0;
  if (p.VL < 32) return (p.VL == 16 ? 21 : p.VL == 8 ? 25 : 29) + (p.Q - 5);
  if (p.W == 1) return p.Q;
  if (p.Q < 5) return 0;
  if (p.W == 2) return 9 + (p.Q - 5);
  if (p.W == 4) return 13 + (p.Q - 5);
  if (p.W == 8) return 17 + (p.Q - 5);
  return 0;
}

// kernel class -> launcher (kernels.h); every launch lands on launch_stream(ctx)
template <bool DUMP>
int launch_class_t(dcpgpu_ctx *ctx, int cls, ScoreArgs const &a)
{
  cudaStream_t const st = launch_stream(ctx);
  cudaError_t e;
  if (cls >= 1 && cls <= 8 || cls >= 21 && cls <= 32)
  {
    StripArgs sa{};
    sa.s = a;
    int const Q = cls <= 8 ? cls : 5 + (cls - 21) % 4;
    int const SEG = cls <= 8 ? 32 : cls <= 24 ? 16 : cls <= 28 ? 8 : 4;
    // profile-stationary CTAs with the short-code rows and the {null, background} table staged in shared
    // memory by TMA: measured +4..10 % over the plain kernels for every whole-warp shape (profiles/README.md)
    if (!DUMP && ctx->stage && Q >= 5 && !a.pairs && a.nseq >= 4 * (32 / SEG))
      e = launch_row_stage(Q, SEG, ROW_WHOLE, sa, ctx->sm_count, st);
    else
      e = launch_row(Q, SEG, ROW_WHOLE, DUMP, sa, ctx->sm_count, st);
  }
  else if (cls >= 9 && cls <= 20)
    e = launch_reg_multi(5 + (cls - 9) % 4, cls <= 12 ? 2 : cls <= 16 ? 4 : 8, DUMP, a, ctx->sm_count, st);
  else
    return fail(ctx, DCPGPU_EINVAL, "bad kernel class");
  CU(e);
  ctx->launches += 1;
  return 0;
}

int launch_class(dcpgpu_ctx *ctx, int cls, ScoreArgs const &a) { return launch_class_t<false>(ctx, cls, a); }

// ---- segmented profiles: level-by-level launches (row_kernel.cuh: FIRST / MID / LAST) ---------
// kind of a segment launch: -2 = first full segment, -1 = later full segment, 0..15 = tail class
// (0..3: full warp, Q = 5..8; 4..7: 16 lanes; 8..11: 8 lanes; 12..15: 4 lanes).
int tail_class(ProfileDesc const &g)
{
  int const base = g.VL == 32 ? 0 : g.VL == 16 ? 4 : g.VL == 8 ? 8 : 12;
  return base + (g.Q - 5);
}

int launch_segment(dcpgpu_ctx *ctx, int kind, StripArgs const &a, cudaStream_t st)
{
  cudaError_t e;
  // profile-stationary variants (grid mode): first / later full segments and whole-warp tails, each
  // measured +0.5..2.5 % over the plain kernels (profiles/README.md)
  bool const staged = ctx->stage && !a.s.pairs && a.s.nseq >= 4;
  if (kind == -2 && staged) e = launch_row_stage(8, 32, ROW_FIRST, a, ctx->sm_count, st);
  else if (kind == -1 && staged) e = launch_row_stage(8, 32, ROW_MID, a, ctx->sm_count, st);
  else if (kind >= 0 && kind < 16 && kind % 4 != 2 && ctx->stage && !a.s.pairs && a.s.nseq >= 4 * (1 << (kind / 4)))
    e = launch_row_stage(5 + kind % 4, 32 >> (kind / 4), ROW_LAST, a, ctx->sm_count, st);
  else if (kind == -2) e = launch_row(8, 32, ROW_FIRST, false, a, ctx->sm_count, st);
  else if (kind == -1) e = launch_row(8, 32, ROW_MID, false, a, ctx->sm_count, st);
  else if (kind >= 0 && kind < 16) e = launch_row(5 + kind % 4, 32 >> (kind / 4), ROW_LAST, false, a, ctx->sm_count, st);
  else return fail(ctx, DCPGPU_EINVAL, "bad segment kind");
  CU(e);
  ctx->launches += 1;
  return 0;
}

struct SegLaunch
{
  int kind, level;
  size_t off, count; // range of the entry lists
};

struct SegChunk
{
  std::vector<SegLaunch> launches;
  int levels = 0;
};

struct SegPlan
{
  bool grid = false; // entries are profiles (x nseq reads) or explicit pairs
  std::vector<SegChunk> chunks; // each fits the boundary-column budget; they run one after the other
  size_t stride = 0;
};

// entries[r] = {profile id, key}: key = the profile id again (grid) or the pair index (pairs);
// inside a chunk the entry's position is its column rank.  Builds and uploads the per-launch
// lists.
int prepare_segments(dcpgpu_ctx *ctx, bool grid, std::vector<std::pair<int, long long>> const &entries, size_t per_entry,
                     int maxlen, SegPlan *plan)
{
  plan->grid = grid;
  plan->stride = (size_t)std::min(std::max(maxlen, 1), DCPGPU_MAX_WINDOW) + 3; // rows 0..L, read two rows ahead
  size_t chunk_entries = entries.size();
  {
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    size_t const have = ctx->col_cap * sizeof(Mail);
    size_t const budget = std::min<size_t>((fr + have) / 2, ctx->col_budget) / (plan->stride * sizeof(Mail));
    chunk_entries = std::min(chunk_entries, std::max<size_t>(budget / std::max<size_t>(per_entry, 1), 1));
  }
  size_t const ncols = chunk_entries * per_entry;
  if (ncols * plan->stride > ctx->col_cap)
  {
    if (ctx->d_col) CU(cudaFree(ctx->d_col));
    ctx->d_col = nullptr;
    ctx->col_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_col), ncols * plan->stride * sizeof(Mail)));
    ctx->col_cap = ncols * plan->stride;
  }
  std::vector<int> list;
  std::vector<long long> order, colmap;
  for (size_t e0 = 0; e0 < entries.size(); e0 += chunk_entries)
  {
    size_t const e1 = std::min(entries.size(), e0 + chunk_entries);
    SegChunk chunk;
    // bucket the ranks by (level, kind)
    std::vector<std::vector<size_t>> full(dcpgpu_ctx::MAXSEG), tail((size_t)dcpgpu_ctx::MAXSEG * 16);
    for (size_t r = e0; r < e1; ++r)
    {
      int const p = entries[r].first, nseg = ctx->h_seg_count[(size_t)p], first = ctx->h_seg_first[(size_t)p];
      if (first < 0 || nseg < 2 || nseg > dcpgpu_ctx::MAXSEG)
        return fail(ctx, DCPGPU_ESTATE, "segments: profile is not segmented");
      for (int lv = 0; lv + 1 < nseg; ++lv) full[(size_t)lv].push_back(r);
      tail[(size_t)(nseg - 1) * 16 + (size_t)tail_class(ctx->h_segs[(size_t)(first + nseg - 1)])].push_back(r);
      chunk.levels = std::max(chunk.levels, nseg);
    }
    auto add = [&](int kind, int level, std::vector<size_t> const &ranks) {
      if (ranks.empty()) return;
      chunk.launches.push_back(SegLaunch{kind, level, colmap.size(), ranks.size()});
      for (size_t r : ranks)
      {
        list.push_back(entries[r].first);
        order.push_back(entries[r].second);
        colmap.push_back((long long)(r - e0));
      }
    };
    for (int lv = 0; lv < chunk.levels; ++lv)
    {
      for (int tc = 0; tc < 16; ++tc) add(tc, lv, tail[(size_t)lv * 16 + (size_t)tc]);
      add(lv == 0 ? -2 : -1, lv, full[(size_t)lv]);
    }
    plan->chunks.push_back(std::move(chunk));
  }
  int rc;
  size_t const m = colmap.size();
  if ((rc = ensure(ctx, ctx->d_seg_list, ctx->seg_list_cap, m))) return rc;
  if ((rc = ensure(ctx, ctx->d_seg_order, ctx->seg_order_cap, m))) return rc;
  if ((rc = ensure(ctx, ctx->d_seg_colmap, ctx->seg_colmap_cap, m))) return rc;
  CU(counted_copy(ctx, ctx->d_seg_list, list.data(), m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_seg_order, order.data(), m * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_seg_colmap, colmap.data(), m * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream)); // the vectors are locals
  return 0;
}

// Called with the side streams forked.  The full segments of level s run in order on one side
// stream; the tails of level s (they only need the full segments up to s-1) on the other side
// streams.  Chunks share the column buffer, so the streams are joined between them.
int launch_segments(dcpgpu_ctx *ctx, SegPlan const &plan, ScoreArgs const &a0)
{
  cudaStream_t const chain = ctx->side[0];
  int rc;
  for (size_t ci = 0; ci < plan.chunks.size(); ++ci)
  {
    SegChunk const &chunk = plan.chunks[ci];
    if (ci > 0)
    {
      if ((rc = join_streams(ctx))) return rc;
    }
    CU(cudaMemsetAsync(ctx->d_seg_cursor, 0, dcpgpu_ctx::MAXSEG * 17 * sizeof(unsigned long long), ctx->stream));
    if ((rc = fork_streams(ctx))) return rc; // (re)fork: the side streams see the zeroed cursors
    int next_side = 1;
    size_t i = 0;
    for (int lv = 0; lv < chunk.levels; ++lv)
    {
      for (; i < chunk.launches.size() && chunk.launches[i].level == lv; ++i)
      {
        SegLaunch const &L = chunk.launches[i];
        StripArgs sa{};
        sa.s = a0;
        sa.redo = ctx->d_redo;
        sa.nredo = ctx->d_counters + SLOT_NREDO;
        sa.col = ctx->d_col;
        sa.col_stride = plan.stride;
        sa.strip = 0;
        sa.item0 = 0;
        sa.segs = ctx->d_segs;
        sa.seg_first = ctx->d_seg_first;
        sa.level = lv;
        sa.colmap = ctx->d_seg_colmap + L.off;
        sa.s.counter = ctx->d_seg_cursor + (size_t)lv * 17 + (size_t)(L.kind < 0 ? 16 : L.kind);
        if (plan.grid)
        {
          sa.s.class_profiles = ctx->d_seg_list + L.off;
          sa.s.nitems = (unsigned long long)L.count * (unsigned long long)a0.nseq;
        }
        else
        {
          sa.s.order = ctx->d_seg_order + L.off;
          sa.s.nitems = L.count;
        }
        cudaStream_t st = chain;
        if (L.kind >= 0)
        { // a tail: any other side stream, after the previous level's full segments
          st = ctx->side[next_side];
          next_side = next_side % (dcpgpu_ctx::NSIDE - 1) + 1;
          CU(cudaStreamWaitEvent(st, ctx->ev_level[lv - 1], 0));
        }
        if ((rc = launch_segment(ctx, L.kind, sa, st))) return rc;
        if (L.kind < 0) CU(cudaEventRecord(ctx->ev_level[lv], chain));
      }
    }
  }
  return 0;
}

// Pairs whose speculation failed (d_redo[0..n)): run them on the exact multi-warp kernels.
// `to_pair` maps a result index to the window it stands for.
template <class F>
int redo_exact(dcpgpu_ctx *ctx, uint32_t flags, F &&to_pair)
{
  unsigned long long n = 0;
  CU(counted_copy(ctx, &n, ctx->d_counters + SLOT_NREDO, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->last_redo = (int64_t)n;
  if (n == 0) return 0;
  std::vector<long long> redo((size_t)n);
  CU(counted_copy(ctx, redo.data(), ctx->d_redo, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::sort(redo.begin(), redo.end());
  std::vector<Pair> rp((size_t)n);
  std::vector<std::vector<long long>> bucket(NCLASS);
  for (size_t i = 0; i < (size_t)n; ++i)
  {
    rp[i] = to_pair(redo[i]);
    bucket[(size_t)kernel_class(ctx, rp[i].profile)].push_back((long long)i);
  }
  std::vector<long long> order, out_index;
  size_t first[NCLASS + 1];
  for (int c = 0; c < NCLASS; ++c)
  {
    first[c] = order.size();
    for (long long i : bucket[(size_t)c])
    {
      order.push_back(i);
      out_index.push_back(redo[(size_t)i]);
    }
  }
  first[NCLASS] = order.size();
  int rc;
  if ((rc = ensure(ctx, ctx->d_redo_pairs, ctx->redo_pairs_cap, (size_t)n))) return rc;
  if ((rc = ensure(ctx, ctx->d_redo_order, ctx->redo_order_cap, (size_t)n))) return rc;
  if ((rc = ensure(ctx, ctx->d_redo_out, ctx->redo_out_cap, (size_t)n))) return rc;
  CU(counted_copy(ctx, ctx->d_redo_pairs, rp.data(), (size_t)n * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_redo_order, order.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_redo_out, out_index.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_counters, 0, NCLASS * sizeof(unsigned long long), ctx->stream));
  if ((rc = fork_streams(ctx))) return rc;
  for (int c = 9; c <= 20; ++c)
  {
    size_t const m = first[c + 1] - first[c];
    if (!m) continue;
    ScoreArgs a{};
    a.profiles = ctx->d_profiles;
    a.reads = reads_view(ctx);
    a.xt = ctx->d_xt[flags & 3u];
    a.pairs = ctx->d_redo_pairs;
    a.order = ctx->d_redo_order + first[c];
    a.out_index = ctx->d_redo_out + first[c];
    a.nitems = m;
    a.counter = ctx->d_counters + c;
    a.out = ctx->d_out;
    a.nhits = ctx->d_counters + SLOT_NHITS;
    if ((rc = launch_class(ctx, c, a))) return rc;
  }
  if ((rc = join_streams(ctx))) return rc;
  CU(cudaStreamSynchronize(ctx->stream)); // host vectors above
  return 0;
}

// floats of scratch a generic-kernel launch over `nitems` pairs needs, and its grid
template <bool TRACE>
int generic_plan(dcpgpu_ctx *ctx, unsigned long long nitems, int max_K, unsigned *grid, size_t *floats)
{
  int const KG = (max_K + 31) & ~31;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic_kernel<TRACE>, GEN_THREADS, 0));
  if (per_sm < 1) per_sm = 1;
  unsigned long long const want = (nitems + GEN_WARPS - 1) / GEN_WARPS;
  *grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)per_sm * ctx->sm_count);
  *floats = (size_t)19 * KG * GEN_WARPS * *grid;
  return 0;
}

template <bool TRACE>
int launch_generic(dcpgpu_ctx *ctx, GenArgs a, int max_K)
{
  unsigned grid = 0;
  size_t need = 0;
  int rc = generic_plan<TRACE>(ctx, a.s.nitems, max_K, &grid, &need);
  if (rc) return rc;
  if ((rc = ensure(ctx, ctx->d_scratch, ctx->scratch_cap, need))) return rc;
  a.scratch = ctx->d_scratch;
  a.scratch_stride = (size_t)19 * ((max_K + 31) & ~31);
  generic_kernel<TRACE><<<grid, GEN_THREADS, 0, launch_stream(ctx)>>>(a);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return 0;
}

template <int NW>
int trace_cta_plan(dcpgpu_ctx *ctx, unsigned long long nitems, int max_K, unsigned *grid, size_t *floats)
{
  int const KG = (max_K + 31) & ~31;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_cta_kernel<NW>, 32 * NW, 0));
  if (per_sm < 1) per_sm = 1;
  *grid = (unsigned)std::min<unsigned long long>(nitems, (unsigned long long)per_sm * ctx->sm_count);
  *floats = (size_t)19 * KG * *grid;
  return 0;
}

template <int NW>
int launch_trace_cta(dcpgpu_ctx *ctx, GenArgs a, int max_K, unsigned grid, float *scratch)
{
  a.scratch = scratch;
  a.scratch_stride = (size_t)19 * ((max_K + 31) & ~31);
  trace_cta_kernel<NW><<<grid, 32 * NW, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return 0;
}

ReadsView reads_view(dcpgpu_ctx const *ctx)
{
  ReadsView r;
  r.words = ctx->d_words;
  r.hist = ctx->d_hist;
  r.seq_word = ctx->d_seq_word;
  r.seq_len = ctx->d_seq_len;
  r.nseq = ctx->nseq;
  r.eight = 8;
  r.nwords = ctx->nwords;
  return r;
}

int begin_pass(dcpgpu_ctx *ctx, size_t npairs)
{
  int rc = sync_profiles(ctx);
  if (rc) return rc;
  if ((rc = ensure(ctx, ctx->d_out, ctx->out_cap, npairs))) return rc;
  CU(cudaMemsetAsync(ctx->d_counters, 0, NSLOTS * sizeof(unsigned long long), ctx->stream));
  ctx->forked = false;
  ctx->pinned = nullptr;
  ctx->last_cells = 0;
  ctx->last_redo = 0;
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}

int end_pass(dcpgpu_ctx *ctx, int64_t npairs)
{
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->last_n = npairs;
  return 0;
}

} // namespace

template <int MODE>
static int run_alu_peak(dcpgpu_ctx *ctx, double *tops)
{
  int const iters = 4096, threads = 256;
  int const blocks = ctx->sm_count * 8;
  float *d = nullptr;
  CU(cudaMalloc(reinterpret_cast<void **>(&d), (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
  {
    cudaEventRecord(a, ctx->stream);
    alu_peak_kernel<MODE><<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0f + rep);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep && ms < best) best = ms;
    ctx->launches += 1;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaError_t e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "alu_peak");
  // fp32 lane-operations per iteration and thread.  A three-input min counts as two
  // operations (it replaces two two-input mins of the recurrence), an f32x2 add as two.
  double per_iter = 16.0;                 // modes 0, 1, 2: 16 instructions, 16 ops
  if (MODE == 3) per_iter = 32.0;         // 16 FMNMX3
  if (MODE == 4) per_iter = 32.0;         // 16 FADD2
  if (MODE == 5) per_iter = 32.0;         // 16 FADD + 8 FMNMX3
  if (MODE == 6) per_iter = 32.0;         // 16 FADD + 8 VIMNMX3
  double const ops = per_iter * iters * (double)blocks * threads;
  *tops = ops / (best * 1e-3) / 1e12;
  return 0;
}


// ---- C ABI -------------------------------------------------------------------------------

extern "C" {

char const *dcpgpu_strerror(int code)
{
  switch (code)
  {
  case DCPGPU_OK: return "ok";
  case DCPGPU_ENODEVICE: return "no CUDA device available (deciphon_b200 has no CPU fallback)";
  case DCPGPU_ECUDA: return "CUDA call failed";
  case DCPGPU_ENOMEM: return "out of memory";
  case DCPGPU_EINVAL: return "invalid argument";
  case DCPGPU_ESTATE: return "call made in the wrong state";
  case DCPGPU_EDECODE: return "could not decode a fragment into a codon";
  default: return "unknown dcpgpu error";
  }
}

char const *dcpgpu_last_error(dcpgpu_ctx const *ctx) { return ctx ? ctx->err.c_str() : ""; }

int dcpgpu_open(dcpgpu_ctx **out, int device)
{
  if (!out) return DCPGPU_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return DCPGPU_ENODEVICE;
  if (device < 0 || device >= ndev) return DCPGPU_EINVAL;
  dcpgpu_ctx *ctx = new (std::nothrow) dcpgpu_ctx;
  if (!ctx) return DCPGPU_ENOMEM;
  ctx->device = device;
  {
    char const *v = std::getenv("DCPGPU_SUBWARP");
    ctx->subwarp = !(v && v[0] == '0');
    if ((v = std::getenv("DCPGPU_COL_BUDGET_MB")) && std::atoll(v) > 0) ctx->col_budget = (size_t)std::atoll(v) << 20;
    if ((v = std::getenv("DCPGPU_LZ_SLACK"))) ctx->lz_slack = std::atoll(v);
    if ((v = std::getenv("DCPGPU_STAGE"))) ctx->stage = std::atoi(v);
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  for (int i = 0; i < dcpgpu_ctx::NSIDE; ++i)
  {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < dcpgpu_ctx::MAXSEG; ++i)
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_level[i], cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_seg_cursor), dcpgpu_ctx::MAXSEG * 17 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_counters), NSLOTS * sizeof(unsigned long long));
  if (e != cudaSuccess)
  {
    dcpgpu_close(ctx); // destroys whatever was created (every handle starts out null)
    return DCPGPU_ECUDA;
  }
  *out = ctx;
  return 0;
}

void dcpgpu_close(dcpgpu_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto &s : ctx->slabs) cudaFree(s.base);
  for (auto &b : ctx->pool) { cudaFree(b.em); cudaFree(b.trans); }
  cudaFree(ctx->d_profiles);
  cudaFree(ctx->d_stage);
  cudaFree(ctx->d_words);
  cudaFree(ctx->d_hist);
  cudaFree(ctx->d_seq_word);
  cudaFree(ctx->d_seq_len);
  for (auto p : ctx->d_xt) cudaFree(p);
  cudaFree(ctx->d_out);
  cudaFree(ctx->d_counters);
  cudaFree(ctx->d_class_profiles);
  cudaFree(ctx->d_pairs);
  cudaFree(ctx->d_order);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_tpairs);
  cudaFree(ctx->d_decoders);
  cudaFree(ctx->d_m_hit); cudaFree(ctx->d_m_start); cudaFree(ctx->d_m_stop); cudaFree(ctx->d_m_begin); cudaFree(ctx->d_m_end);
  cudaFree(ctx->d_m_bad); cudaFree(ctx->d_m_len); cudaFree(ctx->d_m_off); cudaFree(ctx->d_m_codon); cudaFree(ctx->d_m_amino);
  cudaFree(ctx->d_m_text);
  cudaFree(ctx->d_xnodes);
  cudaFree(ctx->d_nodes);
  cudaFree(ctx->d_xnode_off);
  cudaFree(ctx->d_node_off);
  cudaFree(ctx->d_step_off);
  cudaFree(ctx->d_lz_ids);
  cudaFree(ctx->d_lz_sz);
  cudaFree(ctx->d_lz_off);
  cudaFree(ctx->d_nsteps);
  cudaFree(ctx->d_tout);
  cudaFree(ctx->d_step_ids);
  cudaFree(ctx->d_dump);
  cudaFree(ctx->d_hit_idx);
  cudaFree(ctx->d_redo);
  cudaFree(ctx->d_segs);
  cudaFree(ctx->d_seg_first);
  cudaFree(ctx->d_seg_cursor);
  cudaFree(ctx->d_seg_list);
  cudaFree(ctx->d_seg_order);
  cudaFree(ctx->d_seg_colmap);
  for (auto e : ctx->ev_level)
    if (e) cudaEventDestroy(e);
  cudaFree(ctx->d_col);
  cudaFree(ctx->d_redo_pairs);
  cudaFree(ctx->d_redo_order);
  cudaFree(ctx->d_redo_out);
  cudaFree(ctx->d_dump_off);
  cudaFree(ctx->d_tile_off);
  cudaFree(ctx->d_step_sz);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int i = 0; i < dcpgpu_ctx::NSIDE; ++i)
  {
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int dcpgpu_set_stream(dcpgpu_ctx *ctx, void *cuda_stream)
{
  if (!ctx) return DCPGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) CU(cudaStreamDestroy(ctx->stream));
  if (cuda_stream)
  {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
  }
  else
  {
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return 0;
}

int dcpgpu_sync(dcpgpu_ctx *ctx)
{
  if (!ctx) return DCPGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int32_t dcpgpu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int64_t dcpgpu_device_info(dcpgpu_ctx const *ctx, int what)
{
  if (!ctx) return -1;
  size_t fr = 0, tot = 0;
  switch (what)
  {
  case 0: return ctx->sm_count;
  case 1:
  case 2:
    cudaSetDevice(ctx->device);
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return -1;
    return (int64_t)(what == 1 ? tot : fr);
  case 3: return (int64_t)ctx->profile_bytes;
  default: return -1;
  }
}

int dcpgpu_pool_add(dcpgpu_ctx *ctx, int nnodes, float const *emission, float const *trans,
                    int64_t *first_node_id)
{
  if (!ctx || nnodes <= 0 || !emission || !trans) return fail(ctx, DCPGPU_EINVAL, "pool_add: bad argument");
  CU(cudaSetDevice(ctx->device));
  PoolBlock b{nullptr, nullptr, ctx->pool_nodes, nnodes};
  size_t const eb = (size_t)nnodes * NCODES * sizeof(float), tb = (size_t)nnodes * 7 * sizeof(float);
  CU(cudaMalloc(reinterpret_cast<void **>(&b.em), eb));
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&b.trans), tb);
  if (e != cudaSuccess)
  {
    cudaFree(b.em);
    return fail_cuda(ctx, e, "cudaMalloc(pool trans)");
  }
  e = counted_copy(ctx, b.em, emission, eb, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = counted_copy(ctx, b.trans, trans, tb, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess)
  { // the block never becomes part of the pool
    cudaFree(b.em);
    cudaFree(b.trans);
    return fail_cuda(ctx, e, "pool_add: upload");
  }
  ctx->pool.push_back(b);
  ctx->pool_nodes += nnodes;
  if (first_node_id) *first_node_id = b.first;
  return 0;
}

int dcpgpu_pool_release(dcpgpu_ctx *ctx)
{
  if (!ctx) return DCPGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (auto &b : ctx->pool)
  {
    cudaFree(b.em);
    cudaFree(b.trans);
  }
  ctx->pool.clear(); // node ids keep growing so that stale ids stay invalid
  return 0;
}

int dcpgpu_profile_add(dcpgpu_ctx *ctx, int K, int64_t const *node_ids, int64_t first_node_id,
                       float const *BMk, float const *null_emission, float const *bg_emission,
                       int32_t *profile_index)
{
  if (!ctx || K < 1 || K > DCPGPU_MAX_CORE_SIZE || !BMk || !null_emission || !bg_emission)
    return fail(ctx, DCPGPU_EINVAL, "profile_add: bad argument");
  CU(cudaSetDevice(ctx->device));

  // resolve pool node ids to device rows
  std::vector<NodeRef> refs((size_t)K);
  size_t hint = 0;
  for (int k = 0; k < K; ++k)
  {
    int64_t const id = node_ids ? node_ids[k] : first_node_id + k;
    PoolBlock const *blk = nullptr;
    if (hint < ctx->pool.size() && id >= ctx->pool[hint].first && id < ctx->pool[hint].first + ctx->pool[hint].n)
      blk = &ctx->pool[hint];
    else
      for (size_t i = 0; i < ctx->pool.size(); ++i)
        if (id >= ctx->pool[i].first && id < ctx->pool[i].first + ctx->pool[i].n)
        {
          blk = &ctx->pool[i];
          hint = i;
          break;
        }
    if (!blk) return fail(ctx, DCPGPU_EINVAL, "profile_add: node id not in the pool");
    refs[k].em = blk->em + (size_t)(id - blk->first) * NCODES;
    refs[k].trans = blk->trans + (size_t)(id - blk->first) * 7;
  }

  ProfileDesc d;
  d.K = K;
  layout_shape(K, &d.Q, &d.W, &d.VL, ctx->subwarp);
  d.Kpad = d.VL * d.Q;
  d.Kfull = K;
  void *pem = nullptr, *pcore = nullptr, *pnb = nullptr;
  int rc;
  if ((rc = arena_alloc(ctx, (size_t)NCODES * d.Kpad * sizeof(float), &pem))) return rc;
  if ((rc = arena_alloc(ctx, (size_t)C_ROWS * d.Kpad * sizeof(float), &pcore))) return rc;
  if ((rc = arena_alloc(ctx, (size_t)NCODES * sizeof(float2), &pnb))) return rc;
  d.em = static_cast<float *>(pem);
  d.core = static_cast<float *>(pcore);
  d.nulbg = static_cast<float2 *>(pnb);

  // staging: refs | BMk | null | bg
  size_t const o_refs = 0;
  size_t const o_bm = (o_refs + (size_t)K * sizeof(NodeRef) + 255) & ~size_t(255);
  size_t const o_nul = (o_bm + (size_t)K * sizeof(float) + 255) & ~size_t(255);
  size_t const o_bg = o_nul + (size_t)NCODES * sizeof(float);
  size_t const total = o_bg + (size_t)NCODES * sizeof(float);
  if ((rc = ensure(ctx, ctx->d_stage, ctx->d_stage_cap, total))) return rc;
  CU(counted_copy(ctx, ctx->d_stage + o_refs, refs.data(), (size_t)K * sizeof(NodeRef), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_stage + o_bm, BMk, (size_t)K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_stage + o_nul, null_emission, NCODES * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_stage + o_bg, bg_emission, NCODES * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));

  NodeRef const *drefs = reinterpret_cast<NodeRef const *>(ctx->d_stage + o_refs);
  int const VL = d.VL;
  dim3 const grid((d.Kpad + 127) / 128, NCODES);
  int *d_bad = reinterpret_cast<int *>(ctx->d_counters + SLOT_BAD);
  CU(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  pack_em_kernel<<<grid, 128, 0, ctx->stream>>>(drefs, 0, K, d.Q, VL, d.Kpad, static_cast<float *>(pem), d_bad);
  pack_core_kernel<<<(d.Kpad + 127) / 128, 128, 0, ctx->stream>>>(
      drefs, reinterpret_cast<float const *>(ctx->d_stage + o_bm), 0, K, K, d.Q, VL, d.Kpad, static_cast<float *>(pcore), d_bad);
  pack_nulbg_kernel<<<(NCODES + 127) / 128, 128, 0, ctx->stream>>>(
      reinterpret_cast<float const *>(ctx->d_stage + o_nul), reinterpret_cast<float const *>(ctx->d_stage + o_bg),
      static_cast<float2 *>(pnb), d_bad);
  CU(cudaGetLastError());
  // Profiles the strip kernels run are also stored as segments: 256 nodes each in the Q = 8
  // full-warp layout, the tail in the layout a profile of its size would get (strip_kernel.cuh).
  int seg_first = -1, seg_count = 0;
  if (d.W > 1 && d.W <= 8 && d.Q <= MAXQ_REG) // exactly the strip classes 9..20
  {
    seg_first = (int)ctx->h_segs.size();
    for (int k0 = 0; k0 < K; k0 += 256)
    {
      ProfileDesc g;
      g.K = std::min(256, K - k0);
      g.Kfull = K;
      if (k0 + 256 < K) { g.Q = 8; g.W = 1; g.VL = 32; }
      else layout_shape(g.K, &g.Q, &g.W, &g.VL, true);
      g.Kpad = g.VL * g.Q;
      void *sem = nullptr, *score = nullptr;
      if ((rc = arena_alloc(ctx, (size_t)NCODES * g.Kpad * sizeof(float), &sem))) return rc;
      if ((rc = arena_alloc(ctx, (size_t)C_ROWS * g.Kpad * sizeof(float), &score))) return rc;
      g.em = static_cast<float *>(sem);
      g.core = static_cast<float *>(score);
      g.nulbg = d.nulbg;
      dim3 const sgrid((g.Kpad + 127) / 128, NCODES);
      pack_em_kernel<<<sgrid, 128, 0, ctx->stream>>>(drefs, k0, g.K, g.Q, g.VL, g.Kpad, static_cast<float *>(sem), d_bad);
      pack_core_kernel<<<(g.Kpad + 127) / 128, 128, 0, ctx->stream>>>(
          drefs, reinterpret_cast<float const *>(ctx->d_stage + o_bm), k0, g.K, K, g.Q, g.VL, g.Kpad,
          static_cast<float *>(score), d_bad);
      CU(cudaGetLastError());
      ctx->launches += 2;
      ctx->h_segs.push_back(g);
      ++seg_count;
    }
  }
  int h_bad = 0;
  CU(counted_copy(ctx, &h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->launches += 3;
  // the staging buffer is reused by the next call: finish the pack first
  CU(cudaStreamSynchronize(ctx->stream));

  ctx->h_profiles.push_back(d);
  ctx->h_unsafe.push_back(h_bad != 0);
  ctx->h_seg_first.push_back(h_bad ? -1 : seg_first);
  ctx->h_seg_count.push_back(h_bad ? 0 : seg_count);
  ctx->profiles_dirty = true;
  if (profile_index) *profile_index = (int32_t)ctx->h_profiles.size() - 1;
  return 0;
}

int dcpgpu_profile_count(dcpgpu_ctx const *ctx) { return ctx ? (int)ctx->h_profiles.size() : 0; }

int dcpgpu_profile_core_size(dcpgpu_ctx const *ctx, int32_t profile)
{
  if (!ctx || profile < 0 || (size_t)profile >= ctx->h_profiles.size()) return -1;
  return ctx->h_profiles[(size_t)profile].K;
}

namespace dcp {
// Ten-bit history of every position of the packed read stream: hist[g] = nucleotides g-4..g, the
// most recent in the low two bits (code of the t-mer ending at g = off[t] + (hist[g] & (4^t - 1))).
// Positions before a window's start only ever meet +INF states, so no per-sequence logic is needed.
__global__ void hist_kernel(uint32_t const *__restrict__ words, long long nwords, uint16_t *__restrict__ hist,
                            long long nhist)
{
  long long const g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= nhist) return;
  unsigned h = 0;
  if (g < nwords * 16)
  {
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
      long long const j = g - i;
      if (j >= 0) h |= ((__ldg(words + (j >> 4)) >> (2 * (int)(j & 15))) & 3u) << (2 * i);
    }
  }
  hist[g] = (uint16_t)h;
}

} // namespace dcp

int dcpgpu_reads_set(dcpgpu_ctx *ctx, int32_t nseq, uint8_t const *symbols, int64_t const *offsets)
{
  if (!ctx || nseq < 0 || (nseq > 0 && (!symbols || !offsets))) return fail(ctx, DCPGPU_EINVAL, "reads_set: bad argument");
  CU(cudaSetDevice(ctx->device));
  std::vector<long long> seq_word((size_t)nseq);
  std::vector<int> seq_len((size_t)nseq);
  long long nwords = 0;
  int maxlen = 0;
  for (int s = 0; s < nseq; ++s)
  {
    int64_t const len = offsets[s + 1] - offsets[s];
    if (len < 0 || len > INT32_MAX) return fail(ctx, DCPGPU_EINVAL, "reads_set: bad offsets");
    seq_word[s] = nwords;
    seq_len[s] = (int)len;
    nwords += (len + 15) / 16 + 1; // +1: the kernels prefetch one word past the end
    maxlen = std::max(maxlen, (int)len);
  }
  std::vector<uint32_t> words((size_t)nwords + 1, 0u);
  for (int s = 0; s < nseq; ++s)
  {
    uint8_t const *x = symbols + offsets[s];
    uint32_t *w = words.data() + seq_word[s];
    for (int i = 0; i < seq_len[s]; ++i)
    {
      if (x[i] > 3) return fail(ctx, DCPGPU_EINVAL, "reads_set: symbol outside 0..3");
      w[i >> 4] |= (uint32_t)x[i] << (2 * (i & 15));
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_words) CU(cudaFree(ctx->d_words));
  if (ctx->d_hist) CU(cudaFree(ctx->d_hist));
  ctx->d_hist = nullptr;
  if (ctx->d_seq_word) CU(cudaFree(ctx->d_seq_word));
  if (ctx->d_seq_len) CU(cudaFree(ctx->d_seq_len));
  ctx->d_words = nullptr;
  ctx->d_seq_word = nullptr;
  ctx->d_seq_len = nullptr;
  CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_words), words.size() * sizeof(uint32_t)));
  CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_seq_word), std::max<size_t>(1, seq_word.size()) * sizeof(long long)));
  CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_seq_len), std::max<size_t>(1, seq_len.size()) * sizeof(int)));
  CU(counted_copy(ctx, ctx->d_words, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  {
    long long const nhist = (long long)words.size() * 16 + HIST_SLACK;
    CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_hist), (size_t)nhist * sizeof(uint16_t)));
    hist_kernel<<<(unsigned)((nhist + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_words, (long long)words.size(), ctx->d_hist, nhist);
    CU(cudaGetLastError());
    ctx->launches += 1;
  }
  if (nseq)
  {
    CU(counted_copy(ctx, ctx->d_seq_word, seq_word.data(), seq_word.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(counted_copy(ctx, ctx->d_seq_len, seq_len.data(), seq_len.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->h_seq_len.swap(seq_len);
  ctx->nseq = nseq;
  ctx->nwords = (long long)words.size();
  ctx->maxlen = std::min(maxlen, DCPGPU_MAX_WINDOW);
  return 0;
}

int dcpgpu_reads_count(dcpgpu_ctx const *ctx) { return ctx ? ctx->nseq : 0; }

static int check_pairs(dcpgpu_ctx *ctx, int64_t npairs, dcpgpu_pair const *pairs, int *maxlen)
{
  int ml = 1;
  for (int64_t i = 0; i < npairs; ++i)
  {
    dcpgpu_pair const &p = pairs[i];
    if (p.profile < 0 || (size_t)p.profile >= ctx->h_profiles.size() || p.seq < 0 || p.seq >= ctx->nseq ||
        p.start < 0 || p.len < 1 || p.len > DCPGPU_MAX_WINDOW ||
        (int64_t)p.start + p.len > ctx->h_seq_len[(size_t)p.seq])
      return fail(ctx, DCPGPU_EINVAL, "pair out of range");
    ml = std::max(ml, p.len);
  }
  *maxlen = ml;
  return 0;
}

int dcpgpu_score_pairs(dcpgpu_ctx *ctx, int64_t npairs, dcpgpu_pair const *pairs, uint32_t flags,
                       float *null_cost, float *alt_cost)
{
  if (!ctx || npairs < 0 || (npairs && !pairs)) return fail(ctx, DCPGPU_EINVAL, "score_pairs: bad argument");
  CU(cudaSetDevice(ctx->device));
  int rc, maxlen = 1;
  if ((rc = check_pairs(ctx, npairs, pairs, &maxlen))) return rc;
  if ((rc = ensure_xt(ctx, flags, maxlen))) return rc;
  if ((rc = begin_pass(ctx, (size_t)npairs))) return rc;
  if (npairs == 0) return end_pass(ctx, 0);

  // bucket by kernel class, keeping the caller's order inside a class
  std::vector<std::vector<long long>> bucket(NCLASS);
  int maxK_generic = 1;
  double cells = 0;
  for (int64_t i = 0; i < npairs; ++i)
  {
    ProfileDesc const &d = ctx->h_profiles[(size_t)pairs[i].profile];
    int const c = kernel_class(ctx, pairs[i].profile);
    bucket[(size_t)c].push_back(i);
    if (c == 0) maxK_generic = std::max(maxK_generic, d.K);
    cells += (double)pairs[i].len * d.K;
  }
  std::vector<long long> order;
  order.reserve((size_t)npairs);
  size_t first[NCLASS + 1];
  for (int c = 0; c < NCLASS; ++c)
  {
    first[c] = order.size();
    order.insert(order.end(), bucket[(size_t)c].begin(), bucket[(size_t)c].end());
  }
  first[NCLASS] = order.size();

  if ((rc = ensure(ctx, ctx->d_pairs, ctx->pairs_cap, (size_t)npairs))) return rc;
  if ((rc = ensure(ctx, ctx->d_order, ctx->order_cap, (size_t)npairs))) return rc;
  if ((rc = ensure(ctx, ctx->d_redo, ctx->redo_cap, (size_t)npairs))) return rc;
  // profiles of more than 256 nodes: segment by segment
  SegPlan seg;
  bool seg_ok = false;
  {
    std::vector<std::pair<int, long long>> entries;
    for (int c = 9; c <= 20; ++c)
      for (long long i : bucket[(size_t)c]) entries.push_back({pairs[i].profile, i});
    if (!entries.empty())
    {
      if ((rc = prepare_segments(ctx, false, entries, 1, maxlen, &seg))) return rc;
      seg_ok = true;
    }
  }
  bool any_strip = false;
  CU(counted_copy(ctx, ctx->d_pairs, pairs, (size_t)npairs * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_order, order.data(), (size_t)npairs * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));

  if ((rc = fork_streams(ctx))) return rc;
  if (seg_ok)
  {
    ScoreArgs a{};
    a.profiles = ctx->d_profiles;
    a.reads = reads_view(ctx);
    a.xt = ctx->d_xt[flags & 3u];
    a.pairs = ctx->d_pairs;
    a.out = ctx->d_out;
    a.nhits = ctx->d_counters + SLOT_NHITS;
    if ((rc = launch_segments(ctx, seg, a))) return rc;
    any_strip = true;
  }
  for (int ci = 0; ci < NCLASS; ++ci)
  {
    int const c = launch_order(ci); // longest pairs first: their tail hides under the rest
    size_t const n = first[c + 1] - first[c];
    if (!n || (c >= 9 && c <= 20)) continue; // 9..20 ran as segments above
    ScoreArgs a{};
    a.profiles = ctx->d_profiles;
    a.reads = reads_view(ctx);
    a.xt = ctx->d_xt[flags & 3u];
    a.pairs = ctx->d_pairs;
    a.order = ctx->d_order + first[c];
    a.nitems = n;
    a.counter = ctx->d_counters + c;
    a.out = ctx->d_out;
    a.nhits = ctx->d_counters + SLOT_NHITS;
    if (c == 0)
    {
      GenArgs g{};
      g.s = a;
      if ((rc = launch_generic<false>(ctx, g, maxK_generic))) return rc;
    }
    else if ((rc = launch_class(ctx, c, a)))
      return rc;
  }
  if ((rc = join_streams(ctx))) return rc;
  if (any_strip && (rc = redo_exact(ctx, flags, [&](long long oidx) {
        dcpgpu_pair const &q = pairs[oidx];
        return Pair{q.profile, q.seq, q.start, q.len};
      })))
    return rc;
  ctx->last_cells = cells;
  ctx->total_cells += cells;
  if ((rc = end_pass(ctx, npairs))) return rc;
  return dcpgpu_scores_fetch(ctx, npairs, null_cost, alt_cost);
}

int dcpgpu_score_grid(dcpgpu_ctx *ctx, int32_t prof0, int32_t prof1, int32_t seq0, int32_t seq1, uint32_t flags)
{
  if (!ctx || prof0 < 0 || prof1 < prof0 || (size_t)prof1 > ctx->h_profiles.size() || seq0 < 0 || seq1 < seq0 ||
      seq1 > ctx->nseq)
    return fail(ctx, DCPGPU_EINVAL, "score_grid: bad range");
  CU(cudaSetDevice(ctx->device));
  int rc;
  int const nprof = prof1 - prof0, nseq = seq1 - seq0;
  size_t const npairs = (size_t)nprof * (size_t)nseq;
  if ((rc = ensure_xt(ctx, flags, std::max(ctx->maxlen, 1)))) return rc;
  if ((rc = begin_pass(ctx, npairs))) return rc;
  if (npairs == 0) return end_pass(ctx, 0);

  std::vector<std::vector<int>> bucket(NCLASS);
  int maxK_generic = 1;
  double cells = 0;
  // cells = sum over pairs of min(50K, 100000, len) * K
  std::vector<int> lens(ctx->h_seq_len.begin() + seq0, ctx->h_seq_len.begin() + seq1);
  std::sort(lens.begin(), lens.end());
  std::vector<double> prefix(lens.size() + 1, 0.0);
  for (size_t i = 0; i < lens.size(); ++i) prefix[i + 1] = prefix[i] + lens[i];
  for (int p = prof0; p < prof1; ++p)
  {
    ProfileDesc const &d = ctx->h_profiles[(size_t)p];
    int const c = kernel_class(ctx, p);
    bucket[(size_t)c].push_back(p);
    if (c == 0) maxK_generic = std::max(maxK_generic, d.K);
    int const w = std::min(d.K * 50, DCPGPU_MAX_WINDOW);
    size_t const nle = std::upper_bound(lens.begin(), lens.end(), w) - lens.begin();
    cells += (prefix[nle] + (double)(lens.size() - nle) * w) * d.K;
  }
  std::vector<int> flat;
  flat.reserve((size_t)nprof);
  size_t first[NCLASS + 1];
  for (int c = 0; c < NCLASS; ++c)
  {
    first[c] = flat.size();
    flat.insert(flat.end(), bucket[(size_t)c].begin(), bucket[(size_t)c].end());
  }
  first[NCLASS] = flat.size();
  if ((rc = ensure(ctx, ctx->d_class_profiles, ctx->class_profiles_cap, flat.size()))) return rc;
  {
    size_t strip_items = 0;
    for (int c = 9; c <= 20; ++c) strip_items += (first[c + 1] - first[c]) * (size_t)nseq;
    if ((rc = ensure(ctx, ctx->d_redo, ctx->redo_cap, strip_items))) return rc;
  }
  SegPlan seg;
  bool seg_ok = false;
  {
    std::vector<std::pair<int, long long>> entries;
    for (int c = 9; c <= 20; ++c)
      for (int p : bucket[(size_t)c]) entries.push_back({p, (long long)p});
    if (!entries.empty())
    {
      if ((rc = prepare_segments(ctx, true, entries, (size_t)nseq, ctx->maxlen, &seg))) return rc;
      seg_ok = true;
    }
  }
  bool any_strip = false;
  CU(counted_copy(ctx, ctx->d_class_profiles, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream)); // flat is a local
  CU(cudaEventRecord(ctx->ev0, ctx->stream));

  // longest pairs first (generic, then W = 8 .. 1): their tail hides under the classes that follow
  if ((rc = fork_streams(ctx))) return rc;
  if (seg_ok)
  {
    ScoreArgs a{};
    a.profiles = ctx->d_profiles;
    a.reads = reads_view(ctx);
    a.xt = ctx->d_xt[flags & 3u];
    a.prof0 = prof0;
    a.seq0 = seq0;
    a.nseq = nseq;
    a.out = ctx->d_out;
    a.nhits = ctx->d_counters + SLOT_NHITS;
    if ((rc = launch_segments(ctx, seg, a))) return rc;
    any_strip = true;
  }
  for (int ci = 0; ci < NCLASS; ++ci)
  {
    int const c = launch_order(ci);
    size_t const n = first[c + 1] - first[c];
    if (!n || (c >= 9 && c <= 20)) continue; // 9..20 ran as segments above
    ScoreArgs a{};
    a.profiles = ctx->d_profiles;
    a.reads = reads_view(ctx);
    a.xt = ctx->d_xt[flags & 3u];
    a.class_profiles = ctx->d_class_profiles + first[c];
    a.prof0 = prof0;
    a.seq0 = seq0;
    a.nseq = nseq;
    a.nitems = (unsigned long long)n * (unsigned long long)nseq;
    a.counter = ctx->d_counters + c;
    a.out = ctx->d_out;
    a.nhits = ctx->d_counters + SLOT_NHITS;
    if (c == 0)
    {
      GenArgs g{};
      g.s = a;
      if ((rc = launch_generic<false>(ctx, g, maxK_generic))) return rc;
    }
    else if ((rc = launch_class(ctx, c, a)))
      return rc;
  }
  if ((rc = join_streams(ctx))) return rc;
  if (any_strip && (rc = redo_exact(ctx, flags, [&](long long oidx) {
        int const p = prof0 + (int)(oidx / nseq), sq = seq0 + (int)(oidx % nseq);
        int const w = std::min(ctx->h_profiles[(size_t)p].K * 50, DCPGPU_MAX_WINDOW);
        return Pair{p, sq, 0, std::min(w, ctx->h_seq_len[(size_t)sq])};
      })))
    return rc;
  ctx->last_cells = cells;
  ctx->total_cells += cells;
  return end_pass(ctx, (int64_t)npairs);
}

int dcpgpu_scores_fetch(dcpgpu_ctx *ctx, int64_t npairs, float *null_cost, float *alt_cost)
{
  if (!ctx || npairs < 0 || npairs > ctx->last_n) return fail(ctx, DCPGPU_EINVAL, "scores_fetch: bad count");
  CU(cudaSetDevice(ctx->device));
  if (npairs && (null_cost || alt_cost))
  {
    std::vector<float2> h((size_t)npairs);
    CU(counted_copy(ctx, h.data(), ctx->d_out, (size_t)npairs * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < npairs; ++i)
    {
      if (null_cost) null_cost[i] = h[(size_t)i].x;
      if (alt_cost) alt_cost[i] = h[(size_t)i].y;
    }
  }
  else
    CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dcpgpu_hits_fetch(dcpgpu_ctx *ctx, int64_t cap, int64_t *hit_index, int64_t *nhits)
{
  if (!ctx || cap < 0 || !nhits) return fail(ctx, DCPGPU_EINVAL, "hits_fetch: bad argument");
  CU(cudaSetDevice(ctx->device));
  unsigned long long n = 0;
  CU(counted_copy(ctx, &n, ctx->d_counters + SLOT_NHITS, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *nhits = (int64_t)n;
  if (!hit_index || cap == 0 || n == 0) return 0;
  int rc_e = ensure(ctx, ctx->d_hit_idx, ctx->hit_idx_cap, (size_t)n);
  if (rc_e) return rc_e;
  long long *d_idx = ctx->d_hit_idx;
  CU(cudaMemsetAsync(ctx->d_counters + SLOT_FILL, 0, sizeof(unsigned long long), ctx->stream));
  long long const N = ctx->last_n;
  hits_fill_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_out, N, ctx->d_counters + SLOT_FILL,
                                                                       (long long)n, d_idx);
  ctx->launches += 1;
  std::vector<long long> h((size_t)n);
  cudaError_t e = counted_copy(ctx, h.data(), d_idx, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "hits_fetch");
  std::sort(h.begin(), h.end());
  for (int64_t i = 0; i < std::min<int64_t>(cap, (int64_t)n); ++i)
    hit_index[i] = h[(size_t)i];
  return 0;
}

int dcpgpu_scores_gather(dcpgpu_ctx *ctx, int64_t n, int64_t const *index, float *null_cost, float *alt_cost)
{
  if (!ctx || n < 0 || (n && !index)) return fail(ctx, DCPGPU_EINVAL, "scores_gather: bad argument");
  if (n == 0) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (index[i] < 0 || index[i] >= ctx->last_n) return fail(ctx, DCPGPU_EINVAL, "scores_gather: index out of range");
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure(ctx, ctx->d_hit_idx, ctx->hit_idx_cap, (size_t)n))) return rc;
  if ((rc = ensure(ctx, ctx->d_tout, ctx->tout_cap, (size_t)n))) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "index type");
  CU(counted_copy(ctx, ctx->d_hit_idx, index, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  gather_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_out, ctx->d_hit_idx, (long long)n, ctx->d_tout);
  CU(cudaGetLastError());
  ctx->launches += 1;
  std::vector<float2> h((size_t)n);
  CU(counted_copy(ctx, h.data(), ctx->d_tout, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < n; ++i)
  {
    if (null_cost) null_cost[i] = h[(size_t)i].x;
    if (alt_cost) alt_cost[i] = h[(size_t)i].y;
  }
  return 0;
}

double dcpgpu_last_cells(dcpgpu_ctx const *ctx) { return ctx ? ctx->last_cells : 0.0; }
int64_t dcpgpu_last_redo(dcpgpu_ctx const *ctx) { return ctx ? ctx->last_redo : 0; }
int64_t dcpgpu_launch_count(dcpgpu_ctx const *ctx) { return ctx ? ctx->launches : 0; }

double dcpgpu_counter(dcpgpu_ctx const *ctx, int what)
{
  if (!ctx) return 0.0;
  switch (what)
  {
  case 0: return (double)ctx->h2d_bytes;
  case 1: return (double)ctx->d2h_bytes;
  case 2: return (double)ctx->launches;
  case 3: return ctx->total_cells;
  default: return 0.0;
  }
}

float dcpgpu_last_kernel_ms(dcpgpu_ctx *ctx)
{
  if (!ctx || !ctx->timed) return -1.0f;
  cudaSetDevice(ctx->device);
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

int dcpgpu_trace_pairs(dcpgpu_ctx *ctx, int64_t npairs, dcpgpu_pair const *pairs, uint32_t flags,
                       float *alt_cost, int32_t *nsteps)
{
  if (!ctx || npairs < 0 || (npairs && !pairs)) return fail(ctx, DCPGPU_EINVAL, "trace_pairs: bad argument");
  CU(cudaSetDevice(ctx->device));
  TracePhases ph;
  int rc, maxlen = 1;
  ctx->forked = false; // a failed pass may have left the side streams forked: never route to them unjoined
  ctx->pinned = nullptr;
  ctx->traced = false;
  ctx->matched = false;
  ctx->steps_compact = false;
  if ((rc = check_pairs(ctx, npairs, pairs, &maxlen))) return rc;
  if ((rc = ensure_xt(ctx, flags, maxlen))) return rc;
  if ((rc = sync_profiles(ctx))) return rc;
  ctx->t_pairs.assign(reinterpret_cast<Pair const *>(pairs), reinterpret_cast<Pair const *>(pairs) + npairs);
  ctx->t_xnode_off.assign((size_t)npairs + 1, 0);
  ctx->t_node_off.assign((size_t)npairs + 1, 0);
  ctx->t_nsteps.assign((size_t)npairs, 0);
  if (npairs == 0)
  {
    ctx->traced = true;
    return 0;
  }
  // The full trellis is only materialised on request (DCPGPU_KEEP_TRELLIS) and for profiles the
  // register kernels cannot run; all other pairs are walked straight from the dumped values.
  bool const keep = (flags & DCPGPU_KEEP_TRELLIS) != 0;
  std::vector<long long> slot_off((size_t)npairs + 1, 0); // lazily walked paths: generous slots
  for (int64_t i = 0; i < npairs; ++i)
  {
    int const K = ctx->h_profiles[(size_t)pairs[i].profile].K;
    bool const trellis = keep || kernel_class(ctx, pairs[i].profile) == 0;
    ctx->t_xnode_off[(size_t)i + 1] = ctx->t_xnode_off[(size_t)i] + (trellis ? pairs[i].len + 1 : 0);
    ctx->t_node_off[(size_t)i + 1] = ctx->t_node_off[(size_t)i] + (trellis ? (long long)(pairs[i].len + 1) * K : 0);
    slot_off[(size_t)i + 1] =
        slot_off[(size_t)i] + (trellis ? 0 : std::max<long long>(4, (long long)pairs[i].len + 2 * (long long)K + ctx->lz_slack));
  }
  size_t const n = (size_t)npairs;
  if ((rc = ensure(ctx, ctx->d_tpairs, ctx->tpairs_cap, n))) return rc;
  if ((rc = ensure(ctx, ctx->d_xnodes, ctx->xnodes_cap, (size_t)ctx->t_xnode_off[n]))) return rc;
  if ((rc = ensure(ctx, ctx->d_nodes, ctx->nodes_cap, (size_t)ctx->t_node_off[n]))) return rc;
  if ((rc = ensure(ctx, ctx->d_xnode_off, ctx->xnode_off_cap, n + 1))) return rc;
  if ((rc = ensure(ctx, ctx->d_node_off, ctx->node_off_cap, n + 1))) return rc;
  if ((rc = ensure(ctx, ctx->d_nsteps, ctx->nsteps_cap, n))) return rc;
  if ((rc = ensure(ctx, ctx->d_tout, ctx->tout_cap, n))) return rc;
  if ((rc = ensure(ctx, ctx->d_lz_off, ctx->lz_off_cap, n + 1))) return rc;
  if (slot_off[n])
  {
    if ((rc = ensure(ctx, ctx->d_lz_ids, ctx->lz_ids_cap, (size_t)slot_off[n]))) return rc;
    if ((rc = ensure(ctx, ctx->d_lz_sz, ctx->lz_sz_cap, (size_t)slot_off[n]))) return rc;
  }
  CU(counted_copy(ctx, ctx->d_lz_off, slot_off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_counters + SLOT_OVERFLOW, 0, sizeof(unsigned long long), ctx->stream));
  CU(counted_copy(ctx, ctx->d_tpairs, pairs, n * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_xnode_off, ctx->t_xnode_off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, ctx->d_node_off, ctx->t_node_off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_counters + SLOT_TRACE_CUR, 0, 5 * sizeof(unsigned long long), ctx->stream));
  ph.lap(0, ctx->stream);

  // ---- fast route: pairs whose profile runs on a register kernel -------------------------------
  // value dump by score_reg_kernel<Q,W,DUMP> + parallel argmin kernel (trace_argmin.cuh), in chunks
  // bounded by the dump size (12 bytes per DP cell).
  std::vector<long long> fast, slow;
  for (int64_t i = 0; i < npairs; ++i)
    (kernel_class(ctx, pairs[i].profile) != 0 ? fast : slow).push_back(i);
  if (!fast.empty())
  {
    if ((rc = ensure(ctx, ctx->d_dump_off, ctx->dump_off_cap, n))) return rc;
    // class-major: every launch then holds many pairs of ONE kernel class
    std::vector<std::vector<long long>> by_class(NCLASS);
    for (long long i : fast) by_class[(size_t)kernel_class(ctx, pairs[i].profile)].push_back(i);
    // Size the dump buffer ONCE (a full round, or the largest single pair): reallocating tens of GB
    // between launches costs more than the kernels.  The device is asked for its free memory only
    // when the buffer has to grow (the query itself costs tens of ms on a busy context).
    size_t budget = ctx->dump_cap; // floats a round may use
    {
      size_t single = 0, all = 0;
      for (long long i : fast)
      {
        size_t const f = DumpView::floats(pairs[i].len, ctx->h_profiles[(size_t)pairs[i].profile].Kpad);
        single = std::max(single, f);
        all += f;
      }
      if (all > ctx->dump_cap)
      {
        size_t fr = 0, tot = 0;
        CU(cudaMemGetInfo(&fr, &tot));
        size_t const have = ctx->dump_cap * sizeof(float);
        size_t const room = std::min<size_t>((fr + have) / 2, size_t(24) << 30) / sizeof(float);
        size_t need = std::max(single, std::min(all, room));
        // a trace that large is a burst of hits and bursts come in all sizes: take the whole budget
        // once instead of growing (and re-allocating tens of GB) burst after burst
        if (need > (size_t(1) << 28)) need = std::max(need, room);
        if (need > ctx->dump_cap)
        {
          if (ctx->d_dump) CU(cudaFree(ctx->d_dump));
          ctx->d_dump = nullptr;
          ctx->dump_cap = 0;
          CU(cudaMalloc(reinterpret_cast<void **>(&ctx->d_dump), need * sizeof(float)));
          ctx->dump_cap = need;
        }
        budget = ctx->dump_cap;
      }
    }
    // Rounds bounded by the dump budget; inside a round every kernel class present gets its own
    // side stream (dump kernel, then the walk or argmin kernels behind it), so the low-occupancy
    // launches of the large-profile classes overlap instead of adding up their tails.
    std::vector<long long> sorted; // pairs of the fast route, largest kernel class first
    for (int cls = NCLASS - 1; cls >= 1; --cls)
      sorted.insert(sorted.end(), by_class[(size_t)cls].begin(), by_class[(size_t)cls].end());
    auto cls_of = [&](long long i) { return kernel_class(ctx, pairs[i].profile); };
    auto run_fast = [&]() -> int {
      size_t c0 = 0;
      while (c0 < sorted.size())
      {
        // round [c0, c1)
        size_t c1 = c0, floats = 0;
        std::vector<long long> doff;
        while (c1 < sorted.size())
        {
          dcpgpu_pair const &pr = pairs[sorted[c1]];
          size_t const f = DumpView::floats(pr.len, ctx->h_profiles[(size_t)pr.profile].Kpad);
          if (c1 > c0 && floats + f > budget) break;
          doff.push_back((long long)floats);
          floats += f;
          ++c1;
        }
        size_t const cn = c1 - c0;
        if (floats > ctx->dump_cap) return fail(ctx, DCPGPU_ESTATE, "trace: dump buffer too small (internal error)");
        // groups of one class: [g0, g1) relative to c0; tile prefix sums restart per group
        std::vector<size_t> gstart;
        std::vector<long long> tile_off;
        std::vector<size_t> tstart;
        for (size_t i = 0; i < cn; ++i)
        {
          if (i == 0 || cls_of(sorted[c0 + i]) != cls_of(sorted[c0 + i - 1]))
          {
            gstart.push_back(i);
            tstart.push_back(tile_off.size());
            tile_off.push_back(0);
          }
          tile_off.push_back(tile_off.back() + (pairs[sorted[c0 + i]].len + ARGMIN_ROWS - 1) / ARGMIN_ROWS);
        }
        gstart.push_back(cn);
        if ((rc = ensure(ctx, ctx->d_order, ctx->order_cap, cn))) return rc;
        if ((rc = ensure(ctx, ctx->d_tile_off, ctx->tile_off_cap, tile_off.size()))) return rc;
        CU(counted_copy(ctx, ctx->d_order, sorted.data() + c0, cn * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        CU(counted_copy(ctx, ctx->d_tile_off, tile_off.data(), tile_off.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        // dump offsets are indexed by the item's position in its launch
        CU(counted_copy(ctx, ctx->d_dump_off, doff.data(), cn * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_counters, 0, NCLASS * sizeof(unsigned long long), ctx->stream));
        if ((rc = fork_streams(ctx))) return rc;
        for (size_t gi = 0; gi + 1 < gstart.size(); ++gi)
        {
          size_t const g0 = gstart[gi], gn = gstart[gi + 1] - g0;
          int const cls = cls_of(sorted[c0 + g0]);
          cudaStream_t const st = ctx->side[gi % dcpgpu_ctx::NSIDE];
          ctx->pinned = st; // the group's kernels depend on each other: one stream
          ScoreArgs sa{};
          sa.profiles = ctx->d_profiles;
          sa.reads = reads_view(ctx);
          sa.xt = ctx->d_xt[flags & 3u];
          sa.pairs = ctx->d_tpairs;
          sa.order = ctx->d_order + g0;
          sa.nitems = gn;
          sa.counter = ctx->d_counters + cls;
          sa.out = ctx->d_tout;
          sa.nhits = ctx->d_counters + SLOT_TRACE_NHITS;
          sa.dump = ctx->d_dump;
          sa.dump_off = ctx->d_dump_off + g0;
          rc = launch_class_t<true>(ctx, cls, sa);
          ctx->pinned = nullptr;
          if (rc) return rc;
          if (!keep)
          { // paths only: walk the dumped values directly
            LazyWalkArgs z{};
            z.profiles = ctx->d_profiles;
            z.reads = reads_view(ctx);
            z.xt = ctx->d_xt[flags & 3u];
            z.pairs = ctx->d_tpairs;
            z.order = ctx->d_order + g0;
            z.nitems = (long long)gn;
            z.dump = ctx->d_dump;
            z.dump_off = ctx->d_dump_off + g0;
            z.nsteps = ctx->d_nsteps;
            z.slot_off = ctx->d_lz_off;
            z.overflow = ctx->d_counters + SLOT_OVERFLOW;
            z.ids = ctx->d_lz_ids;
            z.sizes = ctx->d_lz_sz;
            lazy_walk_kernel<<<(unsigned)((gn + LAZY_WARPS - 1) / LAZY_WARPS), 32 * LAZY_WARPS, 0, st>>>(z);
            CU(cudaGetLastError());
            ctx->launches += 1;
            continue;
          }
          long long const *toff = ctx->d_tile_off + tstart[gi];
          ArgminArgs g{};
          g.profiles = ctx->d_profiles;
          g.reads = reads_view(ctx);
          g.xt = ctx->d_xt[flags & 3u];
          g.pairs = ctx->d_tpairs;
          g.order = ctx->d_order + g0;
          g.tile_off = toff;
          g.nitems = (long long)gn;
          g.dump = ctx->d_dump;
          g.dump_off = ctx->d_dump_off + g0;
          g.xnodes = ctx->d_xnodes;
          g.nodes = ctx->d_nodes;
          g.xnode_off = ctx->d_xnode_off;
          g.node_off = ctx->d_node_off;
          trace_argmin_kernel<<<(unsigned)tile_off[tstart[gi] + gn], ARGMIN_THREADS, 0, st>>>(g);
          CU(cudaGetLastError());
          WalkCountArgs w{};
          w.profiles = ctx->d_profiles;
          w.pairs = ctx->d_tpairs;
          w.order = ctx->d_order + g0;
          w.nitems = (long long)gn;
          w.xnodes = ctx->d_xnodes;
          w.nodes = ctx->d_nodes;
          w.xnode_off = ctx->d_xnode_off;
          w.node_off = ctx->d_node_off;
          w.nsteps = ctx->d_nsteps;
          walk_count_kernel<<<(unsigned)((gn + 63) / 64), 64, 0, st>>>(w);
          CU(cudaGetLastError());
          ctx->launches += 2;
        }
        if ((rc = join_streams(ctx))) return rc;
        // host vectors and the device order/tile/offset buffers are reused by the next round
        CU(cudaStreamSynchronize(ctx->stream));
        c0 = c1;
      }
      return 0;
    };
    ph.lap(1, ctx->stream);
    if ((rc = run_fast())) return rc;
    ph.lap(2, ctx->stream);
    if (!keep)
    { // a path that outgrew its slot was only counted: give every pair its exact size and redo
      unsigned long long over = 0;
      CU(counted_copy(ctx, &over, ctx->d_counters + SLOT_OVERFLOW, sizeof over, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if (over)
      {
        CU(counted_copy(ctx, ctx->t_nsteps.data(), ctx->d_nsteps, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < n; ++i)
        {
          long long const have = slot_off[i + 1] - slot_off[i];
          slot_off[i + 1] = slot_off[i] + (have ? std::max<long long>(have, ctx->t_nsteps[i]) : 0);
        }
        if ((rc = ensure(ctx, ctx->d_lz_ids, ctx->lz_ids_cap, (size_t)slot_off[n]))) return rc;
        if ((rc = ensure(ctx, ctx->d_lz_sz, ctx->lz_sz_cap, (size_t)slot_off[n]))) return rc;
        CU(counted_copy(ctx, ctx->d_lz_off, slot_off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_counters + SLOT_OVERFLOW, 0, sizeof(unsigned long long), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // slot_off is re-read by nothing after this, but keep it simple
        if ((rc = run_fast())) return rc;
      }
    }
  }

  ph.lap(3, ctx->stream);
  // ---- slow route (profiles the register kernels cannot run: K > 2048 or a negative cost) ----
  // Classes by profile size: one warp per pair up to K = 256 (throughput), a CTA of 2/4/8 warps
  // per pair above (the pass would otherwise last as long as its largest profile).
  std::vector<long long> tb[4];
  int tmaxK[4] = {1, 1, 1, 1};
  for (long long i : slow)
  {
    int const K = ctx->h_profiles[(size_t)pairs[i].profile].K;
    int const c = K <= 256 ? 0 : K <= 512 ? 1 : K <= 1024 ? 2 : 3;
    tb[c].push_back(i);
    tmaxK[c] = std::max(tmaxK[c], K);
  }
  std::vector<long long> torder;
  size_t tfirst[5];
  for (int c = 0; c < 4; ++c)
  {
    tfirst[c] = torder.size();
    torder.insert(torder.end(), tb[c].begin(), tb[c].end());
  }
  tfirst[4] = torder.size();
  if (!torder.empty())
  {
    if ((rc = ensure(ctx, ctx->d_order, ctx->order_cap, torder.size()))) return rc;
    CU(counted_copy(ctx, ctx->d_order, torder.data(), torder.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  }

  unsigned tgrid[4] = {0, 0, 0, 0};
  size_t tfloats[4] = {0, 0, 0, 0};
  if (!tb[0].empty() && (rc = generic_plan<true>(ctx, tb[0].size(), tmaxK[0], &tgrid[0], &tfloats[0]))) return rc;
  if (!tb[1].empty() && (rc = trace_cta_plan<2>(ctx, tb[1].size(), tmaxK[1], &tgrid[1], &tfloats[1]))) return rc;
  if (!tb[2].empty() && (rc = trace_cta_plan<4>(ctx, tb[2].size(), tmaxK[2], &tgrid[2], &tfloats[2]))) return rc;
  if (!tb[3].empty() && (rc = trace_cta_plan<8>(ctx, tb[3].size(), tmaxK[3], &tgrid[3], &tfloats[3]))) return rc;
  if ((rc = ensure(ctx, ctx->d_scratch, ctx->scratch_cap, tfloats[0] + tfloats[1] + tfloats[2] + tfloats[3]))) return rc;

  // largest profiles first: they bound the duration of the pass
  size_t sc_off = 0;
  static int const cursor_slot[4] = {SLOT_TRACE_CUR, SLOT_TRACE_CUR + 1, SLOT_TRACE_CUR + 2, SLOT_TRACE_CUR + 3};
  for (int c = 3; c >= 0; --c)
  {
    if (tb[c].empty()) continue;
    GenArgs g{};
    g.s.profiles = ctx->d_profiles;
    g.s.reads = reads_view(ctx);
    g.s.xt = ctx->d_xt[flags & 3u];
    g.s.pairs = ctx->d_tpairs;
    g.s.order = ctx->d_order + tfirst[c];
    g.s.nitems = tb[c].size();
    g.s.counter = ctx->d_counters + cursor_slot[c];
    g.s.out = ctx->d_tout;
    g.s.nhits = ctx->d_counters + SLOT_TRACE_NHITS;
    g.xnodes = ctx->d_xnodes;
    g.nodes = ctx->d_nodes;
    g.xnode_off = ctx->d_xnode_off;
    g.node_off = ctx->d_node_off;
    g.nsteps = ctx->d_nsteps;
    float *scr = ctx->d_scratch + sc_off;
    sc_off += tfloats[c];
    if (c == 0)
    {
      g.scratch = scr;
      g.scratch_stride = (size_t)19 * ((tmaxK[0] + 31) & ~31);
      generic_kernel<true><<<tgrid[0], GEN_THREADS, 0, ctx->stream>>>(g);
      CU(cudaGetLastError());
      ctx->launches += 1;
    }
    else if (c == 1) { if ((rc = launch_trace_cta<2>(ctx, g, tmaxK[1], tgrid[1], scr))) return rc; }
    else if (c == 2) { if ((rc = launch_trace_cta<4>(ctx, g, tmaxK[2], tgrid[2], scr))) return rc; }
    else { if ((rc = launch_trace_cta<8>(ctx, g, tmaxK[3], tgrid[3], scr))) return rc; }
  }

  ph.lap(4, ctx->stream);
  std::vector<float2> h(n);
  CU(counted_copy(ctx, h.data(), ctx->d_tout, n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, ctx->t_nsteps.data(), ctx->d_nsteps, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ph.lap(5, ctx->stream);
  if (ph.on)
    fprintf(stderr, "[trace_pairs %lld pairs] prep+H2D %.4f  dump sizing %.4f  dump+walk rounds %.4f  overflow redo %.4f  slow route %.4f  D2H %.4f s\n",
            (long long)npairs, ph.t[0], ph.t[1], ph.t[2], ph.t[3], ph.t[4], ph.t[5]);
  for (size_t i = 0; i < n; ++i)
  {
    if (alt_cost) alt_cost[i] = h[i].y;
    if (nsteps) nsteps[i] = ctx->t_nsteps[i];
    if (ctx->t_nsteps[i] <= 0) return fail(ctx, DCPGPU_ESTATE, "trace: corrupt trellis (internal error)");
  }
  ctx->traced = true;
  return 0;
}

// Compact device layout of the traced paths: path i at [off[i], off[i] + nsteps[i]) in path order.
static int compact_steps(dcpgpu_ctx *ctx, std::vector<long long> *off_out)
{
  size_t const n = ctx->t_pairs.size();
  std::vector<long long> off(n + 1, 0);
  for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + ctx->t_nsteps[i];
  if (off_out) *off_out = off;
  if (ctx->steps_compact) return 0;
  size_t const total = (size_t)off[n];
  int rc;
  if ((rc = ensure(ctx, ctx->d_step_off, ctx->step_off_cap, n + 1))) return rc;
  if ((rc = ensure(ctx, ctx->d_step_ids, ctx->step_ids_cap, total))) return rc;
  if ((rc = ensure(ctx, ctx->d_step_sz, ctx->step_sz_cap, total))) return rc;
  uint16_t *d_ids = ctx->d_step_ids;
  uint8_t *d_sz = ctx->d_step_sz;
  CU(counted_copy(ctx, ctx->d_step_off, off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream)); // off is a local
  if (ctx->t_node_off[n] > 0)
  { // pairs traced through a trellis: second back-walk, in path order
    WalkArgs w{};
    w.profiles = ctx->d_profiles;
    w.pairs = ctx->d_tpairs;
    w.npairs = (long long)n;
    w.xnodes = ctx->d_xnodes;
    w.nodes = ctx->d_nodes;
    w.xnode_off = ctx->d_xnode_off;
    w.node_off = ctx->d_node_off;
    w.nsteps = ctx->d_nsteps;
    w.step_off = ctx->d_step_off;
    w.ids = d_ids;
    w.sizes = d_sz;
    walk_write_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(w);
    CU(cudaGetLastError());
    ctx->launches += 1;
  }
  { // pairs walked from the dumped values: move each path to its place
    GatherArgs g{};
    g.npairs = (long long)n;
    g.nsteps = ctx->d_nsteps;
    g.src_off = ctx->d_lz_off;
    g.dst_off = ctx->d_step_off;
    g.src_ids = ctx->d_lz_ids;
    g.src_sizes = ctx->d_lz_sz;
    g.ids = d_ids;
    g.sizes = d_sz;
    gather_steps_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(g);
    CU(cudaGetLastError());
    ctx->launches += 1;
  }
  ctx->steps_compact = true;
  return 0;
}

int dcpgpu_trace_fetch(dcpgpu_ctx *ctx, int64_t const *offsets, uint16_t *state_ids, uint8_t *seqsizes)
{
  if (!ctx || !ctx->traced) return fail(ctx, DCPGPU_ESTATE, "trace_fetch before trace_pairs");
  size_t const n = ctx->t_pairs.size();
  if (n == 0) return 0;
  if (!offsets || !state_ids || !seqsizes) return fail(ctx, DCPGPU_EINVAL, "trace_fetch: bad argument");
  CU(cudaSetDevice(ctx->device));
  std::vector<long long> off;
  int rc;
  if ((rc = compact_steps(ctx, &off))) return rc;
  size_t const total = (size_t)off[n];
  uint16_t *d_ids = ctx->d_step_ids;
  uint8_t *d_sz = ctx->d_step_sz;
  bool compact = true; // the usual placement: paths back to back in pair order
  for (size_t i = 0; i < n && compact; ++i) compact = offsets[i] == offsets[0] + off[i];
  if (compact)
  {
    CU(counted_copy(ctx, state_ids + offsets[0], d_ids, total * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(counted_copy(ctx, seqsizes + offsets[0], d_sz, total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  std::vector<uint16_t> h_ids(total);
  std::vector<uint8_t> h_sz(total);
  CU(counted_copy(ctx, h_ids.data(), d_ids, total * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, h_sz.data(), d_sz, total, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; ++i)
  {
    std::memcpy(state_ids + offsets[i], h_ids.data() + off[i], (size_t)ctx->t_nsteps[i] * sizeof(uint16_t));
    std::memcpy(seqsizes + offsets[i], h_sz.data() + off[i], (size_t)ctx->t_nsteps[i]);
  }
  return 0;
}

int dcpgpu_trace_trellis(dcpgpu_ctx *ctx, int64_t i, uint32_t *xnodes, uint16_t *nodes)
{
  if (!ctx || !ctx->traced) return fail(ctx, DCPGPU_ESTATE, "trace_trellis before trace_pairs");
  if (i < 0 || (size_t)i >= ctx->t_pairs.size()) return fail(ctx, DCPGPU_EINVAL, "trace_trellis: bad index");
  CU(cudaSetDevice(ctx->device));
  size_t const nx = (size_t)(ctx->t_xnode_off[(size_t)i + 1] - ctx->t_xnode_off[(size_t)i]);
  if (nx == 0) return fail(ctx, DCPGPU_ESTATE, "trace_trellis: trace_pairs ran without DCPGPU_KEEP_TRELLIS");
  size_t const nn = (size_t)(ctx->t_node_off[(size_t)i + 1] - ctx->t_node_off[(size_t)i]);
  if (xnodes)
    CU(counted_copy(ctx, xnodes, ctx->d_xnodes + ctx->t_xnode_off[(size_t)i], nx * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (nodes)
    CU(counted_copy(ctx, nodes, ctx->d_nodes + ctx->t_node_off[(size_t)i], nn * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dcpgpu_alu_peak(dcpgpu_ctx *ctx, int mode, double *tera_ops_per_s)
{
  if (!ctx || !tera_ops_per_s) return DCPGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  switch (mode)
  {
  case 0: return run_alu_peak<0>(ctx, tera_ops_per_s);
  case 1: return run_alu_peak<1>(ctx, tera_ops_per_s);
  case 2: return run_alu_peak<2>(ctx, tera_ops_per_s);
  case 3: return run_alu_peak<3>(ctx, tera_ops_per_s);
  case 4: return run_alu_peak<4>(ctx, tera_ops_per_s);
  case 5: return run_alu_peak<5>(ctx, tera_ops_per_s);
  case 6: return run_alu_peak<6>(ctx, tera_ops_per_s);
  default: return fail(ctx, DCPGPU_EINVAL, "alu_peak: bad mode");
  }
}

int dcpgpu_profile_set_decoder(dcpgpu_ctx *ctx, int32_t profile, float const *node_dists, float const *null_dist,
                               float const *bg_dist, char const *gencode64)
{
  if (!ctx || profile < 0 || (size_t)profile >= ctx->h_profiles.size() || !node_dists || !null_dist || !bg_dist ||
      !gencode64)
    return fail(ctx, DCPGPU_EINVAL, "profile_set_decoder: bad argument");
  CU(cudaSetDevice(ctx->device));
  int const K = ctx->h_profiles[(size_t)profile].K;
  void *mem = nullptr;
  int rc;
  if ((rc = arena_alloc(ctx, (size_t)(K + 2) * DIST_FLOATS * sizeof(float), &mem))) return rc;
  float *d = static_cast<float *>(mem);
  CU(counted_copy(ctx, d, node_dists, (size_t)K * DIST_FLOATS * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, d + (size_t)K * DIST_FLOATS, null_dist, DIST_FLOATS * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, d + (size_t)(K + 1) * DIST_FLOATS, bg_dist, DIST_FLOATS * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream)); // the caller's buffers may be reused right away
  ctx->h_decoders.resize(ctx->h_profiles.size(), DecoderDesc{nullptr, {0}});
  DecoderDesc &dd = ctx->h_decoders[(size_t)profile];
  dd.dists = d;
  std::memcpy(dd.gencode, gencode64, 64);
  ctx->profiles_dirty = true;
  return 0;
}

int dcpgpu_match_build(dcpgpu_ctx *ctx, float epsilon, int is_rna, int32_t *hit, int32_t *hit_start, int32_t *hit_stop,
                       int64_t *text_off)
{
  if (!ctx || !ctx->traced) return fail(ctx, DCPGPU_ESTATE, "match_build before trace_pairs");
  size_t const n = ctx->t_pairs.size();
  if (text_off) text_off[0] = 0;
  ctx->m_text_off.assign(n + 1, 0);
  ctx->matched = true;
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = sync_profiles(ctx))) return rc;
  std::vector<long long> off;
  if ((rc = compact_steps(ctx, &off))) return rc;
  size_t const total = (size_t)off[n];
  if ((rc = ensure(ctx, ctx->d_m_hit, ctx->m_cap[0], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_start, ctx->m_cap[1], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_stop, ctx->m_cap[2], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_begin, ctx->m_cap[3], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_end, ctx->m_cap[4], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_len, ctx->m_cap[5], n))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_off, ctx->m_cap[6], n + 1))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_bad, ctx->m_bad_cap, 1))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_codon, ctx->m_codon_cap, total))) return rc;
  if ((rc = ensure(ctx, ctx->d_m_amino, ctx->m_amino_cap, total))) return rc;
  CU(cudaMemsetAsync(ctx->d_m_bad, 0, sizeof(int), ctx->stream));
  MatchArgs a{};
  a.pairs = ctx->d_tpairs;
  a.npairs = (long long)n;
  a.nsteps = ctx->d_nsteps;
  a.step_off = ctx->d_step_off;
  a.ids = ctx->d_step_ids;
  a.sizes = ctx->d_step_sz;
  a.profiles = ctx->d_profiles;
  a.decoders = ctx->d_decoders;
  a.reads = reads_view(ctx);
  a.eps = (double)epsilon;
  a.is_rna = is_rna;
  a.extent_only = text_off ? 0 : 1;
  a.hit = ctx->d_m_hit;
  a.hit_start = ctx->d_m_start;
  a.hit_stop = ctx->d_m_stop;
  a.seg_begin = ctx->d_m_begin;
  a.seg_end = ctx->d_m_end;
  a.text_len = ctx->d_m_len;
  a.text_off = ctx->d_m_off;
  a.codon = ctx->d_m_codon;
  a.amino = ctx->d_m_amino;
  a.bad = ctx->d_m_bad;
  unsigned const grid = (unsigned)((n + MATCH_WARPS - 1) / MATCH_WARPS);
  match_kernel<false><<<grid, 32 * MATCH_WARPS, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches += 1;
  std::vector<long long> len(n);
  std::vector<int> h_hit(n), h_start(n), h_stop(n);
  int bad = 0;
  CU(counted_copy(ctx, len.data(), ctx->d_m_len, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, h_hit.data(), ctx->d_m_hit, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, h_start.data(), ctx->d_m_start, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, h_stop.data(), ctx->d_m_stop, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(counted_copy(ctx, &bad, ctx->d_m_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (bad) return fail(ctx, DCPGPU_EDECODE, "match_build: a fragment could not be decoded into a codon (or no decoder tables)");
  for (size_t i = 0; i < n; ++i)
  {
    ctx->m_text_off[i + 1] = ctx->m_text_off[i] + len[i];
    if (hit) hit[i] = h_hit[i];
    if (hit_start) hit_start[i] = h_start[i];
    if (hit_stop) hit_stop[i] = h_stop[i];
    if (text_off) text_off[i + 1] = (int64_t)ctx->m_text_off[i + 1];
  }
  size_t const bytes = (size_t)ctx->m_text_off[n];
  if (a.extent_only) return 0;
  if ((rc = ensure(ctx, ctx->d_m_text, ctx->m_text_cap, bytes + 1))) return rc;
  CU(counted_copy(ctx, ctx->d_m_off, ctx->m_text_off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  a.text = ctx->d_m_text;
  match_kernel<true><<<grid, 32 * MATCH_WARPS, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches += 1;
  return 0;
}

int dcpgpu_match_fetch(dcpgpu_ctx *ctx, char *text)
{
  if (!ctx || !ctx->matched) return fail(ctx, DCPGPU_ESTATE, "match_fetch before match_build");
  size_t const bytes = ctx->m_text_off.empty() ? 0 : (size_t)ctx->m_text_off.back();
  if (bytes == 0) return 0;
  if (!text) return fail(ctx, DCPGPU_EINVAL, "match_fetch: bad argument");
  CU(cudaSetDevice(ctx->device));
  CU(counted_copy(ctx, text, ctx->d_m_text, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dcpgpu_frame_tables(dcpgpu_ctx *ctx, int32_t nstates, float const *nuclt_lprobs, float const *codon_marg_lprobs,
                        float epsilon, float *emission)
{
  if (!ctx || nstates < 0 || (nstates && (!nuclt_lprobs || !codon_marg_lprobs || !emission)) || !(epsilon >= 0.f) ||
      !(epsilon < 1.f))
    return fail(ctx, DCPGPU_EINVAL, "frame_tables: bad argument");
  if (nstates == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  size_t const n = (size_t)nstates;
  size_t const in_floats = n * (4 + 125), out_floats = n * NCODES;
  int rc;
  if ((rc = ensure(ctx, ctx->d_scratch, ctx->scratch_cap, in_floats + out_floats))) return rc;
  float *d_nuclt = ctx->d_scratch, *d_marg = d_nuclt + n * 4, *d_out = d_marg + n * 125;
  CU(counted_copy(ctx, d_nuclt, nuclt_lprobs, n * 4 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(counted_copy(ctx, d_marg, codon_marg_lprobs, n * 125 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  FrameArgs a{d_nuclt, d_marg, d_out, epsilon, nstates};
  frame_table_kernel<<<(unsigned)nstates, 256, 0, ctx->stream>>>(a);
  CU(cudaGetLastError());
  ctx->launches += 1;
  CU(counted_copy(ctx, emission, d_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dcpgpu_xtrans(int window_len, uint32_t flags, float out[13])
{
  if (window_len < 1 || !out) return DCPGPU_EINVAL;
  float tmp[X_STRIDE];
  host_xtrans(window_len, flags & DCPGPU_MULTI_HITS, flags & DCPGPU_HMMER3_COMPAT, tmp);
  std::memcpy(out, tmp, 13 * sizeof(float));
  return 0;
}

} // extern "C"
