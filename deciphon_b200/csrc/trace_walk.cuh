// Trace pass, fast route, paths only: back-walk with argmins decided on demand.
//
// trellis_unzip (c-core/trellis.c:147-167) visits one trellis word per path step -- about L of
// the (L+1)*K words viterbi_path (viterbi.c:726-732) fills in.  An argmin depends only on the
// VALUES of its candidates, and score_reg_kernel<Q,W,DUMP=true> leaves exactly those values
// (bit-identical to the reference's) in global memory.  So instead of materialising the whole
// bit matrix (trace_argmin.cuh: 12 bytes read + 2 written per DP cell), one warp per pair walks
// T@L -> S@0 and decides each visited word on the spot: lane j evaluates the j-th candidate in
// the reference's order with the reference's arithmetic ((state + transition) + emission,
// viterbi.c:201-212, 485-586), one REDUX.MIN over the bit patterns (all costs are >= +0 on this
// route) finds the minimum and the lowest lane holding it is the reference's first-wins,
// strict-less argmin.  The decision is the trellis field the reference would have stored
// (trellis.h:42-56), so the step decoding below is trellis_step (generic_kernel.cuh) verbatim.
//
// The steps are written in walk order (T first) into a per-pair slot sized by a generous bound
// and reversed into path order by gather_steps_kernel; a path that outgrows its slot is only
// counted and the host reruns with exact sizes.  The kernel that keeps the full trellis
// (DCPGPU_KEEP_TRELLIS) stays available for inspection and is tested equal to this one.
#pragma once
#include "generic_kernel.cuh"
#include "score_kernel.cuh"

namespace dcp {

constexpr int LAZY_WARPS = 4; // pairs per CTA

struct LazyWalkArgs
{
  ProfileDesc const *profiles;
  ReadsView reads;
  float const *xt;
  Pair const *pairs;
  long long const *order;    // [nitems] pair indices handled by this launch
  long long nitems;
  float const *dump;
  long long const *dump_off; // per item of this launch
  int *nsteps;                  // [pair] steps of the path, -1 if no finite path exists
  long long const *slot_off;    // [pair+1] slots of the pairs in ids/sizes (empty: not on this route)
  unsigned long long *overflow; // paths that did not fit their slot
  uint16_t *ids;                // steps in walk order (T first, S last)
  uint8_t *sizes;
};

struct LazyCtx
{
  DumpView dv;
  ProfileDesc pd;
  float const *xt;
  uint32_t const *words;
  int start, L, lane;
};

// lowest lane whose candidate equals the minimum; false if no candidate is finite
__device__ __forceinline__ bool lazy_pick(float cand, bool has, int tag, int &wtag)
{
  unsigned const bits = has ? __float_as_uint(cand) : 0xFFFFFFFFu;
  unsigned const m = __reduce_min_sync(FULL_MASK, bits);
  if (m >= 0x7F800000u) return false;
  int const win = __ffs(__ballot_sync(FULL_MASK, bits == m)) - 1;
  wtag = __shfl_sync(FULL_MASK, tag, win);
  return true;
}

// One step back from (state, stage); warp-uniform arguments and result.
__device__ __forceinline__ bool lazy_step(LazyCtx const &c, int &state, int &stage, int &size)
{
  float const INF = CUDART_INF_F;
  int const lane = c.lane, l = stage, Kpad = c.pd.Kpad;
  int const msb = state & (3 << 14);
  if (state == ST_B && l == 0)
  { // stage 0 holds all-zero fields (before(), viterbi.c:602-629): B <- S
    state = ST_S;
    size = 0;
    return true;
  }
  if (l < 1) return false;

  // codes of the emission lengths ending at row l
  unsigned hist = 0;
  for (int i = max(0, l - 5); i < l; ++i)
  {
    int const g = c.start + i;
    hist = (hist << 2) | ((c.words[g >> 4] >> (2 * (g & 15))) & 3u);
  }
  auto code = [&](int t) {
    return t == 1 ? (int)(hist & 3) : t == 2 ? 4 + (int)(hist & 15) : t == 3 ? 20 + (int)(hist & 63)
           : t == 4 ? 84 + (int)(hist & 255) : 340 + (int)(hist & 1023);
  };
  auto X = [&](int lz, int j) { return lz >= 1 ? c.dv.xs[(size_t)(lz - 1) * 8 + j] : INF; }; // N,B,J,E,C
  float const *xt = c.xt;

  float cand = INF;
  bool has = false;
  int tag = 0, wtag = 0, prev;

  if (msb == ST_X)
  {
    if (state == ST_E)
    { // first-wins over M_0, D_0, M_1, D_1, ... = smallest (value, 2k+isD)  (viterbi.c:540-558)
      float ev = INF;
      int ei = 0;
      size_t const row = (size_t)(l - 1) * Kpad;
      for (int k = lane; k < c.pd.K; k += 32)
      {
        int const pos = layout_pos(k, c.pd.Q, c.pd.VL);
        DCP_UPD(ev, c.dv.M[row + pos], ei, 2 * k + 0);
        DCP_UPD(ev, c.dv.D[row + pos], ei, 2 * k + 1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        float const ov = __shfl_xor_sync(FULL_MASK, ev, o);
        int const oi = __shfl_xor_sync(FULL_MASK, ei, o);
        if (ov < ev || (ov == ev && oi < ei))
        {
          ev = ov;
          ei = oi;
        }
      }
      if (!(ev < INF)) return false;
      size = 0;
      prev = ((ei & 1) ? ST_D : ST_M) | (ei / 2 + 1);
    }
    else if (state == ST_T)
    {
      if (lane < 2)
      {
        cand = lane == 0 ? X(l, 3) + xt[X_ET] : X(l, 4) + xt[X_CT];
        tag = lane;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = 0;
      prev = wtag ? ST_C : ST_E;
    }
    else if (state == ST_B)
    { // S+SB is +INF for l >= 1 (viterbi.c:495)
      if (lane < 3)
      {
        cand = lane == 0 ? X(l, 0) + xt[X_NB] : lane == 1 ? X(l, 3) + xt[X_EB] : X(l, 2) + xt[X_JB];
        tag = lane + 1;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = 0;
      prev = wtag == 1 ? ST_N : wtag == 2 ? ST_E : ST_J;
    }
    else if (state == ST_C || state == ST_J || state == ST_N)
    { // emission length 5..1; the entering transition before the self loop
      if (lane < 10)
      {
        int const t = 5 - (lane >> 1), w = lane & 1, lz = l - t;
        float const nil = __ldg(&c.pd.nulbg[code(t)]).x;
        float sv, tr;
        if (state == ST_C)
        {
          sv = w ? X(lz, 4) : X(lz, 3);
          tr = w ? xt[X_CC] : xt[X_EC];
        }
        else if (state == ST_J)
        {
          sv = w ? X(lz, 2) : X(lz, 3);
          tr = w ? xt[X_JJ] : xt[X_EJ];
        }
        else
        {
          sv = w ? X(lz, 0) : (lz == 0 ? 0.0f : INF);
          tr = w ? xt[X_NN] : xt[X_SN];
        }
        cand = sv + tr + nil;
        tag = w * 5 + t - 1;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = wtag % 5 + 1;
      int const self = state, enter = state == ST_N ? ST_S : ST_E;
      prev = wtag / 5 ? self : enter;
    }
    else
      return false;
  }
  else
  {
    int const k = (state & 0x3fff) - 1; // state_core_idx, state.c:25
    if (k < 0 || k >= c.pd.K) return false;
    int const pk = layout_pos(k, c.pd.Q, c.pd.VL);
    int const pk1 = k > 0 ? layout_pos(k - 1, c.pd.Q, c.pd.VL) : 0; // node k-1 in a dumped row
    float const *core = c.pd.core;
    if (msb == ST_M)
    { // emission length 5..1; BM, MM, IM, DM  (viterbi.c:485-530)
      if (lane < 20)
      {
        int const t = 5 - (lane >> 2), src = lane & 3, lz = l - t;
        float sv;
        if (src == 0)
          sv = lz < 0 ? INF : lz == 0 ? xt[X_SB] : X(lz, 1);
        else if (k > 0 && lz >= 1)
        {
          size_t const at = (size_t)(lz - 1) * Kpad + pk1;
          sv = src == 1 ? c.dv.M[at] : src == 2 ? c.dv.I[at] : c.dv.D[at];
        }
        else
          sv = INF;
        int const ct = src == 0 ? C_BM : src == 1 ? C_MM : src == 2 ? C_IM : C_DM;
        float const tr = __ldg(core + (size_t)ct * Kpad + pk);
        float const e = __ldg(c.pd.em + (size_t)code(t) * Kpad + pk);
        cand = (sv + tr) + e;
        tag = src * 5 + t - 1;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = wtag % 5 + 1;
      int const src = wtag / 5;
      if (src && k <= 0) return false;
      prev = src == 0 ? ST_B : ((src == 1 ? ST_M : src == 2 ? ST_I : ST_D) | k);
    }
    else if (msb == ST_I)
    { // II before MI (viterbi.c:535-536)
      if (lane < 10)
      {
        int const t = 5 - (lane >> 1), w = lane & 1, lz = l - t;
        float sv = INF;
        if (lz >= 1)
        {
          size_t const at = (size_t)(lz - 1) * Kpad + pk;
          sv = w ? c.dv.M[at] : c.dv.I[at];
        }
        float const tr = __ldg(core + (size_t)(w ? C_MI : C_II) * Kpad + pk);
        float const b = __ldg(&c.pd.nulbg[code(t)]).y;
        cand = (sv + tr) + b;
        tag = (w ? 0 : 5) + t - 1;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = wtag % 5 + 1;
      prev = (wtag / 5 ? ST_I : ST_M) | (k + 1);
    }
    else
    { // D_k(l) <- M_{k-1}(l), D_{k-1}(l)  (viterbi.c:538, 552-580)
      if (k <= 0) return false;
      if (lane < 2)
      {
        size_t const at = (size_t)(l - 1) * Kpad + pk1;
        cand = lane == 0 ? c.dv.M[at] + __ldg(core + (size_t)C_MD * Kpad + pk)
                         : c.dv.D[at] + __ldg(core + (size_t)C_DD * Kpad + pk);
        tag = lane;
        has = true;
      }
      if (!lazy_pick(cand, has, tag, wtag)) return false;
      size = 0;
      prev = (wtag ? ST_D : ST_M) | k;
    }
  }
  state = prev;
  stage -= size;
  return stage >= 0;
}

__global__ void __launch_bounds__(32 * LAZY_WARPS) lazy_walk_kernel(LazyWalkArgs a)
{
  int const lane = threadIdx.x & 31;
  long long const item = (long long)blockIdx.x * LAZY_WARPS + (threadIdx.x >> 5);
  if (item >= a.nitems) return;
  long long const oidx = a.order[item];
  Pair const pr = a.pairs[oidx];
  ProfileDesc const pd = a.profiles[pr.profile];
  int const L = pr.len;
  LazyCtx const c{DumpView(const_cast<float *>(a.dump) + a.dump_off[item], L, pd.Kpad),
                  pd,
                  a.xt + (size_t)L * X_STRIDE,
                  a.reads.words + a.reads.seq_word[pr.seq],
                  pr.start,
                  L,
                  lane};
  long long const guard = (long long)(L + 2) * (pd.K + 4) + 8;
  long long const base = a.slot_off[oidx], cap = a.slot_off[oidx + 1] - base;
  int state = ST_T, stage = L;
  long long count = 0;
  while (state != ST_S || stage)
  {
    int const cur = state;
    int size = 0;
    if (!lazy_step(c, state, stage, size) || count >= guard)
    {
      if (lane == 0) a.nsteps[oidx] = -1;
      return;
    }
    if (lane == 0 && count < cap)
    {
      a.ids[base + count] = (uint16_t)cur;
      a.sizes[base + count] = (uint8_t)size;
    }
    ++count;
  }
  if (lane == 0)
  {
    if (count < cap)
    {
      a.ids[base + count] = (uint16_t)ST_S;
      a.sizes[base + count] = 0;
    }
    else
      atomicAdd(a.overflow, 1ULL);
    a.nsteps[oidx] = (int)count + 1;
  }
}

// path i of the lazy walk, reversed -> its place in the compact, pair-ordered layout (one warp
// per pair)
struct GatherArgs
{
  long long npairs;
  int const *nsteps;
  long long const *src_off; // [pair+1] slots; empty: the pair was traced through a kept trellis
  long long const *dst_off;
  uint16_t const *src_ids;
  uint8_t const *src_sizes;
  uint16_t *ids;
  uint8_t *sizes;
};

__global__ void gather_steps_kernel(GatherArgs a)
{
  long long const i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.npairs) return;
  long long const s = a.src_off[i];
  if (a.src_off[i + 1] == s) return;
  long long const d = a.dst_off[i];
  int const n = a.nsteps[i];
  for (int j = threadIdx.x & 31; j < n; j += 32)
  {
    a.ids[d + j] = a.src_ids[s + (n - 1 - j)];
    a.sizes[d + j] = a.src_sizes[s + (n - 1 - j)];
  }
}

} // namespace dcp
