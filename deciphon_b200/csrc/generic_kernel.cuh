// Generic pass for any core size K (1..16384) and the TRACE pass for hits.
//
// One warp per pair, nodes interleaved across lanes (k = chunk*32 + lane), the 6-row ring of
// M/I/D (the reference's time frame, c-core/viterbi.c:12,160-161) kept in a per-warp
// scratch area in global memory (L1/L2 resident).  Candidates are evaluated with the
// reference's own arithmetic and order, (state + transition) + emission, emission length
// t = 5..1 and BM,MM,IM,DM / II,MI within it (viterbi.c:485-536), with strict-less updates,
// i.e. first candidate wins (viterbi.c:201-212, intrinsics.h:144-149).
//
// With TRACE the kernel also writes the reference's bit-packed trellis (trellis.h:12-56,
// packing of viterbi.c:631-692) and walks it back T@L -> S@0 (trellis.c:147-167) to count
// the path's steps; walk_write_kernel then emits the steps.
#pragma once
#include "score_kernel.cuh"

namespace dcp {

constexpr int GEN_THREADS = 128;
constexpr int GEN_WARPS = GEN_THREADS / 32;

// state ids, c-core/state.h:7-25
enum : int
{
  ST_M = 0 << 14,
  ST_I = 1 << 14,
  ST_D = 2 << 14,
  ST_X = 3 << 14,
  ST_S = ST_X | 3,
  ST_N = ST_X | 4,
  ST_B = ST_X | 5,
  ST_E = ST_X | 6,
  ST_J = ST_X | 7,
  ST_C = ST_X | 8,
  ST_T = ST_X | 9,
};

struct GenArgs
{
  ScoreArgs s;
  float *scratch;         // per-warp scratch, scratch_stride floats each
  size_t scratch_stride;  // >= 19 * KG(max)
  // trace outputs (TRACE only), per pair
  uint32_t *xnodes;       // pool
  uint16_t *nodes;        // pool
  long long const *xnode_off;
  long long const *node_off;
  int *nsteps;
};

#define DCP_UPD(val, cand, ptr, tag)                                                             \
  do                                                                                             \
  {                                                                                              \
    float c_ = (cand);                                                                           \
    if (c_ < (val))                                                                              \
    {                                                                                            \
      (val) = c_;                                                                                \
      (ptr) = (tag);                                                                             \
    }                                                                                            \
  } while (0)

// previous_state + emission_size of c-core/trellis.c:51-113 for one step of the back-walk.
__device__ __forceinline__ bool trellis_step(int K, uint32_t const *xnodes, uint16_t const *nodes,
                                             int &state, int &stage, int &size)
{
  int const msb = state & (3 << 14);
  int prev;
  if (msb == ST_X)
  {
    uint32_t const xn = xnodes[stage];
    unsigned const vN = xn & 15, vB = (xn >> 4) & 3, vE = (xn >> 6) & 0x7fff, vC = (xn >> 21) & 15,
                   vT = (xn >> 25) & 1, vJ = (xn >> 26) & 15;
    if (state == ST_T) { size = 0; prev = vT ? ST_C : ST_E; }
    else if (state == ST_E) { size = 0; prev = ((vE & 1) ? ST_D : ST_M) | (int)(vE / 2 + 1); }
    else if (state == ST_C) { size = vC % 5 + 1; prev = vC / 5 ? ST_C : ST_E; }
    else if (state == ST_J) { size = vJ % 5 + 1; prev = vJ / 5 ? ST_J : ST_E; }
    else if (state == ST_N) { size = vN % 5 + 1; prev = vN / 5 ? ST_N : ST_S; }
    else if (state == ST_B) { size = 0; prev = vB == 0 ? ST_S : vB == 1 ? ST_N : vB == 2 ? ST_E : ST_J; }
    else return false;
  }
  else
  {
    int const k = (state & 0x3fff) - 1; // state_core_idx, state.c:25
    uint16_t const nd = nodes[(size_t)stage * K + k];
    unsigned const vM = nd & 31, vD = (nd >> 5) & 1, vI = (nd >> 6) & 15;
    if (msb == ST_M)
    {
      size = vM % 5 + 1;
      unsigned const src = vM / 5;
      if (src && k <= 0) return false;
      prev = src == 0 ? ST_B : ((src == 1 ? ST_M : src == 2 ? ST_I : ST_D) | k);
    }
    else if (msb == ST_I) { size = vI % 5 + 1; prev = (vI / 5 ? ST_I : ST_M) | (k + 1); }
    else { size = 0; if (k <= 0) return false; prev = (vD ? ST_D : ST_M) | k; }
  }
  state = prev;
  stage -= size;
  return stage >= 0;
}

// Walk T@L -> S@0.  Returns the number of steps (including S), or -1 on a corrupt trellis.
// If ids != nullptr the steps are written in path order into ids/sz[0..n).
__device__ inline int trellis_walk(int K, int L, uint32_t const *xnodes, uint16_t const *nodes,
                                   int n, uint16_t *ids, uint8_t *sz)
{
  int state = ST_T, stage = L, count = 0;
  long long const guard = (long long)(L + 2) * (K + 4) + 8;
  while (state != ST_S || stage)
  {
    int const cur = state;
    int size = 0;
    if (!trellis_step(K, xnodes, nodes, state, stage, size)) return -1;
    if (ids)
    {
      ids[n - 1 - count] = (uint16_t)cur;
      sz[n - 1 - count] = (uint8_t)size;
    }
    if (++count > guard) return -1;
  }
  if (ids)
  {
    ids[n - 1 - count] = (uint16_t)ST_S;
    sz[n - 1 - count] = 0;
  }
  return count + 1;
}

template <bool TRACE>
__global__ void __launch_bounds__(GEN_THREADS, 5) generic_kernel(GenArgs a)
{
  // special-state ring: xs[slot][0..6] = S, N, B, J, E, C, R
  __shared__ float xs_all[GEN_WARPS][6][8];
  int const lane = threadIdx.x & 31;
  int const warp = threadIdx.x >> 5;
  float(*xs)[8] = xs_all[warp];
  float *scratch = a.scratch + (size_t)(blockIdx.x * GEN_WARPS + warp) * a.scratch_stride;
  float const INF = CUDART_INF_F;

  for (;;)
  {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.s.counter, 1ULL);
    item = __shfl_sync(FULL_MASK, item, 0);
    if (item >= a.s.nitems) break;

    int p, sq, start, L;
    long long oidx;
    if (a.s.pairs)
    {
      oidx = a.s.order ? a.s.order[item] : (long long)item;
      Pair const pr = a.s.pairs[oidx];
      p = pr.profile; sq = pr.seq; start = pr.start; L = pr.len;
    }
    else
    {
      int const pi = (int)(item / (unsigned)a.s.nseq);
      int const si = (int)(item - (unsigned long long)pi * (unsigned)a.s.nseq);
      p = a.s.class_profiles[pi];
      sq = a.s.seq0 + si;
      start = 0;
      oidx = (long long)(p - a.s.prof0) * a.s.nseq + si;
      L = -1;
    }
    ProfileDesc const pd = a.s.profiles[p];
    int const K = pd.K, Kpad = pd.Kpad;
    if (L < 0) L = min(min(K * 50, 100000), a.s.reads.seq_len[sq]);
    int const KG = (K + 31) & ~31;
    float *rM = scratch, *rI = scratch + 6 * (size_t)KG, *rD = scratch + 12 * (size_t)KG;
    int *posk = reinterpret_cast<int *>(scratch + 18 * (size_t)KG);

    float const *xt = a.s.xt + (size_t)L * X_STRIDE;
    float const RR = xt[X_RR], SN = xt[X_SN], NN = xt[X_NN], SB = xt[X_SB], NB = xt[X_NB],
                EB = xt[X_EB], JB = xt[X_JB], EJ = xt[X_EJ], JJ = xt[X_JJ], EC = xt[X_EC],
                CC = xt[X_CC], ET = xt[X_ET], CT = xt[X_CT];

    uint32_t *xnodes = nullptr;
    uint16_t *nodes = nullptr;
    if (TRACE)
    {
      xnodes = a.xnodes + a.xnode_off[oidx];
      nodes = a.nodes + a.node_off[oidx];
    }

    // row 0 (viterbi.c:472-474; before() writes all-zero trellis fields, :602-629)
    for (int k = lane; k < KG; k += 32)
    {
      posk[k] = k < K ? layout_pos(k, pd.Q, pd.VL) : 0;
      for (int sl = 0; sl < 6; ++sl)
      {
        rM[sl * (size_t)KG + k] = INF;
        rI[sl * (size_t)KG + k] = INF;
        rD[sl * (size_t)KG + k] = INF;
      }
      if (TRACE && k < K) nodes[k] = 0;
    }
    if (lane < 8)
      for (int sl = 0; sl < 6; ++sl)
        xs[sl][lane] = INF;
    __syncwarp();
    if (lane == 0)
    {
      xs[0][0] = 0.0f; // S
      xs[0][2] = SB;   // B
      xs[0][6] = -RR;  // R(0), viterbi.c:703
      if (TRACE) xnodes[0] = 0;
    }
    __syncwarp();

    uint32_t const *wp = a.s.reads.words + a.s.reads.seq_word[sq] + (start >> 4);
    uint32_t word = __ldg(wp) >> (2 * (start & 15));
    int left = 16 - (start & 15);
    unsigned hist = 0;
    float Tv = INF, Rv = INF;

    for (int l = 1; l <= L; ++l)
    {
      hist = ((hist << 2) | (word & 3u)) & 1023u;
      word >>= 2;
      if (--left == 0) { word = __ldg(++wp); left = 16; }
      int const sl = l % 6;
      int const T = l < 5 ? l : 5;
      int code[6];
      code[1] = hist & 3;
      code[2] = 4 + (hist & 15);
      code[3] = 20 + (hist & 63);
      code[4] = 84 + (hist & 255);
      code[5] = 340 + (hist & 1023);

      // special states (viterbi.c:492-502) and the null model (viterbi.c:704-717)
      float N = INF, Jv = INF, C = INF, R = INF;
      int pN = 0, pJ = 0, pC = 0;
      for (int t = T; t >= 1; --t)
      {
        int const z = (l - t) % 6;
        float const nil = __ldg(&pd.nulbg[code[t]]).x;
        DCP_UPD(N, xs[z][0] + SN + nil, pN, 0 + t - 1);
        DCP_UPD(N, xs[z][1] + NN + nil, pN, 5 + t - 1);
        DCP_UPD(Jv, xs[z][4] + EJ + nil, pJ, 0 + t - 1);
        DCP_UPD(Jv, xs[z][3] + JJ + nil, pJ, 5 + t - 1);
        DCP_UPD(C, xs[z][4] + EC + nil, pC, 0 + t - 1);
        DCP_UPD(C, xs[z][5] + CC + nil, pC, 5 + t - 1);
        R = fminf(R, xs[z][6] + RR + nil);
      }

      float carryM = INF, carryD = INF;
      float ev = INF;
      int ei = 0;
      for (int c0 = 0; c0 < KG; c0 += 32)
      {
        int const k = c0 + lane;
        bool const valid = k < K;
        int const pk = posk[k];
        float bm = INF, mm = INF, mi = INF, md = INF, im = INF, ii = INF, dm = INF, dd = INF;
        if (valid)
        {
          bm = __ldg(pd.core + C_BM * Kpad + pk);
          mm = __ldg(pd.core + C_MM * Kpad + pk);
          mi = __ldg(pd.core + C_MI * Kpad + pk);
          md = __ldg(pd.core + C_MD * Kpad + pk);
          im = __ldg(pd.core + C_IM * Kpad + pk);
          ii = __ldg(pd.core + C_II * Kpad + pk);
          dm = __ldg(pd.core + C_DM * Kpad + pk);
          dd = __ldg(pd.core + C_DD * Kpad + pk);
        }
        float M = INF, I = INF;
        int mp = 0, ip = 0;
        auto candidates = [&](int t) {
          size_t const z = (size_t)((l - t) % 6) * KG;
          float const e = valid ? __ldg(pd.em + (size_t)code[t] * Kpad + pk) : INF;
          float const b = __ldg(&pd.nulbg[code[t]]).y;
          float const Bz = xs[(l - t) % 6][2];
          float const pm = k > 0 ? rM[z + k - 1] : INF;
          float const pi = k > 0 ? rI[z + k - 1] : INF;
          float const pdv = k > 0 ? rD[z + k - 1] : INF;
          DCP_UPD(M, (Bz + bm) + e, mp, 0 + t - 1);
          DCP_UPD(M, (pm + mm) + e, mp, 5 + t - 1);
          DCP_UPD(M, (pi + im) + e, mp, 10 + t - 1);
          DCP_UPD(M, (pdv + dm) + e, mp, 15 + t - 1);
          DCP_UPD(I, (rI[z + k] + ii) + b, ip, 5 + t - 1); // II before MI, viterbi.c:535-536
          DCP_UPD(I, (rM[z + k] + mi) + b, ip, 0 + t - 1);
        };
        if (l >= 5)
        { // steady state: all five emission lengths, unrolled so that the loads overlap
#pragma unroll
          for (int t = 5; t >= 1; --t)
            candidates(t);
        }
        else
          for (int t = T; t >= 1; --t)
            candidates(t);
        // delete chain inside the chunk, carried across chunks (viterbi.c:538,561-580)
        float mprev = __shfl_up_sync(FULL_MASK, M, 1);
        if (lane == 0) mprev = carryM;
        float D = mprev + md;
        int dbit = 0;
        for (;;)
        {
          float dp = __shfl_up_sync(FULL_MASK, D, 1);
          if (lane == 0) dp = carryD;
          float const c = dp + dd;
          bool const imp = c < D;
          if (!__any_sync(FULL_MASK, imp)) break;
          if (imp) { D = c; dbit = 1; }
        }
        carryM = __shfl_sync(FULL_MASK, M, 31);
        carryD = __shfl_sync(FULL_MASK, D, 31);
        DCP_UPD(ev, M, ei, 2 * k + 0); // E candidates in node order (viterbi.c:540-541)
        DCP_UPD(ev, D, ei, 2 * k + 1);
        rM[(size_t)sl * KG + k] = M;
        rI[(size_t)sl * KG + k] = I;
        rD[(size_t)sl * KG + k] = D;
        if (TRACE && valid)
          nodes[(size_t)l * K + k] = (uint16_t)((unsigned)mp | ((unsigned)dbit << 5) | ((unsigned)ip << 6));
      }
      // E: smallest (value, 2k+isD) over all lanes
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        float const ov = __shfl_xor_sync(FULL_MASK, ev, o);
        int const oi = __shfl_xor_sync(FULL_MASK, ei, o);
        if (ov < ev || (ov == ev && oi < ei)) { ev = ov; ei = oi; }
      }
      float B = INF;
      int pB = 0, pT = 0;
      DCP_UPD(B, N + NB, pB, 1); // S+SB is +INF for l >= 1 (viterbi.c:495)
      DCP_UPD(B, ev + EB, pB, 2);
      DCP_UPD(B, Jv + JB, pB, 3);
      Tv = INF;
      DCP_UPD(Tv, ev + ET, pT, 0);
      DCP_UPD(Tv, C + CT, pT, 1);
      Rv = R;
      __syncwarp();
      if (lane == 0)
      {
        xs[sl][0] = INF;
        xs[sl][1] = N;
        xs[sl][2] = B;
        xs[sl][3] = Jv;
        xs[sl][4] = ev;
        xs[sl][5] = C;
        xs[sl][6] = R;
        if (TRACE)
          xnodes[l] = (uint32_t)pN | ((uint32_t)pB << 4) | ((uint32_t)ei << 6) | ((uint32_t)pC << 21) |
                      ((uint32_t)pT << 25) | ((uint32_t)pJ << 26);
      }
      __syncwarp();
    }

    if (lane == 0)
    {
      a.s.out[oidx] = make_float2(Rv, Tv);
      float const d = Tv - Rv;
      if (d <= 0.0f && d > -INF) atomicAdd(a.s.nhits, 1ULL);
    }
    if (TRACE)
    {
      __threadfence();
      __syncwarp();
      if (lane == 0) a.nsteps[oidx] = trellis_walk(K, L, xnodes, nodes, 0, nullptr, nullptr);
    }
    __syncwarp();
  }
}

// ---- trace pass for large profiles: one CTA of NW warps per pair -----------------------------
// generic_kernel<true> walks the K/32 node chunks of a row serially in one warp, so the pass
// lasts as long as its largest profile (measured: the trace launch equals the time of the one
// K ~ 1600 hit in it).  Here the chunks of a row are spread over the NW warps of a CTA for the
// M/I candidates (phase 1, no dependency between chunks), then warp 0 alone resolves the delete
// chain, E, the special states and the trellis xnode of the row (phase 2, serial in k).  Same
// arithmetic, candidate order and strict-less updates as generic_kernel.
template <int NW>
__global__ void __launch_bounds__(32 * NW) trace_cta_kernel(GenArgs a)
{
  __shared__ float xs[6][8]; // special-state ring: S, N, B, J, E, C, R
  __shared__ unsigned long long next_item;
  int const lane = threadIdx.x & 31;
  int const warp = threadIdx.x >> 5;
  float *scratch = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
  float const INF = CUDART_INF_F;

  for (;;)
  {
    if (threadIdx.x == 0) next_item = atomicAdd(a.s.counter, 1ULL);
    __syncthreads();
    unsigned long long const item = next_item;
    __syncthreads();
    if (item >= a.s.nitems) break;
    long long const oidx = a.s.order ? a.s.order[item] : (long long)item;
    Pair const pr = a.s.pairs[oidx];
    int const L = pr.len;
    ProfileDesc const pd = a.s.profiles[pr.profile];
    int const K = pd.K, Kpad = pd.Kpad;
    int const KG = (K + 31) & ~31;
    float *rM = scratch, *rI = scratch + 6 * (size_t)KG, *rD = scratch + 12 * (size_t)KG;
    int *posk = reinterpret_cast<int *>(scratch + 18 * (size_t)KG);
    float const *xt = a.s.xt + (size_t)L * X_STRIDE;
    float const RR = xt[X_RR], SN = xt[X_SN], NN = xt[X_NN], SB = xt[X_SB], NB = xt[X_NB],
                EB = xt[X_EB], JB = xt[X_JB], EJ = xt[X_EJ], JJ = xt[X_JJ], EC = xt[X_EC],
                CC = xt[X_CC], ET = xt[X_ET], CT = xt[X_CT];
    uint32_t *xnodes = a.xnodes + a.xnode_off[oidx];
    uint16_t *nodes = a.nodes + a.node_off[oidx];

    // row 0 (viterbi.c:472-474; before() writes all-zero trellis fields, :602-629)
    for (int k = threadIdx.x; k < KG; k += 32 * NW)
    {
      posk[k] = k < K ? layout_pos(k, pd.Q, pd.VL) : 0;
      for (int sl = 0; sl < 6; ++sl)
      {
        rM[sl * (size_t)KG + k] = INF;
        rI[sl * (size_t)KG + k] = INF;
        rD[sl * (size_t)KG + k] = INF;
      }
      if (k < K) nodes[k] = 0;
    }
    if (threadIdx.x < 48) xs[threadIdx.x >> 3][threadIdx.x & 7] = INF;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      xs[0][0] = 0.0f; // S
      xs[0][2] = SB;   // B
      xs[0][6] = -RR;  // R(0), viterbi.c:703
      xnodes[0] = 0;
    }
    __syncthreads();

    uint32_t const *wp = a.s.reads.words + a.s.reads.seq_word[pr.seq] + (pr.start >> 4);
    uint32_t word = __ldg(wp) >> (2 * (pr.start & 15));
    int left = 16 - (pr.start & 15);
    unsigned hist = 0;
    float Tv = INF, Rv = INF;

    for (int l = 1; l <= L; ++l)
    {
      hist = ((hist << 2) | (word & 3u)) & 1023u;
      word >>= 2;
      if (--left == 0) { word = __ldg(++wp); left = 16; }
      int const sl = l % 6;
      int const T = l < 5 ? l : 5;
      int code[6];
      code[1] = hist & 3;
      code[2] = 4 + (hist & 15);
      code[3] = 20 + (hist & 63);
      code[4] = 84 + (hist & 255);
      code[5] = 340 + (hist & 1023);

      // ---- phase 1: M and I of every chunk, chunks spread over the warps ----
      for (int c0 = 32 * warp; c0 < KG; c0 += 32 * NW)
      {
        int const k = c0 + lane;
        bool const valid = k < K;
        int const pk = posk[k];
        float bm = INF, mm = INF, mi = INF, im = INF, ii = INF, dm = INF;
        if (valid)
        {
          bm = __ldg(pd.core + C_BM * Kpad + pk);
          mm = __ldg(pd.core + C_MM * Kpad + pk);
          mi = __ldg(pd.core + C_MI * Kpad + pk);
          im = __ldg(pd.core + C_IM * Kpad + pk);
          ii = __ldg(pd.core + C_II * Kpad + pk);
          dm = __ldg(pd.core + C_DM * Kpad + pk);
        }
        float M = INF, I = INF;
        int mp = 0, ip = 0;
        auto candidates = [&](int t) {
          size_t const z = (size_t)((l - t) % 6) * KG;
          float const e = valid ? __ldg(pd.em + (size_t)code[t] * Kpad + pk) : INF;
          float const b = __ldg(&pd.nulbg[code[t]]).y;
          float const Bz = xs[(l - t) % 6][2];
          float const pm = k > 0 ? rM[z + k - 1] : INF;
          float const pi = k > 0 ? rI[z + k - 1] : INF;
          float const pdv = k > 0 ? rD[z + k - 1] : INF;
          DCP_UPD(M, (Bz + bm) + e, mp, 0 + t - 1);
          DCP_UPD(M, (pm + mm) + e, mp, 5 + t - 1);
          DCP_UPD(M, (pi + im) + e, mp, 10 + t - 1);
          DCP_UPD(M, (pdv + dm) + e, mp, 15 + t - 1);
          DCP_UPD(I, (rI[z + k] + ii) + b, ip, 5 + t - 1); // II before MI, viterbi.c:535-536
          DCP_UPD(I, (rM[z + k] + mi) + b, ip, 0 + t - 1);
        };
        if (l >= 5)
        {
#pragma unroll
          for (int t = 5; t >= 1; --t)
            candidates(t);
        }
        else
          for (int t = T; t >= 1; --t)
            candidates(t);
        rM[(size_t)sl * KG + k] = M;
        rI[(size_t)sl * KG + k] = I;
        if (valid) nodes[(size_t)l * K + k] = (uint16_t)((unsigned)mp | ((unsigned)ip << 6));
      }
      __syncthreads();

      // ---- phase 2 (warp 0): special states, delete chain, E, B, T, trellis xnode ----
      if (warp == 0)
      {
        float N = INF, Jv = INF, C = INF, R = INF;
        int pN = 0, pJ = 0, pC = 0;
        for (int t = T; t >= 1; --t)
        {
          int const z = (l - t) % 6;
          float const nil = __ldg(&pd.nulbg[code[t]]).x;
          DCP_UPD(N, xs[z][0] + SN + nil, pN, 0 + t - 1);
          DCP_UPD(N, xs[z][1] + NN + nil, pN, 5 + t - 1);
          DCP_UPD(Jv, xs[z][4] + EJ + nil, pJ, 0 + t - 1);
          DCP_UPD(Jv, xs[z][3] + JJ + nil, pJ, 5 + t - 1);
          DCP_UPD(C, xs[z][4] + EC + nil, pC, 0 + t - 1);
          DCP_UPD(C, xs[z][5] + CC + nil, pC, 5 + t - 1);
          R = fminf(R, xs[z][6] + RR + nil);
        }
        float carryM = INF, carryD = INF;
        float ev = INF;
        int ei = 0;
        for (int c0 = 0; c0 < KG; c0 += 32)
        {
          int const k = c0 + lane;
          bool const valid = k < K;
          int const pk = posk[k];
          float const md = valid ? __ldg(pd.core + C_MD * Kpad + pk) : INF;
          float const dd = valid ? __ldg(pd.core + C_DD * Kpad + pk) : INF;
          float const M = rM[(size_t)sl * KG + k];
          float mprev = __shfl_up_sync(FULL_MASK, M, 1);
          if (lane == 0) mprev = carryM;
          float D = mprev + md;
          int dbit = 0;
          for (;;)
          {
            float dp = __shfl_up_sync(FULL_MASK, D, 1);
            if (lane == 0) dp = carryD;
            float const c = dp + dd;
            bool const imp = c < D;
            if (!__any_sync(FULL_MASK, imp)) break;
            if (imp) { D = c; dbit = 1; }
          }
          carryM = __shfl_sync(FULL_MASK, M, 31);
          carryD = __shfl_sync(FULL_MASK, D, 31);
          DCP_UPD(ev, M, ei, 2 * k + 0);
          DCP_UPD(ev, D, ei, 2 * k + 1);
          rD[(size_t)sl * KG + k] = D;
          if (valid && dbit) nodes[(size_t)l * K + k] |= (uint16_t)(1u << 5);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          float const ov = __shfl_xor_sync(FULL_MASK, ev, o);
          int const oi = __shfl_xor_sync(FULL_MASK, ei, o);
          if (ov < ev || (ov == ev && oi < ei)) { ev = ov; ei = oi; }
        }
        float B = INF;
        int pB = 0, pT = 0;
        DCP_UPD(B, N + NB, pB, 1);
        DCP_UPD(B, ev + EB, pB, 2);
        DCP_UPD(B, Jv + JB, pB, 3);
        Tv = INF;
        DCP_UPD(Tv, ev + ET, pT, 0);
        DCP_UPD(Tv, C + CT, pT, 1);
        Rv = R;
        if (lane == 0)
        {
          xs[sl][0] = INF;
          xs[sl][1] = N;
          xs[sl][2] = B;
          xs[sl][3] = Jv;
          xs[sl][4] = ev;
          xs[sl][5] = C;
          xs[sl][6] = R;
          xnodes[l] = (uint32_t)pN | ((uint32_t)pB << 4) | ((uint32_t)ei << 6) | ((uint32_t)pC << 21) |
                      ((uint32_t)pT << 25) | ((uint32_t)pJ << 26);
        }
      }
      __syncthreads();
    }

    if (threadIdx.x == 0)
    {
      a.s.out[oidx] = make_float2(Rv, Tv);
      __threadfence();
      a.nsteps[oidx] = trellis_walk(K, L, xnodes, nodes, 0, nullptr, nullptr);
    }
    __syncthreads();
  }
}

struct WalkArgs
{
  ProfileDesc const *profiles;
  Pair const *pairs;
  long long npairs;
  uint32_t const *xnodes;
  uint16_t const *nodes;
  long long const *xnode_off;
  long long const *node_off;
  int const *nsteps;
  long long const *step_off;
  uint16_t *ids;
  uint8_t *sizes;
};

// One thread per traced pair: second back-walk, writing the steps in path order.
__global__ void walk_write_kernel(WalkArgs a)
{
  long long const i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= a.npairs) return;
  int const n = a.nsteps[i];
  if (n <= 0 || a.node_off[i + 1] == a.node_off[i]) return; // no trellis kept: walked lazily
  Pair const pr = a.pairs[i];
  int const K = a.profiles[pr.profile].K;
  trellis_walk(K, pr.len, a.xnodes + a.xnode_off[i], a.nodes + a.node_off[i], n,
               a.ids + a.step_off[i], a.sizes + a.step_off[i]);
}

} // namespace dcp

// ---- FP32 ALU issue-rate microbenchmark (the roofline denominator of the score kernel) ----
// Every operation is pinned with asm volatile so that ptxas cannot fuse or drop any of them.
// MODE 0: FADD + FMNMX 1:1 (the recurrence's own mix: 17 add : 16 two-input min per cell)
// MODE 1: FADD only   MODE 2: FMNMX only   MODE 3: FMNMX3 only   MODE 4: FADD2 (f32x2) only
// MODE 5: FADD + FMNMX3 2:1 (a cell written with three-input mins: 18 add : 9 min3)
namespace dcp {
#define DCP_FADD(d, a, b) asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define DCP_FMIN(d, a, b) asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define DCP_FMIN3(d, a, b, c) asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
template <int MODE>
__global__ void __launch_bounds__(256) alu_peak_kernel(float *out, int iters, float seed)
{
  float a[8], b[8], c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
  {
    a[j] = seed + (float)(threadIdx.x + j);
    b[j] = seed * 0.5f + (float)j;
    c[j] = seed * 0.25f + (float)(j * 3);
  }
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      if (MODE == 0) { DCP_FADD(a[j], a[j], b[j]); DCP_FMIN(c[j], c[j], a[j]); }
      if (MODE == 1) { DCP_FADD(a[j], a[j], b[j]); DCP_FADD(c[j], c[j], b[j]); }
      if (MODE == 2) { DCP_FMIN(a[j], a[j], b[j]); DCP_FMIN(c[j], c[j], b[j]); }
      if (MODE == 3) { DCP_FMIN3(a[j], a[j], b[j], c[j]); DCP_FMIN3(b[j], b[j], c[j], a[j]); }
      if (MODE == 5) { DCP_FADD(a[j], a[j], b[j]); DCP_FADD(b[j], b[j], c[j]); DCP_FMIN3(c[j], c[j], a[j], b[j]); }
      if (MODE == 6)
      { // the score row's own mix: two FADD per three-input INTEGER min (VIMNMX3, row_kernel.cuh)
        DCP_FADD(a[j], a[j], b[j]);
        DCP_FADD(b[j], b[j], c[j]);
        int m;
        asm volatile("min.s32 %0, %1, %2;" : "=r"(m) : "r"(__float_as_int(c[j])), "r"(__float_as_int(a[j])));
        c[j] = __int_as_float(min(m, __float_as_int(b[j])));
      }
    }
    if (MODE == 4)
    {
#pragma unroll
      for (int j = 0; j < 8; j += 2)
      {
        unsigned long long x, y;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[j]), "f"(a[j + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[j]), "f"(b[j + 1]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[j]), "=f"(a[j + 1]) : "l"(x));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    s += a[j] + b[j] + c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace dcp
