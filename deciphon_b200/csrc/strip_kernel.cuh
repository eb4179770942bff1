// Score pass for profiles of more than 256 nodes: speculative column strips.
//
// score_reg_kernel<Q,W> (W > 1) pays two CTA barriers and a second delete-chain round per DP
// row because B(l) = min(N(l)+NB, E(l)+EB, J(l)+JB) (c-core/viterbi.c:495-496,582-583) needs
// E(l), a minimum over ALL nodes of the row, before any node of row l+1 can start.  But for a
// pair without a domain the E and J terms essentially never win that minimum (measured: 98 % of
// random (read, profile) pairs have B(l) == N(l)+NB on every row, median margin 4 nats), and N(l)
// depends on the read alone.
//
// So a profile is cut into W STRIPS of 32*Q nodes and every strip runs the whole window with the
// single-warp software-pipelined row, ASSUMING B(l) = N(l)+NB.  The only coupling left is
// one-directional: strip w needs M, I, D (final) of the last node of strip w-1 for the same row.
// Strip w of every pair of a kernel class is ONE LAUNCH and strip w+1 the next one on the same
// stream; a warp runs one strip of one pair and the boundary {M, I, D, running min of E} of
// every row travels through a per-pair column in global memory (16 bytes a row, fetched one row
// ahead, overwritten in place).  The running minimum of E travels with it, so the LAST strip
// sees the true E(l): it carries the J and C states and checks, row by row, that the assumption
// held.  If it did (bit-for-bit, the minimum is the same float), every value computed is the
// exact one and alt = T(L) is written; otherwise the pair is queued for the exact multi-warp
// kernel.  Results are therefore always bit-identical to the reference -- speculation only
// decides which kernel produces them.
//
// Measured alternatives (profiles/README.md): the W strips of a pair as the W warps of one CTA,
// coupled through a shared-memory ring with producer/consumer row counters (15-25 % slower:
// polling and reconvergence instructions, two strips' short-code rows competing for L1), and one
// warp walking the strips of a pair inside one kernel (register-capped schedule).
#pragma once
#include "score_kernel.cuh"

namespace dcp {

struct StripArgs
{
  ScoreArgs s;
  long long *redo;               // pairs whose speculation failed
  unsigned long long *nredo;
  // one-strip-per-launch kernels: boundary column of launch item i at col + i * col_stride
  Mail *col;
  unsigned long long col_stride;
  int strip;                     // which strip of the pair this launch runs
  unsigned long long item0;      // first item of this launch's range ([item0, s.nitems))
  // segmented profiles (below): descriptor of segment `level` of profile p at segs[seg_first[p] + level];
  // column of an item: colmap[item] (explicit pairs) or colmap[profile slot] * nseq + read (grid)
  ProfileDesc const *segs;
  int const *seg_first;
  int level;
  long long const *colmap;
};

// Segmented profiles.  With one strip per launch nothing forces the strips of a profile to share
// a shape, so a profile of more than 256 nodes is ALSO stored as segments: 256 nodes each in the
// Q = 8 full-warp layout (the most efficient row there is), and a tail of 1..256 nodes in the
// layout a stand-alone profile of that size would get (sub-warp for <= 128 nodes).  The score
// pass runs the segments level by level; the exact kernels (redo, trace) keep using the
// whole-profile layout.  K = 300: 363 + 330/4 instructions per row instead of 2 x 290.
__device__ __forceinline__ ProfileDesc strip_desc(StripArgs const &a, int p)
{
  return a.segs ? a.segs[a.seg_first[p] + a.level] : a.s.profiles[p];
}

// d_lazy with a fixed incoming value for the head lane (the previous strip's final D)
template <int Q>
__device__ __forceinline__ float d_lazy_in(Lane<Q> const &s, float (&D)[Q], bool head, float head_in)
{
  float din;
  for (;;)
  {
    din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din = head_in;
    float const c = din + s.DD[0];
    if (!__any_sync(FULL_MASK, c < D[0])) break;
    D[0] = fminf(D[0], c);
    d_sweep<Q>(s, D);
  }
  return din;
}

// No spin waits, no co-scheduling: every launch has the L1 footprint and the register budget of
// score_reg_kernel<Q,1>; the column costs 32 bytes of HBM traffic per row against 33*32*Q flops.
// FIRST/LAST are compile-time: the first strip has no boundary to read, only the last one
// carries J and C, verifies the speculated B and writes the result (or queues the pair).
constexpr int LSTRIP_WARPS = 4; // independent (pair, strip) items per CTA

template <int Q, int W, int J, bool FIRST, bool LAST>
__device__ __forceinline__ void lstrip_row(Lane<Q> &s, float (&Mp)[Q], float (&Ip)[Q], float &xp, ProfileDesc const &pd,
                                           RowBase<Q, 32 * W> const &rb, unsigned hist, unsigned hist1, int lane,
                                           float NB, float EB, float JB, float4 &bnext, Mail *slot, float &E,
                                           float &x, bool &ok)
{
  constexpr int s1 = (J + 4) % 5, s2 = (J + 3) % 5, s3 = (J + 2) % 5, s4 = (J + 1) % 5;
  uint32_t const rowb = (uint32_t)pd.Kpad * 4u;

  // boundary of row l (requested during row l-1); request row l+1's
  float4 const b = bnext;
  if constexpr (!FIRST) bnext = __ldcg(reinterpret_cast<float4 const *>(slot + 1));

  // (A) finish row l with the one-nucleotide term (needs P(l-1), Q(l-1))
  float M[Q], I[Q];
  float xacc;
  {
    int const c1 = hist & 3;
    float2 const nb = ldg_nulbg(pd.nulbg, c1);
    float e[Q];
    rb.load(e, (uint32_t)c1 * rowb);
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      M[q] = fminf(Mp[q], s.P[s1][q] + e[q]);
      I[q] = fminf(Ip[q], s.Qv[s1][q] + nb.y);
    }
    xacc = fminf(xp, s.px[s1] + nb.x);
  }

  // emission rows of row l+1 for t = 2..5
  int const c2 = 4 + (hist1 & 15), c3 = 20 + (hist1 & 63), c4 = 84 + (hist1 & 255), c5 = 340 + (hist1 & 1023);
  float2 const nb2 = ldg_nulbg(pd.nulbg, c2), nb3 = ldg_nulbg(pd.nulbg, c3), nb4 = ldg_nulbg(pd.nulbg, c4),
               nb5 = ldg_nulbg(pd.nulbg, c5);
  float e2[Q], e3[Q], e4[Q], e5[Q];
  rb.load(e2, (uint32_t)c2 * rowb);
  rb.load(e3, (uint32_t)c3 * rowb);
  rb.load(e4, (uint32_t)c4 * rowb);
  rb.load(e5, (uint32_t)c5 * rowb);

  bool const head = !FIRST && lane == 0;

  // delete chain of row l (viterbi.c:538, 552-580); the head lane's predecessor is the boundary
  float mprev = __shfl_up_sync(FULL_MASK, M[Q - 1], 1);
  float iprev = __shfl_up_sync(FULL_MASK, I[Q - 1], 1);
  if (head)
  {
    mprev = b.x;
    iprev = b.y;
  }
  float D[Q];
  D[0] = mprev + s.MD[0];
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = M[q - 1] + s.MD[q];
  {
    float din0 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din0 = b.z;
    D[0] = fminf(D[0], din0 + s.DD[0]);
    d_sweep<Q>(s, D);
  }

  // row l+1, t = 2..5 (rows l-1..l-4 = ring slots s1..s4), in the shadow of the sweeps
#pragma unroll
  for (int q = 0; q < Q; ++q)
  {
    Mp[q] = fminf(min3(s.P[s1][q] + e2[q], s.P[s2][q] + e3[q], s.P[s3][q] + e4[q]), s.P[s4][q] + e5[q]);
    Ip[q] = fminf(min3(s.Qv[s1][q] + nb2.y, s.Qv[s2][q] + nb3.y, s.Qv[s3][q] + nb4.y), s.Qv[s4][q] + nb5.y);
  }
  xp = fminf(min3(s.px[s1] + nb2.x, s.px[s2] + nb3.x, s.px[s3] + nb4.x), s.px[s4] + nb5.x);

  {
    float din1 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din1 = b.z;
    D[0] = fminf(D[0], din1 + s.DD[0]);
    d_sweep<Q>(s, D);
  }
  float dprev;
  if constexpr (FIRST) dprev = d_lazy<Q>(s, D, false);
  else dprev = d_lazy_in<Q>(s, D, head, b.z);

  // running E over the strips so far; the last strip holds E(l) of the whole row
  float e = e_partial<Q>(M, D);
  if constexpr (!FIRST) e = fminf(e, b.w);
  if constexpr (!LAST)
  {
    if (lane == 31) __stcg(reinterpret_cast<float4 *>(slot), make_float4(M[Q - 1], I[Q - 1], D[Q - 1], e));
  }
  E = e;

  // special states.  Lane 0: N, lane 3: R in every strip; lanes 1, 2 (J, C) only mean something
  // in the last strip, which also verifies the speculated B.
  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0);
  float const B = N + NB;
  if constexpr (LAST)
  {
    float const Jv = __shfl_sync(FULL_MASK, x, 1);
    float const Btrue = min3(B, e + EB, Jv + JB); // viterbi.c:495-496,582-583
    ok = ok && (Btrue == B);
    s.px[J] = fminf(e + s.xa, x + s.xb);
  }
  else
    s.px[J] = x + s.xb;

  s.P[J][0] = fminf(min3(B + s.BM[0], mprev + s.MM[0], iprev + s.IM[0]), dprev + s.DM[0]);
#pragma unroll
  for (int q = 1; q < Q; ++q)
    s.P[J][q] = fminf(min3(B + s.BM[q], M[q - 1] + s.MM[q], I[q - 1] + s.IM[q]), D[q - 1] + s.DM[q]);
#pragma unroll
  for (int q = 0; q < Q; ++q)
    s.Qv[J][q] = fminf(I[q] + s.II[q], M[q] + s.MI[q]);
}

template <int Q, int W, bool FIRST, bool LAST>
__global__ void __launch_bounds__(32 * LSTRIP_WARPS, Q >= 6 ? 2 : 3) score_lstrip_kernel(StripArgs a)
{
  constexpr int VL = 32 * W;
  int const lane = threadIdx.x & 31;

  for (;;)
  {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.s.counter, 1ULL);
    item = __shfl_sync(FULL_MASK, item, 0);
    size_t colidx = (size_t)item; // columns are indexed inside the range unless a map is given
    item += a.item0;
    if (item >= a.s.nitems) break;

    int p, sq, start, L;
    long long oidx;
    if (a.s.pairs)
    {
      oidx = a.s.order[item];
      Pair const pr = a.s.pairs[oidx];
      p = pr.profile; sq = pr.seq; start = pr.start; L = pr.len;
      if (a.colmap) colidx = (size_t)a.colmap[item];
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
      if (a.colmap) colidx = (size_t)a.colmap[pi] * (unsigned)a.s.nseq + si;
    }
    ProfileDesc const pd = strip_desc(a, p);
    if (L < 0) L = min(min(pd.Kfull * 50, 100000), a.s.reads.seq_len[sq]);
    float const *xt = a.s.xt + (size_t)L * X_STRIDE;
    Mail *const col = a.col + colidx * a.col_stride;

    Lane<Q> s;
    int const Kpad = pd.Kpad;
    int const vl = a.strip * 32 + lane;
    RowBase<Q, VL> const rb(pd.em, vl);
    load_chunks<Q, VL>(s.BM, pd.core + C_BM * Kpad, vl);
    load_chunks<Q, VL>(s.MM, pd.core + C_MM * Kpad, vl);
    load_chunks<Q, VL>(s.MI, pd.core + C_MI * Kpad, vl);
    load_chunks<Q, VL>(s.MD, pd.core + C_MD * Kpad, vl);
    load_chunks<Q, VL>(s.IM, pd.core + C_IM * Kpad, vl);
    load_chunks<Q, VL>(s.II, pd.core + C_II * Kpad, vl);
    load_chunks<Q, VL>(s.DM, pd.core + C_DM * Kpad, vl);
    load_chunks<Q, VL>(s.DD, pd.core + C_DD * Kpad, vl);
    float const RR = xt[X_RR], SN = xt[X_SN], NN = xt[X_NN], SB = xt[X_SB], NB = xt[X_NB], EB = xt[X_EB],
                JB = xt[X_JB], EJ = xt[X_EJ], JJ = xt[X_JJ], EC = xt[X_EC], CC = xt[X_CC], ET = xt[X_ET],
                CT = xt[X_CT];
#pragma unroll
    for (int j = 0; j < 5; ++j)
    {
#pragma unroll
      for (int q = 0; q < Q; ++q)
      {
        s.P[j][q] = CUDART_INF_F;
        s.Qv[j][q] = CUDART_INF_F;
      }
      s.px[j] = CUDART_INF_F;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
      s.P[0][q] = SB + s.BM[q]; // row 0: B = SB (viterbi.c:472-473)
    s.xa = lane == 1 ? EJ : lane == 2 ? EC : CUDART_INF_F;
    s.xb = lane == 0 ? NN : lane == 1 ? JJ : lane == 2 ? CC : lane == 3 ? RR : CUDART_INF_F;
    s.px[0] = lane == 0 ? (0.0f + SN) : lane == 3 ? ((-RR) + RR) : CUDART_INF_F;

    // nucleotide stream, six positions ahead of the DP row (see score_one)
    uint32_t const *wp = a.s.reads.words + a.s.reads.seq_word[sq] + (start >> 4);
    uint32_t word = __ldg(wp) >> (2 * (start & 15));
    int left = 16 - (start & 15);
    unsigned H = 0;
#define DCP_NEXT_NT()                                                                            \
  {                                                                                              \
    H = ((H << 2) | (word & 3u)) & 0x3FFFFFu;                                                    \
    word >>= 2;                                                                                  \
    if (--left == 0)                                                                             \
    {                                                                                            \
      word = __ldg(++wp);                                                                        \
      left = 16;                                                                                 \
    }                                                                                            \
  }
#pragma unroll
    for (int i = 0; i < 6; ++i)
      DCP_NEXT_NT()
    float E = CUDART_INF_F, x = CUDART_INF_F;
    bool ok = true;
    float Mp[Q], Ip[Q], xp = CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      Mp[q] = CUDART_INF_F;
      Ip[q] = CUDART_INF_F;
    }
    float4 bnext = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    if constexpr (!FIRST) bnext = __ldcg(reinterpret_cast<float4 const *>(col + 1));
#define DCP_ROW(JJ_)                                                                             \
  {                                                                                              \
    if (l > L) break;                                                                            \
    DCP_NEXT_NT()                                                                                \
    lstrip_row<Q, W, JJ_, FIRST, LAST>(s, Mp, Ip, xp, pd, rb, (H >> 12) & 1023u, (H >> 10) & 1023u, lane, NB, EB,  \
                                       JB, bnext, col + l, E, x, ok);                            \
    ++l;                                                                                         \
  }
    int l = 1;
    for (;;)
    {
      DCP_ROW(1)
      DCP_ROW(2)
      DCP_ROW(3)
      DCP_ROW(4)
      DCP_ROW(0)
    }
#undef DCP_ROW
#undef DCP_NEXT_NT

    if constexpr (LAST)
    {
      float const C = __shfl_sync(FULL_MASK, x, 2);
      float const R = __shfl_sync(FULL_MASK, x, 3);
      if (lane == 0)
      {
        float const alt = fminf(E + ET, C + CT); // viterbi.c:585-586, 599
        if (ok)
        {
          a.s.out[oidx] = make_float2(R, alt);
          float const d = alt - R;
          if (d <= 0.0f && d > -CUDART_INF_F) atomicAdd(a.s.nhits, 1ULL);
        }
        else
          a.redo[atomicAdd(a.nredo, 1ULL)] = oidx; // the exact kernel will produce this pair
      }
    }
    __syncwarp();
  }
}

} // namespace dcp
