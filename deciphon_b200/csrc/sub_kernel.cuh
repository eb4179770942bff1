// Score (and dump) kernel for profiles of at most 128 nodes: G = 2, 4 or 8 pairs share a warp.
//
// The single-warp kernel spends a fixed number of instructions per DP row on things that do
// not scale with the nodes a lane holds -- the k-1 shuffles, the delete-chain votes, the E
// reduction, the special states, the nucleotide stream, the row addressing -- so at Q = 1..4
// nodes per lane (K <= 128) it ran at 140-380 GCUPS against 470 at Q = 8.  Here a pair gets
// SEG = 32/G lanes with Q = 5..8 nodes each (layout.cuh: VL = 16/8/4), and the G pairs of a warp
// run the SAME instruction stream: one width-SEG shuffle moves the k-1 values of all of them,
// one vote closes all their delete chains, N, J, C and the null model's R live on the first four
// lanes of every segment.  Everything that was warp-uniform in score_kernel.cuh (pair, profile
// pointers, window length, special transitions, nucleotide history) is simply per lane here.
// The recurrence, its operation order and therefore every bit of the results are those of
// dp_row (score_kernel.cuh); rows past a pair's own window keep running on whatever the stream
// holds (its result was captured at its last row) until the longest window of the warp is done.
#pragma once
#include "score_kernel.cuh"

namespace dcp {

constexpr int SUB_GROUPS = 4; // warps per CTA, each with its own G pairs

template <int SEG>
__device__ __forceinline__ float seg_min(float v)
{
#pragma unroll
  for (int o = SEG / 2; o > 0; o >>= 1)
    v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o, SEG));
  return v;
}

template <int Q, int SEG>
__device__ __forceinline__ float d_lazy_seg(Lane<Q> const &s, float (&D)[Q])
{
  float din;
  for (;;)
  {
    din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    float const c = din + s.DD[0];
    if (!__any_sync(FULL_MASK, c < D[0])) break;
    D[0] = fminf(D[0], c);
    d_sweep<Q>(s, D);
  }
  return din;
}

// One software-pipelined DP row for the G pairs of a warp; see dp_row for the structure.
template <int Q, int SEG, int J, bool DUMP>
__device__ __forceinline__ void sub_row(Lane<Q> &s, float (&Mp)[Q], float (&Ip)[Q], float &xp,
                                        RowBase<Q, SEG> const &rb, float2 const *nulbg, uint32_t rowb, unsigned hist,
                                        unsigned hist1, int sl, float NB, float EB, float JB, float &E, float &x,
                                        DumpRef<DUMP> const &dv, int Kpad, int l, bool in_window)
{
  constexpr int s1 = (J + 4) % 5, s2 = (J + 3) % 5, s3 = (J + 2) % 5, s4 = (J + 1) % 5;

  // (A) finish row l: the t = 1 term needs P(l-1), Q(l-1)
  float M[Q], I[Q];
  float xacc;
  {
    int const c1 = hist & 3;
    float2 const nb = ldg_nulbg(nulbg, c1);
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
  float2 const nb2 = ldg_nulbg(nulbg, c2), nb3 = ldg_nulbg(nulbg, c3), nb4 = ldg_nulbg(nulbg, c4),
               nb5 = ldg_nulbg(nulbg, c5);
  float e2[Q], e3[Q], e4[Q], e5[Q];
  rb.load(e2, (uint32_t)c2 * rowb);
  rb.load(e3, (uint32_t)c3 * rowb);
  rb.load(e4, (uint32_t)c4 * rowb);
  rb.load(e5, (uint32_t)c5 * rowb);

  // Delete chain of row l (viterbi.c:538, 552-567).  The first lane of a segment is node 0,
  // whose incoming transitions are +INF (protein.c:366-370): the value a width-SEG shuffle
  // leaves there (its own) is inert.
  float const mprev = __shfl_up_sync(FULL_MASK, M[Q - 1], 1, SEG);
  float const iprev = __shfl_up_sync(FULL_MASK, I[Q - 1], 1, SEG);
  float D[Q];
  D[0] = mprev + s.MD[0];
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = M[q - 1] + s.MD[q];
  {
    float const din0 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    D[0] = fminf(D[0], din0 + s.DD[0]);
    d_sweep<Q>(s, D);
  }

  // row l+1, t = 2..5: rows l-1, l-2, l-3, l-4 are ring slots s1..s4 of THIS row
#pragma unroll
  for (int q = 0; q < Q; ++q)
  {
    Mp[q] = fminf(min3(s.P[s1][q] + e2[q], s.P[s2][q] + e3[q], s.P[s3][q] + e4[q]), s.P[s4][q] + e5[q]);
    Ip[q] = fminf(min3(s.Qv[s1][q] + nb2.y, s.Qv[s2][q] + nb3.y, s.Qv[s3][q] + nb4.y), s.Qv[s4][q] + nb5.y);
  }
  xp = fminf(min3(s.px[s1] + nb2.x, s.px[s2] + nb3.x, s.px[s3] + nb4.x), s.px[s4] + nb5.x);

  {
    float const din1 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    D[0] = fminf(D[0], din1 + s.DD[0]);
    d_sweep<Q>(s, D);
  }
  float const dprev = d_lazy_seg<Q, SEG>(s, D);

  // E(l) = min_k min(M_k, D_k) over the pair's own lanes (viterbi.c:540-558)
  float e = e_lane<Q>(M, D);
  E = seg_min<SEG>(e);

  // special states: x is N(l) on segment lane 0, J(l) on 1, C(l) on 2, R(l) on 3
  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0, SEG);
  float const Jv = __shfl_sync(FULL_MASK, x, 1, SEG);
  float const B = min3(N + NB, E + EB, Jv + JB); // viterbi.c:495-496,582-583
  s.px[J] = fminf(E + s.xa, x + s.xb);

  if constexpr (DUMP)
  {
    if (in_window)
    { // the row's final values, for the walk / argmin kernels (lane-chunked order, layout.cuh)
      size_t const at = (size_t)(l - 1) * Kpad;
      store_chunks<Q, SEG>(dv.M + at, sl, M);
      store_chunks<Q, SEG>(dv.I + at, sl, I);
      store_chunks<Q, SEG>(dv.D + at, sl, D);
      float *xr = dv.xs + (size_t)(l - 1) * 8;
      if (sl == 0)
      {
        xr[0] = N;
        xr[1] = B;
        xr[2] = Jv;
        xr[3] = E;
      }
      if (sl == 2) xr[4] = x; // C(l)
    }
  }

  // P(l), Q(l) into the slot that held row l-5
  s.P[J][0] = fminf(min3(mprev + s.MM[0], iprev + s.IM[0], dprev + s.DM[0]), B + s.BM[0]); // B last: it arrives last
#pragma unroll
  for (int q = 1; q < Q; ++q)
    s.P[J][q] = fminf(min3(M[q - 1] + s.MM[q], I[q - 1] + s.IM[q], D[q - 1] + s.DM[q]), B + s.BM[q]);
#pragma unroll
  for (int q = 0; q < Q; ++q)
    s.Qv[J][q] = fminf(I[q] + s.II[q], M[q] + s.MI[q]);
}

template <int Q, int G, bool DUMP = false>
__global__ void __launch_bounds__(32 * SUB_GROUPS, Q >= 6 ? 2 : 3) score_sub_kernel(ScoreArgs a)
{
  constexpr int SEG = 32 / G;
  int const lane = threadIdx.x & 31;
  int const seg = lane / SEG, sl = lane % SEG;

  for (;;)
  {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.counter, (unsigned long long)G);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (base >= a.nitems) break;
    bool const active = base + seg < a.nitems;
    unsigned long long const item = active ? base + seg : base; // idle segments shadow segment 0

    int p, sq, start, L;
    long long oidx;
    if (a.pairs)
    {
      oidx = a.order[item];
      Pair const pr = a.pairs[oidx];
      p = pr.profile;
      sq = pr.seq;
      start = pr.start;
      L = pr.len;
      if (a.out_index) oidx = a.out_index[item];
    }
    else
    {
      int const pi = (int)(item / (unsigned)a.nseq);
      int const si = (int)(item - (unsigned long long)pi * (unsigned)a.nseq);
      p = a.class_profiles[pi];
      sq = a.seq0 + si;
      start = 0;
      oidx = (long long)(p - a.prof0) * a.nseq + si;
      L = -1;
    }
    ProfileDesc const pd = a.profiles[p];
    if (L < 0) L = min(min(pd.K * 50, 100000), a.reads.seq_len[sq]);
    int const Lmax = __reduce_max_sync(FULL_MASK, L);
    float const *xt = a.xt + (size_t)L * X_STRIDE;
    DumpRef<DUMP> const dv(DUMP ? a.dump + a.dump_off[item] : nullptr, L, pd.Kpad);

    Lane<Q> s;
    int const Kpad = pd.Kpad;
    load_chunks<Q, SEG>(s.BM, pd.core + C_BM * Kpad, sl);
    load_chunks<Q, SEG>(s.MM, pd.core + C_MM * Kpad, sl);
    load_chunks<Q, SEG>(s.MI, pd.core + C_MI * Kpad, sl);
    load_chunks<Q, SEG>(s.MD, pd.core + C_MD * Kpad, sl);
    load_chunks<Q, SEG>(s.IM, pd.core + C_IM * Kpad, sl);
    load_chunks<Q, SEG>(s.II, pd.core + C_II * Kpad, sl);
    load_chunks<Q, SEG>(s.DM, pd.core + C_DM * Kpad, sl);
    load_chunks<Q, SEG>(s.DD, pd.core + C_DD * Kpad, sl);
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
      s.P[0][q] = SB + s.BM[q]; // row 0: S = 0, B = SB (viterbi.c:472-473)
    s.xa = sl == 1 ? EJ : sl == 2 ? EC : CUDART_INF_F;
    s.xb = sl == 0 ? NN : sl == 1 ? JJ : sl == 2 ? CC : sl == 3 ? RR : CUDART_INF_F;
    s.px[0] = sl == 0 ? (0.0f + SN) : sl == 3 ? ((-RR) + RR) : CUDART_INF_F; // null R(0) = -RR, viterbi.c:703

    // nucleotide stream, six positions ahead of the DP row (see score_one); never past the buffer
    uint32_t const *wp = a.reads.words + a.reads.seq_word[sq] + (start >> 4);
    uint32_t const *const wend = a.reads.words + a.reads.nwords - 1;
    uint32_t word = __ldg(wp) >> (2 * (start & 15));
    int left = 16 - (start & 15);
    unsigned H = 0;
#define DCP_NEXT_NT()                                                                            \
  {                                                                                              \
    H = ((H << 2) | (word & 3u)) & 0x3FFFFFu;                                                    \
    word >>= 2;                                                                                  \
    if (--left == 0)                                                                             \
    {                                                                                            \
      wp = wp < wend ? wp + 1 : wp;                                                              \
      word = __ldg(wp);                                                                          \
      left = 16;                                                                                 \
    }                                                                                            \
  }
#pragma unroll
    for (int i = 0; i < 6; ++i)
      DCP_NEXT_NT()

    RowBase<Q, SEG> const rb(pd.em, sl);
    uint32_t const rowb = (uint32_t)Kpad * 4u;
    float E = CUDART_INF_F, x = CUDART_INF_F, Eres = CUDART_INF_F, xres = CUDART_INF_F;
    float Mp[Q], Ip[Q], xp = CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      Mp[q] = CUDART_INF_F;
      Ip[q] = CUDART_INF_F;
    }
#define DCP_ROW(JJ_)                                                                             \
  {                                                                                              \
    if (l > Lmax) break;                                                                         \
    DCP_NEXT_NT()                                                                                \
    sub_row<Q, SEG, JJ_, DUMP>(s, Mp, Ip, xp, rb, pd.nulbg, rowb, (H >> 12) & 1023u, (H >> 10) & 1023u, sl, NB, EB, \
                               JB, E, x, dv, Kpad, l, l <= L);                                   \
    if (l == L)                                                                                  \
    {                                                                                            \
      Eres = E;                                                                                  \
      xres = x;                                                                                  \
    }                                                                                            \
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

    float const C = __shfl_sync(FULL_MASK, xres, 2, SEG);
    float const R = __shfl_sync(FULL_MASK, xres, 3, SEG);
    if (sl == 0 && active)
    {
      float const alt = fminf(Eres + ET, C + CT); // viterbi.c:585-586, 599
      a.out[oidx] = make_float2(R, alt);           // null cost: viterbi.c:718
      float const d = alt - R;
      if (d <= 0.0f && d > -CUDART_INF_F) atomicAdd(a.nhits, 1ULL);
    }
    __syncwarp();
  }
}

} // namespace dcp
