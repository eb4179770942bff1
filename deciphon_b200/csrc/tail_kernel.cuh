// Last segment of a segmented profile (strip_kernel.cuh) when it has at most 128 nodes: the
// sub-warp row of sub_kernel.cuh (G = 2/4/8 pairs per warp) with the strip kernels' boundary
// input and speculation check.  Each pair's segment lane 0 takes {M, I, D of node k0-1, running
// E} of row l from the pair's boundary column; B(l) = N(l)+NB is assumed and verified against
// the true minimum on every row of the pair's own window.
#pragma once
#include "strip_kernel.cuh"
#include "sub_kernel.cuh"

namespace dcp {

template <int Q, int SEG>
__device__ __forceinline__ float d_lazy_seg_in(Lane<Q> const &s, float (&D)[Q], bool head, float head_in)
{
  float din;
  for (;;)
  {
    din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    if (head) din = head_in;
    float const c = din + s.DD[0];
    if (!__any_sync(FULL_MASK, c < D[0])) break;
    D[0] = fminf(D[0], c);
    d_sweep<Q>(s, D);
  }
  return din;
}

template <int Q, int SEG, int J>
__device__ __forceinline__ void tail_row(Lane<Q> &s, float (&Mp)[Q], float (&Ip)[Q], float &xp,
                                         RowBase<Q, SEG> const &rb, float2 const *nulbg, uint32_t rowb, unsigned hist,
                                         unsigned hist1, int sl, float NB, float EB, float JB, float4 &bnext,
                                         Mail const *slot, float &E, float &x, bool &ok)
{
  constexpr int s1 = (J + 4) % 5, s2 = (J + 3) % 5, s3 = (J + 2) % 5, s4 = (J + 1) % 5;

  // boundary of row l (requested during row l-1); request row l+1's
  float4 const b = bnext;
  bnext = __ldcg(reinterpret_cast<float4 const *>(slot + 1));

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

  int const c2 = 4 + (hist1 & 15), c3 = 20 + (hist1 & 63), c4 = 84 + (hist1 & 255), c5 = 340 + (hist1 & 1023);
  float2 const nb2 = ldg_nulbg(nulbg, c2), nb3 = ldg_nulbg(nulbg, c3), nb4 = ldg_nulbg(nulbg, c4),
               nb5 = ldg_nulbg(nulbg, c5);
  float e2[Q], e3[Q], e4[Q], e5[Q];
  rb.load(e2, (uint32_t)c2 * rowb);
  rb.load(e3, (uint32_t)c3 * rowb);
  rb.load(e4, (uint32_t)c4 * rowb);
  rb.load(e5, (uint32_t)c5 * rowb);

  bool const head = sl == 0; // its predecessor node lives in the previous segment
  float mprev = __shfl_up_sync(FULL_MASK, M[Q - 1], 1, SEG);
  float iprev = __shfl_up_sync(FULL_MASK, I[Q - 1], 1, SEG);
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
    float din0 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    if (head) din0 = b.z;
    D[0] = fminf(D[0], din0 + s.DD[0]);
    d_sweep<Q>(s, D);
  }

#pragma unroll
  for (int q = 0; q < Q; ++q)
  {
    Mp[q] = fminf(min3(s.P[s1][q] + e2[q], s.P[s2][q] + e3[q], s.P[s3][q] + e4[q]), s.P[s4][q] + e5[q]);
    Ip[q] = fminf(min3(s.Qv[s1][q] + nb2.y, s.Qv[s2][q] + nb3.y, s.Qv[s3][q] + nb4.y), s.Qv[s4][q] + nb5.y);
  }
  xp = fminf(min3(s.px[s1] + nb2.x, s.px[s2] + nb3.x, s.px[s3] + nb4.x), s.px[s4] + nb5.x);

  {
    float din1 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    if (head) din1 = b.z;
    D[0] = fminf(D[0], din1 + s.DD[0]);
    d_sweep<Q>(s, D);
  }
  float const dprev = d_lazy_seg_in<Q, SEG>(s, D, head, b.z);

  float e = e_lane<Q>(M, D);
  e = fminf(seg_min<SEG>(e), b.w); // E(l) of the whole row
  E = e;

  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0, SEG);
  float const Jv = __shfl_sync(FULL_MASK, x, 1, SEG);
  float const B = N + NB;
  float const Btrue = min3(B, e + EB, Jv + JB); // viterbi.c:495-496,582-583
  ok = ok && (Btrue == B);
  s.px[J] = fminf(e + s.xa, x + s.xb);

  s.P[J][0] = fminf(min3(B + s.BM[0], mprev + s.MM[0], iprev + s.IM[0]), dprev + s.DM[0]);
#pragma unroll
  for (int q = 1; q < Q; ++q)
    s.P[J][q] = fminf(min3(B + s.BM[q], M[q - 1] + s.MM[q], I[q - 1] + s.IM[q]), D[q - 1] + s.DM[q]);
#pragma unroll
  for (int q = 0; q < Q; ++q)
    s.Qv[J][q] = fminf(I[q] + s.II[q], M[q] + s.MI[q]);
}

template <int Q, int G>
__global__ void __launch_bounds__(32 * SUB_GROUPS, Q >= 6 ? 2 : 3) score_subtail_kernel(StripArgs a)
{
  constexpr int SEG = 32 / G;
  int const lane = threadIdx.x & 31;
  int const seg = lane / SEG, sl = lane % SEG;

  for (;;)
  {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.s.counter, (unsigned long long)G);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (base >= a.s.nitems) break;
    bool const active = base + seg < a.s.nitems;
    unsigned long long const item = active ? base + seg : base; // idle segments shadow segment 0

    int p, sq, start, L;
    long long oidx;
    size_t colidx;
    if (a.s.pairs)
    {
      oidx = a.s.order[item];
      Pair const pr = a.s.pairs[oidx];
      p = pr.profile;
      sq = pr.seq;
      start = pr.start;
      L = pr.len;
      colidx = (size_t)a.colmap[item];
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
      colidx = (size_t)a.colmap[pi] * (unsigned)a.s.nseq + si;
    }
    ProfileDesc const pd = strip_desc(a, p);
    if (L < 0) L = min(min(pd.Kfull * 50, 100000), a.s.reads.seq_len[sq]);
    int const Lmax = __reduce_max_sync(FULL_MASK, L);
    float const *xt = a.s.xt + (size_t)L * X_STRIDE;
    Mail const *const col = a.col + colidx * a.col_stride;

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
      s.P[0][q] = SB + s.BM[q]; // row 0: B = SB (viterbi.c:472-473)
    s.xa = sl == 1 ? EJ : sl == 2 ? EC : CUDART_INF_F;
    s.xb = sl == 0 ? NN : sl == 1 ? JJ : sl == 2 ? CC : sl == 3 ? RR : CUDART_INF_F;
    s.px[0] = sl == 0 ? (0.0f + SN) : sl == 3 ? ((-RR) + RR) : CUDART_INF_F;

    uint32_t const *wp = a.s.reads.words + a.s.reads.seq_word[sq] + (start >> 4);
    uint32_t const *const wend = a.s.reads.words + a.s.reads.nwords - 1;
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
    bool ok = true, okres = true;
    float Mp[Q], Ip[Q], xp = CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      Mp[q] = CUDART_INF_F;
      Ip[q] = CUDART_INF_F;
    }
    float4 bnext = __ldcg(reinterpret_cast<float4 const *>(col + 1));
#define DCP_ROW(JJ_)                                                                             \
  {                                                                                              \
    if (l > Lmax) break;                                                                         \
    DCP_NEXT_NT()                                                                                \
    tail_row<Q, SEG, JJ_>(s, Mp, Ip, xp, rb, pd.nulbg, rowb, (H >> 12) & 1023u, (H >> 10) & 1023u, sl, NB, EB, JB, \
                          bnext, col + l, E, x, ok);                                             \
    if (l == L)                                                                                  \
    {                                                                                            \
      Eres = E;                                                                                  \
      xres = x;                                                                                  \
      okres = ok;                                                                                \
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
      if (okres)
      {
        a.s.out[oidx] = make_float2(R, alt);
        float const d = alt - R;
        if (d <= 0.0f && d > -CUDART_INF_F) atomicAdd(a.s.nhits, 1ULL);
      }
      else
        a.redo[atomicAdd(a.nredo, 1ULL)] = oidx; // the exact kernel will produce this pair
    }
    __syncwarp();
  }
}

} // namespace dcp
