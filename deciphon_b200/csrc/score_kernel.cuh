// Score-only pass: null + alternative model min-cost of every (window, profile) pair.
//
// Replaces, per pair, viterbi_null + viterbi_cost (c-core/viterbi.c:696-724), i.e. the
// cost() recurrence of c-core/viterbi.c:451-600 with path = 0, in the factored form of
// SURVEY App. A.1:
//     P_k(l) = min(B(l)+BM_k, M_{k-1}(l)+MM_k, I_{k-1}(l)+IM_k, D_{k-1}(l)+DM_k)
//     Q_k(l) = min(I_k(l)+II_k, M_k(l)+MI_k)
//     M_k(l) = min_t P_k(l-t) + em_k[code(l-t,t)]      t = 1..5
//     I_k(l) = min_t Q_k(l-t) + bg[code(l-t,t)]
//     D_k(l) = min(M_{k-1}(l)+MD_k, D_{k-1}(l)+DD_k)   (serial in k)
// which is value-exact (bit-identical) with the reference because fp32 rounding is
// monotone: min_i((s_i + tau_i) + e) == (min_i (s_i + tau_i)) + e.
//
// Mapping: ONE WARP PER PAIR.  The K nodes are striped across the 32 lanes like the
// reference stripes them across SIMD lanes (viterbi.c:220-221): lane = k / Q, q = k % Q,
// Q = ceil(K/32) <= 8.  Per lane everything lives in registers:
//   * the 8 transition costs of its Q nodes,
//   * a 5-row ring of P and Q (the reference's 6-slot time frame, viterbi.c:12,160-161;
//     the row being computed needs no slot of its own here), rotated by unrolling the row
//     loop 5x so that ring indices are compile-time constants,
// and the cross-lane k-1 dependency is one __shfl_up per state per row (the reference's
// shift(), intrinsics.h:95-106).  The serial delete chain is resolved like the reference's
// lazy sweeps (viterbi.c:561-580): one in-lane sweep, then boundary propagation repeated
// while any lane still improves (warp vote) -- every candidate is a left-to-right chain
// sum, so the fixed point is bit-identical to the serial recurrence.
// The special states are spread over lanes 0..3 (N, J, C and the null model's R), which
// all run the same "min_t prev[t] + null[code_t]" recurrence.
#pragma once
#include "layout.cuh"
#include <math_constants.h>

namespace dcp {

constexpr int SCORE_THREADS = 128;
constexpr unsigned FULL_MASK = 0xffffffffu;

struct ScoreArgs
{
  ProfileDesc const *profiles;
  ReadsView reads;
  float const *xt; // [maxlen+1][X_STRIDE]
  // grid mode (pairs == nullptr): item -> (class_profiles[item / nseq], seq0 + item % nseq)
  int const *class_profiles;
  int prof0;
  int seq0;
  int nseq;
  // explicit mode: item -> pairs[order[item]]
  Pair const *pairs;
  long long const *order;
  unsigned long long nitems;
  unsigned long long *counter; // work-stealing cursor
  float2 *out;                 // {null cost, alt cost} per pair
  unsigned long long *nhits;
};

template <int Q>
__device__ __forceinline__ void load_chunks(float (&e)[Q], float const *__restrict__ row, int lane)
{
  constexpr int N4 = Q / 4;
#pragma unroll
  for (int c = 0; c < N4; ++c)
  {
    float4 v = __ldg(reinterpret_cast<float4 const *>(row + c * 128) + lane);
    e[4 * c + 0] = v.x;
    e[4 * c + 1] = v.y;
    e[4 * c + 2] = v.z;
    e[4 * c + 3] = v.w;
  }
  if constexpr ((Q & 2) != 0)
  {
    float2 v = __ldg(reinterpret_cast<float2 const *>(row + 32 * (N4 * 4)) + lane);
    e[N4 * 4 + 0] = v.x;
    e[N4 * 4 + 1] = v.y;
  }
  if constexpr ((Q & 1) != 0) e[Q - 1] = __ldg(row + 32 * (Q - 1) + lane);
}

__device__ __forceinline__ float shfl_prev(float v, int lane)
{
  float r = __shfl_up_sync(FULL_MASK, v, 1);
  return lane == 0 ? CUDART_INF_F : r; // shift() fills lane 0 with +INF (intrinsics.h:95-106)
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

template <int Q>
struct Lane
{
  // transitions of this lane's Q nodes
  float BM[Q], MM[Q], MI[Q], MD[Q], IM[Q], II[Q], DM[Q], DD[Q];
  // ring: slot (l % 5) holds row l
  float P[5][Q], Qv[5][Q];
  float px[5]; // lane 0: N, 1: J, 2: C, 3: R (null model); "previous-row term" of each
  float xa, xb; // per-lane coefficients of the special-state update
};

// One DP row l (J = l % 5).  hist = last five nucleotides ending at l-1, 2 bits each.
template <int Q, int J>
__device__ __forceinline__ void dp_row(Lane<Q> &s, ProfileDesc const &pd, unsigned hist, int lane,
                                       float NB, float EB, float JB, float &E, float &x)
{
  int code[5];
  code[0] = hist & 3;
  code[1] = 4 + (hist & 15);
  code[2] = 20 + (hist & 63);
  code[3] = 84 + (hist & 255);
  code[4] = 340 + (hist & 1023);

  float M[Q], I[Q];
  float xacc = CUDART_INF_F;
#pragma unroll
  for (int t = 1; t <= 5; ++t)
  {
    int const slot = (J - t + 10) % 5;
    float2 nb = __ldg(pd.nulbg + code[t - 1]);
    float e[Q];
    load_chunks<Q>(e, pd.em + (size_t)code[t - 1] * pd.Kpad, lane);
    if (t == 1)
    {
#pragma unroll
      for (int q = 0; q < Q; ++q)
      {
        M[q] = s.P[slot][q] + e[q];
        I[q] = s.Qv[slot][q] + nb.y;
      }
      xacc = s.px[slot] + nb.x;
    }
    else
    {
#pragma unroll
      for (int q = 0; q < Q; ++q)
      {
        M[q] = fminf(M[q], s.P[slot][q] + e[q]);
        I[q] = fminf(I[q], s.Qv[slot][q] + nb.y);
      }
      xacc = fminf(xacc, s.px[slot] + nb.x);
    }
  }

  // delete chain (viterbi.c:538, 552-580)
  float D[Q];
  float const mprev = shfl_prev(M[Q - 1], lane);
  D[0] = mprev + s.MD[0];
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = M[q - 1] + s.MD[q];
  float din = shfl_prev(D[Q - 1], lane);
  D[0] = fminf(D[0], din + s.DD[0]);
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = fminf(D[q], D[q - 1] + s.DD[q]);
  for (;;)
  {
    din = shfl_prev(D[Q - 1], lane);
    float const c = din + s.DD[0];
    if (!__any_sync(FULL_MASK, c < D[0])) break;
    D[0] = fminf(D[0], c);
#pragma unroll
    for (int q = 1; q < Q; ++q)
      D[q] = fminf(D[q], D[q - 1] + s.DD[q]);
  }
  // din now holds the final D of node k-1 for q = 0

  // E(l) = min_k min(M_k, D_k)  (viterbi.c:540-558)
  float e = fminf(M[0], D[0]);
#pragma unroll
  for (int q = 1; q < Q; ++q)
    e = fminf(e, fminf(M[q], D[q]));
  E = warp_min(e);

  // special states: x is N(l) on lane 0, J(l) on lane 1, C(l) on lane 2, R(l) on lane 3
  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0);
  float const Jv = __shfl_sync(FULL_MASK, x, 1);
  float const B = fminf(fminf(N + NB, E + EB), Jv + JB); // viterbi.c:495-496,582-583
  s.px[J] = fminf(E + s.xa, x + s.xb);

  float const iprev = shfl_prev(I[Q - 1], lane);
  // P(l), Q(l) into the slot that held row l-5
  s.P[J][0] = fminf(fminf(B + s.BM[0], mprev + s.MM[0]), fminf(iprev + s.IM[0], din + s.DM[0]));
#pragma unroll
  for (int q = 1; q < Q; ++q)
    s.P[J][q] = fminf(fminf(B + s.BM[q], M[q - 1] + s.MM[q]), fminf(I[q - 1] + s.IM[q], D[q - 1] + s.DM[q]));
#pragma unroll
  for (int q = 0; q < Q; ++q)
    s.Qv[J][q] = fminf(I[q] + s.II[q], M[q] + s.MI[q]);
}

template <int Q>
__device__ __forceinline__ void score_one(ProfileDesc const &pd, uint32_t const *__restrict__ words,
                                          int start, int L, float const *__restrict__ xt, int lane,
                                          float &null_cost, float &alt_cost)
{
  Lane<Q> s;
  int const Kpad = pd.Kpad;
  load_chunks<Q>(s.BM, pd.core + C_BM * Kpad, lane);
  load_chunks<Q>(s.MM, pd.core + C_MM * Kpad, lane);
  load_chunks<Q>(s.MI, pd.core + C_MI * Kpad, lane);
  load_chunks<Q>(s.MD, pd.core + C_MD * Kpad, lane);
  load_chunks<Q>(s.IM, pd.core + C_IM * Kpad, lane);
  load_chunks<Q>(s.II, pd.core + C_II * Kpad, lane);
  load_chunks<Q>(s.DM, pd.core + C_DM * Kpad, lane);
  load_chunks<Q>(s.DD, pd.core + C_DD * Kpad, lane);

  float const RR = xt[X_RR], SN = xt[X_SN], NN = xt[X_NN], SB = xt[X_SB], NB = xt[X_NB],
              EB = xt[X_EB], JB = xt[X_JB], EJ = xt[X_EJ], JJ = xt[X_JJ], EC = xt[X_EC],
              CC = xt[X_CC], ET = xt[X_ET], CT = xt[X_CT];

  // row 0: S = 0, B = SB (viterbi.c:472-473); null R(0) = -RR (viterbi.c:703)
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
    s.P[0][q] = SB + s.BM[q];
  s.xa = lane == 1 ? EJ : lane == 2 ? EC : CUDART_INF_F;
  s.xb = lane == 0 ? NN : lane == 1 ? JJ : lane == 2 ? CC : lane == 3 ? RR : CUDART_INF_F;
  s.px[0] = lane == 0 ? (0.0f + SN) : lane == 3 ? ((-RR) + RR) : CUDART_INF_F;

  // nucleotide stream
  int g = start;
  uint32_t const *wp = words + (g >> 4);
  uint32_t word = __ldg(wp) >> (2 * (g & 15));
  int left = 16 - (g & 15);
  unsigned hist = 0;
  float E = CUDART_INF_F, x = CUDART_INF_F;

#define DCP_ROW(JJ_)                                                                             \
  {                                                                                              \
    if (l > L) break;                                                                            \
    hist = ((hist << 2) | (word & 3u)) & 1023u;                                                  \
    word >>= 2;                                                                                  \
    if (--left == 0)                                                                             \
    {                                                                                            \
      word = __ldg(++wp);                                                                        \
      left = 16;                                                                                 \
    }                                                                                            \
    dp_row<Q, JJ_>(s, pd, hist, lane, NB, EB, JB, E, x);                                         \
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

  float const C = __shfl_sync(FULL_MASK, x, 2);
  float const R = __shfl_sync(FULL_MASK, x, 3);
  alt_cost = fminf(E + ET, C + CT); // viterbi.c:585-586, 599
  null_cost = R;                    // viterbi.c:718
}

template <int Q>
__global__ void __launch_bounds__(SCORE_THREADS) score_reg_kernel(ScoreArgs a)
{
  int const lane = threadIdx.x & 31;
  for (;;)
  {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1ULL);
    item = __shfl_sync(FULL_MASK, item, 0);
    if (item >= a.nitems) break;

    int p, sq, start, len;
    long long oidx;
    if (a.pairs)
    {
      oidx = a.order[item];
      Pair const pr = a.pairs[oidx];
      p = pr.profile;
      sq = pr.seq;
      start = pr.start;
      len = pr.len;
    }
    else
    {
      int const pi = (int)(item / (unsigned)a.nseq);
      int const si = (int)(item - (unsigned long long)pi * (unsigned)a.nseq);
      p = a.class_profiles[pi];
      sq = a.seq0 + si;
      start = 0;
      oidx = (long long)(p - a.prof0) * a.nseq + si;
      len = -1;
    }
    ProfileDesc const pd = a.profiles[p];
    if (len < 0)
    { // first window of window.c:13-37: [0, min(50K, 100000, |seq|))
      int const w = min(pd.K * 50, 100000);
      len = min(w, a.reads.seq_len[sq]);
    }
    float nul, alt;
    score_one<Q>(pd, a.reads.words + a.reads.seq_word[sq], start, len, a.xt + (size_t)len * X_STRIDE,
                 lane, nul, alt);
    if (lane == 0)
    {
      a.out[oidx] = make_float2(nul, alt);
      float const d = alt - nul; // lrt = -2*((-nul) - (-alt)) >= 0  <=>  alt - nul <= 0
      if (d <= 0.0f && d > -CUDART_INF_F) atomicAdd(a.nhits, 1ULL);
    }
  }
}

} // namespace dcp
