// The single-warp DP row (v2) and the one kernel template built on it.
//
// Replaces, per (window, profile) pair, viterbi_null + viterbi_cost (c-core/viterbi.c:696-724),
// i.e. the cost() recurrence of c-core/viterbi.c:451-600 with path = 0, in the factored form of
// SURVEY App. A.1 (see score_kernel.cuh for the algebra and the bit-exactness argument).
//
// One template, score_row_kernel<Q, SEG, MODE, DUMP>, covers every single-warp shape:
//   SEG = 32        one pair per warp, Q = 1..8 nodes per lane             (128 < K <= 256, and K <= 128 when
//                                                                            sub-warp layouts are switched off)
//   SEG = 16/8/4    G = 32/SEG pairs per warp, Q = 5..8 nodes per lane      (K <= 128)
//   MODE = WHOLE    a whole profile: true B(l), result written
//          FIRST    first 256-node segment of a profile of more than 256 nodes: B(l) = N(l)+NB
//                   speculated, boundary column {M, I, D of the last node, running E} written
//          MID      later full segment: boundary column read (one row ahead) and rewritten in place
//          LAST     tail segment: boundary read, J and C carried, the speculated B verified row by
//                   row; the pair's result is written or the pair is queued for the exact kernel
//   DUMP            (WHOLE only) every row's M, I, D, N, B, J, E, C streamed out for the trace pass
//
// What changed against the round-1 row (score_kernel.cuh:dp_row and its three copies), all
// value-exact:
//  * Instruction selection follows tools/alu_probe.cu on B200: the two-input FMNMX is HALF rate
//    (0.50 warp-instructions/clk/SMSP, like FMNMX3), and the best scalar mix for this recurrence is
//    FADD + three-input INTEGER min 2:1 (VIMNMX3, 0.91 IPC against 0.78 for the FADD:FMNMX:FMNMX3
//    4:2:1 mix of round 1).  All DP values are >= +0 or +INF (checked per profile at upload,
//    score_kernel.cuh), so the signed-integer order of the bit patterns is the float order and
//    min3 on the patterns returns the same bits as fminf.  Every min is a three-input one where
//    the recurrence allows: M and I take {partial of t = 2..4, t = 1 term, t = 5 term}.
//  * E(l) = min_k M_k(l): every D_k is a chain M_j + (costs >= 0), and fp32 addition of a
//    non-negative term never decreases a value, so min_k D_k >= min_k M_k and the delete states
//    cannot change the VALUE of E.  E no longer waits for the delete chain.
//  * The second and later delete-chain sweeps carry only the incoming chain: c_q = c_{q-1} + DD_q,
//    D_q = min(D_q, c_q).  Same left-to-right sums, same minima as a full re-sweep (D is already
//    closed inside the lane), half the dependent latency.
//  * The ten-bit nucleotide history of every read position is precomputed once per batch
//    (hist_kernel, 2 bytes per nucleotide): a row loads one u16 instead of shifting a bit stream,
//    and table rows are addressed with one IMAD.WIDE (fma pipe) per code class and immediates.
#pragma once
#include "score_kernel.cuh"
#include "strip_kernel.cuh"

namespace dcp {

enum RowMode { ROW_WHOLE = 0, ROW_FIRST = 1, ROW_MID = 2, ROW_LAST = 3 };

constexpr int ROW_WARPS = 4;
// The staged kernels also keep the profile's {null, background} table (float2 per code) in shared
// memory: all 1364 codes where that fits beside the rows at the kernel's CTAs per SM, the 340 codes of
// the 1..4-mers at Q = 6 (3 CTAs x (64.5 + 10.9 + 1) KB would exceed the SM's 227 KB).
template <int Q>
__host__ __device__ constexpr int stage_nulbg_codes()
{
  return Q == 6 ? 340 : NCODES;
}         // warps per CTA, each with its own pair(s)
constexpr int HIST_SLACK = 100032;   // u16 entries past the last read (a sub-warp lane runs to its warp's longest window)

// ---- integer three-input min on non-negative floats -------------------------------------------
__device__ __forceinline__ float imin3(float a, float b, float c)
{
  return __int_as_float(min(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c)));
}
__device__ __forceinline__ float imin2(float a, float b)
{
  return __int_as_float(min(__float_as_int(a), __float_as_int(b)));
}
// DCP_LT must be the order DCP_MIN2 minimises in: a sub-warp lane that runs past its own window
// reads stale boundary rows, and a negative float there orders differently as an integer -- with a
// float compare and an integer min the lazy delete-chain loop below would never settle.
#ifndef DCP_ROW_FLOAT_MIN
#define DCP_MIN3(a, b, c) imin3(a, b, c)
#define DCP_MIN2(a, b) imin2(a, b)
#define DCP_LT(a, b) (__float_as_int(a) < __float_as_int(b))
#else
#define DCP_MIN3(a, b, c) min3(a, b, c)
#define DCP_MIN2(a, b) fminf(a, b)
#define DCP_LT(a, b) ((a) < (b))
#endif

// base + code * stride (bytes).  The stride is a run-time operand (mad.wide.u32) only with
// DCP_ROW_RT_STRIDE: measured here, ptxas then emits IMAD.WIDE + IADD3 + IADD3.X (the base pointers
// do not land in aligned register pairs at 255 registers) against LEA + LEA.HI.X for a
// power-of-two immediate, so the immediate form is the default.
__device__ __forceinline__ char const *mad_ptr(char const *base, uint32_t code, uint32_t stride)
{
#ifdef DCP_ROW_RT_STRIDE
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(code), "r"(stride), "l"((unsigned long long)base));
  return reinterpret_cast<char const *>(r);
#else
  return base + (size_t)code * stride;
#endif
}

// one probe of an mbarrier phase; the caller loops in C++ so that the control flow stays structured
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}

template <int OFF>
__device__ __forceinline__ float2 ld_nulbg(float2 const *nulbg, uint32_t idx, uint32_t eight);

constexpr int STAGE_ROWS = 84; // code rows of the 1-, 2- and 3-mers: contiguous at the head of em[1364][Kpad]

// Emission rows of a profile striped over SEG lanes with Q nodes each (layout.cuh): per-lane base
// pointers of the float4 / float2 / float chunks; a code row is base + code * ROWB + constant.
// STAGE: the 84 short-code rows are read from the CTA's shared-memory copy (same layout, staged by
// one TMA bulk copy per profile), the 4-/5-mer rows from global memory as before.
template <int Q, int SEG, bool STAGE = false>
struct EmRows
{
  static constexpr uint32_t ROWB = 4u * SEG * Q; // bytes per code row (Kpad = SEG * Q)
  static constexpr int N4 = Q / 4;
  char const *b4, *b2, *b1;
  uint32_t s4, s2, s1; // shared-memory addresses of the same chunks (STAGE)
  uint32_t snb;        // shared-memory address of the staged {null, background} table, 0 = not staged
  uint32_t rowb; // == ROWB, as a run-time value (see mad_ptr)
  __device__ __forceinline__ uint32_t stride() const
  {
#ifdef DCP_ROW_RT_STRIDE
    return rowb;
#else
    return ROWB;
#endif
  }
  __device__ __forceinline__ EmRows(float const *em, int sl, uint32_t rowb_, uint32_t stage = 0, uint32_t stage_nb = 0)
      : snb(stage_nb), rowb(rowb_)
  {
    b4 = reinterpret_cast<char const *>(em) + (size_t)sl * 16;
    b2 = reinterpret_cast<char const *>(em + SEG * (N4 * 4)) + (size_t)sl * 8;
    b1 = reinterpret_cast<char const *>(em + SEG * (Q - 1)) + (size_t)sl * 4;
    s4 = stage + (uint32_t)sl * 16u;
    s2 = stage + 4u * SEG * (N4 * 4) + (uint32_t)sl * 8u;
    s1 = stage + 4u * SEG * (Q - 1) + (uint32_t)sl * 4u;
  }
  // {null, background} emission costs of code OFF + idx
  template <int OFF>
  __device__ __forceinline__ float2 nulbg(float2 const *table, uint32_t idx, uint32_t eight) const
  {
    if constexpr (STAGE && OFF < stage_nulbg_codes<Q>())
    {
      float2 v;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(snb + idx * 8u + (uint32_t)OFF * 8u));
      return v;
    }
    else
      return ld_nulbg<OFF>(table, idx, eight);
  }
  // row of code OFF + idx
  template <int OFF>
  __device__ __forceinline__ void load(float (&e)[Q], uint32_t idx) const
  {
    constexpr uint32_t C = (uint32_t)OFF * ROWB;
    if constexpr (STAGE && OFF < STAGE_ROWS)
    {
      uint32_t const r = idx * ROWB + C;
      if constexpr (N4 > 0)
      {
#pragma unroll
        for (int c = 0; c < N4; ++c)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(e[4 * c + 0]), "=f"(e[4 * c + 1]), "=f"(e[4 * c + 2]), "=f"(e[4 * c + 3])
                       : "r"(s4 + r + (uint32_t)c * 16u * SEG));
      }
      if constexpr ((Q & 2) != 0)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e[N4 * 4 + 0]), "=f"(e[N4 * 4 + 1]) : "r"(s2 + r));
      if constexpr ((Q & 1) != 0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e[Q - 1]) : "r"(s1 + r));
      return;
    }
    if constexpr (N4 > 0)
    {
      char const *p = mad_ptr(b4, idx, stride());
#pragma unroll
      for (int c = 0; c < N4; ++c)
      {
        float4 v = __ldg(reinterpret_cast<float4 const *>(p + C + c * 16 * SEG));
        e[4 * c + 0] = v.x;
        e[4 * c + 1] = v.y;
        e[4 * c + 2] = v.z;
        e[4 * c + 3] = v.w;
      }
    }
    if constexpr ((Q & 2) != 0)
    {
      float2 v = __ldg(reinterpret_cast<float2 const *>(mad_ptr(b2, idx, stride()) + C));
      e[N4 * 4 + 0] = v.x;
      e[N4 * 4 + 1] = v.y;
    }
    if constexpr ((Q & 1) != 0) e[Q - 1] = __ldg(reinterpret_cast<float const *>(mad_ptr(b1, idx, stride()) + C));
  }
};

template <int OFF>
__device__ __forceinline__ float2 ld_nulbg(float2 const *nulbg, uint32_t idx, uint32_t eight)
{
#ifndef DCP_ROW_RT_STRIDE
  eight = 8u;
#endif
  return __ldg(reinterpret_cast<float2 const *>(mad_ptr(reinterpret_cast<char const *>(nulbg), idx, eight) + OFF * 8));
}

// min over this lane's match states (E(l) = min_k M_k(l), see the header)
template <int Q>
__device__ __forceinline__ float m_lane(float const (&M)[Q])
{
  float a = M[0];
  if constexpr (Q == 1) return a;
  if constexpr (Q == 2) return DCP_MIN2(a, M[1]);
  a = DCP_MIN3(a, M[1], M[2]);
#pragma unroll
  for (int q = 3; q + 1 < Q; q += 2)
    a = DCP_MIN3(a, M[q], M[q + 1]);
  if constexpr (Q > 3 && (Q & 1) == 0) a = DCP_MIN2(a, M[Q - 1]);
  return a;
}

template <int SEG>
__device__ __forceinline__ float seg_min_nonneg(float v)
{
  if constexpr (SEG == 32) return warp_min_nonneg(v);
  else
  {
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1)
      v = DCP_MIN2(v, __shfl_xor_sync(FULL_MASK, v, o, SEG));
    return v;
  }
}

// State a pair carries from row to row besides Lane<Q>: the partial minima of the NEXT row.
template <int Q>
struct Partial
{
  float Mp[Q]; // min over t = 2..4 of P(l+1-t) + em[code_t]
  float M5[Q]; // P(l+1-5) + em[code_5]
  float Ip[Q]; // min over t = 2..4 of Q(l+1-t) + bg[code_t]
  float xp;    // special states, t = 2..4
};

// One DP row l (J = l % 5).  h = history ending at nucleotide l-1 (codes of row l), hn = ending at
// nucleotide l (codes of row l+1).
template <int Q, int SEG, int MODE, bool DUMP, int J, class EM>
__device__ __forceinline__ void row_v2(Lane<Q> &s, Partial<Q> &pt, EM const &em, float2 const *nulbg,
                                       uint32_t eight, uint32_t h, uint32_t hn, int sl, float NB, float EB, float JB, float4 &bnext,
                                       float4 &bnext2, Mail *slot, float &E, float &x, bool &ok, DumpRef<DUMP> const &dv, int l,
                                       bool in_window, bool own)
{
  constexpr int s1 = (J + 4) % 5, s2 = (J + 3) % 5, s3 = (J + 2) % 5, s4 = (J + 1) % 5;
  constexpr bool HEAD_IN = MODE == ROW_MID || MODE == ROW_LAST;  // lane 0's predecessor is the previous segment
  constexpr bool COL_OUT = MODE == ROW_FIRST || MODE == ROW_MID;
  constexpr bool SPEC_B = MODE != ROW_WHOLE;

  // boundary of row l (requested during row l-2: the column left L2 long ago and a DRAM round trip
  // under load outlasts one row); request row l+2's
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (HEAD_IN)
  {
    b = bnext;
    bnext = bnext2;
    bnext2 = __ldcg(reinterpret_cast<float4 const *>(slot + 2));
  }

  // (A) finish row l: the t = 1 term needs P(l-1), Q(l-1); the t = 5 terms read the ring slot this
  // row overwrites at its end (row l-5)
  float M[Q], I[Q];
  float xacc;
  {
    float2 const nb1 = em.template nulbg<0>(nulbg, h & 3u, eight);
    float2 const nb5 = em.template nulbg<340>(nulbg, h & 1023u, eight);
    float e[Q];
    em.template load<0>(e, h & 3u);
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      M[q] = DCP_MIN3(pt.Mp[q], pt.M5[q], s.P[s1][q] + e[q]);
      I[q] = DCP_MIN3(pt.Ip[q], s.Qv[J][q] + nb5.y, s.Qv[s1][q] + nb1.y);
    }
    xacc = DCP_MIN3(pt.xp, s.px[J] + nb5.x, s.px[s1] + nb1.x);
  }

  // E(l) = min_k M_k(l); with segments, the running minimum over the segments so far
  float e = seg_min_nonneg<SEG>(m_lane<Q>(M));
  if constexpr (HEAD_IN) e = DCP_MIN2(e, b.w);
  E = e;

  // emission rows of row l+1 for t = 2..5
  float2 const nb2 = em.template nulbg<4>(nulbg, hn & 15u, eight), nb3 = em.template nulbg<20>(nulbg, hn & 63u, eight),
               nb4 = em.template nulbg<84>(nulbg, hn & 255u, eight);
  float e2[Q], e3[Q], e4[Q], e5[Q];
  em.template load<4>(e2, hn & 15u);
  em.template load<20>(e3, hn & 63u);
  em.template load<84>(e4, hn & 255u);
  em.template load<340>(e5, hn & 1023u);

  // Delete chain of row l (viterbi.c:538, 552-580).  The first lane of a whole profile is node 0,
  // whose incoming transitions are +INF (protein.c:366-370): the value a shuffle leaves there is
  // inert.  With a boundary, the head lane's predecessor is the previous segment's last node.
  bool const head = HEAD_IN && sl == 0;
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
    D[0] = DCP_MIN2(D[0], din0 + s.DD[0]);
#pragma unroll
    for (int q = 1; q < Q; ++q)
      D[q] = DCP_MIN2(D[q], D[q - 1] + s.DD[q]);
  }

  // special states (lane 0: N, 1: J, 2: C, 3: R of its pair) and B(l)
  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0, SEG);
  float B = N + NB;
  float Jv = 0.f;
  if constexpr (MODE == ROW_WHOLE || MODE == ROW_LAST) Jv = __shfl_sync(FULL_MASK, x, 1, SEG);
  if constexpr (MODE == ROW_WHOLE) B = DCP_MIN3(B, e + EB, Jv + JB); // viterbi.c:495-496,582-583
  if constexpr (MODE == ROW_LAST)
  {
    float const Btrue = DCP_MIN3(B, e + EB, Jv + JB);
    ok = ok && (Btrue == B);
  }

  // row l+1, t = 2..5: rows l-1, l-2, l-3, l-4 are ring slots s1..s4 of THIS row
#pragma unroll
  for (int q = 0; q < Q; ++q)
  {
    pt.Mp[q] = DCP_MIN3(s.P[s1][q] + e2[q], s.P[s2][q] + e3[q], s.P[s3][q] + e4[q]);
    pt.M5[q] = s.P[s4][q] + e5[q];
    pt.Ip[q] = DCP_MIN3(s.Qv[s1][q] + nb2.y, s.Qv[s2][q] + nb3.y, s.Qv[s3][q] + nb4.y);
  }
  pt.xp = DCP_MIN3(s.px[s1] + nb2.x, s.px[s2] + nb3.x, s.px[s3] + nb4.x);

  // second sweep unconditionally (one extra sweep per row on average), then lazily while any lane
  // still improves: only the incoming chain is carried (see the header)
  float dprev;
  {
    float din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
    if (head) din = b.z;
    float c = din + s.DD[0];
    D[0] = DCP_MIN2(D[0], c);
#pragma unroll
    for (int q = 1; q < Q; ++q)
    {
      c = c + s.DD[q];
      D[q] = DCP_MIN2(D[q], c);
    }
    for (;;)
    {
      din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1, SEG);
      if (head) din = b.z;
      c = din + s.DD[0];
      if (!__any_sync(FULL_MASK, DCP_LT(c, D[0]))) break;
      D[0] = DCP_MIN2(D[0], c);
#pragma unroll
      for (int q = 1; q < Q; ++q)
      {
        c = c + s.DD[q];
        D[q] = DCP_MIN2(D[q], c);
      }
    }
    dprev = din;
  }

  if constexpr (COL_OUT)
  { // (a lane group that only shadows another pair must not touch that pair's column: a later full
    // segment rewrites it in place while its owner is still reading ahead)
    if (own && sl == SEG - 1) __stcg(reinterpret_cast<float4 *>(slot), make_float4(M[Q - 1], I[Q - 1], D[Q - 1], e));
  }

  if constexpr (DUMP)
  {
    if (in_window)
    { // the row's final values, for the walk / argmin kernels (lane-chunked order, layout.cuh)
      size_t const at = (size_t)(l - 1) * (SEG * Q);
      store_chunks<Q, SEG>(dv.M + at, sl, M);
      store_chunks<Q, SEG>(dv.I + at, sl, I);
      store_chunks<Q, SEG>(dv.D + at, sl, D);
      float *xr = dv.xs + (size_t)(l - 1) * 8;
      if (sl == 0)
      {
        xr[0] = N;
        xr[1] = B;
        xr[2] = Jv;
        xr[3] = e;
      }
      if (sl == 2) xr[4] = x; // C(l)
    }
  }

  // P(l), Q(l) into the slot that held row l-5
  if constexpr (SPEC_B && MODE != ROW_LAST) s.px[J] = x + s.xb;
  else s.px[J] = DCP_MIN2(e + s.xa, x + s.xb);
  {
    float const t0 = DCP_MIN3(mprev + s.MM[0], iprev + s.IM[0], dprev + s.DM[0]);
    float t[Q];
    t[0] = t0;
#pragma unroll
    for (int q = 1; q < Q; ++q)
      t[q] = DCP_MIN3(M[q - 1] + s.MM[q], I[q - 1] + s.IM[q], D[q - 1] + s.DM[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      s.P[J][q] = DCP_MIN2(t[q], B + s.BM[q]);
      s.Qv[J][q] = DCP_MIN2(I[q] + s.II[q], M[q] + s.MI[q]);
    }
  }
}

// Resident CTAs per SM asked of ptxas.  Q = 6 whole-warp: 168 registers with 52 bytes of spills and
// 12 warps/SM beat 209 registers and 8 warps/SM (K = 192: 516 -> 558 GCUPS); the same trade loses
// for Q = 7, 8 (hundreds of spill bytes) and for every sub-warp shape (r2_qp_mb.log).
template <int Q, int SEG, int MODE, bool DUMP>
constexpr int row_min_blocks()
{
  if (Q == 6 && SEG == 32 && MODE == ROW_WHOLE && !DUMP) return 3;
  return Q >= 6 ? 2 : Q >= 4 ? 3 : Q == 3 ? 4 : Q == 2 ? 5 : 6;
}

// STAGE (grid mode only): profile-stationary CTAs.  A CTA claims 4 x (32 / SEG) reads of ONE profile at a
// time (one per lane group), and when the profile changes one elected thread stages its 84 short-code
// emission rows -- contiguous at the head of the table -- and its {null, background} table into shared
// memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); rows then read them with LDS and
// 32-bit addresses.
template <int Q, int SEG, int MODE, bool DUMP = false, bool STAGE = false>
__global__ void __launch_bounds__(32 * ROW_WARPS, row_min_blocks<Q, SEG, MODE, DUMP>()) score_row_kernel(StripArgs a)
{
  constexpr int G = 32 / SEG;
  static_assert(!DUMP || MODE == ROW_WHOLE, "the value dump runs on whole profiles");
  int const lane = threadIdx.x & 31;
  int const seg = lane / SEG, sl = lane % SEG;
  constexpr uint32_t STAGE_BYTES = (uint32_t)STAGE_ROWS * EmRows<Q, SEG>::ROWB;
  constexpr uint32_t STAGE_NB = (uint32_t)stage_nulbg_codes<Q>() * 8u;
  [[maybe_unused]] uint32_t stage_base = 0, stage_bar = 0, stage_phase = 0;
  [[maybe_unused]] int staged_profile = -1;
  __shared__ unsigned long long s_item;
  if constexpr (STAGE)
  {
    extern __shared__ __align__(128) unsigned char stage_mem[];
    stage_base = tma::smem_u32(stage_mem);
    stage_bar = stage_base + STAGE_BYTES + STAGE_NB;
    if (threadIdx.x == 0)
    {
      tma::mbar_init(stage_bar, 1);
      tma::fence_barrier_init();
    }
    __syncthreads();
  }

  for (;;)
  {
    unsigned long long item;
    bool active;
    if constexpr (STAGE)
    {
      __syncthreads(); // every warp is done with the previous claim: s_item and the staged rows may change
      if (threadIdx.x == 0) s_item = atomicAdd(a.s.counter, 1ULL);
      __syncthreads();
      // (the broadcast tells the compiler the claim is warp-uniform: no divergence guards around the row's shuffles)
      unsigned long long const claim = __shfl_sync(FULL_MASK, s_item, 0);
      constexpr unsigned PER_CLAIM = ROW_WARPS * G; // reads of the profile a CTA takes at a time
      unsigned const quads = ((unsigned)a.s.nseq + PER_CLAIM - 1) / PER_CLAIM;
      if (claim >= (a.s.nitems / (unsigned)a.s.nseq) * quads) break;
      unsigned const cpi = (unsigned)(claim / quads), cq = (unsigned)(claim - (unsigned long long)cpi * quads);
      // (the warp's index through a broadcast: the compiler then knows the pair, hence the row loop's
      // trip count, is warp-uniform and puts no divergence guards around the row's shuffles)
      int const csi = (int)(cq * PER_CLAIM) + __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0) * G + seg;
      active = csi < a.s.nseq;
      item = (unsigned long long)cpi * (unsigned)a.s.nseq + (active ? csi : 0);
      int const cp = a.s.class_profiles[cpi];
      if (cp != staged_profile)
      {
        ProfileDesc const cd = MODE == ROW_WHOLE ? a.s.profiles[cp] : a.segs[a.seg_first[cp] + a.level];
        if (threadIdx.x == 0)
        {
          tma::mbar_arrive_expect_tx(stage_bar, STAGE_BYTES + STAGE_NB);
          tma::bulk_g2s(stage_base, cd.em, STAGE_BYTES, stage_bar);
          tma::bulk_g2s(stage_base + STAGE_BYTES, cd.nulbg, STAGE_NB, stage_bar);
        }
        if (threadIdx.x == 0) // one waiter (a structured loop: no opaque branches inside inline asm), then a CTA
          while (!mbar_try_wait(stage_bar, stage_phase)) {} // barrier: the compiler sees converged warps below
        __syncthreads();
        stage_phase ^= 1u;
        staged_profile = cp;
      }
      // (a warp without a read of its own -- nseq not a multiple of four -- shadows read 0 of the profile
      // and writes no result: no divergent control flow around the row's shuffles)
    }
    else
    {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.s.counter, (unsigned long long)G);
      base = __shfl_sync(FULL_MASK, base, 0);
      if (base >= a.s.nitems) break;
      active = base + seg < a.s.nitems;
      item = active ? base + seg : base; // idle segments shadow segment 0
    }

    int p, sq, start, L;
    long long oidx;
    size_t colidx = 0;
    if (a.s.pairs)
    {
      oidx = a.s.order[item];
      Pair const pr = a.s.pairs[oidx];
      p = pr.profile;
      sq = pr.seq;
      start = pr.start;
      L = pr.len;
      if constexpr (MODE == ROW_WHOLE)
      {
        if (a.s.out_index) oidx = a.s.out_index[item];
      }
      else
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
      if constexpr (MODE != ROW_WHOLE) colidx = (size_t)a.colmap[pi] * (unsigned)a.s.nseq + si;
    }
    ProfileDesc const pd = MODE == ROW_WHOLE ? a.s.profiles[p] : a.segs[a.seg_first[p] + a.level];
    if (L < 0) L = min(min(pd.Kfull * 50, 100000), a.s.reads.seq_len[sq]); // first window of window.c:13-37
    int const Lmax = SEG == 32 ? L : __reduce_max_sync(FULL_MASK, L);
    float const *xt = a.s.xt + (size_t)L * X_STRIDE;
    Mail *const col = MODE == ROW_WHOLE ? nullptr : a.col + colidx * a.col_stride;
    DumpRef<DUMP> const dv(DUMP ? a.s.dump + a.s.dump_off[item] : nullptr, L, SEG * Q);

    Lane<Q> s;
    constexpr int Kpad = SEG * Q;
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

    Partial<Q> pt;
#pragma unroll
    for (int q = 0; q < Q; ++q)
    {
      pt.Mp[q] = CUDART_INF_F;
      pt.M5[q] = CUDART_INF_F;
      pt.Ip[q] = CUDART_INF_F;
    }
    pt.xp = CUDART_INF_F;

    // history of row l = hist[first + l - 1]; rows before the window (t > l) meet +INF states only
    uint16_t const *hp = a.s.reads.hist + (a.s.reads.seq_word[sq] * 16 + start);
    uint32_t h = __ldg(hp);
    EmRows<Q, SEG, STAGE> const em(pd.em, sl, (uint32_t)pd.Kpad * 4u, stage_base, stage_base + STAGE_BYTES);
    uint32_t const eight = (uint32_t)a.s.reads.eight;
    float E = CUDART_INF_F, x = CUDART_INF_F, Eres = CUDART_INF_F, xres = CUDART_INF_F;
    bool ok = true, okres = true;
    float4 bnext = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F), bnext2 = bnext;
    if constexpr (MODE == ROW_MID || MODE == ROW_LAST)
    {
      bnext = __ldcg(reinterpret_cast<float4 const *>(col + 1));
      bnext2 = __ldcg(reinterpret_cast<float4 const *>(col + 2));
    }
#define DCP_ROW(JJ_, I_)                                                                         \
  {                                                                                              \
    if (l > Lmax) break;                                                                         \
    uint32_t const hn = __ldg(hp + (I_) + 1);                                                    \
    row_v2<Q, SEG, MODE, DUMP, JJ_>(s, pt, em, pd.nulbg, eight, h, hn, sl, NB, EB, JB, bnext, bnext2, col + l, E, x, ok, dv, l,  \
                                    l <= L, active);                                             \
    if (SEG != 32 && l == L)                                                                     \
    {                                                                                            \
      Eres = E;                                                                                  \
      xres = x;                                                                                  \
      okres = ok;                                                                                \
    }                                                                                            \
    h = hn;                                                                                      \
    ++l;                                                                                         \
  }
    int l = 1;
    for (;;)
    {
      DCP_ROW(1, 0)
      DCP_ROW(2, 1)
      DCP_ROW(3, 2)
      DCP_ROW(4, 3)
      DCP_ROW(0, 4)
      hp += 5;
    }
#undef DCP_ROW
    if constexpr (SEG == 32)
    {
      Eres = E;
      xres = x;
      okres = ok;
    }

    if constexpr (MODE == ROW_WHOLE || MODE == ROW_LAST)
    {
      float const C = __shfl_sync(FULL_MASK, xres, 2, SEG);
      float const R = __shfl_sync(FULL_MASK, xres, 3, SEG);
      if (sl == 0 && active)
      {
        float const alt = fminf(Eres + ET, C + CT); // viterbi.c:585-586, 599
        if (MODE == ROW_WHOLE || okres)
        {
          a.s.out[oidx] = make_float2(R, alt); // null cost: viterbi.c:718
          float const d = alt - R;             // lrt = -2*((-nul) - (-alt)) >= 0  <=>  alt - nul <= 0
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
