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
// This file is the round-1 DP row and the exact kernel with W WARPS PER PAIR.  Since round 2 the
// score pass runs the single-warp row of row_kernel.cuh for every profile of up to 256 nodes and
// for the 256-node segments of larger ones (speculative B); profiles of 257..2048 nodes come
// here -- W = 2/4/8 -- only when that speculation fails, and for the trace pass's value dump
// (DUMP).  The shared types (Lane, Mail, ScoreArgs, DumpView, the TMA helpers) also live here.
// The K nodes are
// striped across the VL = 32*W "virtual lanes" like the reference stripes them across SIMD
// lanes (viterbi.c:220-221): vl = k / Q, q = k % Q, Q <= 8.  Per lane everything lives in
// registers:
//   * the 8 transition costs of its Q nodes,
//   * a 5-row ring of P and Q (the reference's 6-slot time frame, viterbi.c:12,160-161;
//     the row being computed needs no slot of its own here), rotated by unrolling the row
//     loop 5x so that ring indices are compile-time constants,
// and the cross-lane k-1 dependency is one __shfl_up per state per row (the reference's
// shift(), intrinsics.h:95-106).  The serial delete chain is resolved like the reference's
// lazy sweeps (viterbi.c:561-580): one in-lane sweep, then boundary propagation repeated
// while any lane still improves (warp vote) -- every candidate is a left-to-right chain
// sum, so the fixed point is bit-identical to the serial recurrence.  With W > 1 the warp
// boundary values (M, I, D of a warp's last node and its partial E) cross through a
// double-buffered shared-memory mailbox, two CTA barriers per row.
// The special states are spread over lanes 0..3 (N, J, C and the null model's R), which
// all run the same "min_t prev[t] + null[code_t]" recurrence.
//
// Emission tables: the rows of the three short code classes (1-, 2-, 3-mers: 84 rows, L1
// resident) are read with coalesced 128-bit loads.  With W > 1 the rows of the 4- and 5-mers
// (1280 rows, L2/HBM resident) are staged into shared memory by TMA (cp.async.bulk + mbarrier
// complete_tx), five DP rows ahead, through a 5-stage ring whose stage index is the row's
// compile-time ring slot -- so the L2 latency never sits on the row's critical path.  With
// W = 1 (1 KB rows per warp) the same ring measured slower than plain loads and is off.
//
// Instruction selection (measured on B200 with dcpgpu_alu_peak): FADD and the
// two-input FMNMX both issue at ~1 warp-instruction/clk/SMSP, the three-input FMNMX3 at 1/2
// but it replaces two mins and frees an issue slot for the other pipe, so every pair of mins
// is written as one min3 (min.f32 d,a,b,c); E is reduced with one CREDUX.MIN on the bit
// patterns (all costs are >= +0, checked when a profile is uploaded).
#pragma once
#include "layout.cuh"
#include <math_constants.h>

namespace dcp {

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
  long long const *out_index;  // optional: where item's result goes (default: its pair index)
  unsigned long long nitems;
  unsigned long long *counter; // work-stealing cursor
  float2 *out;                 // {null cost, alt cost} per pair
  unsigned long long *nhits;
  // DUMP variant (trace pass): every row's M, I, D and special-state values go to global memory
  float *dump;                 // pool
  long long const *dump_off;   // per item of this launch: floats offset of its block (see DumpView)
};

// Layout of one traced pair's value dump: rows l = 1..L (row 0 is all +INF except B = SB).
//   M, I, D : [L][Kpad] each, node k at column layout_pos(k) (the lane-chunked order of layout.cuh)
//   xs      : [L][8] = N, B, J, E, C
// Dumped DP values of one pair: M, I, D as [L][Kpad] with node k of a row at layout_pos(k)
// (the order the lanes hold them, so a row is written with full-width stores), then
// N, B, J, E, C of every row as [L][8].
struct DumpView
{
  float *M, *I, *D, *xs;
  __host__ __device__ static size_t floats(int L, int Kpad) { return (size_t)L * (3 * (size_t)Kpad + 8); }
  __host__ __device__ DumpView(float *base, int L, int Kpad)
  {
    M = base;
    I = M + (size_t)L * Kpad;
    D = I + (size_t)L * Kpad;
    xs = D + (size_t)L * Kpad;
  }
};

struct __align__(16) Mail
{
  float M, I, D, E;
};

__device__ __forceinline__ float min3(float a, float b, float c)
{
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

namespace tma {
__device__ __forceinline__ uint32_t smem_u32(void const *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, void const *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile("{\n\t"
               ".reg .pred P1;\n\t"
               "LAB_WAIT:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra DONE;\n\t"
               "bra LAB_WAIT;\n\t"
               "DONE:\n\t"
               "}" ::"r"(bar),
               "r"(parity)
               : "memory");
}
} // namespace tma

// Row of one code inside the emission table: Q values of virtual lane vl (see layout.cuh).
template <int Q, int VL>
__device__ __forceinline__ void load_chunks(float (&e)[Q], float const *__restrict__ row, int vl)
{
  constexpr int N4 = Q / 4;
#pragma unroll
  for (int c = 0; c < N4; ++c)
  {
    float4 v = __ldg(reinterpret_cast<float4 const *>(row + c * 4 * VL) + vl);
    e[4 * c + 0] = v.x;
    e[4 * c + 1] = v.y;
    e[4 * c + 2] = v.z;
    e[4 * c + 3] = v.w;
  }
  if constexpr ((Q & 2) != 0)
  {
    float2 v = __ldg(reinterpret_cast<float2 const *>(row + VL * (N4 * 4)) + vl);
    e[N4 * 4 + 0] = v.x;
    e[N4 * 4 + 1] = v.y;
  }
  if constexpr ((Q & 1) != 0) e[Q - 1] = __ldg(row + VL * (Q - 1) + vl);
}

// The same layout, written: Q values of virtual lane vl into one row (full-width coalesced stores).
template <int Q, int VL>
__device__ __forceinline__ void store_chunks(float *row, int vl, float const (&v)[Q])
{
  constexpr int N4 = Q / 4;
#pragma unroll
  for (int c = 0; c < N4; ++c)
    __stcs(reinterpret_cast<float4 *>(row + c * 4 * VL) + vl, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
  if constexpr ((Q & 2) != 0)
    __stcs(reinterpret_cast<float2 *>(row + VL * (N4 * 4)) + vl, make_float2(v[N4 * 4], v[N4 * 4 + 1]));
  if constexpr ((Q & 1) != 0) __stcs(row + VL * (Q - 1) + vl, v[Q - 1]);
}

// Per-lane base pointers into a profile's emission table: a code row is then reached with ONE
// 32-bit byte offset (code * 4*Kpad < 2^32) added to a 64-bit base, instead of a 64-bit
// multiply-add per access.
template <int Q, int VL>
struct RowBase
{
  char const *b4; // first float4 chunk of this lane
  char const *b2; // the float2 chunk (Q & 2)
  char const *b1; // the scalar chunk (Q & 1)
  __device__ __forceinline__ RowBase(float const *em, int vl)
  {
    constexpr int N4 = Q / 4;
    b4 = reinterpret_cast<char const *>(em) + (size_t)vl * 16;
    b2 = reinterpret_cast<char const *>(em + VL * (N4 * 4)) + (size_t)vl * 8;
    b1 = reinterpret_cast<char const *>(em + VL * (Q - 1)) + (size_t)vl * 4;
  }
  __device__ __forceinline__ void load(float (&e)[Q], uint32_t byte_off) const
  {
    constexpr int N4 = Q / 4;
#pragma unroll
    for (int c = 0; c < N4; ++c)
    {
      float4 v = __ldg(reinterpret_cast<float4 const *>(b4 + byte_off + c * 16 * VL));
      e[4 * c + 0] = v.x;
      e[4 * c + 1] = v.y;
      e[4 * c + 2] = v.z;
      e[4 * c + 3] = v.w;
    }
    if constexpr ((Q & 2) != 0)
    {
      float2 v = __ldg(reinterpret_cast<float2 const *>(b2 + byte_off));
      e[N4 * 4 + 0] = v.x;
      e[N4 * 4 + 1] = v.y;
    }
    if constexpr ((Q & 1) != 0) e[Q - 1] = __ldg(reinterpret_cast<float const *>(b1 + byte_off));
  }
};

__device__ __forceinline__ float2 ldg_nulbg(float2 const *nulbg, int code)
{
  return __ldg(reinterpret_cast<float2 const *>(reinterpret_cast<char const *>(nulbg) + (uint32_t)code * 8u));
}

// Same row layout, read from a TMA-filled shared-memory stage (conflict-free LDS.128/64/32).
template <int Q, int VL>
__device__ __forceinline__ void load_chunks_smem(float (&e)[Q], float const *row, int vl)
{
  constexpr int N4 = Q / 4;
#pragma unroll
  for (int c = 0; c < N4; ++c)
  {
    float4 v = *(reinterpret_cast<float4 const *>(row + c * 4 * VL) + vl);
    e[4 * c + 0] = v.x;
    e[4 * c + 1] = v.y;
    e[4 * c + 2] = v.z;
    e[4 * c + 3] = v.w;
  }
  if constexpr ((Q & 2) != 0)
  {
    float2 v = *(reinterpret_cast<float2 const *>(row + VL * (N4 * 4)) + vl);
    e[N4 * 4 + 0] = v.x;
    e[N4 * 4 + 1] = v.y;
  }
  if constexpr ((Q & 1) != 0) e[Q - 1] = row[VL * (Q - 1) + vl];
}

// The TMA ring of one pair: 5 stages (ring slot J) x 2 rows (4-mer, 5-mer) x ROW floats.
// What the row code sees of the dump: nothing at all unless DUMP.
template <bool DUMP>
struct DumpRef
{
  __device__ __forceinline__ DumpRef(float *, int, int) {}
};
template <>
struct DumpRef<true> : DumpView
{
  __device__ __forceinline__ DumpRef(float *base, int L, int Kpad) : DumpView(base, L, Kpad) {}
};

template <int Q, int W>
struct Ring
{
  static constexpr int ROW = 32 * W * Q;          // floats per code row
  static constexpr uint32_t ROW_BYTES = ROW * 4u; // multiple of 128
  float *stage;                                   // [5][2][ROW]
  uint32_t stage_addr;                            // shared-space address of stage
  uint32_t bar_addr;                              // shared-space address of full[5]
  // issue the copies of the row whose last five nucleotides are `h` into ring slot j
  __device__ __forceinline__ void fill(ProfileDesc const &pd, int j, unsigned h) const
  {
    uint32_t const bar = bar_addr + 8u * (uint32_t)j;
    uint32_t const dst = stage_addr + (uint32_t)j * 2u * ROW_BYTES;
    tma::mbar_arrive_expect_tx(bar, 2u * ROW_BYTES);
    tma::bulk_g2s(dst, pd.em + (size_t)(84 + (h & 255u)) * ROW, ROW_BYTES, bar);
    tma::bulk_g2s(dst + ROW_BYTES, pd.em + (size_t)(340 + (h & 1023u)) * ROW, ROW_BYTES, bar);
  }
};

// E partial of a warp: min over its lanes.  Values are >= +0 (or +INF), so the unsigned
// order of the bit patterns equals the float order and one REDUX.MIN does the reduction.
__device__ __forceinline__ float warp_min_nonneg(float v)
{
  return __uint_as_float(__reduce_min_sync(FULL_MASK, __float_as_uint(v)));
}

template <int W>
struct ScoreCfg
{
  static constexpr int GROUPS = W == 1 ? 4 : 1; // independent pairs per CTA
  static constexpr int THREADS = 32 * W * GROUPS;
  // CTAs per SM the register allocation is told to allow: stating the occupancy each class
  // reaches anyway lets ptxas spend the registers up to that step on a better schedule
  template <int Q>
  static constexpr int min_blocks()
  {
    return W != 1 ? 1 : Q >= 6 ? 2 : Q >= 4 ? 3 : Q == 3 ? 4 : Q == 2 ? 5 : 6;
  }
  // Long-code rows through the TMA ring?  Measured on B200 (profiles/): CTA-wide bulk copies
  // (W > 1, 2..8 KB each) beat per-lane loads, 1 KB per-warp copies (W = 1) do not.
  static constexpr bool TMA_RING = W > 1;
};

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

template <int Q>
__device__ __forceinline__ void d_sweep(Lane<Q> const &s, float (&D)[Q])
{
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = fminf(D[q], D[q - 1] + s.DD[q]);
}

// Lazy boundary propagation of the delete chain inside a warp (viterbi.c:569-580).
// `head`: this lane's predecessor node is not in this warp (its incoming D is handled by
// the caller).  Returns the final D of node k-1 as seen by q = 0 of every lane.
template <int Q>
__device__ __forceinline__ float d_lazy(Lane<Q> const &s, float (&D)[Q], bool head)
{
  float din;
  for (;;)
  {
    din = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din = CUDART_INF_F;
    float const c = din + s.DD[0];
    if (!__any_sync(FULL_MASK, c < D[0])) break;
    D[0] = fminf(D[0], c);
    d_sweep<Q>(s, D);
  }
  return din;
}

// min over this lane's nodes of min(M, D): two interleaved chains instead of one (the value
// feeds E -> B -> P of the next row, the longest dependency of the row's tail)
template <int Q>
__device__ __forceinline__ float e_lane(float const (&M)[Q], float const (&D)[Q])
{
  float a = fminf(M[0], D[0]);
  if constexpr (Q == 1) return a;
  float b = fminf(M[1], D[1]);
#pragma unroll
  for (int q = 2; q + 1 < Q; q += 2)
  {
    a = min3(a, M[q], D[q]);
    b = min3(b, M[q + 1], D[q + 1]);
  }
  if constexpr ((Q & 1) != 0 && Q > 1) a = min3(a, M[Q - 1], D[Q - 1]);
  return fminf(a, b);
}

template <int Q>
__device__ __forceinline__ float e_partial(float const (&M)[Q], float const (&D)[Q])
{
  float const e = e_lane<Q>(M, D);
  return warp_min_nonneg(e);
}

// One software-pipelined DP row l (J = l % 5).
//
// Row l+1's accumulation over the emission lengths t = 2..5 reads only rows l-1..l-4, so it does
// not depend on anything row l computes.  It is issued in the shadow of row l's serial tail (the
// delete-chain sweeps, whose dependent FADD/FMNMX chain would otherwise leave the issue slots
// idle); row l itself then only has to add the one-nucleotide term (which needs P(l-1)) to the
// partial accumulators Mp/Ip/xp it inherits before its own delete chain starts.
//   hist  = the five nucleotides ending at l-1 (codes of row l), hist1 = ending at l (row l+1),
//   hist6 = ending at l+5 (row l+6: its long-code rows are requested from the TMA ring now).
template <int Q, int W, int J, bool DUMP>
__device__ __forceinline__ void dp_row(Lane<Q> &s, float (&Mp)[Q], float (&Ip)[Q], float &xp, ProfileDesc const &pd,
                                       Ring<Q, W> const &ring, unsigned hist, unsigned hist1, unsigned hist6,
                                       bool have_next, bool refill, unsigned &phase, int lane, int warp, float NB,
                                       float EB, float JB, Mail *mail, int volatile *flags, int par, float &E,
                                       float &x, DumpRef<DUMP> const &dv, int l)
{
  constexpr int VL = 32 * W;
  constexpr int s1 = (J + 4) % 5, s2 = (J + 3) % 5, s3 = (J + 2) % 5, s4 = (J + 1) % 5;
  constexpr int JN = (J + 1) % 5; // ring slot of DP row l+1
  int const vl = warp * 32 + lane;
  uint32_t const rowb = (uint32_t)pd.Kpad * 4u; // bytes per code row
  RowBase<Q, VL> const rb(pd.em, vl);

  // (A) finish row l: the t = 1 term needs P(l-1), Q(l-1)
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
  if constexpr (ScoreCfg<W>::TMA_RING)
  { // the 4- and 5-mer rows of DP row l+1 were requested by TMA five rows ago
    if (have_next)
    {
      tma::mbar_wait(ring.bar_addr + 8u * JN, (phase >> JN) & 1u);
      phase ^= 1u << JN;
    }
    load_chunks_smem<Q, VL>(e4, ring.stage + (JN * 2 + 0) * Ring<Q, W>::ROW, vl);
    load_chunks_smem<Q, VL>(e5, ring.stage + (JN * 2 + 1) * Ring<Q, W>::ROW, vl);
  }
  else
  {
    rb.load(e4, (uint32_t)c4 * rowb);
    rb.load(e5, (uint32_t)c5 * rowb);
  }

  // Delete chain of row l: source terms + first sweep (viterbi.c:538, 552-567).  Lane 0 of the
  // first warp is node 0, whose incoming transitions are +INF (protein.c:366-370), so the
  // wrapped shuffle value is inert there; lane 0 of a later warp gets its predecessor through
  // the mailbox below.
  bool const head = W > 1 && lane == 0 && warp > 0;
  float mprev = __shfl_up_sync(FULL_MASK, M[Q - 1], 1);
  float iprev = __shfl_up_sync(FULL_MASK, I[Q - 1], 1);
  if (head) mprev = CUDART_INF_F;
  float D[Q];
  D[0] = mprev + s.MD[0];
#pragma unroll
  for (int q = 1; q < Q; ++q)
    D[q] = M[q - 1] + s.MD[q];
  {
    float din0 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din0 = CUDART_INF_F;
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

  // second sweep unconditionally (measured: one extra sweep per row on average), then lazy
  {
    float din1 = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) din1 = CUDART_INF_F;
    D[0] = fminf(D[0], din1 + s.DD[0]);
    d_sweep<Q>(s, D);
  }
  float dprev = d_lazy<Q>(s, D, head);

  // E(l) = min_k min(M_k, D_k)  (viterbi.c:540-558)
  float e = e_partial<Q>(M, D);
  if constexpr (W == 1) { E = e; }
  else
  {
    Mail *box = mail + par * W;
    if (lane == 31) box[warp] = Mail{M[Q - 1], I[Q - 1], D[Q - 1], e};
    __syncthreads();
    // every warp has consumed ring slot JN (row l+1): request row l+6 into it
    if (ScoreCfg<W>::TMA_RING && refill && threadIdx.x == 0) ring.fill(pd, JN, hist6);
    if (warp > 0)
    {
      Mail const pm = box[warp - 1];
      float c = CUDART_INF_F;
      if (lane == 0)
      {
        mprev = pm.M;
        iprev = pm.I;
        c = fminf(pm.M + s.MD[0], pm.D + s.DD[0]);
      }
      if (__any_sync(FULL_MASK, c < D[0]))
      {
        float const dlast = D[Q - 1];
        D[0] = fminf(D[0], c);
        d_sweep<Q>(s, D);
        d_lazy<Q>(s, D, head);
        float const e2n = e_partial<Q>(M, D);
        if (lane == 31 && (D[Q - 1] != dlast || e2n != e))
        {
          box[warp].D = D[Q - 1];
          box[warp].E = e2n;
          flags[par] = 1;
        }
        e = e2n;
      }
    }
    for (;;)
    { // confirm: almost always one barrier; a warp's last D changing after the exchange is rare
      __syncthreads();
      if (flags[par] == 0) break;
      __syncthreads();
      if (threadIdx.x == 0) flags[par] = 0;
      __syncthreads();
      if (warp > 0)
      {
        Mail const pm = box[warp - 1];
        float const c = lane == 0 ? pm.D + s.DD[0] : CUDART_INF_F;
        if (__any_sync(FULL_MASK, c < D[0]))
        {
          float const dlast = D[Q - 1];
          D[0] = fminf(D[0], c);
          d_sweep<Q>(s, D);
          d_lazy<Q>(s, D, head);
          float const e2n = e_partial<Q>(M, D);
          if (lane == 31 && (D[Q - 1] != dlast || e2n != e))
          {
            box[warp].D = D[Q - 1];
            box[warp].E = e2n;
            flags[par] = 1;
          }
          e = e2n;
        }
      }
    }
    dprev = __shfl_up_sync(FULL_MASK, D[Q - 1], 1);
    if (head) dprev = box[warp - 1].D;
    float ee = box[0].E;
#pragma unroll
    for (int w = 1; w < W; ++w)
      ee = fminf(ee, box[w].E);
    E = ee;
  }

  // special states: x is N(l) on lane 0, J(l) on lane 1, C(l) on lane 2, R(l) on lane 3
  x = xacc;
  float const N = __shfl_sync(FULL_MASK, x, 0);
  float const Jv = __shfl_sync(FULL_MASK, x, 1);
  float const B = min3(N + NB, E + EB, Jv + JB); // viterbi.c:495-496,582-583
  s.px[J] = fminf(E + s.xa, x + s.xb);

  if constexpr (DUMP)
  { // trace pass: the row's final values, for the argmin kernels
    size_t const at = (size_t)(l - 1) * pd.Kpad; // rows in the emission tables' lane-chunked order
    store_chunks<Q, VL>(dv.M + at, vl, M);
    store_chunks<Q, VL>(dv.I + at, vl, I);
    store_chunks<Q, VL>(dv.D + at, vl, D);
    if (warp == 0)
    {
      float *xr = dv.xs + (size_t)(l - 1) * 8;
      if (lane == 0)
      {
        xr[0] = N;
        xr[1] = B;
        xr[2] = Jv;
        xr[3] = E;
      }
      if (lane == 2) xr[4] = x; // C(l)
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

template <int Q, int W, bool DUMP>
__device__ __forceinline__ void score_one(ProfileDesc const &pd, Ring<Q, W> const &ring, unsigned &phase,
                                          uint32_t const *__restrict__ words, int start, int L,
                                          float const *__restrict__ xt, int lane, int warp, Mail *mail,
                                          int volatile *flags, float &null_cost, float &alt_cost,
                                          DumpRef<DUMP> const &dv)
{
  constexpr int VL = 32 * W;
  Lane<Q> s;
  int const Kpad = pd.Kpad;
  int const vl = warp * 32 + lane;
  load_chunks<Q, VL>(s.BM, pd.core + C_BM * Kpad, vl);
  load_chunks<Q, VL>(s.MM, pd.core + C_MM * Kpad, vl);
  load_chunks<Q, VL>(s.MI, pd.core + C_MI * Kpad, vl);
  load_chunks<Q, VL>(s.MD, pd.core + C_MD * Kpad, vl);
  load_chunks<Q, VL>(s.IM, pd.core + C_IM * Kpad, vl);
  load_chunks<Q, VL>(s.II, pd.core + C_II * Kpad, vl);
  load_chunks<Q, VL>(s.DM, pd.core + C_DM * Kpad, vl);
  load_chunks<Q, VL>(s.DD, pd.core + C_DD * Kpad, vl);

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

  // Nucleotide stream, running six positions ahead of the DP row: H holds the last eleven
  // nucleotides; at row l bits 12..21 = the five ending at l-1 (codes of row l), bits 10..19 =
  // ending at l (row l+1), bits 0..9 = ending at l+5 (row l+6, requested from the TMA ring).
  int g = start;
  uint32_t const *wp = words + (g >> 4);
  uint32_t word = __ldg(wp) >> (2 * (g & 15));
  int left = 16 - (g & 15);
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
  // prologue: long-code rows of DP rows 2..6 (ring slots 2,3,4,0,1); row r ends at nucleotide r-1.
  // Row 1 has no predecessors for t >= 2, so its slot is never read and never filled.
  if (ScoreCfg<W>::TMA_RING && threadIdx.x == 0)
  {
#pragma unroll
    for (int r = 2; r <= 6; ++r)
      if (r <= L) ring.fill(pd, r % 5, (H >> (2 * (6 - r))) & 1023u);
  }
  float E = CUDART_INF_F, x = CUDART_INF_F;

  // partial accumulators of the next row (t = 2..5); row 1 has no such predecessors
  float Mp[Q], Ip[Q], xp = CUDART_INF_F;
#pragma unroll
  for (int q = 0; q < Q; ++q)
  {
    Mp[q] = CUDART_INF_F;
    Ip[q] = CUDART_INF_F;
  }
#define DCP_ROW(JJ_)                                                                             \
  {                                                                                              \
    if (l > L) break;                                                                            \
    DCP_NEXT_NT()                                                                                \
    dp_row<Q, W, JJ_, DUMP>(s, Mp, Ip, xp, pd, ring, (H >> 12) & 1023u, (H >> 10) & 1023u, H & 1023u,      \
                            l + 1 <= L, l + 6 <= L, phase, lane, warp, NB, EB, JB, mail, flags, l & 1, E, x,  \
                            dv, l);                                                              \
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

  float const C = __shfl_sync(FULL_MASK, x, 2);
  float const R = __shfl_sync(FULL_MASK, x, 3);
  alt_cost = fminf(E + ET, C + CT); // viterbi.c:585-586, 599
  null_cost = R;                    // viterbi.c:718
}

template <int Q, int W>
constexpr size_t score_smem_bytes()
{
  return ScoreCfg<W>::TMA_RING ? (size_t)ScoreCfg<W>::GROUPS * (5 * 2 * Ring<Q, W>::ROW * sizeof(float) + 64) : 256;
}

template <int Q, int W, bool DUMP = false>
__global__ void __launch_bounds__(ScoreCfg<W>::THREADS, ScoreCfg<W>::template min_blocks<Q>()) score_reg_kernel(ScoreArgs a)
{
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ Mail mail[2 * W];
  __shared__ int flags[2];
  __shared__ unsigned long long next_item;
  int const lane = threadIdx.x & 31;
  int const group = W == 1 ? (int)(threadIdx.x >> 5) : 0; // independent pair slot inside the CTA
  int const warp = W == 1 ? 0 : (int)(threadIdx.x >> 5);  // warp index inside the pair

  // carve the TMA ring of this group: stages first (128-byte aligned), then the mbarriers
  constexpr size_t STAGE_BYTES = 5 * 2 * Ring<Q, W>::ROW * sizeof(float);
  Ring<Q, W> ring;
  ring.stage = reinterpret_cast<float *>(dyn_smem + (size_t)group * STAGE_BYTES);
  ring.stage_addr = tma::smem_u32(ring.stage);
  ring.bar_addr = tma::smem_u32(dyn_smem + (size_t)ScoreCfg<W>::GROUPS * STAGE_BYTES + (size_t)group * 64);
  if (ScoreCfg<W>::TMA_RING && (W == 1 ? lane : (int)threadIdx.x) == 0)
  {
#pragma unroll
    for (int j = 0; j < 5; ++j)
      tma::mbar_init(ring.bar_addr + 8u * j, 1);
    tma::fence_barrier_init();
  }
  unsigned phase = 0; // parity of the next completion of each ring slot's mbarrier
  if (W > 1)
  {
    if (threadIdx.x < 2) flags[threadIdx.x] = 0;
    __syncthreads();
  }
  else
    __syncwarp();
  for (;;)
  {
    unsigned long long item = 0;
    if constexpr (W == 1)
    {
      if (lane == 0) item = atomicAdd(a.counter, 1ULL);
      item = __shfl_sync(FULL_MASK, item, 0);
    }
    else
    {
      if (threadIdx.x == 0) next_item = atomicAdd(a.counter, 1ULL);
      __syncthreads();
      item = next_item;
      __syncthreads();
    }
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
      len = -1;
    }
    ProfileDesc const pd = a.profiles[p];
    if (len < 0)
    { // first window of window.c:13-37: [0, min(50K, 100000, |seq|))
      int const w = min(pd.K * 50, 100000);
      len = min(w, a.reads.seq_len[sq]);
    }
    float nul, alt;
    DumpRef<DUMP> const dv(DUMP ? a.dump + a.dump_off[item] : nullptr, len, pd.Kpad);
    score_one<Q, W, DUMP>(pd, ring, phase, a.reads.words + a.reads.seq_word[sq], start, len,
                          a.xt + (size_t)len * X_STRIDE, lane, warp, mail, flags, nul, alt, dv);
    if (lane == 0 && warp == 0)
    {
      a.out[oidx] = make_float2(nul, alt);
      float const d = alt - nul; // lrt = -2*((-nul) - (-alt)) >= 0  <=>  alt - nul <= 0
      if (d <= 0.0f && d > -CUDART_INF_F) atomicAdd(a.nhits, 1ULL);
    }
  }
}

} // namespace dcp
