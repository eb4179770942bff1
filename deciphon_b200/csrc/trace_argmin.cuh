// Trace pass, fast route: trellis words from dumped DP values.
//
// The reference's viterbi_path (c-core/viterbi.c:726-732) reruns the DP keeping, per state and
// row, the first-wins argmin of its candidates (viterbi.c:201-212, 485-586) and packs it into the
// trellis (viterbi.c:631-692, trellis.h:42-56).  An argmin depends only on the VALUES of the
// predecessor states, and score_reg_kernel<Q,W,DUMP=true> produces exactly those values (bit
// identical to the reference's, see score_kernel.cuh) at register speed.  So the trace pass is:
//   1. rerun the pairs that passed the LRT gate with the DUMP variant: M, I, D of every cell and
//      N, B, J, E, C of every row go to global memory (12 bytes per cell, streamed);
//   2. this kernel: every trellis word is computed INDEPENDENTLY from the dumped values with the
//      reference's own candidate arithmetic ((state + transition) + emission), candidate order
//      (emission length 5..1; BM,MM,IM,DM; II before MI; MD,DD; SN,NN; EJ,JJ; EC,CC; NB,EB,JB;
//      ET,CT) and strict-less updates -- no serial dependency left, fully parallel over cells;
//   3. trellis_walk / walk_write_kernel as before (trellis.c:147-167).
#pragma once
#include "generic_kernel.cuh"

namespace dcp {

constexpr int ARGMIN_ROWS = 8;     // rows per CTA = warps per CTA
constexpr int ARGMIN_THREADS = 32 * ARGMIN_ROWS;

struct ArgminArgs
{
  ProfileDesc const *profiles;
  ReadsView reads;
  float const *xt;
  Pair const *pairs;
  long long const *order;    // [nitems] pair indices handled by this launch
  long long const *tile_off; // [nitems+1] prefix sum of ceil(len / ARGMIN_ROWS)
  long long nitems;
  float const *dump;
  long long const *dump_off; // per item of this launch
  uint32_t *xnodes;
  uint16_t *nodes;
  long long const *xnode_off;
  long long const *node_off;
};

__global__ void __launch_bounds__(ARGMIN_THREADS, 2) trace_argmin_kernel(ArgminArgs a)
{
  __shared__ int codes[ARGMIN_ROWS][6];
  __shared__ float bgs[ARGMIN_ROWS][6];
  __shared__ float Bw[ARGMIN_ROWS + 5];
  float const INF = CUDART_INF_F;
  long long const tile = blockIdx.x;
  // which pair does this tile belong to?
  long long lo = 0, hi = a.nitems;
  while (hi - lo > 1)
  {
    long long const mid = (lo + hi) >> 1;
    if (a.tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  long long const oidx = a.order[lo];
  Pair const pr = a.pairs[oidx];
  ProfileDesc const pd = a.profiles[pr.profile];
  int const K = pd.K, Kpad = pd.Kpad, L = pr.len;
  int const r0 = 1 + (int)(tile - a.tile_off[lo]) * ARGMIN_ROWS; // first row of this tile
  int const nrows = min(ARGMIN_ROWS, L - r0 + 1);
  DumpView const dv(const_cast<float *>(a.dump) + a.dump_off[lo], L, Kpad);
  float const *xt = a.xt + (size_t)L * X_STRIDE;
  float const SN = xt[X_SN], NN = xt[X_NN], SB = xt[X_SB], NB = xt[X_NB], EB = xt[X_EB], JB = xt[X_JB],
              EJ = xt[X_EJ], JJ = xt[X_JJ], EC = xt[X_EC], CC = xt[X_CC], ET = xt[X_ET], CT = xt[X_CT];
  uint32_t *xnodes = a.xnodes + a.xnode_off[oidx];
  uint16_t *nodes = a.nodes + a.node_off[oidx];
  uint32_t const *words = a.reads.words + a.reads.seq_word[pr.seq];

  // codes of the five emission lengths for each row of the tile (thread r handles row r0 + r)
  if (threadIdx.x < nrows)
  {
    int const l = r0 + threadIdx.x;
    unsigned hist = 0;
    for (int i = max(0, l - 5); i < l; ++i)
    {
      int const g = pr.start + i;
      hist = (hist << 2) | ((words[g >> 4] >> (2 * (g & 15))) & 3u);
    }
    codes[threadIdx.x][1] = hist & 3;
    codes[threadIdx.x][2] = 4 + (hist & 15);
    codes[threadIdx.x][3] = 20 + (hist & 63);
    codes[threadIdx.x][4] = 84 + (hist & 255);
    codes[threadIdx.x][5] = 340 + (hist & 1023);
  }
  if (tile == a.tile_off[lo])
  { // stage 0 of this pair: all-zero fields (before(), viterbi.c:602-629)
    for (int k = threadIdx.x; k < K; k += ARGMIN_THREADS) nodes[k] = 0;
    if (threadIdx.x == 0) xnodes[0] = 0;
  }
  __syncthreads();

  int const LQ = pd.Q, LVL = pd.VL; // dumped rows are in layout_pos order
  auto valM = [&](int l, int k) { return l >= 1 ? dv.M[(size_t)(l - 1) * Kpad + layout_pos(k, LQ, LVL)] : INF; };
  auto valD = [&](int l, int k) { return l >= 1 ? dv.D[(size_t)(l - 1) * Kpad + layout_pos(k, LQ, LVL)] : INF; };
  auto valX = [&](int l, int j) { return dv.xs[(size_t)(l - 1) * 8 + j]; }; // l >= 1: N,B,J,E,C

  // per (row, t): background emission; per window row: B
  if (threadIdx.x < ARGMIN_ROWS * 5)
  {
    int const r = threadIdx.x / 5, t = 1 + threadIdx.x % 5;
    bgs[r][t] = r < nrows ? __ldg(&pd.nulbg[codes[r][t]]).y : INF;
  }
  if (threadIdx.x < ARGMIN_ROWS + 5)
  {
    int const lz = r0 - 5 + (int)threadIdx.x; // window row
    Bw[threadIdx.x] = lz < 0 ? INF : lz == 0 ? SB : lz <= L ? valX(lz, 1) : INF; // B(0) = SB, viterbi.c:473
  }
  __syncthreads();

  // ---- node words: a thread owns node k for all rows of the tile and slides a window of the
  //      dumped values of the five previous rows through registers (each value loaded once) ----
  for (int k = threadIdx.x; k < K; k += ARGMIN_THREADS)
  {
    int const pk = layout_pos(k, pd.Q, pd.VL);
    int const pk1 = k > 0 ? layout_pos(k - 1, pd.Q, pd.VL) : 0;
    float const bm = __ldg(pd.core + C_BM * Kpad + pk), mm = __ldg(pd.core + C_MM * Kpad + pk),
                mi = __ldg(pd.core + C_MI * Kpad + pk), md = __ldg(pd.core + C_MD * Kpad + pk),
                im = __ldg(pd.core + C_IM * Kpad + pk), ii = __ldg(pd.core + C_II * Kpad + pk),
                dm = __ldg(pd.core + C_DM * Kpad + pk), dd = __ldg(pd.core + C_DD * Kpad + pk);
    // window row j <-> DP row r0 - 5 + j;  values: M,I,D of node k-1 and M,I of node k
    float wM1[ARGMIN_ROWS + 5], wI1[ARGMIN_ROWS + 5], wD1[ARGMIN_ROWS + 5], wM[ARGMIN_ROWS + 5], wI[ARGMIN_ROWS + 5];
#pragma unroll
    for (int j = 0; j < ARGMIN_ROWS + 5; ++j)
    {
      int const lz = r0 - 5 + j;
      bool const ok = lz >= 1 && lz <= L;
      size_t const row = ok ? (size_t)(lz - 1) * Kpad : 0;
      wM[j] = ok ? dv.M[row + pk] : INF;
      wI[j] = ok ? dv.I[row + pk] : INF;
      wM1[j] = ok && k > 0 ? dv.M[row + pk1] : INF;
      wI1[j] = ok && k > 0 ? dv.I[row + pk1] : INF;
      wD1[j] = ok && k > 0 ? dv.D[row + pk1] : INF;
    }
#pragma unroll
    for (int r = 0; r < ARGMIN_ROWS; ++r)
    {
      if (r < nrows)
      {
        int const l = r0 + r;
        float M = INF, I = INF;
        int mp = 0, ip = 0;
#pragma unroll
        for (int t = 5; t >= 1; --t)
        { // rows before 0 hold +INF, so emission lengths t > l never win (viterbi.c:485)
          int const j = r + 5 - t; // window row of DP row l - t
          float const e = __ldg(pd.em + (size_t)codes[r][t] * Kpad + pk);
          float const b = bgs[r][t];
          DCP_UPD(M, (Bw[j] + bm) + e, mp, 0 + t - 1);
          DCP_UPD(M, (wM1[j] + mm) + e, mp, 5 + t - 1);
          DCP_UPD(M, (wI1[j] + im) + e, mp, 10 + t - 1);
          DCP_UPD(M, (wD1[j] + dm) + e, mp, 15 + t - 1);
          DCP_UPD(I, (wI[j] + ii) + b, ip, 5 + t - 1); // II before MI, viterbi.c:535-536
          DCP_UPD(I, (wM[j] + mi) + b, ip, 0 + t - 1);
        }
        float D = INF;
        int dbit = 0;
        DCP_UPD(D, wM1[r + 5] + md, dbit, 0); // M_{k-1}(l), D_{k-1}(l); +INF for k = 0
        DCP_UPD(D, wD1[r + 5] + dd, dbit, 1);
        nodes[(size_t)l * K + k] = (uint16_t)((unsigned)mp | ((unsigned)dbit << 5) | ((unsigned)ip << 6));
      }
    }
  }

  // ---- xnode words: warp w handles row r0 + w ----
  int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < nrows)
  {
    int const l = r0 + warp;
    int const T = l < 5 ? l : 5;
    // E: first-wins over M_0, D_0, M_1, D_1, ... = smallest (value, 2k+isD)
    float ev = INF;
    int ei = 0;
    for (int k = lane; k < K; k += 32)
    {
      DCP_UPD(ev, valM(l, k), ei, 2 * k + 0);
      DCP_UPD(ev, valD(l, k), ei, 2 * k + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      float const ov = __shfl_xor_sync(FULL_MASK, ev, o);
      int const oi = __shfl_xor_sync(FULL_MASK, ei, o);
      if (ov < ev || (ov == ev && oi < ei)) { ev = ov; ei = oi; }
    }
    if (lane == 0)
    {
      float N = INF, Jv = INF, C = INF;
      int pN = 0, pJ = 0, pC = 0;
      for (int t = T; t >= 1; --t)
      {
        int const lz = l - t;
        float const nil = __ldg(&pd.nulbg[codes[warp][t]]).x;
        float const Sz = lz == 0 ? 0.0f : INF;
        float const Nz = lz >= 1 ? valX(lz, 0) : INF, Jz = lz >= 1 ? valX(lz, 2) : INF,
                    Ez = lz >= 1 ? valX(lz, 3) : INF, Cz = lz >= 1 ? valX(lz, 4) : INF;
        DCP_UPD(N, Sz + SN + nil, pN, 0 + t - 1);
        DCP_UPD(N, Nz + NN + nil, pN, 5 + t - 1);
        DCP_UPD(Jv, Ez + EJ + nil, pJ, 0 + t - 1);
        DCP_UPD(Jv, Jz + JJ + nil, pJ, 5 + t - 1);
        DCP_UPD(C, Ez + EC + nil, pC, 0 + t - 1);
        DCP_UPD(C, Cz + CC + nil, pC, 5 + t - 1);
      }
      float const El = valX(l, 3), Nl = valX(l, 0), Jl = valX(l, 2), Cl = valX(l, 4);
      float B = INF, Tv = INF;
      int pB = 0, pT = 0;
      DCP_UPD(B, Nl + NB, pB, 1); // S+SB is +INF for l >= 1 (viterbi.c:495)
      DCP_UPD(B, El + EB, pB, 2);
      DCP_UPD(B, Jl + JB, pB, 3);
      DCP_UPD(Tv, El + ET, pT, 0);
      DCP_UPD(Tv, Cl + CT, pT, 1);
      (void)N; (void)Jv; (void)C; (void)ev;
      xnodes[l] = (uint32_t)pN | ((uint32_t)pB << 4) | ((uint32_t)ei << 6) | ((uint32_t)pC << 21) |
                  ((uint32_t)pT << 25) | ((uint32_t)pJ << 26);
    }
  }
}

// one thread per traced pair: count the steps of its path (first back-walk)
struct WalkCountArgs
{
  ProfileDesc const *profiles;
  Pair const *pairs;
  long long const *order;
  long long nitems;
  uint32_t const *xnodes;
  uint16_t const *nodes;
  long long const *xnode_off;
  long long const *node_off;
  int *nsteps;
};

__global__ void walk_count_kernel(WalkCountArgs a)
{
  long long const i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= a.nitems) return;
  long long const oidx = a.order[i];
  Pair const pr = a.pairs[oidx];
  int const K = a.profiles[pr.profile].K;
  a.nsteps[oidx] = trellis_walk(K, pr.len, a.xnodes + a.xnode_off[oidx], a.nodes + a.node_off[oidx], 0, nullptr, nullptr);
}

} // namespace dcp
