/*
 * dcp_oracle.c -- TEST INFRASTRUCTURE ONLY (the "oracle").
 *
 * A plain scalar C restatement of Deciphon's scan hot path, written from the
 * behaviour of the reference (never copied): every function cites the reference
 * file:line it follows.  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA path; nothing in the product
 * (deciphon_b200/) may import, link or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the reference's golden snap.dcs rows (LRT 291.6 / 349.3 / 360.4 and the
 *       full M1..MK paths) using parameters from the golden minifam.dcp, and
 *   (2) the reference's own viterbi.c/trellis.c compiled into oracle/_ref
 *       (bit-for-bit on scores, step-for-step on paths, random inputs).
 *
 * All arithmetic is fp32 min-plus over costs (= -log prob), +INF = impossible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NCODES 1364 /* c-core/viterbi.c:13 TABLE_SIZE = 4+16+64+256+1024 */

/* state ids, c-core/state.h:7-25 */
enum
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

/* order of the 13 special-transition costs used everywhere in this repo */
enum { X_RR, X_SN, X_NN, X_SB, X_NB, X_EB, X_JB, X_EJ, X_JJ, X_EC, X_CC, X_ET, X_CT, X_N };
/* order of the 8 per-node transition cost arrays (each [K], destination-indexed) */
enum { C_BM, C_MM, C_MI, C_MD, C_IM, C_II, C_DM, C_DD, C_N };

static inline float fmin2(float a, float b) { return fminf(a, b); }

/*
 * Special transitions for one window of `window_len` nucleotides.
 * Follows c-core/thread.c:111-112 (seq_size = max(L/3, 1), integer division) and
 * c-core/xtrans.c:21-51 (log-probs; the log() calls are double precision and the
 * results are stored to float) then c-core/xtrans.c:53-68 (negation into costs).
 */
void orc_xtrans(int window_len, int multi_hits, int hmmer3_compat, float out[X_N])
{
  int seq_size = window_len / 3;
  if (seq_size < 1) seq_size = 1;
  float L = (float)seq_size;
  float q = 0.0;
  float log_q = -INFINITY;
  if (multi_hits)
  {
    q = 0.5;
    log_q = log(0.5);
  }
  float lp = log(L) - log(L + 2 + q / (1 - q));
  float l1p = log(2 + q / (1 - q)) - log(L + 2 + q / (1 - q));
  float lr = log(L) - log(L + 1);

  float NN = lp, CC = lp, JJ = lp;
  float NB = l1p, CT = l1p, JB = l1p;
  float RR = lr;
  float EJ = log_q;
  float EC = log(1 - q);
  if (hmmer3_compat) NN = CC = JJ = logf(1);

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

/*
 * Table index of the len-mer starting at pos (x holds symbols 0..3 = ACGT).
 * Restates the third-party imm_eseq_get(seq,pos,len,1) called at
 * c-core/thread.c:92-96; convention pinned by the golden LRTs:
 * big-endian base 4 plus offsets {0,4,20,84,340}.
 */
int orc_code(uint8_t const *x, int pos, int len)
{
  static int const off[6] = {0, 0, 4, 20, 84, 340};
  int c = 0;
  for (int i = 0; i < len; ++i)
    c = c * 4 + x[pos + i];
  return off[len] + c;
}

/*
 * Profile log-probs (.dcp form) -> destination-indexed costs, the layout
 * transform of c-core/protein.c:353-383: BM_k = -BMk[k]; MM,MD,IM,DM,DD of node
 * k are stored at k+1, MI,II at k; node 0's incoming and node K-1's MI/II = +INF.
 * trans is [(K+1)][7] in the order MM,MI,MD,IM,II,DM,DD (c-core/trans.h).
 * out is [C_N][K].
 */
void orc_core_costs(int K, float const *BMk, float const *trans, float *out)
{
  for (int i = 0; i < C_N * K; ++i)
    out[i] = INFINITY;
  for (int k = 0; k < K; ++k)
    out[C_BM * K + k] = -BMk[k];
  for (int k = 0; k + 1 < K; ++k)
  {
    float const *t = trans + 7 * k;
    out[C_MM * K + k + 1] = -t[0];
    out[C_MI * K + k + 0] = -t[1];
    out[C_MD * K + k + 1] = -t[2];
    out[C_IM * K + k + 1] = -t[3];
    out[C_II * K + k + 0] = -t[4];
    out[C_DM * K + k + 1] = -t[5];
    out[C_DD * K + k + 1] = -t[6];
  }
}

/* Null model, c-core/viterbi.c:696-719: R(0) = -RR; R(l) = min_t R(l-t)+RR+nul. */
float orc_null(float const *nul, float const *xt, uint8_t const *x, int L)
{
  float RR = xt[X_RR];
  float R[6];
  for (int i = 0; i < 6; ++i)
    R[i] = INFINITY;
  R[0] = -RR; /* R[j] holds row l-1-j before row l */
  for (int l = 1; l <= L; ++l)
  {
    float r = INFINITY;
    int T = l < 5 ? l : 5;
    for (int t = T; t >= 1; --t)
      r = fmin2(r, R[t - 1] + RR + nul[orc_code(x, l - t, t)]);
    for (int j = 5; j > 0; --j)
      R[j] = R[j - 1];
    R[0] = r;
  }
  return R[0];
}

/*
 * Alternative model, score only: the recurrence of c-core/viterbi.c:451-600 in
 * factored form.  P_k(l') = min over the four predecessors of
 * M_k, Q_k(l') = min over the two predecessors of I_k; both are hoisted out of
 * the emission-length loop, which is value-exact because fp32 rounding is
 * monotone.  Memory is a 6-row ring like the reference's (viterbi.c:160-161).
 *
 * nul,bg: [1364] costs.  em: [K][1364] costs (node-major).  ct: [C_N][K].
 */
float orc_alt(int K, float const *nul, float const *bg, float const *em, float const *ct,
              float const *xt, uint8_t const *x, int L)
{
  float const *BM = ct + C_BM * K, *MM = ct + C_MM * K, *MI = ct + C_MI * K,
              *MD = ct + C_MD * K, *IM = ct + C_IM * K, *II = ct + C_II * K,
              *DM = ct + C_DM * K, *DD = ct + C_DD * K;
  /* ring over rows: slot (l % 6) */
  float *P = malloc(sizeof(float) * 6 * K), *Q = malloc(sizeof(float) * 6 * K);
  float *M = malloc(sizeof(float) * K), *I = malloc(sizeof(float) * K),
        *D = malloc(sizeof(float) * K);
  float N[6], J[6], C[6], E[6], S[6], B;
  float Tfinal = INFINITY;

  for (int l = 0; l <= L; ++l)
  {
    int s = l % 6;
    for (int k = 0; k < K; ++k)
      M[k] = I[k] = D[k] = INFINITY;
    N[s] = J[s] = C[s] = E[s] = INFINITY;
    S[s] = l ? INFINITY : 0.f; /* viterbi.c:472 */
    if (l > 0)
    {
      int T = l < 5 ? l : 5;
      for (int t = T; t >= 1; --t)
      {
        int z = (l - t) % 6;
        int code = orc_code(x, l - t, t);
        float nil = nul[code], b = bg[code];
        N[s] = fmin2(N[s], fmin2(S[z] + xt[X_SN] + nil, N[z] + xt[X_NN] + nil)); /* :492-493 */
        J[s] = fmin2(J[s], fmin2(E[z] + xt[X_EJ] + nil, J[z] + xt[X_JJ] + nil)); /* :498-499 */
        C[s] = fmin2(C[s], fmin2(E[z] + xt[X_EC] + nil, C[z] + xt[X_CC] + nil)); /* :501-502 */
        for (int k = 0; k < K; ++k)
        {
          M[k] = fmin2(M[k], P[z * K + k] + em[(size_t)k * NCODES + code]); /* :526-529 */
          I[k] = fmin2(I[k], Q[z * K + k] + b);                             /* :535-536 */
        }
      }
      for (int k = 1; k < K; ++k) /* :538, :561-580 fixed point */
        D[k] = fmin2(M[k - 1] + MD[k], D[k - 1] + DD[k]);
      float e = INFINITY; /* :540-558 */
      for (int k = 0; k < K; ++k)
        e = fmin2(e, fmin2(M[k], D[k]));
      E[s] = e;
    }
    /* :495-496, :582-583 (S+SB only finite at l = 0, :473) */
    B = fmin2(fmin2(S[s] + xt[X_SB], N[s] + xt[X_NB]), fmin2(E[s] + xt[X_EB], J[s] + xt[X_JB]));
    Tfinal = fmin2(E[s] + xt[X_ET], C[s] + xt[X_CT]); /* :585-586 */
    for (int k = 0; k < K; ++k)
    {
      float pm = k ? M[k - 1] : INFINITY, pi = k ? I[k - 1] : INFINITY, pd = k ? D[k - 1] : INFINITY;
      P[s * K + k] = fmin2(fmin2(B + BM[k], pm + MM[k]), fmin2(pi + IM[k], pd + DM[k]));
      Q[s * K + k] = fmin2(I[k] + II[k], M[k] + MI[k]);
    }
  }
  free(P); free(Q); free(M); free(I); free(D);
  return Tfinal; /* :599 */
}

/* strict-less update == first candidate wins (viterbi.c:201-212, intrinsics.h:144-149) */
#define UPD(val, ptr, cand, tag)                                               \
  do                                                                           \
  {                                                                            \
    float c_ = (cand);                                                         \
    if (c_ < (val))                                                            \
    {                                                                          \
      (val) = c_;                                                              \
      (ptr) = (tag);                                                           \
    }                                                                          \
  } while (0)

/*
 * Trace pass: same DP, recording the first-wins argmin of every state in the
 * reference's candidate order (viterbi.c:485-586) and packing
 * it exactly like after() (viterbi.c:631-692) / trellis_set (trellis.h:42-56):
 *   xnodes[l]: N bits 0-3, B 4-5, E 6-20, C 21-24, T 25, J 26-29
 *   nodes[l*K+k]: M bits 0-4, D bit 5, I bits 6-9
 * Candidates use the reference's unfactored arithmetic (s + tau) + e.
 * xnodes has L+1 entries, nodes (L+1)*K.  Returns the alt cost T(L).
 */
float orc_trace(int K, float const *nul, float const *bg, float const *em, float const *ct,
                float const *xt, uint8_t const *x, int L, uint32_t *xnodes, uint16_t *nodes)
{
  float const *BM = ct + C_BM * K, *MM = ct + C_MM * K, *MI = ct + C_MI * K,
              *MD = ct + C_MD * K, *IM = ct + C_IM * K, *II = ct + C_II * K,
              *DM = ct + C_DM * K, *DD = ct + C_DD * K;
  size_t W = (size_t)K;
  float *M = malloc(sizeof(float) * 6 * W), *I = malloc(sizeof(float) * 6 * W),
        *D = malloc(sizeof(float) * 6 * W);
  float N[6], J[6], C[6], E[6], S[6], B[6];
  float Tl = INFINITY;

  for (int l = 0; l <= L; ++l)
  {
    int s = l % 6;
    float *Ms = M + s * W, *Is = I + s * W, *Ds = D + s * W;
    for (int k = 0; k < K; ++k)
      Ms[k] = Is[k] = Ds[k] = INFINITY;
    N[s] = J[s] = C[s] = E[s] = B[s] = INFINITY;
    S[s] = l ? INFINITY : 0.f;
    Tl = INFINITY;
    int pN = 0, pB = 0, pE = 0, pJ = 0, pC = 0, pT = 0;
    uint16_t *node = nodes + (size_t)l * W;
    if (l == 0)
    {
      B[0] = xt[X_SB]; /* viterbi.c:473; before() writes all-zero fields :602-629 */
      xnodes[0] = 0;
      memset(node, 0, sizeof(uint16_t) * W);
      continue;
    }
    int T = l < 5 ? l : 5;
    for (int k = 0; k < K; ++k)
      node[k] = 0;
    for (int t = T; t >= 1; --t)
    {
      int z = (l - t) % 6;
      float const *Mz = M + z * W, *Iz = I + z * W, *Dz = D + z * W;
      int code = orc_code(x, l - t, t);
      float nil = nul[code], b = bg[code];
      UPD(N[s], pN, S[z] + xt[X_SN] + nil, 0 + t - 1);
      UPD(N[s], pN, N[z] + xt[X_NN] + nil, 5 + t - 1);
      UPD(J[s], pJ, E[z] + xt[X_EJ] + nil, 0 + t - 1);
      UPD(J[s], pJ, J[z] + xt[X_JJ] + nil, 5 + t - 1);
      UPD(C[s], pC, E[z] + xt[X_EC] + nil, 0 + t - 1);
      UPD(C[s], pC, C[z] + xt[X_CC] + nil, 5 + t - 1);
      for (int k = 0; k < K; ++k)
      {
        float e = em[(size_t)k * NCODES + code];
        float pm = k ? Mz[k - 1] : INFINITY, pi = k ? Iz[k - 1] : INFINITY,
              pd = k ? Dz[k - 1] : INFINITY;
        int m = node[k] & 31, i = (node[k] >> 6) & 15;
        UPD(Ms[k], m, (B[z] + BM[k]) + e, 0 + t - 1);
        UPD(Ms[k], m, (pm + MM[k]) + e, 5 + t - 1);
        UPD(Ms[k], m, (pi + IM[k]) + e, 10 + t - 1);
        UPD(Ms[k], m, (pd + DM[k]) + e, 15 + t - 1);
        UPD(Is[k], i, (Iz[k] + II[k]) + b, 5 + t - 1); /* II before MI, viterbi.c:535-536 */
        UPD(Is[k], i, (Mz[k] + MI[k]) + b, 0 + t - 1);
        node[k] = (uint16_t)((node[k] & (1u << 5)) | (unsigned)m | ((unsigned)i << 6));
      }
    }
    for (int k = 1; k < K; ++k)
    {
      int d = 0;
      UPD(Ds[k], d, Ms[k - 1] + MD[k], 0);
      UPD(Ds[k], d, Ds[k - 1] + DD[k], 1);
      node[k] = (uint16_t)(node[k] | (d << 5));
    }
    for (int k = 0; k < K; ++k)
    {
      UPD(E[s], pE, Ms[k], 2 * k + 0);
      UPD(E[s], pE, Ds[k], 2 * k + 1);
    }
    /* B: S+SB is +INF for l >= 1; order SB, NB, EB, JB (viterbi.c:495-496,582-583) */
    UPD(B[s], pB, N[s] + xt[X_NB], 1);
    UPD(B[s], pB, E[s] + xt[X_EB], 2);
    UPD(B[s], pB, J[s] + xt[X_JB], 3);
    UPD(Tl, pT, E[s] + xt[X_ET], 0);
    UPD(Tl, pT, C[s] + xt[X_CT], 1);
    /* node K-1's I field is never written by after() (viterbi.c:651-673) */
    node[K - 1] = (uint16_t)(node[K - 1] & ~(15u << 6));
    xnodes[l] = (uint32_t)pN | ((uint32_t)pB << 4) | ((uint32_t)pE << 6) | ((uint32_t)pC << 21) |
                ((uint32_t)pT << 25) | ((uint32_t)pJ << 26);
  }
  free(M); free(I); free(D);
  return Tl;
}

/*
 * Back-walk T@L -> S@0, c-core/trellis.c:147-167 with previous_state (:51-98)
 * and emission_size (:100-113).  Writes steps in path order (S first).
 * Returns the number of steps, or -1 if cap is too small.
 */
int orc_unzip(int K, int L, uint32_t const *xnodes, uint16_t const *nodes, uint16_t *state_ids,
              uint8_t *sizes, int cap)
{
  int n = 0;
  int state = ST_T, stage = L;
  while (state != ST_S || stage)
  {
    int size = 0, prev = 0;
    int msb = state & (3 << 14);
    if (msb == ST_X)
    {
      uint32_t xn = xnodes[stage];
      unsigned vN = xn & 15, vB = (xn >> 4) & 3, vE = (xn >> 6) & 0x7fff, vC = (xn >> 21) & 15,
               vT = (xn >> 25) & 1, vJ = (xn >> 26) & 15;
      if (state == ST_T) { size = 0; prev = vT ? ST_C : ST_E; }
      else if (state == ST_E) { size = 0; prev = (vE & 1) ? (ST_D | (vE / 2 + 1)) : (ST_M | (vE / 2 + 1)); }
      else if (state == ST_C) { size = vC % 5 + 1; prev = vC / 5 ? ST_C : ST_E; }
      else if (state == ST_J) { size = vJ % 5 + 1; prev = vJ / 5 ? ST_J : ST_E; }
      else if (state == ST_N) { size = vN % 5 + 1; prev = vN / 5 ? ST_N : ST_S; }
      else if (state == ST_B) { size = 0; prev = (int[]){ST_S, ST_N, ST_E, ST_J}[vB]; }
      else return -2;
    }
    else
    {
      int k = (state & 0x3fff) - 1; /* state_core_idx, state.c:25 */
      uint16_t nd = nodes[(size_t)stage * K + k];
      unsigned vM = nd & 31, vD = (nd >> 5) & 1, vI = (nd >> 6) & 15;
      if (msb == ST_M)
      {
        size = vM % 5 + 1;
        int src = vM / 5;
        prev = src == 0 ? ST_B : src == 1 ? (ST_M | k) : src == 2 ? (ST_I | k) : (ST_D | k);
        if (src && k <= 0) return -3;
      }
      else if (msb == ST_I) { size = vI % 5 + 1; prev = vI / 5 ? (ST_I | (k + 1)) : (ST_M | (k + 1)); }
      else { size = 0; if (k <= 0) return -3; prev = vD ? (ST_D | k) : (ST_M | k); }
    }
    if (n >= cap) return -1;
    state_ids[n] = (uint16_t)state;
    sizes[n] = (uint8_t)size;
    ++n;
    state = prev;
    stage -= size;
    if (stage < 0) return -4;
  }
  if (n >= cap) return -1;
  state_ids[n] = (uint16_t)ST_S;
  sizes[n] = 0;
  ++n;
  for (int i = 0, j = n - 1; i < j; ++i, --j)
  {
    uint16_t a = state_ids[i]; state_ids[i] = state_ids[j]; state_ids[j] = a;
    uint8_t b = sizes[i]; sizes[i] = sizes[j]; sizes[j] = b;
  }
  return n;
}

/* lrt.h:6-9 applied to log-likelihoods (= -cost), thread.c:114-119 */
float orc_lrt(float null_cost, float alt_cost)
{
  float null = -null_cost, alt = -alt_cost;
  return -2 * (null - alt);
}

/*
 * Hit extent of a decoded path, c-core/thread.c:130-166: window-relative
 * [hit_start, hit_stop) from the first B to the last E, and the step range
 * [begin, end) = first B .. step after the last E.  Returns 0 if the path has
 * no B..E segment.
 */
int orc_hit_extent(int nsteps, uint16_t const *state_ids, uint8_t const *sizes, int *hit_start,
                   int *hit_stop, int *begin, int *end)
{
  int pos = 0, i = 0;
  while (i < nsteps && state_ids[i] != ST_B)
    pos += sizes[i++];
  if (i >= nsteps) return 0;
  *hit_start = pos;
  *begin = i;
  int stop = pos, e = -1;
  for (int j = i; j < nsteps; ++j)
  {
    if (state_ids[j] == ST_E) { stop = pos; e = j + 1; }
    pos += sizes[j];
  }
  if (e < 0) return 0;
  *hit_stop = stop;
  *end = e;
  return 1;
}

/*
 * Window iteration, c-core/window.c:7-37.  st = {start, stop, idx, last_hit_pos}
 * initialised to {-1, 0, -1, -1}.  Returns 0 when the sequence is exhausted.
 */
int orc_window_next(int st[4], int seq_len, int core_size)
{
  if (st[1] == seq_len) return 0;
  int stop_miss = st[1] + 1;
  int a = st[0] + 1, b = st[0] + st[3] + 1;
  int start_miss = a > b ? a : b;
  int c = stop_miss - core_size * 4;
  if (c > start_miss) start_miss = c;
  int w = core_size * 50 < 100000 ? core_size * 50 : 100000;
  st[0] = start_miss;
  st[1] = start_miss + w;
  if (st[1] > seq_len) st[1] = seq_len;
  st[2] += 1;
  return 1;
}

/* state.c:47-90 */
int orc_state_name(int id, char *name)
{
  int msb = id & (3 << 14);
  if (msb == ST_X)
  {
    char const *t = "FRGSNBEJCT";
    int i = id & 0x3fff;
    if (i > 9) return 1;
    name[0] = t[i];
    name[1] = 0;
    return 0;
  }
  name[0] = msb == ST_M ? 'M' : msb == ST_I ? 'I' : 'D';
  int v = id & 0x3fff, n = 0;
  char tmp[8];
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  for (int i = 0; i < n; ++i)
    name[1 + i] = tmp[n - 1 - i];
  name[1 + n] = 0;
  return 0;
}
