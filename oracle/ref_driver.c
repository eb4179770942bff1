/*
 * ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin driver around the REFERENCE's own hot path, compiled where it lies
 * (/root/reference/c-core/{viterbi,trellis,state,error,loglevel,xrealloc,bug}.c)
 * into oracle/_ref/libdcpref_*.so by oracle/Makefile.  It only calls the
 * reference's public functions (c-core/viterbi.h, c-core/trellis.h); nothing of
 * the reference is copied.  Used to (a) validate oracle/dcp_oracle.c and the CUDA
 * path bit-for-bit and (b) time the reference CPU scan (bench.py --impl reference).
 *
 * The special transitions of ref_scan come from the reference's own c-core/xtrans.c
 * (compiled with two-macro stand-ins for imm_lprob.h / imm_dump.h, oracle/ref_shim/);
 * ref_set_xtrans still lets a test load any 13 costs.
 */
#include "imm_path.h"
#include "trellis.h"
#include "viterbi.h"
#include "xtrans.h"
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NCODES 1364
enum { C_BM, C_MM, C_MI, C_MD, C_IM, C_II, C_DM, C_DD, C_N };

int orc_window_next(int st[4], int seq_len, int core_size);
int orc_hit_extent(int nsteps, uint16_t const *state_ids, uint8_t const *sizes, int *hit_start,
                   int *hit_stop, int *begin, int *end);

struct ref_profile
{
  int K;
  struct viterbi *v;
};

struct code_arg
{
  uint8_t const *x;
};

static int const code_off[6] = {0, 0, 4, 20, 84, 340};

/* the viterbi_code_fn the reference calls back (c-core/viterbi.h:33, thread.c:92-96) */
static int code_fn(int pos, int len, void *arg)
{
  uint8_t const *x = ((struct code_arg *)arg)->x;
  int c = 0;
  for (int i = 0; i < len; ++i)
    c = c * 4 + x[pos + i];
  return code_off[len] + c;
}

/* Load a profile given in cost form, through the reference's own setters
 * (the calls protein_setup_viterbi makes, c-core/protein.c:353-394). */
struct ref_profile *ref_profile_new(int K, float const *nul, float const *bg, float const *em,
                                    float const *ct)
{
  struct ref_profile *p = malloc(sizeof(*p));
  if (!p) return NULL;
  p->K = K;
  p->v = viterbi_new();
  if (!p->v || viterbi_setup(p->v, K))
  {
    free(p);
    return NULL;
  }
  static enum core_trans_id const ids[C_N] = {CORE_TRANS_BM, CORE_TRANS_MM, CORE_TRANS_MI,
                                               CORE_TRANS_MD, CORE_TRANS_IM, CORE_TRANS_II,
                                               CORE_TRANS_DM, CORE_TRANS_DD};
  for (int j = 0; j < C_N; ++j)
    for (int k = 0; k < K; ++k)
      viterbi_set_core_trans(p->v, ids[j], ct[j * K + k], k);
  for (int i = 0; i < NCODES; ++i)
  {
    viterbi_set_null(p->v, nul[i], i);
    viterbi_set_background(p->v, bg[i], i);
    for (int k = 0; k < K; ++k)
      viterbi_set_match(p->v, em[(size_t)k * NCODES + i], k, i);
  }
  return p;
}

void ref_profile_del(struct ref_profile *p)
{
  if (!p) return;
  viterbi_del(p->v);
  free(p);
}

static void set_xtrans(struct viterbi *v, float const *xt)
{
  for (int i = 0; i < 13; ++i)
    viterbi_set_extr_trans(v, (enum extr_trans_id)i, xt[i]); /* same order as viterbi.h:4-19 */
}

void ref_set_xtrans(struct ref_profile *p, float const *xt) { set_xtrans(p->v, xt); }

float ref_null(struct ref_profile *p, uint8_t const *x, int L)
{
  struct code_arg a = {x};
  return viterbi_null(p->v, L, code_fn, &a);
}

float ref_cost(struct ref_profile *p, uint8_t const *x, int L)
{
  struct code_arg a = {x};
  return viterbi_cost(p->v, L, code_fn, &a);
}

/* viterbi_path + trellis_unzip; optionally copies the raw trellis words out. */
int ref_path(struct ref_profile *p, uint8_t const *x, int L, uint16_t *state_ids, uint8_t *sizes,
             int cap, uint32_t *xnodes_out, uint16_t *nodes_out)
{
  struct code_arg a = {x};
  if (viterbi_path(p->v, L, code_fn, &a)) return -10;
  struct trellis *tr = viterbi_trellis(p->v);
  if (xnodes_out) memcpy(xnodes_out, tr->xnodes, sizeof(uint32_t) * (size_t)(L + 1));
  if (nodes_out) memcpy(nodes_out, tr->nodes, sizeof(uint16_t) * (size_t)(L + 1) * (size_t)p->K);
  struct imm_path path = {0, 0, NULL};
  if (trellis_unzip(tr, L, &path))
  {
    free(path.steps);
    return -11;
  }
  int n = path.nsteps;
  if (n > cap)
  {
    free(path.steps);
    return -1;
  }
  for (int i = 0; i < n; ++i)
  {
    state_ids[i] = path.steps[i].state_id;
    sizes[i] = (uint8_t)path.steps[i].seqsize;
  }
  free(path.steps);
  return n;
}

/*
 * The reference CPU scan minus HMMER and file output:
 * OpenMP over contiguous profile partitions (c-core/scan.c:188-208,
 * partition_size.c:13-16); each thread loops profiles -> reads -> windows
 * (thread.c:49-86) and per window does xtrans, viterbi_null, viterbi_cost and,
 * for lrt >= 0, viterbi_path + trellis_unzip (thread.c:98-128) and the hit
 * extent that feeds the next window (thread.c:130-166).
 *
 * profs[nprof]; reads are symbols 0..3, concatenated, read r = x[off[r]..off[r+1]).
 * out_null/out_alt (optional) receive the FIRST window's costs per (profile, read).
 * thread_seconds (optional, nthreads entries) receives each partition's busy time, so the
 * caller can report the parallel efficiency of the count partitioning.
 * Returns wall seconds; *cells = sum of L*K over windows, *nhits = windows with lrt >= 0.
 */
double ref_scan(int nprof, struct ref_profile **profs, int nreads, uint8_t const *x,
                int64_t const *off, int multi_hits, int hmmer3_compat, int nthreads,
                float *out_null, float *out_alt, double *cells, int64_t *nhits,
                double *thread_seconds)
{
  if (nthreads > nprof) nthreads = nprof;
  if (nthreads < 1) nthreads = 1;
  double tot_cells = 0;
  int64_t tot_hits = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for num_threads(nthreads) schedule(static, 1) reduction(+ : tot_cells, tot_hits)
  for (int part = 0; part < nthreads; ++part)
  {
    int start = 0;
    for (int i = 0; i < part; ++i)
      start += (int)(((long)(nprof - i > 0 ? nprof - i : 0) + nthreads - 1) / nthreads);
    int size = (int)(((long)(nprof - part > 0 ? nprof - part : 0) + nthreads - 1) / nthreads);
    struct timespec p0, p1;
    clock_gettime(CLOCK_MONOTONIC, &p0);
    int cap = 0;
    uint16_t *ids = NULL;
    uint8_t *szs = NULL;
    for (int pi = start; pi < start + size && pi < nprof; ++pi)
    {
      struct ref_profile *p = profs[pi];
      for (int r = 0; r < nreads; ++r)
      {
        int len = (int)(off[r + 1] - off[r]);
        int st[4] = {-1, 0, -1, -1};
        while (orc_window_next(st, len, p->K))
        {
          int L = st[1] - st[0];
          uint8_t const *w = x + off[r] + st[0];
          struct xtrans xt; /* work_reset(work, max(L / 3, 1)), c-core/thread.c:112 */
          xtrans_init(&xt);
          xtrans_setup(&xt, multi_hits != 0, hmmer3_compat != 0, L / 3 > 1 ? L / 3 : 1);
          xtrans_setup_viterbi(&xt, p->v);
          float nul = ref_null(p, w, L);
          float alt = ref_cost(p, w, L);
          tot_cells += (double)L * p->K;
          if (st[2] == 0 && out_null) out_null[(size_t)pi * nreads + r] = nul;
          if (st[2] == 0 && out_alt) out_alt[(size_t)pi * nreads + r] = alt;
          float lrt = -2 * ((-nul) - (-alt));
          if (!isfinite(lrt) || lrt < 0) continue;
          tot_hits += 1;
          int need = L + 2 * p->K + 64;
          if (need > cap)
          {
            cap = need * 2;
            ids = realloc(ids, sizeof(*ids) * (size_t)cap);
            szs = realloc(szs, (size_t)cap);
          }
          int n = ref_path(p, w, L, ids, szs, cap, NULL, NULL);
          int hs, he, b, e;
          if (n > 0 && orc_hit_extent(n, ids, szs, &hs, &he, &b, &e)) st[3] = he - 1;
        }
      }
    }
    free(ids);
    free(szs);
    clock_gettime(CLOCK_MONOTONIC, &p1);
    if (thread_seconds)
      thread_seconds[part] = (double)(p1.tv_sec - p0.tv_sec) + 1e-9 * (double)(p1.tv_nsec - p0.tv_nsec);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (cells) *cells = tot_cells;
  if (nhits) *nhits = tot_hits;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int ref_num_lanes(void)
{
#if __AVX512F__
  return 16;
#else
  return 8;
#endif
}
