/*
 * TEST INFRASTRUCTURE: minimal stand-in for the un-vendored third-party header
 * imm_path.h (EBI-Metagenomics/imm, not present under /root/reference) so that
 * the reference's own c-core/trellis.c compiles unmodified into oracle/_ref.
 * Only what trellis.c:147-167 uses: imm_step(), imm_path_add(), imm_path_reverse().
 */
#ifndef IMM_PATH_H
#define IMM_PATH_H
#include <stdint.h>
#include <stdlib.h>

struct imm_step
{
  uint16_t state_id;
  int8_t seqsize;
  float score;
};

struct imm_path
{
  int nsteps;
  int capacity;
  struct imm_step *steps;
};

static inline struct imm_step imm_step(int state_id, int seqsize, float score)
{
  struct imm_step s = {(uint16_t)state_id, (int8_t)seqsize, score};
  return s;
}

static inline int imm_path_add(struct imm_path *p, struct imm_step s)
{
  if (p->nsteps == p->capacity)
  {
    int cap = p->capacity ? 2 * p->capacity : 256;
    struct imm_step *n = realloc(p->steps, sizeof(*n) * (size_t)cap);
    if (!n) return 1;
    p->steps = n;
    p->capacity = cap;
  }
  p->steps[p->nsteps++] = s;
  return 0;
}

static inline void imm_path_reverse(struct imm_path *p)
{
  for (int i = 0, j = p->nsteps - 1; i < j; ++i, --j)
  {
    struct imm_step t = p->steps[i];
    p->steps[i] = p->steps[j];
    p->steps[j] = t;
  }
}
#endif
