/*
 * ref_xtrans_capture.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Built together with the reference's own c-core/xtrans.c (compiled where it lies) into
 * oracle/_ref/libdcpref_xtrans.so.  It supplies the one symbol xtrans.c needs from viterbi.c,
 * viterbi_set_extr_trans, as a recorder, so that the 13 special-transition COSTS the reference
 * loads for a window (xtrans_setup + xtrans_setup_viterbi, c-core/xtrans.c:21-68, called from
 * thread.c:112 / work.c) can be compared bit for bit with dcpgpu_xtrans and the oracle.
 */
#include "viterbi.h"
#include "xtrans.h"

static _Thread_local float captured[13];

void viterbi_set_extr_trans(struct viterbi *v, enum extr_trans_id id, float scalar)
{
  (void)v;
  captured[(int)id] = scalar;
}

/* out[13] in the order of enum extr_trans_id (c-core/viterbi.h:4-19) */
void ref_xtrans(int seq_size, int multi_hits, int hmmer3_compat, float *out)
{
  struct xtrans x;
  xtrans_init(&x);
  xtrans_setup(&x, multi_hits != 0, hmmer3_compat != 0, seq_size);
  xtrans_setup_viterbi(&x, (struct viterbi *)0);
  for (int i = 0; i < 13; ++i)
    out[i] = captured[i];
}

/* many window lengths at once: out[n][13] */
void ref_xtrans_many(int n, int const *seq_size, int multi_hits, int hmmer3_compat, float *out)
{
  for (int i = 0; i < n; ++i)
    ref_xtrans(seq_size[i], multi_hits, hmmer3_compat, out + 13 * (long)i);
}
