/*
 * TEST INFRASTRUCTURE: stand-in for the un-vendored third-party header imm_lprob.h
 * (EBI-Metagenomics/imm) so that the reference's own c-core/xtrans.c compiles unmodified
 * into oracle/_ref.  xtrans.c:10-18,30 uses exactly two macros: the log-probabilities of
 * 1 and of 0.
 */
#ifndef IMM_LPROB_H
#define IMM_LPROB_H
#include <math.h>
#define IMM_LPROB_ONE 0.0f
#define IMM_LPROB_ZERO (-INFINITY)
#endif
