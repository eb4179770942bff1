/*
 * TEST INFRASTRUCTURE: stand-in for the un-vendored third-party header imm_dump.h so that
 * c-core/xtrans.c compiles unmodified (only xtrans_dump, a debug printer, uses it).
 */
#ifndef IMM_DUMP_H
#define IMM_DUMP_H
#include <stddef.h>
#include <stdio.h>
static inline void imm_dump_array_f32(size_t n, float const *x, FILE *fp)
{
  fputc('[', fp);
  for (size_t i = 0; i < n; ++i)
    fprintf(fp, i ? ",%g" : "%g", (double)x[i]);
  fputc(']', fp);
}
#endif
