// score_row_kernel<Q, 32, ROW_WHOLE, DUMP>, Q = 1..8: one pair per warp (row_kernel.cuh).
#include "k_common.cuh"

namespace dcp {

template <bool DUMP>
static cudaError_t whole32(int Q, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (Q)
  {
  case 1: return launch_row_t<1, 32, ROW_WHOLE, DUMP>(a, sm_count, st);
  case 2: return launch_row_t<2, 32, ROW_WHOLE, DUMP>(a, sm_count, st);
  case 3: return launch_row_t<3, 32, ROW_WHOLE, DUMP>(a, sm_count, st);
  case 4: return launch_row_t<4, 32, ROW_WHOLE, DUMP>(a, sm_count, st);
  default: return launch_row_q58<32, ROW_WHOLE, DUMP>(Q, a, sm_count, st);
  }
}

cudaError_t launch_row_whole32(int Q, bool dump, StripArgs const &a, int sm_count, cudaStream_t st)
{
  return dump ? whole32<true>(Q, a, sm_count, st) : whole32<false>(Q, a, sm_count, st);
}

cudaError_t launch_row(int Q, int SEG, int mode, bool dump, StripArgs const &a, int sm_count, cudaStream_t st)
{
  if (mode != ROW_WHOLE)
  {
    if (dump) return cudaErrorInvalidValue;
    return launch_row_seg(Q, SEG, mode, a, sm_count, st);
  }
  if (SEG == 32) return launch_row_whole32(Q, dump, a, sm_count, st);
  return launch_row_sub(Q, SEG, dump, a, sm_count, st);
}

} // namespace dcp
