// score_row_kernel<Q, SEG, ROW_WHOLE, true>, SEG = 16/8/4, Q = 5..8: value dump for the trace pass.
#include "k_common.cuh"

namespace dcp {

cudaError_t launch_row_sub_dump(int Q, int SEG, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (SEG)
  {
  case 16: return launch_row_q58<16, ROW_WHOLE, true>(Q, a, sm_count, st);
  case 8: return launch_row_q58<8, ROW_WHOLE, true>(Q, a, sm_count, st);
  case 4: return launch_row_q58<4, ROW_WHOLE, true>(Q, a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

} // namespace dcp
