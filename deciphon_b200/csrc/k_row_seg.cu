// Segment kernels of profiles of more than 256 nodes (row_kernel.cuh): first / later full
// 256-node segment as <8, 32>, the tail segment in the layout of a profile of its size.
#include "k_common.cuh"

namespace dcp {

cudaError_t launch_row_seg(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st)
{
  if (mode == ROW_FIRST) return Q == 8 && SEG == 32 ? launch_row_t<8, 32, ROW_FIRST, false>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode == ROW_MID) return Q == 8 && SEG == 32 ? launch_row_t<8, 32, ROW_MID, false>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode != ROW_LAST) return cudaErrorInvalidValue;
  switch (SEG)
  {
  case 32: return launch_row_q58<32, ROW_LAST, false>(Q, a, sm_count, st);
  case 16: return launch_row_q58<16, ROW_LAST, false>(Q, a, sm_count, st);
  case 8: return launch_row_q58<8, ROW_LAST, false>(Q, a, sm_count, st);
  case 4: return launch_row_q58<4, ROW_LAST, false>(Q, a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

} // namespace dcp
