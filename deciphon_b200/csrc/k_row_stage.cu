// score_row_kernel<Q, 32, MODE, false, STAGE = true>, Q = 5, 6, 8 (FIRST / MID: Q = 8): profile-stationary CTAs with
// the short-code emission rows staged in shared memory by TMA (row_kernel.cuh).
#include "k_common.cuh"

namespace dcp {

cudaError_t launch_row_stage(int Q, int mode, StripArgs const &a, int sm_count, cudaStream_t st)
{
  if (mode == ROW_FIRST) return Q == 8 ? launch_row_stage_t<8, ROW_FIRST>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode == ROW_MID) return Q == 8 ? launch_row_stage_t<8, ROW_MID>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode == ROW_LAST)
    switch (Q)
    {
    case 5: return launch_row_stage_t<5, ROW_LAST>(a, sm_count, st);
    case 6: return launch_row_stage_t<6, ROW_LAST>(a, sm_count, st);
    case 8: return launch_row_stage_t<8, ROW_LAST>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
    }
  if (mode != ROW_WHOLE) return cudaErrorInvalidValue;
  switch (Q)
  {
  case 5: return launch_row_stage_t<5, ROW_WHOLE>(a, sm_count, st);
  case 6: return launch_row_stage_t<6, ROW_WHOLE>(a, sm_count, st);
  case 8: return launch_row_stage_t<8, ROW_WHOLE>(a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

} // namespace dcp
