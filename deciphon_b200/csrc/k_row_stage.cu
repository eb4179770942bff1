// score_row_kernel<Q, 32, MODE, false, STAGE = true>, Q = 5, 6, 8 (FIRST / MID: Q = 8): profile-stationary
// CTAs with the short-code emission rows and the {null, background} table staged in shared memory by TMA
// (row_kernel.cuh).  The sub-warp shapes live in k_row_stage_sub.cu.
#include "k_common.cuh"

namespace dcp {

cudaError_t launch_row_stage_sub(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st);

template <int MODE>
static cudaError_t stage_q(int Q, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (Q)
  {
  case 5: return launch_row_stage_t<5, 32, MODE>(a, sm_count, st);
  case 6: return launch_row_stage_t<6, 32, MODE>(a, sm_count, st);
  case 8: return launch_row_stage_t<8, 32, MODE>(a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_row_stage(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st)
{
  if (SEG != 32) return launch_row_stage_sub(Q, SEG, mode, a, sm_count, st);
  if (mode == ROW_FIRST) return Q == 8 ? launch_row_stage_t<8, 32, ROW_FIRST>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode == ROW_MID) return Q == 8 ? launch_row_stage_t<8, 32, ROW_MID>(a, sm_count, st) : cudaErrorInvalidValue;
  if (mode == ROW_LAST) return stage_q<ROW_LAST>(Q, a, sm_count, st);
  if (mode == ROW_WHOLE) return stage_q<ROW_WHOLE>(Q, a, sm_count, st);
  return cudaErrorInvalidValue;
}

} // namespace dcp
