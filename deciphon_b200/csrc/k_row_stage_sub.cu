// score_row_kernel<Q, SEG, WHOLE / LAST, false, STAGE = true>, SEG = 16/8/4, Q = 5, 6, 8: the sub-warp
// shapes of the profile-stationary kernels (8 / 16 / 32 reads of one profile per CTA).
#include "k_common.cuh"

namespace dcp {

template <int SEG, int MODE>
static cudaError_t stage_sub_q(int Q, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (Q)
  {
  case 5: return launch_row_stage_t<5, SEG, MODE>(a, sm_count, st);
  case 6: return launch_row_stage_t<6, SEG, MODE>(a, sm_count, st);
  case 8: return launch_row_stage_t<8, SEG, MODE>(a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

template <int MODE>
static cudaError_t stage_sub(int Q, int SEG, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (SEG)
  {
  case 16: return stage_sub_q<16, MODE>(Q, a, sm_count, st);
  case 8: return stage_sub_q<8, MODE>(Q, a, sm_count, st);
  case 4: return stage_sub_q<4, MODE>(Q, a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_row_stage_sub(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st)
{
  if (mode == ROW_WHOLE) return stage_sub<ROW_WHOLE>(Q, SEG, a, sm_count, st);
  if (mode == ROW_LAST) return stage_sub<ROW_LAST>(Q, SEG, a, sm_count, st);
  return cudaErrorInvalidValue;
}

} // namespace dcp
