// Shared by the kernel translation units: persistent-grid launch of a row kernel.
#pragma once
#include "kernels.h"
#include "row_kernel.cuh"
#include <algorithm>

namespace dcp {

template <int Q, int SEG, int MODE, bool DUMP>
cudaError_t launch_row_t(StripArgs const &a, int sm_count, cudaStream_t st)
{
  constexpr int T = 32 * ROW_WARPS, PER_CTA = ROW_WARPS * (32 / SEG);
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_row_kernel<Q, SEG, MODE, DUMP>, T, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  unsigned long long const want = (a.s.nitems + PER_CTA - 1) / PER_CTA;
  unsigned const grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)per_sm * sm_count);
  if (grid == 0) return cudaSuccess;
  score_row_kernel<Q, SEG, MODE, DUMP><<<grid, T, 0, st>>>(a);
  return cudaGetLastError();
}

// Profile-stationary variant (grid mode, one pair per warp): the 84 short-code rows of the CTA's
// current profile live in dynamic shared memory (+ 16 bytes for the mbarrier of the TMA bulk copy).
template <int Q, int SEG, int MODE>
cudaError_t launch_row_stage_t(StripArgs const &a, int sm_count, cudaStream_t st)
{
  constexpr int T = 32 * ROW_WARPS, PER_CLAIM = ROW_WARPS * (32 / SEG);
  constexpr size_t SMEM = (size_t)STAGE_ROWS * EmRows<Q, SEG>::ROWB + (size_t)stage_nulbg_codes<Q>() * 8 + 16;
  if (a.s.pairs || a.s.nseq <= 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(score_row_kernel<Q, SEG, MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_row_kernel<Q, SEG, MODE, false, true>, T, SMEM);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  unsigned long long const quads = ((unsigned long long)a.s.nseq + PER_CLAIM - 1) / PER_CLAIM;
  unsigned long long const want = (a.s.nitems / (unsigned long long)a.s.nseq) * quads;
  unsigned const grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)per_sm * sm_count);
  if (grid == 0) return cudaSuccess;
  score_row_kernel<Q, SEG, MODE, false, true><<<grid, T, SMEM, st>>>(a);
  return cudaGetLastError();
}

// Q = 5, 6, 8 switch for one (SEG, MODE, DUMP); layout_shape never hands a single-warp shape Q = 7
template <int SEG, int MODE, bool DUMP>
cudaError_t launch_row_q58(int Q, StripArgs const &a, int sm_count, cudaStream_t st)
{
  switch (Q)
  {
  case 5: return launch_row_t<5, SEG, MODE, DUMP>(a, sm_count, st);
  case 6: return launch_row_t<6, SEG, MODE, DUMP>(a, sm_count, st);
  case 8: return launch_row_t<8, SEG, MODE, DUMP>(a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_row_whole32(int Q, bool dump, StripArgs const &a, int sm_count, cudaStream_t st);
cudaError_t launch_row_sub(int Q, int SEG, bool dump, StripArgs const &a, int sm_count, cudaStream_t st);
cudaError_t launch_row_seg(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st);
cudaError_t launch_row_stage(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st);

} // namespace dcp
