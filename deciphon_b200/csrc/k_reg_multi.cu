// score_reg_kernel<Q, W, DUMP>, W = 2/4/8 warps per pair, Q = 5..8 (score_kernel.cuh): the exact
// kernel for profiles of 257..2048 nodes -- redo of failed speculation and the trace value dump.
#include "kernels.h"
#include <algorithm>

namespace dcp {

template <int Q, int W, bool DUMP>
static cudaError_t reg_t(ScoreArgs const &a, int sm_count, cudaStream_t st)
{
  constexpr int T = ScoreCfg<W>::THREADS, G = ScoreCfg<W>::GROUPS;
  constexpr size_t SMEM = score_smem_bytes<Q, W>();
  // per device and cheap: set on every launch (one process drives several GPUs, one context each)
  cudaError_t e = cudaFuncSetAttribute(score_reg_kernel<Q, W, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_reg_kernel<Q, W, DUMP>, T, SMEM);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  unsigned long long const want = (a.nitems + G - 1) / G;
  unsigned const grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)per_sm * sm_count);
  if (grid == 0) return cudaSuccess;
  score_reg_kernel<Q, W, DUMP><<<grid, T, SMEM, st>>>(a);
  return cudaGetLastError();
}

template <int W, bool DUMP>
static cudaError_t reg_q(int Q, ScoreArgs const &a, int sm_count, cudaStream_t st)
{
  switch (Q)
  {
  case 5: return reg_t<5, W, DUMP>(a, sm_count, st);
  case 6: return reg_t<6, W, DUMP>(a, sm_count, st);
  case 7: return reg_t<7, W, DUMP>(a, sm_count, st);
  case 8: return reg_t<8, W, DUMP>(a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

template <bool DUMP>
static cudaError_t reg_w(int Q, int W, ScoreArgs const &a, int sm_count, cudaStream_t st)
{
  switch (W)
  {
  case 2: return reg_q<2, DUMP>(Q, a, sm_count, st);
  case 4: return reg_q<4, DUMP>(Q, a, sm_count, st);
  case 8: return reg_q<8, DUMP>(Q, a, sm_count, st);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_reg_multi(int Q, int W, bool dump, ScoreArgs const &a, int sm_count, cudaStream_t st)
{
  return dump ? reg_w<true>(Q, W, a, sm_count, st) : reg_w<false>(Q, W, a, sm_count, st);
}

} // namespace dcp
