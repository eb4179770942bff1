// Host-callable launchers of the score kernels.  The kernels are instantiated in their own
// translation units (k_*.cu, compiled in parallel); dcpgpu.cu only sees these functions.
// Every launcher sizes a persistent grid (resident CTAs per SM x SMs, never more than the work),
// launches on `st` and returns the CUDA status; cudaErrorInvalidValue for a shape that is not built.
#pragma once
#include "score_kernel.cuh"
#include "strip_kernel.cuh"

namespace dcp {

// score_row_kernel<Q, SEG, MODE, DUMP> (row_kernel.cuh): SEG = 32 with Q = 1..8, SEG = 16/8/4 with
// Q = 5..8; MODE = RowMode; DUMP only with ROW_WHOLE; FIRST/MID only as <8, 32>.
cudaError_t launch_row(int Q, int SEG, int mode, bool dump, StripArgs const &a, int sm_count, cudaStream_t st);
// The profile-stationary variants (grid mode only; Q = 5, 6, 8; FIRST / MID as <8, 32>): the short-code rows
// and the {null, background} table of a CTA's current profile staged in shared memory by TMA bulk copies.
cudaError_t launch_row_stage(int Q, int SEG, int mode, StripArgs const &a, int sm_count, cudaStream_t st);

// score_reg_kernel<Q, W, DUMP> (score_kernel.cuh), W = 2/4/8 warps per pair, Q = 5..8: the exact
// kernel for profiles of more than 256 nodes (redo of failed speculation, trace value dump).
cudaError_t launch_reg_multi(int Q, int W, bool dump, ScoreArgs const &a, int sm_count, cudaStream_t st);

} // namespace dcp
