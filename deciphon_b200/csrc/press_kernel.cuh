// Frame-state emission tables on the device: the press path's hot loop.
//
// For every state (a match state of a node, or a profile's null / background state) the reference
// fills a 1364-entry table with the log-probability of emitting each 1..5-nucleotide fragment
// (imm_score_table_scores over an imm_frame_state, called at c-core/protein.c:102,
// protein_null.c:24 and protein_background.c:19; the state is set up at model.c:318-336 from a
// base distribution, codon marginals and the indel rate epsilon).  The third-party formula is
// stated in DESIGN.md and pinned by the tests on the 576 nodes of the reference's golden minifam.dcp;
// this kernel evaluates the same closed form in fp32: one CTA per state, the 125 codon marginals
// and 4 base probabilities staged in shared memory, one thread per code.
//
// HBM-bound by its output (5,456 bytes per state against ~50 flops per entry).
#pragma once
#include "layout.cuh"
#include <math_constants.h>

namespace dcp {

struct FrameArgs
{
  float const *nuclt; // [nstates][4]   base log-probs
  float const *marg;  // [nstates][125] codon marginal log-probs, index 4 = any base
  float *out;         // [nstates][1364] log-probs
  float eps;
  int nstates;
};

__global__ void __launch_bounds__(256) frame_table_kernel(FrameArgs a)
{
  __shared__ float P[125];
  __shared__ float b[4];
  int const st = blockIdx.x;
  for (int i = threadIdx.x; i < 125; i += blockDim.x) P[i] = expf(a.marg[(size_t)st * 125 + i]);
  if (threadIdx.x < 4) b[threadIdx.x] = expf(a.nuclt[(size_t)st * 4 + threadIdx.x]);
  __syncthreads();
  float const e = a.eps, f = 1.0f - a.eps;
  auto M = [&](int x, int y, int z) { return P[x * 25 + y * 5 + z]; };
  auto single = [&](int x) { return M(x, 4, 4) + M(4, x, 4) + M(4, 4, x); };
  auto pair = [&](int x, int y) { return M(4, x, y) + M(x, 4, y) + M(x, y, 4); };
  for (int code = threadIdx.x; code < NCODES; code += blockDim.x)
  {
    int const n = code < 4 ? 1 : code < 20 ? 2 : code < 84 ? 3 : code < 340 ? 4 : 5;
    int const off = n == 1 ? 0 : n == 2 ? 4 : n == 3 ? 20 : n == 4 ? 84 : 340;
    int z[5];
    int c = code - off;
    for (int i = n - 1; i >= 0; --i)
    {
      z[i] = c & 3;
      c >>= 2;
    }
    float v;
    if (n == 1) v = e * e * f * f / 3.0f * single(z[0]);
    else if (n == 2)
      v = 2.0f * e * f * f * f / 3.0f * pair(z[0], z[1]) +
          e * e * e * f / 3.0f * (b[z[0]] * single(z[1]) + b[z[1]] * single(z[0]));
    else if (n == 3)
      v = f * f * f * f * M(z[0], z[1], z[2]) +
          4.0f * e * e * f * f / 9.0f *
              (b[z[0]] * pair(z[1], z[2]) + b[z[1]] * pair(z[0], z[2]) + b[z[2]] * pair(z[0], z[1])) +
          e * e * e * e / 9.0f *
              (b[z[1]] * b[z[2]] * single(z[0]) + b[z[0]] * b[z[2]] * single(z[1]) + b[z[0]] * b[z[1]] * single(z[2]));
    else if (n == 4)
    {
      float one = 0.f, two = 0.f;
      for (int i = 0; i < 4; ++i)
      {
        int r[3], m = 0;
        for (int k = 0; k < 4; ++k)
          if (k != i) r[m++] = z[k];
        one += b[z[i]] * M(r[0], r[1], r[2]);
      }
      for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
        {
          int r[2], m = 0;
          for (int k = 0; k < 4; ++k)
            if (k != i && k != j) r[m++] = z[k];
          two += b[z[i]] * b[z[j]] * pair(r[0], r[1]);
        }
      v = e * f * f * f / 2.0f * one + e * e * e * f / 9.0f * two;
    }
    else
    {
      float two = 0.f;
      for (int i = 0; i < 5; ++i)
        for (int j = i + 1; j < 5; ++j)
        {
          int r[3], m = 0;
          for (int k = 0; k < 5; ++k)
            if (k != i && k != j) r[m++] = z[k];
          two += b[z[i]] * b[z[j]] * M(r[0], r[1], r[2]);
        }
      v = e * e * f * f / 10.0f * two;
    }
    a.out[(size_t)st * NCODES + code] = v > 0.f ? logf(v) : -CUDART_INF_F;
  }
}

} // namespace dcp
