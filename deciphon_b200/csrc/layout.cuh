// Device data layout of the scan hot path (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcp {

constexpr int NCODES = 1364;   // c-core/viterbi.c:13
constexpr int MAXQ_REG = 8;    // nodes per lane the register-resident kernel supports

// special-transition slots, the order of enum extr_trans_id (c-core/viterbi.h:4-19)
enum { X_RR, X_SN, X_NN, X_SB, X_NB, X_EB, X_JB, X_EJ, X_JJ, X_EC, X_CC, X_ET, X_CT, X_STRIDE = 16 };
// per-node transition rows, destination-indexed (c-core/protein.c:361-383)
enum { C_BM, C_MM, C_MI, C_MD, C_IM, C_II, C_DM, C_DD, C_ROWS };

// A profile resident in HBM.
//   em    [1364][Kpad]  match-emission costs, one row per code
//   core  [8][Kpad]     transition costs
//   nulbg [1364]        {null, background} emission costs
// Node k lives at "virtual lane" vl = k / Q, slot q = k % Q (the reference's striping,
// c-core/viterbi.c:220-221, with VL = 32*W virtual lanes instead of 8/16 SIMD lanes).
// Inside a row the Q slots of a lane are split into chunks of 4, then 2, then 1 floats
// so that a warp fetches each chunk with one coalesced 128/64/32-bit load per lane:
//   pos(vl, q) = VL*q0 + vl*w + (q - q0),  chunk [q0, q0+w) containing q.
struct ProfileDesc
{
  float const *em;
  float const *core;
  float2 const *nulbg;
  int K;
  int Q;    // nodes per lane
  int W;    // warps per pair
  int Kpad; // VL * Q
  int VL;   // virtual lanes a row is striped over: 32 * W, or 16/8/4 for profiles of at most
            // 128/64/32 nodes (two/four/eight pairs share a warp, row_kernel.cuh)
  int Kfull; // nodes of the whole profile when this describes one segment of it (strip_kernel.cuh), else K
};

__host__ __device__ inline int layout_pos(int k, int Q, int VL)
{
  int vl = k / Q, q = k - vl * Q;
  int n4 = Q & ~3;
  int q0, w;
  if (q < n4) { q0 = q & ~3; w = 4; }
  else if ((Q & 2) && q < n4 + 2) { q0 = n4; w = 2; }
  else { q0 = n4 + (Q & 2); w = 1; }
  return VL * q0 + vl * w + (q - q0);
}

inline void layout_shape(int K, int *Q, int *W, int *VL, bool subwarp = true)
{
  if (subwarp && K <= 4 * MAXQ_REG * 4)
  { // sub-warp: Q = 5..8 nodes on 4, 8 or 16 lanes (K <= 16 pads up to Q = 5 on 4 lanes)
    int vl = K <= 4 * MAXQ_REG ? 4 : K <= 8 * MAXQ_REG ? 8 : 16;
    int q = (K + vl - 1) / vl;
    if (q < 5) q = 5;
    if (q == 7) q = 8; // measured: the 4+2+1 chunk row (three loads per code) is slower than Q = 8 padded
    *W = 1;
    *VL = vl;
    *Q = q;
    return;
  }
  int w = 1;
  while (32 * w * MAXQ_REG < K) w *= 2;
  *W = w;
  *VL = 32 * w;
  *Q = (K + 32 * w - 1) / (32 * w);
  if (w == 1 && *Q == 7) *Q = 8; // same measurement (K = 224: 27.8 ms at Q = 7, 26.8 ms at Q = 8)
}

// One (window, profile) unit of work; mirrors dcpgpu_pair.
struct Pair
{
  int profile;
  int seq;
  int start;
  int len;
};

// Packed reads: 2 bits per nucleotide, 16 per word, nucleotide i of a sequence at bits
// 2*(i%16) of word seq_word[s] + i/16.
struct ReadsView
{
  uint32_t const *words;
  uint16_t const *hist;      // [16 * nwords + slack] ten-bit history ending at each position (row_kernel.cuh)
  long long const *seq_word; // [nseq] first word of each sequence
  int const *seq_len;        // [nseq]
  int nseq;
  int eight;                 // the value 8 as a run-time operand (row_kernel.cuh:mad_ptr)
  long long nwords;          // words in the buffer (bounds of the packed stream)
};

} // namespace dcp
