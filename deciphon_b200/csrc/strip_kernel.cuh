// Score pass for profiles of more than 256 nodes: speculative column strips.
//
// score_reg_kernel<Q,W> (W > 1) pays two CTA barriers and a second delete-chain round per DP
// row because B(l) = min(N(l)+NB, E(l)+EB, J(l)+JB) (c-core/viterbi.c:495-496,582-583) needs
// E(l), a minimum over ALL nodes of the row, before any node of row l+1 can start.  But for a
// pair without a domain the E and J terms essentially never win that minimum (measured: 98 % of
// random (read, profile) pairs have B(l) == N(l)+NB on every row, median margin 4 nats), and N(l)
// depends on the read alone.
//
// So a profile is cut into W STRIPS of 32*Q nodes and every strip runs the whole window with the
// single-warp software-pipelined row, ASSUMING B(l) = N(l)+NB.  The only coupling left is
// one-directional: strip w needs M, I, D (final) of the last node of strip w-1 for the same row.
// Strip w of every pair of a kernel class is ONE LAUNCH and strip w+1 the next one on the same
// stream; a warp runs one strip of one pair and the boundary {M, I, D, running min of E} of
// every row travels through a per-pair column in global memory (16 bytes a row, fetched one row
// ahead, overwritten in place).  The running minimum of E travels with it, so the LAST strip
// sees the true E(l): it carries the J and C states and checks, row by row, that the assumption
// held.  If it did (bit-for-bit, the minimum is the same float), every value computed is the
// exact one and alt = T(L) is written; otherwise the pair is queued for the exact multi-warp
// kernel.  Results are therefore always bit-identical to the reference -- speculation only
// decides which kernel produces them.
//
// Measured alternatives (profiles/README.md): the W strips of a pair as the W warps of one CTA,
// coupled through a shared-memory ring with producer/consumer row counters (15-25 % slower:
// polling and reconvergence instructions, two strips' short-code rows competing for L1), and one
// warp walking the strips of a pair inside one kernel (register-capped schedule).
#pragma once
#include "score_kernel.cuh"

namespace dcp {

struct StripArgs
{
  ScoreArgs s;
  long long *redo;               // pairs whose speculation failed
  unsigned long long *nredo;
  // one-strip-per-launch kernels: boundary column of launch item i at col + i * col_stride
  Mail *col;
  unsigned long long col_stride;
  int strip;                     // which strip of the pair this launch runs
  unsigned long long item0;      // first item of this launch's range ([item0, s.nitems))
  // segmented profiles (below): descriptor of segment `level` of profile p at segs[seg_first[p] + level];
  // column of an item: colmap[item] (explicit pairs) or colmap[profile slot] * nseq + read (grid)
  ProfileDesc const *segs;
  int const *seg_first;
  int level;
  long long const *colmap;
};

} // namespace dcp
