// Post-processing of traced paths on the device: hit extent, fragment -> codon -> amino acid and the
// bytes of the "match" column of a products.tsv row.
//
// What c-core does per hit on the host after trellis_unzip:
//   thread.c:130-166   hit extent: first B .. last E of the path, window-relative positions
//   match.c:66-90,     per step "<fragment>,<state name>,<codon>,<amino>", steps joined by ';'
//   product_thread.c:112-148
//   decoder.c:38-58    codon of a fragment = imm_frame_cond_decode over the state's base
//                      distribution and codon marginals (M_k: node k, I_k: background, N/J/C: null)
//   state.c:47-90      state names
// Here: one warp per traced pair, two passes (sizes, then bytes) around a host prefix sum over
// the pairs; the decoded codon and amino letter of every step are cached between the passes.
#pragma once
#include "../host/gencode.h"
#include "layout.cuh"

namespace dcp {

constexpr int MATCH_WARPS = 4;
constexpr int DIST_FLOATS = 129; // 4 base log-probs + 125 codon marginal log-probs

struct DecoderDesc
{
  float const *dists; // [K + 2][129]: nodes 0..K-1, then null, then background; nullptr: none uploaded
  char gencode[64];   // amino letters in TCAG order
};

struct MatchArgs
{
  Pair const *pairs;
  long long npairs;
  int const *nsteps;
  long long const *step_off; // [npairs + 1] compact path layout
  uint16_t const *ids;
  uint8_t const *sizes;
  ProfileDesc const *profiles;
  DecoderDesc const *decoders;
  ReadsView reads;
  double eps;
  int is_rna;
  int extent_only; // stop after the hit extents (no decode tables needed)
  // per pair
  int *hit;        // 1: the path has a B..E segment
  int *hit_start;  // window-relative (thread.c:140-158)
  int *hit_stop;
  int *seg_begin;  // step index of the first B
  int *seg_end;    // one past the last E
  long long *text_len;
  long long const *text_off; // pass 2
  // per step
  uint8_t *codon; // a*16 + b*4 + c (ACGT indices); 255: mute step
  char *amino;
  char *text;
  int *bad;       // a fragment no codon can have produced (DCP_EDECODON)
};

enum { MST_M = 0 << 14, MST_I = 1 << 14, MST_D = 2 << 14, MST_X = 3 << 14, MST_S = MST_X | 3, MST_N = MST_X | 4,
       MST_B = MST_X | 5, MST_E = MST_X | 6, MST_J = MST_X | 7, MST_C = MST_X | 8, MST_T = MST_X | 9 };

__device__ __forceinline__ bool match_mute(int id)
{ // state.c:17-23
  int const msb = id & (3 << 14);
  if (msb == MST_X) return id == MST_S || id == MST_B || id == MST_E || id == MST_T;
  return msb == MST_D;
}

// characters of a state name (state.c:47-90): one letter, plus the node number for core states
__device__ __forceinline__ int match_name_len(int id)
{
  if ((id & (3 << 14)) == MST_X) return 1;
  int const k = id & 0x3fff;
  return 1 + (k >= 10000 ? 5 : k >= 1000 ? 4 : k >= 100 ? 3 : k >= 10 ? 2 : 1);
}

__device__ __forceinline__ int match_name_put(int id, char *out)
{
  int const msb = id & (3 << 14);
  if (msb == MST_X)
  {
    char const names[] = "FRGSNBEJCT";
    int const i = id & 0x3fff;
    out[0] = i <= 9 ? names[i] : '?';
    return 1;
  }
  out[0] = msb == MST_M ? 'M' : msb == MST_I ? 'I' : 'D';
  int k = id & 0x3fff, n = match_name_len(id) - 1;
  for (int i = n; i >= 1; --i)
  {
    out[i] = (char)('0' + k % 10);
    k /= 10;
  }
  return n + 1;
}

__device__ __forceinline__ int read_nt(ReadsView const &r, long long first_word, int pos)
{
  return (int)((__ldg(r.words + first_word + (pos >> 4)) >> (2 * (pos & 15))) & 3u);
}

// warp-wide inclusive scan of an int
__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    int const t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Pass 1 (TEXT = false): extents, per-step codon/amino, text length of every pair.
// Pass 2 (TEXT = true): the bytes, at text_off[pair].
template <bool TEXT>
__global__ void __launch_bounds__(32 * MATCH_WARPS) match_kernel(MatchArgs a)
{
  long long const pi = blockIdx.x * (long long)MATCH_WARPS + (threadIdx.x >> 5);
  int const lane = threadIdx.x & 31;
  if (pi >= a.npairs) return;
  Pair const pr = a.pairs[pi];
  int const ns = a.nsteps[pi];
  uint16_t const *ids = a.ids + a.step_off[pi];
  uint8_t const *szs = a.sizes + a.step_off[pi];
  long long const w0 = a.reads.seq_word[pr.seq];

  int b, end, hit_start;
  if constexpr (!TEXT)
  {
    // first B and last E at or after it, with the window position in front of each (thread.c:130-166)
    int first_b = ns, pos_b = 0, last_e = -1, pos_e = 0, base = 0;
    for (int j0 = 0; j0 < ns; j0 += 32)
    {
      int const j = j0 + lane;
      int const sz = j < ns ? szs[j] : 0;
      int const id = j < ns ? ids[j] : 0;
      int const incl = warp_incl_scan(sz, lane);
      int const pos = base + incl - sz; // position in front of step j
      unsigned const mb = __ballot_sync(0xffffffffu, j < ns && id == MST_B);
      if (first_b == ns && mb)
      {
        int const l = __ffs(mb) - 1;
        first_b = j0 + l;
        pos_b = __shfl_sync(0xffffffffu, pos, l);
      }
      unsigned me = __ballot_sync(0xffffffffu, j < ns && id == MST_E && j >= first_b);
      if (me)
      {
        int const l = 31 - __clz(me);
        last_e = j0 + l;
        pos_e = __shfl_sync(0xffffffffu, pos, l);
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    bool const hit = first_b < ns && last_e >= 0;
    b = first_b;
    end = last_e + 1;
    hit_start = pos_b;
    if (lane == 0)
    {
      a.hit[pi] = hit ? 1 : 0;
      a.hit_start[pi] = pos_b;
      a.hit_stop[pi] = pos_e;
      a.seg_begin[pi] = b;
      a.seg_end[pi] = end;
      if (!hit || a.extent_only) a.text_len[pi] = 0;
    }
    if (!hit || a.extent_only) return;
  }
  else
  {
    if (!a.hit[pi]) return;
    b = a.seg_begin[pi];
    end = a.seg_end[pi];
    hit_start = a.hit_start[pi];
  }

  DecoderDesc const &dec = a.decoders[pr.profile];
  int const K = a.profiles[pr.profile].Kfull;
  char const *sym = a.is_rna ? "ACGU" : "ACGT";
  char *const out = TEXT ? a.text + a.text_off[pi] : nullptr;
  long long total = 0;
  int pos_base = hit_start;
  for (int j0 = b; j0 < end; j0 += 32)
  {
    int const j = j0 + lane;
    bool const on = j < end;
    int const sz = on ? szs[j] : 0;
    int const id = on ? ids[j] : 0;
    int const incl = warp_incl_scan(sz, lane);
    int const pos = pos_base + incl - sz;
    bool const mute = match_mute(id);
    int len = 0;
    if (on) len = (j > b ? 1 : 0) + sz + 1 + match_name_len(id) + 1 + (mute ? 1 : 5);
    if constexpr (!TEXT)
    {
      if (on)
      {
        uint8_t cod = 255;
        char am = 0;
        if (!mute)
        {
          int z[5];
          for (int t = 0; t < sz; ++t) z[t] = read_nt(a.reads, w0, pr.start + pos + t);
          int const msb = id & (3 << 14);
          int const row = msb == MST_I ? K + 1 : msb == MST_M ? (id & 0x3fff) - 1 : K; // decoder.c:43-49
          dcpb::NucltDist const &nd = *reinterpret_cast<dcpb::NucltDist const *>(dec.dists + (size_t)row * DIST_FLOATS);
          int c3[3];
          if (dec.dists && dcpb::frame_decode(nd, a.eps, z, sz, c3))
          {
            cod = (uint8_t)(c3[0] * 16 + c3[1] * 4 + c3[2]);
            am = dcpb::codon_amino(dec.gencode, c3[0], c3[1], c3[2]);
          }
          else
            atomicExch(a.bad, 1);
        }
        a.codon[a.step_off[pi] + j] = cod;
        a.amino[a.step_off[pi] + j] = am;
      }
    }
    int const lincl = warp_incl_scan(len, lane);
    if constexpr (TEXT)
    {
      if (on)
      {
        char *p = out + total + lincl - len;
        if (j > b) *p++ = ';';
        for (int t = 0; t < sz; ++t) *p++ = sym[read_nt(a.reads, w0, pr.start + pos + t)];
        *p++ = ',';
        p += match_name_put(id, p);
        *p++ = ',';
        if (!mute)
        {
          uint8_t const cod = a.codon[a.step_off[pi] + j];
          *p++ = sym[(cod >> 4) & 3];
          *p++ = sym[(cod >> 2) & 3];
          *p++ = sym[cod & 3];
          *p++ = ',';
          *p++ = a.amino[a.step_off[pi] + j];
        }
        else
          *p++ = ',';
      }
    }
    total += __shfl_sync(0xffffffffu, lincl, 31); // bytes of this chunk of steps (uniform across the warp)
    pos_base += __shfl_sync(0xffffffffu, incl, 31);
  }
  if constexpr (!TEXT)
  {
    if (lane == 0) a.text_len[pi] = total;
  }
}

} // namespace dcp
