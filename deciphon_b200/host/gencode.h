// Genetic codes (NCBI translation tables) and the frame-state codon decoder, shared by the scan
// (amino-acid column of a match, decoder.c:38-58) and the press (model.c:390-441).
//
// The reference takes both from the third-party imm library (imm_gencode_get,
// imm_frame_cond_decode), which is not in the reference tree; the tables are NCBI's, the decoder is
// the frame-state model stated in DESIGN.md (press), pinned on the golden minifam.dcp tables by the tests.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define DCPB_HD __host__ __device__
#else
#define DCPB_HD
#endif

namespace dcpb {

// id -> 64 amino letters in TCAG x TCAG x TCAG codon order (NCBI gc.prt); nullptr: unknown id.
inline char const *gencode_table(int id)
{
  switch (id)
  {
  case 1: return "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 2: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG";
  case 3: return "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 4: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 5: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG";
  case 6: return "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 9: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
  case 10: return "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 11: return "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 12: return "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 13: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG";
  case 14: return "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
  case 15: return "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 16: return "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 21: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG";
  case 22: return "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 23: return "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 24: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG";
  case 25: return "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 26: return "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 27: return "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 28: return "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 29: return "FFLLSSSSYYYYCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 30: return "FFLLSSSSYYEECC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 31: return "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 33: return "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG";
  default: return nullptr;
  }
}

// amino letter of codon (a, b, c) given as ACGT indices
DCPB_HD inline char codon_amino(char const *table, int a, int b, int c)
{
  // A, C, G, T -> position in T, C, A, G
  int const ta = a == 0 ? 2 : a == 1 ? 1 : a == 2 ? 3 : 0, tb = b == 0 ? 2 : b == 1 ? 1 : b == 2 ? 3 : 0,
            tc = c == 0 ? 2 : c == 1 ? 1 : c == 2 ? 3 : 0;
  return table[ta * 16 + tb * 4 + tc];
}

// Base distribution and codon marginals of one state, as stored in a .dcp record (log-probs).
struct NucltDist
{
  float nuclt[4];
  float codon[125]; // [a][b][c], index 4 = any base (imm_codon_marg)
};

// p(codon, fragment z[0..n)) of a frame state with indel rate eps, given pc = p(codon) and the base
// probabilities b: the emission formula of the frame table (csrc/press_kernel.cuh) with every
// codon marginal p(pattern) replaced by p(codon) * [codon matches pattern] (third-party
// imm_frame_cond_lprob).
DCPB_HD inline double frame_joint_prob(double pc, double const b[4], double eps, int const codon[3], int const *z, int n)
{
  double const e = eps, f = 1.0 - eps;
  auto M = [&](int x, int y, int w) {
    return ((x == 4 || x == codon[0]) && (y == 4 || y == codon[1]) && (w == 4 || w == codon[2])) ? pc : 0.0;
  };
  auto single = [&](int x) { return M(x, 4, 4) + M(4, x, 4) + M(4, 4, x); };
  auto pair = [&](int x, int y) { return M(4, x, y) + M(x, 4, y) + M(x, y, 4); };
  double v = 0;
  if (n == 1) v = e * e * f * f / 3 * single(z[0]);
  else if (n == 2)
    v = 2 * e * f * f * f / 3 * pair(z[0], z[1]) + e * e * e * f / 3 * (b[z[0]] * single(z[1]) + b[z[1]] * single(z[0]));
  else if (n == 3)
    v = f * f * f * f * M(z[0], z[1], z[2]) +
        4 * e * e * f * f / 9 * (b[z[0]] * pair(z[1], z[2]) + b[z[1]] * pair(z[0], z[2]) + b[z[2]] * pair(z[0], z[1])) +
        e * e * e * e / 9 *
            (b[z[1]] * b[z[2]] * single(z[0]) + b[z[0]] * b[z[2]] * single(z[1]) + b[z[0]] * b[z[1]] * single(z[2]));
  else if (n == 4)
  {
    double one = 0, two = 0;
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
    v = e * f * f * f / 2 * one + e * e * e * f / 9 * two;
  }
  else if (n == 5)
  {
    double two = 0;
    for (int i = 0; i < 5; ++i)
      for (int j = i + 1; j < 5; ++j)
      {
        int r[3], m = 0;
        for (int k = 0; k < 5; ++k)
          if (k != i && k != j) r[m++] = z[k];
        two += b[z[i]] * b[z[j]] * M(r[0], r[1], r[2]);
      }
    v = e * e * f * f / 10 * two;
  }
  return v;
}

// Most likely codon of a 1..5-nt fragment: argmax of p(codon, fragment) over the 64 codons in
// ACGT-major order, the first maximum wins (imm_frame_cond_decode as called at decoder.c:38-58).
// False when no codon can have produced the fragment (the caller reports DCP_EDECODON like
// decoder.c:52-56).
DCPB_HD inline bool frame_decode(NucltDist const &d, double eps, int const *z, int n, int out[3])
{
  double b[4];
  for (int i = 0; i < 4; ++i) b[i] = exp((double)d.nuclt[i]);
  double best = 0.0;
  bool found = false;
  for (int a = 0; a < 4; ++a)
    for (int bb = 0; bb < 4; ++bb)
      for (int c = 0; c < 4; ++c)
      {
        double const lp = d.codon[a * 25 + bb * 5 + c];
        if (!(lp > -INFINITY)) continue;
        int const codon[3] = {a, bb, c};
        double const v = frame_joint_prob(exp(lp), b, eps, codon, z, n);
        if (v > best)
        {
          best = v;
          out[0] = a;
          out[1] = bb;
          out[2] = c;
          found = true;
        }
      }
  return found;
}

} // namespace dcpb
