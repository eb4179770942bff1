// deciphon_b200.cpp -- the reference's scan API (include/deciphon_b200.h) over the GPU C ABI.
//
// Host side of the drop-in: what c-core does per thread, per profile, per sequence, per
// window (scan.c:167-216 -> thread.c:49-208) is done here as WAVES of batched GPU passes:
//   wave w = the w-th window.c window of every still-active (sequence, profile) pair:
//     score pass (viterbi_null + viterbi_cost)  -> lrt gate (thread.c:119-121)
//     trace pass (viterbi_path + trellis_unzip) -> hit extent (thread.c:130-166)
//     -> window_set_last_hit_position -> next window (window.c:13-37)
// Everything that is not arithmetic on the DP -- .dcp parsing, sequence clean-up, windowing,
// hit extents, row formatting -- follows the reference's semantics and cites it.
#include "../../include/deciphon_b200.h"
#include "../../include/dcpgpu.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace {

constexpr int NCODES = DCPGPU_NUM_CODES;

// state ids, c-core/state.h:7-25
enum : int
{
  ST_M = 0 << 14, ST_I = 1 << 14, ST_D = 2 << 14, ST_X = 3 << 14,
  ST_S = ST_X | 3, ST_N = ST_X | 4, ST_B = ST_X | 5, ST_E = ST_X | 6,
  ST_J = ST_X | 7, ST_C = ST_X | 8, ST_T = ST_X | 9,
};

bool state_is_mute(int id)
{ // c-core/state.c:17-23
  int const msb = id & (3 << 14);
  if (msb == ST_X) return id == ST_S || id == ST_B || id == ST_E || id == ST_T;
  return msb == ST_D;
}

void state_name(int id, char *out)
{ // c-core/state.c:47-90
  int const msb = id & (3 << 14);
  if (msb == ST_X)
  {
    static char const names[] = "FRGSNBEJCT";
    int const i = id & 0x3fff;
    out[0] = i <= 9 ? names[i] : '?';
    out[1] = 0;
    return;
  }
  snprintf(out, 16, "%c%d", msb == ST_M ? 'M' : msb == ST_I ? 'I' : 'D', id & 0x3fff);
}

int map_gpu_error(int rc)
{
  switch (rc)
  {
  case DCPGPU_OK: return 0;
  case DCPGPU_ENODEVICE: return DCP_EGPUNODEVICE;
  case DCPGPU_ECUDA: return DCP_EGPUFAIL;
  case DCPGPU_ENOMEM: return DCP_EGPUNOMEM;
  default: return DCP_EGPUINTERNAL;
  }
}

// ---- .dcp reader (database_reader.c:26-80, protein.c:283-351; layout in SURVEY App. A.7) ----

struct Reader
{
  std::vector<unsigned char> buf;
  size_t p = 0;
  bool ok = true;

  bool need(size_t n)
  {
    if (p + n > buf.size()) ok = false;
    return ok;
  }
  uint64_t be(int n)
  {
    if (!need((size_t)n)) return 0;
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | buf[p++];
    return v;
  }
  int peek() { return need(1) ? buf[p] : -1; }
  bool map(uint32_t *n)
  {
    int const b = peek();
    if (b < 0) return false;
    if ((b & 0xf0) == 0x80) { *n = b & 15; ++p; return true; }
    if (b == 0xde) { ++p; *n = (uint32_t)be(2); return ok; }
    if (b == 0xdf) { ++p; *n = (uint32_t)be(4); return ok; }
    return ok = false;
  }
  bool array(uint32_t *n)
  {
    int const b = peek();
    if (b < 0) return false;
    if ((b & 0xf0) == 0x90) { *n = b & 15; ++p; return true; }
    if (b == 0xdc) { ++p; *n = (uint32_t)be(2); return ok; }
    if (b == 0xdd) { ++p; *n = (uint32_t)be(4); return ok; }
    return ok = false;
  }
  bool str(std::string *s)
  {
    int const b = peek();
    if (b < 0) return false;
    size_t n;
    if ((b & 0xe0) == 0xa0) { n = b & 31; ++p; }
    else if (b == 0xd9) { ++p; n = be(1); }
    else if (b == 0xda) { ++p; n = be(2); }
    else if (b == 0xdb) { ++p; n = be(4); }
    else return ok = false;
    if (!need(n)) return false;
    s->assign(reinterpret_cast<char const *>(&buf[p]), n);
    p += n;
    return true;
  }
  bool key(char const *want)
  {
    std::string s;
    return str(&s) && (s == want || (ok = false));
  }
  bool integer(int64_t *v)
  {
    int const b = peek();
    if (b < 0) return false;
    ++p;
    if (b <= 0x7f) { *v = b; return true; }
    if (b >= 0xe0) { *v = b - 256; return true; }
    switch (b)
    {
    case 0xcc: *v = (int64_t)be(1); return ok;
    case 0xcd: *v = (int64_t)be(2); return ok;
    case 0xce: *v = (int64_t)be(4); return ok;
    case 0xcf: *v = (int64_t)be(8); return ok;
    case 0xd0: *v = (int8_t)be(1); return ok;
    case 0xd1: *v = (int16_t)be(2); return ok;
    case 0xd2: *v = (int32_t)be(4); return ok;
    case 0xd3: *v = (int64_t)be(8); return ok;
    default: return ok = false;
    }
  }
  bool boolean(bool *v)
  {
    int const b = peek();
    if (b != 0xc2 && b != 0xc3) return ok = false;
    ++p;
    *v = b == 0xc3;
    return true;
  }
  bool f32(float *v)
  {
    int const b = peek();
    if (b == 0xca)
    {
      ++p;
      uint32_t u = (uint32_t)be(4);
      memcpy(v, &u, 4);
      return ok;
    }
    if (b == 0xcb)
    {
      ++p;
      uint64_t u = be(8);
      double d;
      memcpy(&d, &u, 8);
      *v = (float)d;
      return ok;
    }
    return ok = false;
  }
  // bin (current writer, host-endian, write.c:59-66) or ext (golden file, big-endian)
  bool blob(bool *big_endian, size_t *n)
  {
    int const b = peek();
    if (b < 0) return false;
    ++p;
    if (b >= 0xc4 && b <= 0xc6) { *n = be(1 << (b - 0xc4)); *big_endian = false; return need(*n); }
    if (b >= 0xc7 && b <= 0xc9) { *n = be(1 << (b - 0xc7)); be(1); *big_endian = true; return need(*n); }
    if (b >= 0xd4 && b <= 0xd8) { *n = (size_t)1 << (b - 0xd4); be(1); *big_endian = true; return need(*n); }
    return ok = false;
  }
  bool f32array(size_t count, float *out)
  {
    int const b = peek();
    if (b < 0) return false;
    if ((b & 0xf0) == 0x90 || b == 0xdc || b == 0xdd)
    {
      uint32_t n;
      if (!array(&n) || n != count) return ok = false;
      for (uint32_t i = 0; i < n; ++i)
        if (!f32(out + i)) return false;
      return true;
    }
    bool be_;
    size_t n;
    if (!blob(&be_, &n) || n != count * 4) return ok = false;
    if (be_)
      for (size_t i = 0; i < count; ++i)
      {
        uint32_t u = ((uint32_t)buf[p] << 24) | ((uint32_t)buf[p + 1] << 16) | ((uint32_t)buf[p + 2] << 8) | buf[p + 3];
        memcpy(out + i, &u, 4);
        p += 4;
      }
    else
    {
      memcpy(out, &buf[p], n);
      p += n;
    }
    return true;
  }
  // skip any value (alphabet sub-maps whose encoding lives in third-party imm)
  bool skip()
  {
    int const b = peek();
    if (b < 0) return false;
    uint32_t n;
    if ((b & 0xf0) == 0x80 || b == 0xde || b == 0xdf)
    {
      if (!map(&n)) return false;
      for (uint32_t i = 0; i < 2 * n; ++i)
        if (!skip()) return false;
      return true;
    }
    if ((b & 0xf0) == 0x90 || b == 0xdc || b == 0xdd)
    {
      if (!array(&n)) return false;
      for (uint32_t i = 0; i < n; ++i)
        if (!skip()) return false;
      return true;
    }
    if ((b & 0xe0) == 0xa0 || (b >= 0xd9 && b <= 0xdb)) { std::string s; return str(&s); }
    if (b == 0xc0 || b == 0xc2 || b == 0xc3) { ++p; return true; }
    if (b == 0xca || b == 0xcb) { float f; return f32(&f); }
    if ((b >= 0xc4 && b <= 0xc9) || (b >= 0xd4 && b <= 0xd8)) { bool e; size_t m; if (!blob(&e, &m)) return false; p += m; return true; }
    int64_t v;
    return integer(&v);
  }
};

struct NucltDist
{
  float nuclt[4];
  float codon[125]; // [a][b][c], index 4 = any (imm_codon_marg)
};

struct ProfileMeta
{
  std::string accession;
  int gencode = 1;
  int K = 0;
  NucltDist null_dist, bg_dist;
  std::vector<NucltDist> nodes; // K
};

bool read_nuclt_dist(Reader &r, NucltDist *d)
{ // nuclt_dist.c:13-20; golden encoding: array(2){f32[4], f32[125]}
  int const b = r.peek();
  uint32_t n;
  if ((b & 0xf0) == 0x90)
  {
    if (!r.array(&n) || n != 2) return r.ok = false;
    return r.f32array(4, d->nuclt) && r.f32array(125, d->codon);
  }
  // unknown (imm-defined) encoding: walk it and pick the two float arrays by size
  size_t const start = r.p;
  if (!r.skip()) return false;
  size_t const end = r.p;
  bool got4 = false, got125 = false;
  for (size_t q = start; q < end && !(got4 && got125); ++q)
  {
    Reader t;
    t.buf.assign(r.buf.begin() + (long)q, r.buf.begin() + (long)end);
    float tmp[125];
    if (!got4 && t.f32array(4, tmp)) { memcpy(d->nuclt, tmp, 16); got4 = true; q += t.p - 1; continue; }
    t.p = 0; t.ok = true;
    if (!got125 && t.f32array(125, tmp)) { memcpy(d->codon, tmp, 500); got125 = true; q += t.p - 1; }
  }
  return r.ok = got4 && got125;
}

} // namespace

// ---- public objects --------------------------------------------------------------------------

struct Sequence
{
  long id;
  std::string name;
  std::string data; // upper-cased, disambiguated (sequence.c:29-36)
};

struct dcp_batch
{
  std::vector<Sequence> seqs;
};

struct dcp_press
{
  int unused;
};

struct dcp_scan
{
  dcpgpu_ctx *gpu = nullptr;
  bool multi_hits = true, hmmer3_compat = false;
  float epsilon = 0.01f;
  std::string abc_name = "dna";
  bool is_rna = false;
  std::vector<ProfileMeta> profiles;
  void (*callback)(void *) = nullptr;
  void *userdata = nullptr;
  std::atomic<bool> interrupted{false};
  std::atomic<int> done_proteins{0};
};

namespace {

// ---- sequence clean-up: uppercase + disambiguate (sequence.c:29-36, disambiguate.c:37-86) ----

int disambiguate(std::string &s)
{
  size_t count[5] = {0, 0, 0, 0, 0}; // A C G T U
  for (char &c : s)
  {
    if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
    if (c == 'A') count[0]++;
    if (c == 'C') count[1]++;
    if (c == 'G') count[2]++;
    if (c == 'T') count[3]++;
    if (c == 'U') count[4]++;
  }
  if (count[3] > 0 && count[4] > 0) return DCP_ENUCLTSEQTU;
  static char const letters[] = "ACGTU";
  auto best = [&](std::initializer_list<int> idx) {
    int bi = *idx.begin();
    for (int i : idx)
      if (count[i] > count[bi]) bi = i; // first listed wins ties (disambiguate.c:22-35)
    return letters[bi];
  };
  for (char &c : s)
  {
    switch (c)
    {
    case 'R': c = best({0, 2}); break;
    case 'Y': c = best({1, 3}); break;
    case 'M': c = best({0, 1}); break;
    case 'K': c = best({2, 3}); break;
    case 'S': c = best({1, 2}); break;
    case 'W': c = best({0, 3}); break;
    case 'H': c = best({0, 1, 3}); break;
    case 'B': c = best({1, 2, 3}); break;
    case 'V': c = best({0, 1, 2}); break;
    case 'D': c = best({0, 2, 3}); break;
    case 'N': c = best({0, 1, 2, 3}); break;
    case 'X': c = best({0, 1, 2, 3}); break;
    default: break;
    }
  }
  return 0;
}

// ---- genetic code (third-party imm_gencode_decode, pinned by the golden rows) ----------------

char const *gencode_table(int id)
{ // NCBI translation tables, codon order TCAG x TCAG x TCAG
  switch (id)
  {
  case 1: case 11: return "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  case 4: return "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  default: return nullptr;
  }
}

char codon_amino(char const *table, int a, int b, int c)
{ // a,b,c in ACGT order -> TCAG order
  static int const tcag[4] = {2, 1, 3, 0};
  return table[tcag[a] * 16 + tcag[b] * 4 + tcag[c]];
}

// Fragment (1..5 nt) -> most likely codon.  The reference calls the third-party
// imm_frame_cond_decode (decoder.c:38-58), whose body is not in the reference tree.
// PARITY UNPINNED for fragments that are not an admissible 3-mer: this is
// the maximum-probability codon among those reachable with the FEWEST indel events (the
// leading-order term of the frame model); a 3-nt fragment with non-zero codon probability
// decodes to itself, which is what every golden row exercises.
bool decode_codon(NucltDist const &d, int const *z, int n, int out[3])
{
  auto lp = [&](int a, int b, int c) { return d.codon[a * 25 + b * 5 + c]; };
  float best = -INFINITY;
  bool found = false;
  auto consider = [&](int a, int b, int c) {
    float const v = lp(a, b, c);
    if (v > best || !found) { if (v > best || !found) { best = v; out[0] = a; out[1] = b; out[2] = c; found = true; } }
  };
  if (n == 3)
  {
    if (std::isfinite(lp(z[0], z[1], z[2]))) { out[0] = z[0]; out[1] = z[1]; out[2] = z[2]; return true; }
    for (int pos = 0; pos < 3; ++pos)
      for (int x = 0; x < 4; ++x)
      {
        int c[3] = {z[0], z[1], z[2]};
        c[pos] = x;
        consider(c[0], c[1], c[2]);
      }
  }
  else if (n == 2)
  {
    for (int x = 0; x < 4; ++x) { consider(x, z[0], z[1]); consider(z[0], x, z[1]); consider(z[0], z[1], x); }
  }
  else if (n == 1)
  {
    for (int x = 0; x < 4; ++x)
      for (int y = 0; y < 4; ++y) { consider(z[0], x, y); consider(x, z[0], y); consider(x, y, z[0]); }
  }
  else if (n == 4)
  {
    for (int skip = 0; skip < 4; ++skip)
    {
      int c[3], m = 0;
      for (int i = 0; i < 4; ++i) if (i != skip) c[m++] = z[i];
      consider(c[0], c[1], c[2]);
    }
  }
  else if (n == 5)
  {
    for (int s1 = 0; s1 < 5; ++s1)
      for (int s2 = s1 + 1; s2 < 5; ++s2)
      {
        int c[3], m = 0;
        for (int i = 0; i < 5; ++i) if (i != s1 && i != s2) c[m++] = z[i];
        consider(c[0], c[1], c[2]);
      }
  }
  return found && std::isfinite(best);
}

int load_file(char const *path, std::vector<unsigned char> *out)
{
  FILE *fp = fopen(path, "rb");
  if (!fp) return DCP_EOPENDB;
  fseek(fp, 0, SEEK_END);
  long const n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (n < 0) { fclose(fp); return DCP_EFREAD; }
  out->resize((size_t)n);
  size_t const got = n ? fread(out->data(), 1, (size_t)n, fp) : 0;
  fclose(fp);
  return got == (size_t)n ? 0 : DCP_EFREAD;
}

int mkdir_p(std::string const &dir)
{
  if (mkdir(dir.c_str(), 0755) == 0 || errno == EEXIST) return 0;
  return DCP_EMKDIR;
}

struct Row
{
  int profile;
  int seq_order;
  int window;
  std::string text;
};

struct Active
{ // window iteration state of one (profile, sequence), window.c:7-11
  int profile, seq;
  int start, stop, idx, last_hit;
};

bool window_next(Active &w, int seq_len, int K)
{ // window.c:13-37
  if (w.stop == seq_len) return false;
  int const stop_miss = w.stop + 1;
  int start_miss = std::max(w.start + 1, w.start + w.last_hit + 1);
  start_miss = std::max(start_miss, stop_miss - K * 4);
  w.start = start_miss;
  w.stop = std::min(start_miss + std::min(K * 50, 100000), seq_len);
  w.idx += 1;
  return true;
}

struct DbHeader
{
  float epsilon = 0.01f;
  bool is_rna = false;
  bool has_ga = false;
  int entry_dist = 0;
  uint32_t num_proteins = 0;
};

// Streams a .dcp file: header (database_reader.c:26-80) then one callback per protein record
// (protein.c:283-351).  Arrays handed to the callback are in .dcp (log-prob) form.
template <class F>
int parse_db(char const *dbfile, DbHeader *hdr, F &&on_protein)
{
  Reader r;
  int rc = load_file(dbfile, &r.buf);
  if (rc) return rc;
  uint32_t n;
  int64_t iv;
  if (!r.map(&n) || n != 2 || !r.key("header") || !r.map(&n) || n != 8) return DCP_ENOTDBFILE;
  if (!r.key("magic_number") || !r.integer(&iv) || iv != 0xC6F1) return DCP_ENOTDBFILE;
  if (!r.key("version") || !r.integer(&iv)) return DCP_EFDATA;
  if (iv != 1) return DCP_EDBVERSION;
  if (!r.key("entry_dist") || !r.integer(&iv)) return DCP_EFDATA;
  hdr->entry_dist = (int)iv;
  if (!r.key("epsilon") || !r.f32(&hdr->epsilon)) return DCP_EFDATA;
  if (!r.key("abc")) return DCP_EFDATA;
  {
    // alphabet: only the symbols matter here (imm_abc_unpack is third-party)
    size_t const start = r.p;
    if (!r.skip()) return DCP_EFDATA;
    std::string blob(reinterpret_cast<char const *>(&r.buf[start]), r.p - start);
    hdr->is_rna = blob.find("ACGU") != std::string::npos;
    if (blob.find("ACGT") == std::string::npos && !hdr->is_rna) return DCP_ENUCLTNOSUPPORT; // scan.c:107-108
  }
  if (!r.key("amino") || !r.skip()) return DCP_EFDATA;
  if (!r.key("has_ga") || !r.boolean(&hdr->has_ga)) return DCP_EFDATA;
  if (!r.key("protein_sizes") || !r.skip()) return DCP_EFDATA;
  if (!r.key("proteins") || !r.array(&hdr->num_proteins)) return DCP_EFDATA;

  std::vector<float> nul(NCODES), bg(NCODES), emission, trans, bmk;
  uint32_t seen = 0;
  for (uint32_t pi = 0; pi < hdr->num_proteins; ++pi)
  {
    ProfileMeta m;
    std::string consensus;
    if (!r.map(&n) || n != 10) return r.ok ? DCP_EFDATA : DCP_EENDOFFILE;
    if (!r.key("accession") || !r.str(&m.accession)) return DCP_EFDATA;
    if (m.accession.size() >= 32) return DCP_ELONGACCESSION;
    if (!r.key("gencode") || !r.integer(&iv)) return DCP_EFDATA;
    m.gencode = (int)iv;
    if (!gencode_table(m.gencode)) return DCP_EGENCODEID;
    if (!r.key("consensus") || !r.str(&consensus)) return DCP_EFDATA;
    if (!r.key("core_size") || !r.integer(&iv)) return DCP_EFDATA;
    if (iv < 1 || iv > DCPGPU_MAX_CORE_SIZE) return DCP_ELARGECORESIZE;
    int const K = m.K = (int)iv;
    if (!r.key("null_nuclt_dist") || !read_nuclt_dist(r, &m.null_dist)) return DCP_EFDATA;
    if (!r.key("null_emission") || !r.f32array(NCODES, nul.data())) return DCP_EFDATA;
    if (!r.key("bg_nuclt_dist") || !read_nuclt_dist(r, &m.bg_dist)) return DCP_EFDATA;
    if (!r.key("bg_emission") || !r.f32array(NCODES, bg.data())) return DCP_EFDATA;
    if (!r.key("nodes") || !r.map(&n) || n != (uint32_t)(K + 1) * 3) return DCP_EFDATA;
    emission.resize((size_t)(K + 1) * NCODES);
    trans.resize((size_t)(K + 1) * 7);
    m.nodes.resize((size_t)K + 1);
    for (int i = 0; i <= K; ++i)
    {
      if (!r.key("nuclt_dist") || !read_nuclt_dist(r, &m.nodes[(size_t)i])) return DCP_EFDATA;
      if (!r.key("trans") || !r.f32array(7, &trans[(size_t)i * 7])) return DCP_EFDATA;
      if (!r.key("emission") || !r.f32array(NCODES, &emission[(size_t)i * NCODES])) return DCP_EFDATA;
    }
    bmk.resize((size_t)K);
    if (!r.key("BMk") || !r.f32array((size_t)K, bmk.data())) return r.ok ? DCP_EFDATA : DCP_EENDOFFILE;
    m.nodes.resize((size_t)K);
    if ((rc = on_protein(*hdr, m, emission.data(), trans.data(), bmk.data(), nul.data(), bg.data()))) return rc;
    ++seen;
  }
  if (!r.ok) return DCP_EENDOFFILE;
  if (seen != hdr->num_proteins) return DCP_EINVALNUMPROTEINS;
  return 0;
}

} // namespace

extern "C" {

char const *dcp_error_string(int code)
{
  static std::map<int, char const *> const msg = {
      {1, "alphabets differ"}, {2, "could not close file"}, {3, "file holds invalid data"},
      {4, "could not re-open file"}, {5, "could not read file"}, {6, "could not seek in file"},
      {7, "could not query file position"}, {8, "function used incorrectly (or not supported by this build)"},
      {9, "could not write file"}, {10, "could not obtain file path"}, {11, "sequence of length zero"},
      {12, "model of length zero"}, {13, "zero partitions"}, {14, "could not decode fragment into a codon"},
      {15, "model too large"}, {16, "protein too large"}, {17, "could not read HMMER3 profile"},
      {18, "too many partitions"}, {19, "too many transitions"}, {20, "out of memory"},
      {21, "could not open database file"}, {22, "could not open HMM file"}, {23, "could not open temporary file"},
      {24, "file path was truncated"}, {25, "could not unpack DP"}, {26, "could not pack DP"},
      {27, "could not unpack nucleotide distribution"}, {28, "could not pack nucleotide distribution"},
      {29, "could not set transition"}, {30, "could not add state"}, {31, "could not reset DP"},
      {32, "could not stat file"}, {33, "could not open file"}, {34, "file too large"}, {35, "path too long"},
      {36, "could not reset task"}, {37, "could not create task"}, {38, "could not set up task"},
      {39, "could not write product"}, {40, "invalid partition"}, {41, "accession too long"},
      {42, "too many threads"}, {43, "could not create temporary file"}, {44, "could not flush file"},
      {45, "could not create directory"}, {46, "wrong file format"}, {47, "could not remove directory"},
      {48, "could not remove file"}, {49, "genetic code must be set first"}, {50, "invalid genetic code id"},
      {51, "could not dial the HMMER daemon (this build runs without it: use port <= 0)"},
      {52, "could not submit task to the HMMER daemon"}, {53, "could not fetch task from the HMMER daemon"},
      {54, "could not pack HMMER result"}, {55, "HMMER daemon retry limit reached"},
      {56, "could not warm up the HMMER daemon"}, {57, "sequence letter is neither DNA nor RNA"},
      {58, "could not open file descriptor"}, {59, "could not make temporary file"}, {60, "alphabet name too long"},
      {61, "consensus too long"}, {62, "HMMER daemon not dialed"}, {63, "too many core nodes"},
      {64, "invalid state"}, {65, "invalid size"}, {66, "unexpected end of file"}, {67, "unexpected end of nodes"},
      {68, "database version not supported"}, {69, "not a deciphon database file"}, {70, "invalid state id"},
      {71, "nucleotide alphabet must be DNA or RNA"}, {72, "database is DNA, sequence is RNA"},
      {73, "database is RNA, sequence is DNA"}, {74, "sequence mixes T and U"}, {75, "no hit found"},
      {76, "could not open file"}, {77, "could not close file"}, {78, "could not duplicate descriptor"},
      {79, "too many proteins"}, {80, "invalid number of proteins"},
      {81, "no CUDA device available (deciphon_b200 has no CPU fallback)"}, {82, "CUDA call failed"},
      {83, "GPU memory exhausted"}, {84, "internal error in the GPU layer"}};
  auto it = msg.find(code);
  if (it != msg.end()) return it->second;
  static thread_local char unknown[40];
  snprintf(unknown, sizeof unknown, "unknown error #%d", code);
  return unknown;
}

// ---- batch (batch.c:14-58, sequence.c:15-45) ---------------------------------------------------

struct dcp_batch *dcp_batch_new(void) { return new (std::nothrow) dcp_batch; }

void dcp_batch_del(struct dcp_batch *x) { delete x; }

int dcp_batch_add(struct dcp_batch *x, long id, char const *name, char const *data)
{
  if (!x || !name || !data) return DCP_EFUNCUSE;
  Sequence s{id, name, data};
  int const rc = disambiguate(s.data);
  if (rc) return rc;
  x->seqs.push_back(std::move(s));
  return 0;
}

void dcp_batch_reset(struct dcp_batch *x)
{
  if (x) x->seqs.clear();
}

// ---- press: outside the hot path ---------------------------------------------------------------

struct dcp_press *dcp_press_new(void) { return new (std::nothrow) dcp_press{0}; }
int dcp_press_setup(struct dcp_press *, int, float) { return DCP_EFUNCUSE; }
int dcp_press_open(struct dcp_press *, char const *, char const *) { return DCP_EFUNCUSE; }
long dcp_press_nproteins(struct dcp_press const *) { return 0; }
int dcp_press_next(struct dcp_press *) { return DCP_EFUNCUSE; }
bool dcp_press_end(struct dcp_press const *) { return true; }
int dcp_press_close(struct dcp_press *) { return DCP_EFUNCUSE; }
void dcp_press_del(struct dcp_press const *x) { delete x; }

// ---- scan ---------------------------------------------------------------------------------------

struct dcp_scan *dcp_scan_new(void) { return new (std::nothrow) dcp_scan; }

void dcp_scan_del(struct dcp_scan const *scan)
{
  dcp_scan *x = const_cast<dcp_scan *>(scan);
  if (!x) return;
  if (x->gpu) dcpgpu_close(x->gpu);
  delete x;
}

int dcp_scan_setup(struct dcp_scan *x, char const *dbfile, int port, int num_threads, bool multi_hits,
                   bool hmmer3_compat, bool cache, void (*callback)(void *), void *userdata)
{
  (void)cache;
  if (!x || !dbfile) return DCP_EFUNCUSE;
  if (num_threads > 128) return DCP_EMANYTHREADS; // THREAD_MAX, thread.h:7 / scan.c:95
  if (port > 0) return DCP_EH3CDIAL;              // see include/deciphon_b200.h
  x->multi_hits = multi_hits;
  x->hmmer3_compat = hmmer3_compat;
  x->callback = callback;
  x->userdata = userdata;
  x->profiles.clear();
  if (x->gpu) { dcpgpu_close(x->gpu); x->gpu = nullptr; }

  DbHeader hdr;
  int rc = parse_db(dbfile, &hdr, [&](DbHeader const &h, ProfileMeta &m, float const *emission, float const *trans,
                                        float const *bmk, float const *nul, float const *bg) -> int {
    if (!x->gpu)
    { // first protein: the header is known, bring the device up
      x->epsilon = h.epsilon;
      x->is_rna = h.is_rna;
      x->abc_name = h.is_rna ? "rna" : "dna";
      int device = 0;
      if (char const *env = getenv("DCP_GPU_DEVICE")) device = atoi(env);
      int const e = dcpgpu_open(&x->gpu, device);
      if (e) return map_gpu_error(e);
    }
    // the GPU analogue of work_setup/protein_setup_viterbi (work.c:24-46, protein.c:353-394)
    int64_t first = 0;
    int e;
    if ((e = dcpgpu_pool_add(x->gpu, m.K, emission, trans, &first))) return map_gpu_error(e);
    if ((e = dcpgpu_profile_add(x->gpu, m.K, nullptr, first, bmk, nul, bg, nullptr))) return map_gpu_error(e);
    if ((e = dcpgpu_pool_release(x->gpu))) return map_gpu_error(e);
    x->profiles.push_back(std::move(m));
    return 0;
  });
  if (rc) return rc;
  if (!x->gpu)
  { // empty database: still needs a device for dcp_scan_run
    x->is_rna = hdr.is_rna;
    x->abc_name = hdr.is_rna ? "rna" : "dna";
    if ((rc = dcpgpu_open(&x->gpu, 0))) return map_gpu_error(rc);
  }
  return 0;
}

// Parse a database without touching the GPU: number of profiles and total core size.
int dcpb200_db_info(char const *dbfile, int *num_proteins, long *total_core_size, float *epsilon)
{
  if (!dbfile) return DCP_EFUNCUSE;
  DbHeader hdr;
  long total = 0;
  int n = 0;
  int const rc = parse_db(dbfile, &hdr, [&](DbHeader const &, ProfileMeta &m, float const *, float const *,
                                              float const *, float const *, float const *) -> int {
    total += m.K;
    ++n;
    return 0;
  });
  if (rc) return rc;
  if (num_proteins) *num_proteins = n;
  if (total_core_size) *total_core_size = total;
  if (epsilon) *epsilon = hdr.epsilon;
  return 0;
}

void dcp_scan_interrupt(struct dcp_scan *x)
{
  if (x) x->interrupted = true;
}

int dcp_scan_progress(struct dcp_scan const *x)
{ // scan.c:224-227
  if (!x || x->profiles.empty()) return 0;
  return (int)((100L * x->done_proteins) / (long)x->profiles.size());
}

int dcp_scan_run(struct dcp_scan *x, struct dcp_batch *batch, char const *product_dir)
{
  if (!x || !batch || !product_dir || !x->gpu) return DCP_EFUNCUSE;
  x->interrupted = false;
  x->done_proteins = 0;
  int const P = (int)x->profiles.size();
  int const S = (int)batch->seqs.size();
  uint32_t const flags = (x->multi_hits ? DCPGPU_MULTI_HITS : 0u) | (x->hmmer3_compat ? DCPGPU_HMMER3_COMPAT : 0u);

  // batch_encode (batch.c:60-70, sequence.c:47-84)
  std::vector<int64_t> offsets((size_t)S + 1, 0);
  for (int s = 0; s < S; ++s)
    offsets[(size_t)s + 1] = offsets[(size_t)s] + (int64_t)batch->seqs[(size_t)s].data.size();
  std::vector<uint8_t> symbols((size_t)offsets[(size_t)S] + 1);
  for (int s = 0; s < S; ++s)
  {
    std::string const &d = batch->seqs[(size_t)s].data;
    uint8_t *o = symbols.data() + offsets[(size_t)s];
    for (size_t i = 0; i < d.size(); ++i)
    {
      switch (d[i])
      {
      case 'A': o[i] = 0; break;
      case 'C': o[i] = 1; break;
      case 'G': o[i] = 2; break;
      case 'T': if (x->is_rna) return DCP_EDBRNASEQDNA; o[i] = 3; break;
      case 'U': if (!x->is_rna) return DCP_EDBDNASEQRNA; o[i] = 3; break;
      default: return DCP_ESEQABC;
      }
    }
  }
  int rc;
  if ((rc = dcpgpu_reads_set(x->gpu, S, symbols.data(), offsets.data()))) return map_gpu_error(rc);

  // product_open (product.c:15-32)
  std::string const dir(product_dir);
  if ((rc = mkdir_p(dir)) || (rc = mkdir_p(dir + "/hmmer"))) return rc;

  std::vector<Row> rows;
  std::vector<Active> active; // pairs that still have windows to visit after the current wave

  // turn the traced hits of one wave into rows and next-window state
  auto process_hits = [&](std::vector<dcpgpu_pair> const &pairs, std::vector<int> const &win_idx,
                          std::vector<float> const &nulc, std::vector<float> const &altc,
                          std::vector<Active *> const &owner) -> int {
    // trace in chunks bounded by trellis bytes
    size_t i0 = 0;
    while (i0 < pairs.size())
    {
      size_t i1 = i0;
      double bytes = 0;
      while (i1 < pairs.size())
      {
        double const b = (double)(pairs[i1].len + 1) * (2.0 * x->profiles[(size_t)pairs[i1].profile].K + 4.0);
        if (i1 > i0 && bytes + b > 6e9) break;
        bytes += b;
        ++i1;
      }
      size_t const n = i1 - i0;
      std::vector<int32_t> nsteps(n);
      std::vector<float> talt(n);
      if ((rc = dcpgpu_trace_pairs(x->gpu, (int64_t)n, &pairs[i0], flags, talt.data(), nsteps.data())))
        return map_gpu_error(rc);
      std::vector<int64_t> off(n + 1, 0);
      for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + nsteps[i];
      std::vector<uint16_t> ids((size_t)off[n] + 1);
      std::vector<uint8_t> szs((size_t)off[n] + 1);
      if ((rc = dcpgpu_trace_fetch(x->gpu, off.data(), ids.data(), szs.data()))) return map_gpu_error(rc);

      for (size_t i = 0; i < n; ++i)
      {
        dcpgpu_pair const &pr = pairs[i0 + i];
        ProfileMeta const &pm = x->profiles[(size_t)pr.profile];
        Sequence const &sq = batch->seqs[(size_t)pr.seq];
        uint16_t const *sid = &ids[(size_t)off[i]];
        uint8_t const *ssz = &szs[(size_t)off[i]];
        int const ns = nsteps[i];
        // hit extent: first B .. last E (thread.c:130-166), window-relative positions
        int pos = 0, b = 0;
        while (b < ns && sid[b] != ST_B) pos += ssz[b++];
        if (b >= ns) continue;
        int const hit_start = pos;
        int hit_stop = pos, end = -1;
        for (int j = b; j < ns; ++j)
        {
          if (sid[j] == ST_E) { hit_stop = pos; end = j + 1; }
          pos += ssz[j];
        }
        if (end < 0) continue;
        if (owner[i0 + i]) owner[i0 + i]->last_hit = hit_stop - 1; // thread.c:162

        // row (product_thread.c:40-79) and match string (product_thread.c:112-148)
        float const null_ll = -nulc[i0 + i], alt_ll = -altc[i0 + i];
        float const lrt = -2 * (null_ll - alt_ll); // lrt.h:6-9
        char head[256];
        snprintf(head, sizeof head, "%ld\t%d\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%.1f\t%.2g\t", sq.id, win_idx[i0 + i],
                 pr.start, pr.start + pr.len, 0, hit_start, hit_stop, pm.accession.c_str(), x->abc_name.c_str(),
                 (double)lrt, 0.0);
        std::string text(head);
        char const *table = gencode_table(pm.gencode);
        int p2 = hit_start;
        char const *data = sq.data.c_str() + pr.start;
        for (int j = b; j < end; ++j)
        {
          if (j > b) text.push_back(';');
          int const sz = ssz[j];
          text.append(data + p2, (size_t)sz);
          text.push_back(',');
          char nm[16];
          state_name(sid[j], nm);
          text.append(nm);
          text.push_back(',');
          if (!state_is_mute(sid[j]))
          {
            int z[5], cod[3];
            for (int t = 0; t < sz; ++t)
            {
              char const c = data[p2 + t];
              z[t] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
            }
            int const msb = sid[j] & (3 << 14);
            NucltDist const &nd = msb == ST_I   ? pm.bg_dist
                                  : msb == ST_M ? pm.nodes[(size_t)((sid[j] & 0x3fff) - 1)]
                                                : pm.null_dist; // decoder.c:43-49
            if (!decode_codon(nd, z, sz, cod)) return DCP_EDECODON;
            char const sym[] = {'A', 'C', 'G', x->is_rna ? 'U' : 'T'};
            text.push_back(sym[cod[0]]);
            text.push_back(sym[cod[1]]);
            text.push_back(sym[cod[2]]);
            text.push_back(',');
            text.push_back(codon_amino(table, cod[0], cod[1], cod[2]));
          }
          else
            text.push_back(',');
          p2 += sz;
        }
        text.push_back('\n');
        rows.push_back(Row{pr.profile, pr.seq, win_idx[i0 + i], std::move(text)});
      }
      i0 = i1;
    }
    return 0;
  };

  if (P > 0 && S > 0)
  {
    // ---- wave 0: the first window of every (sequence, profile), generated on the device ----
    for (int s = 0; s < S; ++s)
      if (batch->seqs[(size_t)s].data.empty()) return DCP_EZEROSEQ;
    if ((rc = dcpgpu_score_grid(x->gpu, 0, P, 0, S, flags))) return map_gpu_error(rc);
    int64_t nhits = 0;
    if ((rc = dcpgpu_hits_fetch(x->gpu, 0, nullptr, &nhits))) return map_gpu_error(rc);
    std::vector<int64_t> hit((size_t)nhits);
    if (nhits && (rc = dcpgpu_hits_fetch(x->gpu, nhits, hit.data(), &nhits))) return map_gpu_error(rc);
    std::vector<float> nul0((size_t)P * S), alt0((size_t)P * S);
    if ((rc = dcpgpu_scores_fetch(x->gpu, (int64_t)P * S, nul0.data(), alt0.data()))) return map_gpu_error(rc);
    if (x->callback) x->callback(x->userdata);
    if (x->interrupted) return 0;

    // pairs whose sequence is longer than the first window keep iterating (window.c:27-31)
    for (int p = 0; p < P; ++p)
    {
      int const w = std::min(x->profiles[(size_t)p].K * 50, 100000);
      for (int s = 0; s < S; ++s)
      {
        int const len = (int)batch->seqs[(size_t)s].data.size();
        if (len > w) active.push_back(Active{p, s, 0, w, 0, -1});
      }
    }
    std::map<std::pair<int, int>, Active *> by_pair;
    for (auto &a : active) by_pair[{a.profile, a.seq}] = &a;

    std::vector<dcpgpu_pair> hp((size_t)nhits);
    std::vector<int> widx((size_t)nhits, 0);
    std::vector<float> hn((size_t)nhits), ha((size_t)nhits);
    std::vector<Active *> owner((size_t)nhits, nullptr);
    for (int64_t i = 0; i < nhits; ++i)
    {
      int const p = (int)(hit[(size_t)i] / S), s = (int)(hit[(size_t)i] % S);
      int const len = (int)batch->seqs[(size_t)s].data.size();
      hp[(size_t)i] = dcpgpu_pair{p, s, 0, std::min(len, std::min(x->profiles[(size_t)p].K * 50, 100000))};
      hn[(size_t)i] = nul0[(size_t)hit[(size_t)i]];
      ha[(size_t)i] = alt0[(size_t)hit[(size_t)i]];
      auto it = by_pair.find({p, s});
      if (it != by_pair.end()) owner[(size_t)i] = it->second;
    }
    if ((rc = process_hits(hp, widx, hn, ha, owner))) return rc;

    // ---- waves 1..: explicit windows of the pairs that are still active ----
    while (!active.empty() && !x->interrupted)
    {
      std::vector<Active> next;
      std::vector<dcpgpu_pair> wp;
      for (auto &a : active)
      {
        int const len = (int)batch->seqs[(size_t)a.seq].data.size();
        if (window_next(a, len, x->profiles[(size_t)a.profile].K))
        {
          wp.push_back(dcpgpu_pair{a.profile, a.seq, a.start, a.stop - a.start});
          next.push_back(a);
        }
      }
      if (wp.empty()) break;
      std::vector<float> nulc(wp.size()), altc(wp.size());
      if ((rc = dcpgpu_score_pairs(x->gpu, (int64_t)wp.size(), wp.data(), flags, nulc.data(), altc.data())))
        return map_gpu_error(rc);
      std::vector<dcpgpu_pair> hp2;
      std::vector<int> widx2;
      std::vector<float> hn2, ha2;
      std::vector<Active *> owner2;
      for (size_t i = 0; i < wp.size(); ++i)
      {
        float const lrt = -2 * ((-nulc[i]) - (-altc[i]));
        if (!std::isfinite(lrt) || lrt < 0) continue; // thread.c:121
        hp2.push_back(wp[i]);
        widx2.push_back(next[i].idx);
        hn2.push_back(nulc[i]);
        ha2.push_back(altc[i]);
        owner2.push_back(&next[i]);
      }
      if ((rc = process_hits(hp2, widx2, hn2, ha2, owner2))) return rc;
      active.swap(next);
      if (x->callback) x->callback(x->userdata);
    }
  }
  x->done_proteins = P;

  // product_close (product.c:34-87): header + rows in (profile, batch order, window) order
  std::stable_sort(rows.begin(), rows.end(), [](Row const &a, Row const &b) {
    if (a.profile != b.profile) return a.profile < b.profile;
    if (a.seq_order != b.seq_order) return a.seq_order < b.seq_order;
    return a.window < b.window;
  });
  FILE *fp = fopen((dir + "/products.tsv").c_str(), "wb");
  if (!fp) return DCP_EWRITEPROD;
  bool ok = fputs("sequence\twindow\twindow_start\twindow_stop\thit\thit_start\thit_stop\tprofile\tabc\tlrt\tevalue\tmatch\n", fp) >= 0;
  for (auto const &r : rows) ok = ok && fputs(r.text.c_str(), fp) >= 0;
  ok = (fclose(fp) == 0) && ok;
  return ok ? 0 : DCP_EWRITEPROD;
}

} // extern "C"
