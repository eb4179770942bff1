// press.cpp -- the reference's press API (dcp_press_*, c-core/press.c) over the GPU C ABI.
//
// HMMER3 ASCII profiles -> per-profile quasi-codon tables -> .dcp database:
//   hmm_reader.c:19-103    node 0 carries the begin transitions, then (match emissions,
//                          transitions) per node; file values are -ln p, "*" = probability zero;
//                          the null amino model is HMMER3's Swiss-Prot 50.8 background (:78-103)
//   model.c:62-96          match log-odds = match - null log-probs -> setup_nuclt_dist
//   model.c:390-441        amino -> codon log-probs over the synonymous codons of the genetic code
//                          (stop codons impossible), normalised; base log-probs = mean over the
//                          three codon positions; codon marginals with "any base" wildcards
//   model.c:284-309        occupancy-based entry distribution B -> M_k
//   protein.c:67-120       record node i = match state of node min(i, K-1) with the transitions out
//                          of node i+1's predecessor (alt.trans[min(i+1, K)])
//   protein.c:102          imm_score_table_scores: the 1364-entry frame-state table of every state --
//                          the hot loop, done on the device (dcpgpu_frame_tables, csrc/press_kernel.cuh)
//   database_writer.c, protein.c:234-281, write.c: the file
// The writer emits the current encoding byte for byte in size (test_press.c:26: minifam.hmm ->
// 3,609,858 bytes): float arrays as bin + host-endian floats, nuclt_dist as two arrays of float32
// values, integers in their smallest MessagePack form.
#include "../../include/dcpgpu.h"
#include "../../include/deciphon_b200.h"
#include "dcp_common.h"
#include "gencode.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

constexpr int NCODES = DCPGPU_NUM_CODES;
constexpr int MODEL_MAX = 16384; // model.h:12
char const AMINO[] = "ACDEFGHIKLMNPQRSTVWY"; // imm_amino_iupac = HMMER3 column order

// HMMER3 amino background, Swiss-Prot 50.8 (hmm_reader.c:78-103)
float const NULL_AMINO[20] = {0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f,
                              0.0590092f, 0.0594422f, 0.0963728f, 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f,
                              0.0540978f, 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

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

double logsumexp(std::vector<double> const &v)
{
  double m = -INFINITY;
  for (double x : v) m = std::max(m, x);
  if (!(m > -INFINITY)) return m;
  double s = 0;
  for (double x : v) s += std::exp(x - m);
  return m + std::log(s);
}

// setup_nuclt_dist (model.c:428-441): amino log-probs (or log-odds) -> base log-probs + codon marginals
void nuclt_dist(char const *gencode, double const amino_lprobs[20], dcpb::NucltDist *out)
{
  int count[128] = {0};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      for (int c = 0; c < 4; ++c) count[(int)dcpb::codon_amino(gencode, a, b, c)] += 1;
  double codon[64];
  std::vector<double> all;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      for (int c = 0; c < 4; ++c)
      {
        char const aa = dcpb::codon_amino(gencode, a, b, c);
        char const *pos = aa ? strchr(AMINO, aa) : nullptr;
        // stop codons (and any letter outside the amino alphabet) stay impossible, model.c:401-421
        double const v = pos ? amino_lprobs[pos - AMINO] - std::log((double)count[(int)aa]) : -INFINITY;
        codon[a * 16 + b * 4 + c] = v;
        all.push_back(v);
      }
  double const norm = logsumexp(all); // imm_codon_lprob_normalize
  for (double &v : codon) v -= norm;
  for (int x = 0; x < 4; ++x)
  { // model.c:366-388: each codon position contributes lprob - log 3
    std::vector<double> terms;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int c = 0; c < 4; ++c)
        {
          double const v = codon[a * 16 + b * 4 + c];
          if (!(v > -INFINITY)) continue;
          int const n = (a == x) + (b == x) + (c == x);
          for (int i = 0; i < n; ++i) terms.push_back(v - std::log(3.0));
        }
    out->nuclt[x] = (float)logsumexp(terms);
  }
  for (int a = 0; a < 5; ++a)
    for (int b = 0; b < 5; ++b)
      for (int c = 0; c < 5; ++c)
      { // imm_codon_marg: index 4 sums the position out
        double s = 0;
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 4; ++k)
              if ((a == 4 || a == i) && (b == 4 || b == j) && (c == 4 || c == k)) s += std::exp(codon[i * 16 + j * 4 + k]);
        out->codon[a * 25 + b * 5 + c] = s > 0 ? (float)std::log(s) : -INFINITY;
      }
}

// calculate_occupancy (model.c:284-309); trans[i] = {MM, MI, MD, IM, II, DM, DD} out of node i
void occupancy(std::vector<float> const &trans, int K, std::vector<float> *bmk)
{
  auto T = [&](int i, int j) { return (double)trans[(size_t)i * 7 + (size_t)j]; };
  auto lae = [](double a, double b) {
    double const m = std::max(a, b);
    if (!(m > -INFINITY)) return m;
    return m + std::log(std::exp(a - m) + std::exp(b - m));
  };
  std::vector<double> locc((size_t)K);
  locc[0] = lae(T(0, 1), T(0, 0));
  for (int i = 1; i < K; ++i)
  {
    double const v0 = locc[(size_t)i - 1] + lae(T(i, 0), T(i, 1));
    double const v1 = std::log1p(-std::exp(locc[(size_t)i - 1])) + T(i, 5);
    locc[(size_t)i] = lae(v0, v1);
  }
  std::vector<double> z;
  for (int i = 0; i < K; ++i) z.push_back(locc[(size_t)i] + std::log((double)(K - i)));
  double const logZ = logsumexp(z);
  bmk->resize((size_t)K);
  for (int i = 0; i < K; ++i) (*bmk)[(size_t)i] = (float)(locc[(size_t)i] - logZ);
}

struct HmmProfile
{
  std::string acc;
  bool has_ga = false;
  int K = 0;
  std::vector<float> match;  // [K][20] log-probs
  std::vector<float> trans;  // [K + 1][7] log-probs, row 0 = begin node
  std::string consensus;
};

// One line of the file into tokens
bool next_line(FILE *fp, std::string *line)
{
  line->clear();
  int c;
  bool any = false;
  while ((c = fgetc(fp)) != EOF)
  {
    any = true;
    if (c == '\n') break;
    line->push_back((char)c);
  }
  return any;
}

std::vector<std::string> split(std::string const &s)
{
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size())
  {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r')) ++i;
    size_t j = i;
    while (j < s.size() && s[j] != ' ' && s[j] != '\t' && s[j] != '\r') ++j;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}

bool number(std::string const &tok, float *out)
{ // file values are -ln p; "*" = probability zero (what hmmer_reader hands to hmm_reader.c)
  if (tok == "*")
  {
    *out = -INFINITY;
    return true;
  }
  char *end = nullptr;
  double const v = strtod(tok.c_str(), &end);
  if (end == tok.c_str() || *end) return false;
  *out = (float)(-v);
  if (*out == 0.0f) *out = 0.0f; // no negative zero
  return true;
}

// Reads the next profile; returns 0, or -1 at end of file, or a DCP_E* code.
int read_profile(FILE *fp, HmmProfile *h)
{
  std::string line;
  *h = HmmProfile{};
  bool in_profile = false;
  while (next_line(fp, &line))
  {
    if (!in_profile)
    {
      if (line.compare(0, 7, "HMMER3/") == 0) in_profile = true;
      continue;
    }
    std::vector<std::string> t = split(line);
    if (t.empty()) continue;
    if (t[0] == "ACC" && t.size() > 1) h->acc = t[1];
    else if (t[0] == "LENG" && t.size() > 1) h->K = atoi(t[1].c_str());
    else if (t[0] == "GA") h->has_ga = t.size() > 1;
    else if (t[0] == "HMM")
      break;
  }
  if (!in_profile) return -1;
  if (h->K <= 0) return DCP_EREADHMMER3;
  if (!next_line(fp, &line)) return DCP_EREADHMMER3; // "m->m m->i ..." header
  if (!next_line(fp, &line)) return DCP_EREADHMMER3;
  std::vector<std::string> t = split(line);
  if (!t.empty() && t[0] == "COMPO")
  {
    if (!next_line(fp, &line)) return DCP_EREADHMMER3;
  }
  // `line` = node 0 insert emissions; then its transitions
  if (!next_line(fp, &line)) return DCP_EREADHMMER3;
  t = split(line);
  if (t.size() < 7) return DCP_EREADHMMER3;
  h->trans.resize(7);
  for (int j = 0; j < 7; ++j)
    if (!number(t[(size_t)j], &h->trans[(size_t)j])) return DCP_EREADHMMER3;
  for (;;)
  {
    if (!next_line(fp, &line)) return DCP_EENDOFNODES;
    t = split(line);
    if (!t.empty() && t[0] == "//") break;
    if (t.size() < 21) return DCP_EREADHMMER3;
    size_t const at = h->match.size();
    h->match.resize(at + 20);
    for (int j = 0; j < 20; ++j)
      if (!number(t[(size_t)j + 1], &h->match[at + (size_t)j])) return DCP_EREADHMMER3;
    h->consensus.push_back(t.size() > 22 && !t[22].empty() ? t[22][0] : '-');
    if (!next_line(fp, &line)) return DCP_EENDOFNODES; // insert emissions
    if (!next_line(fp, &line)) return DCP_EENDOFNODES;
    t = split(line);
    if (t.size() < 7) return DCP_EREADHMMER3;
    size_t const tt = h->trans.size();
    h->trans.resize(tt + 7);
    for (int j = 0; j < 7; ++j)
      if (!number(t[(size_t)j], &h->trans[tt + (size_t)j])) return DCP_EREADHMMER3;
  }
  if ((int)h->consensus.size() != h->K) return DCP_EENDOFNODES;
  return 0;
}

} // namespace

struct dcp_press
{
  int gencode_id = 0;
  char const *gencode = nullptr;
  float epsilon = 0.01f;
  FILE *hmm = nullptr;
  FILE *tmp = nullptr; // the records, appended as they are pressed
  std::string db_path, tmp_path;
  long count = 0;
  long done = 0;
  bool end = false;
  bool has_ga = true;
  dcpgpu_ctx *gpu = nullptr;
  std::vector<uint64_t> sizes;
  dcpb::NucltDist null_dist, bg_dist;
  std::vector<float> null_emission, bg_emission;
};

namespace {

void press_cleanup(dcp_press *x)
{
  if (x->hmm) fclose(x->hmm);
  if (x->tmp) fclose(x->tmp);
  x->hmm = x->tmp = nullptr;
  if (!x->tmp_path.empty()) remove(x->tmp_path.c_str());
  x->tmp_path.clear();
  if (x->gpu) dcpgpu_close(x->gpu);
  x->gpu = nullptr;
}

void pack_nuclt_dist(dcpb::Writer &w, dcpb::NucltDist const &d)
{ // nuclt_dist.c:13-20
  w.array(2);
  w.f32list(d.nuclt, 4);
  w.f32list(d.codon, 125);
}

void pack_abc(dcpb::Writer &w, char const *symbols, int typeid_)
{ // third-party imm_abc_pack: {symbols, idx[94] (symbol -> index, 0x7f = none), any_symbol_id, typeid}
  unsigned char idx[94];
  memset(idx, 0x7f, sizeof idx);
  for (int i = 0; symbols[i]; ++i) idx[symbols[i] - '!'] = (unsigned char)i;
  idx['X' - '!'] = (unsigned char)strlen(symbols);
  w.map(4);
  w.str("symbols");
  w.str(symbols);
  w.str("idx");
  w.byte(0xc7); // ext 8, type 0, as in the golden file
  w.be(94, 1);
  w.byte(0);
  w.out.append(reinterpret_cast<char const *>(idx), 94);
  w.str("any_symbol_id");
  w.uint('X' - '!');
  w.str("typeid");
  w.uint((uint64_t)typeid_);
}

} // namespace

extern "C" {

struct dcp_press *dcp_press_new(void) { return new (std::nothrow) dcp_press; }

int dcp_press_setup(struct dcp_press *x, int gencode_id, float epsilon)
{ // press.c:53-63
  if (!x) return DCP_EFUNCUSE;
  x->gencode = dcpb::gencode_table(gencode_id);
  if (!x->gencode) return DCP_EGENCODEID;
  x->gencode_id = gencode_id;
  x->epsilon = epsilon;
  return 0;
}

int dcp_press_open(struct dcp_press *x, char const *hmm, char const *db)
{ // press.c:65-107
  if (!x || !hmm || !db) return DCP_EFUNCUSE;
  if (!x->gencode) return DCP_ESETGENCODE;
  press_cleanup(x);
  x->hmm = fopen(hmm, "rb");
  if (!x->hmm) return DCP_EOPENHMM;
  x->db_path = db;
  x->tmp_path = x->db_path + ".records.tmp";
  x->tmp = fopen(x->tmp_path.c_str(), "wb+");
  if (!x->tmp)
  {
    press_cleanup(x);
    return DCP_EOPENDB;
  }
  // count_proteins (press.c:111-130)
  x->count = 0;
  {
    char buf[4096];
    while (fgets(buf, sizeof buf, x->hmm))
      if (!strncmp(buf, "HMMER3/f", 8)) ++x->count;
    if (!feof(x->hmm))
    {
      press_cleanup(x);
      return DCP_EFREAD;
    }
    rewind(x->hmm);
  }
  x->done = 0;
  x->end = false;
  x->has_ga = true;
  x->sizes.clear();

  int device = 0;
  if (char const *env = getenv("DCP_GPU_DEVICE")) device = atoi(env);
  int rc = dcpgpu_open(&x->gpu, device);
  if (rc)
  {
    press_cleanup(x);
    return map_gpu_error(rc);
  }
  // null and background states are the same for every profile (model.c:142-155)
  double null_lp[20], zeros[20];
  for (int i = 0; i < 20; ++i)
  {
    null_lp[i] = (double)logf(NULL_AMINO[i]);
    zeros[i] = 0;
  }
  nuclt_dist(x->gencode, null_lp, &x->null_dist);
  nuclt_dist(x->gencode, zeros, &x->bg_dist);
  float nu[8], cm[250];
  memcpy(nu, x->null_dist.nuclt, 16);
  memcpy(nu + 4, x->bg_dist.nuclt, 16);
  memcpy(cm, x->null_dist.codon, 500);
  memcpy(cm + 125, x->bg_dist.codon, 500);
  std::vector<float> em(2 * (size_t)NCODES);
  if ((rc = dcpgpu_frame_tables(x->gpu, 2, nu, cm, x->epsilon, em.data())))
  {
    press_cleanup(x);
    return map_gpu_error(rc);
  }
  x->null_emission.assign(em.begin(), em.begin() + NCODES);
  x->bg_emission.assign(em.begin() + NCODES, em.end());
  return 0;
}

long dcp_press_nproteins(struct dcp_press const *x) { return x ? x->count : 0; }

int dcp_press_next(struct dcp_press *x)
{ // press.c:132-142: read the next profile, absorb it, pack its record
  if (!x || !x->hmm || !x->gpu) return DCP_EFUNCUSE;
  HmmProfile h;
  int rc = read_profile(x->hmm, &h);
  if (rc == -1)
  {
    x->end = true;
    return 0;
  }
  if (rc) return rc;
  int const K = h.K;
  if (K > MODEL_MAX) return DCP_ELARGEMODEL;
  if (h.acc.size() >= 32) return DCP_ELONGACCESSION;
  if (!h.has_ga) x->has_ga = false;

  // model_add_node (model.c:62-96): log-odds against the null amino model -> nuclt_dist
  std::vector<dcpb::NucltDist> nd((size_t)K);
  std::vector<float> nu((size_t)K * 4), cm((size_t)K * 125);
  for (int k = 0; k < K; ++k)
  {
    double lodds[20];
    for (int i = 0; i < 20; ++i) lodds[i] = (double)(h.match[(size_t)k * 20 + (size_t)i] - logf(NULL_AMINO[i]));
    nuclt_dist(x->gencode, lodds, &nd[(size_t)k]);
    memcpy(&nu[(size_t)k * 4], nd[(size_t)k].nuclt, 16);
    memcpy(&cm[(size_t)k * 125], nd[(size_t)k].codon, 500);
  }
  // protein_absorb (protein.c:96-107): the frame-state table of every match state, on the device
  std::vector<float> emission((size_t)K * NCODES);
  if ((rc = dcpgpu_frame_tables(x->gpu, K, nu.data(), cm.data(), x->epsilon, emission.data()))) return map_gpu_error(rc);
  std::vector<float> bmk;
  occupancy(h.trans, K, &bmk); // ENTRY_DIST_OCCUPANCY, press.c:60

  // protein_pack (protein.c:234-281)
  dcpb::Writer w;
  w.map(10);
  w.str("accession");
  w.str(h.acc);
  w.str("gencode");
  w.uint((uint64_t)x->gencode_id);
  w.str("consensus");
  w.str(h.consensus);
  w.str("core_size");
  w.uint((uint64_t)K);
  w.str("null_nuclt_dist");
  pack_nuclt_dist(w, x->null_dist);
  w.str("null_emission");
  w.f32bin(x->null_emission.data(), NCODES);
  w.str("bg_nuclt_dist");
  pack_nuclt_dist(w, x->bg_dist);
  w.str("bg_emission");
  w.f32bin(x->bg_emission.data(), NCODES);
  w.str("nodes");
  w.map((uint32_t)(K + 1) * 3);
  for (int i = 0; i <= K; ++i)
  {
    int const node = std::min(i, K - 1), tr = std::min(i + 1, K); // protein.c:99-104
    w.str("nuclt_dist");
    pack_nuclt_dist(w, nd[(size_t)node]);
    w.str("trans");
    w.f32bin(&h.trans[(size_t)tr * 7], 7);
    w.str("emission");
    w.f32bin(&emission[(size_t)node * NCODES], NCODES);
  }
  w.str("BMk");
  w.f32bin(bmk.data(), (size_t)K);
  if (fwrite(w.out.data(), 1, w.out.size(), x->tmp) != w.out.size()) return DCP_EFWRITE;
  x->sizes.push_back((uint64_t)w.out.size());
  x->done += 1;
  return 0;
}

bool dcp_press_end(struct dcp_press const *x) { return !x || x->end; }

int dcp_press_close(struct dcp_press *x)
{ // press.c:149-160, database_writer.c:136-150
  if (!x) return DCP_EFUNCUSE;
  int rc = 0;
  if (x->tmp && x->gencode)
  {
    FILE *out = fopen(x->db_path.c_str(), "wb");
    if (!out) rc = DCP_EOPENDB;
    if (!rc)
    {
      dcpb::Writer w;
      w.map(2);
      w.str("header");
      w.map(8);
      w.str("magic_number");
      w.uint(0xC6F1);
      w.str("version");
      w.uint(1);
      w.str("entry_dist");
      w.uint(2);
      w.str("epsilon");
      w.f32(x->epsilon);
      w.str("abc");
      pack_abc(w, "ACGT", 4);
      w.str("amino");
      pack_abc(w, AMINO, 2);
      w.str("has_ga");
      w.boolean(x->has_ga);
      w.str("protein_sizes");
      w.array((uint32_t)x->sizes.size());
      for (uint64_t s : x->sizes) w.uint(s);
      w.str("proteins");
      w.array((uint32_t)x->sizes.size());
      bool ok = fwrite(w.out.data(), 1, w.out.size(), out) == w.out.size();
      fflush(x->tmp);
      rewind(x->tmp);
      std::vector<char> buf(size_t(1) << 20);
      size_t got;
      while (ok && (got = fread(buf.data(), 1, buf.size(), x->tmp)) > 0) ok = fwrite(buf.data(), 1, got, out) == got;
      ok = (fclose(out) == 0) && ok;
      if (!ok) rc = DCP_EFWRITE;
    }
  }
  press_cleanup(x);
  return rc;
}

void dcp_press_del(struct dcp_press const *x)
{
  if (!x) return;
  press_cleanup(const_cast<dcp_press *>(x));
  delete x;
}

} // extern "C"
