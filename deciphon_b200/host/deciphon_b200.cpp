// deciphon_b200.cpp -- the reference's scan API (include/deciphon_b200.h) over the GPU C ABI.
//
// Host side of the drop-in.  The reference partitions the database over num_threads OpenMP
// threads (scan.c:95-152, protein_reader.c:112-128), each looping profiles -> sequences ->
// windows (thread.c:49-208).  Here the partitions are SHARDS, one per GPU (contiguous profile
// ranges balanced by record size, i.e. by core size), each driven by its own host thread; a shard
// walks its profiles in CHUNKS (a profile range worth a few 1e10 DP cells) and runs every chunk as
// WAVES of batched GPU passes:
//   wave w = the w-th window.c window of every still-active (sequence, profile) pair of the chunk:
//     score pass (viterbi_null + viterbi_cost)  -> lrt gate (thread.c:119-121)
//     trace pass (viterbi_path + trellis_unzip) -> hit extent, codons, aminos and the bytes of the
//     match column on the device (thread.c:130-180, match.c, decoder.c -> csrc/match_kernel.cuh)
//     -> window_set_last_hit_position -> next window (window.c:13-37)
// After every chunk done_proteins advances, the callback fires (shard 0, like the reference's
// rank 0, scan.c:196-198) and `interrupted` is honoured (thread.c:74-79).  Rows are merged in
// profile order like product_close concatenates the per-thread files (product.c:63-80).
#include "../../include/deciphon_b200.h"
#include "../../include/dcpgpu.h"
#include "dcp_common.h"
#include "gencode.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <thread>
#include <vector>

namespace {

constexpr int NCODES = DCPGPU_NUM_CODES;

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

struct ProfileInfo
{
  std::string accession;
  int gencode = 1;
  int K = 0;
};

struct Shard
{ // one GPU and the contiguous profile range [p0, p1) resident on it
  dcpgpu_ctx *gpu = nullptr;
  int device = 0;
  int p0 = 0, p1 = 0;
  bool stagger = false; // not the first shard of its GPU
};

struct dcp_scan
{
  std::vector<Shard> shards;
  bool multi_hits = true, hmmer3_compat = false;
  float epsilon = 0.01f;
  std::string abc_name = "dna";
  bool is_rna = false;
  std::vector<ProfileInfo> profiles;
  void (*callback)(void *) = nullptr;
  void *userdata = nullptr;
  std::atomic<bool> interrupted{false};
  std::atomic<int> done_proteins{0};
  std::atomic<long> windows{0}, lrt_windows{0}; // cumulative: windows scored / with lrt >= 0 (dcpb200_scan_counter)
  std::atomic<long> speculative_windows{0};     // planned and scored, then re-planned after a hit changed the chain
  std::atomic<long long> cells{0};              // DP cells of the committed windows
  double chunk_cells = 4e11; // DP cells of first windows per chunk (DCP_CHUNK_CELLS)
  bool write_aminos = false; // DCP_WRITE_AMINOS: product_dir/aminos.fa, the hits' amino-acid sequences
};

namespace {

using dcpb::NucltDist;
using dcpb::Reader;

void close_shards(dcp_scan *x)
{
  for (auto &sh : x->shards)
    if (sh.gpu) dcpgpu_close(sh.gpu);
  x->shards.clear();
}

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

int mkdir_p(std::string const &dir)
{
  if (mkdir(dir.c_str(), 0755) == 0 || errno == EEXIST) return 0;
  return DCP_EMKDIR;
}

// DCP_TIMING=1: per-phase wall seconds of a dcp_scan_run on stderr (development aid)
struct PhaseTimer
{
  double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point mark = std::chrono::steady_clock::now();
  void lap(int i)
  {
    auto const now = std::chrono::steady_clock::now();
    t[i] += std::chrono::duration<double>(now - mark).count();
    mark = now;
  }
};

struct Row
{
  int profile;
  int seq_order;
  int window;
  std::string text;
  std::string amino; // FASTA record of the hit's amino-acid sequence (only with DCP_WRITE_AMINOS)
};

// The amino-acid sequence the reference hands to HMMER for a hit (thread.c:168-180): the amino letter
// of every non-mute step of the match column "<frag>,<state>,<codon>,<amino>;...".
std::string amino_of_match(char const *text, size_t n)
{
  std::string out;
  size_t i = 0;
  while (i < n)
  {
    int commas = 0;
    size_t j = i;
    for (; j < n && text[j] != ';'; ++j)
      if (text[j] == ',' && ++commas == 3 && j + 1 < n && text[j + 1] != ';') out.push_back(text[j + 1]);
    i = j + 1;
  }
  return out;
}

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
  std::vector<uint64_t> protein_sizes; // bytes of every record (database_writer.c:60-92)
};

struct ProteinRecord
{
  ProfileInfo info;
  std::vector<float> nul, bg, emission, trans, bmk; // .dcp (log-prob) form
  std::vector<float> dists;                         // [K + 2][129]: nodes, null, background
};

// Streams a .dcp file: header (database_reader.c:26-80), then one callback per protein record
// (protein.c:283-351), each record read from the file on its own (protein_sizes says how many
// bytes it takes), so memory stays at one record whatever the size of the database.
template <class F>
int parse_db(char const *dbfile, DbHeader *hdr, bool want_tables, F &&on_protein)
{
  FILE *fp = fopen(dbfile, "rb");
  if (!fp) return DCP_EOPENDB;
  struct Closer
  {
    FILE *f;
    ~Closer() { fclose(f); }
  } closer{fp};
  Reader r;
  // the header is small (5 bytes per protein at most): read a prefix and grow it if it is cut short
  size_t want = size_t(1) << 20;
  long header_end = -1;
  for (;;)
  {
    r.buf.resize(want);
    if (fseek(fp, 0, SEEK_SET)) return DCP_EFREAD;
    size_t const got = fread(r.buf.data(), 1, want, fp);
    r.buf.resize(got);
    r.p = 0;
    r.ok = true;
    uint32_t n;
    int64_t iv;
    int rc = 0;
    do
    {
      if (!r.map(&n) || n != 2 || !r.key("header") || !r.map(&n) || n != 8) { rc = DCP_ENOTDBFILE; break; }
      if (!r.key("magic_number") || !r.integer(&iv) || iv != 0xC6F1) { rc = DCP_ENOTDBFILE; break; }
      if (!r.key("version") || !r.integer(&iv)) { rc = DCP_EFDATA; break; }
      if (iv != 1) { rc = DCP_EDBVERSION; break; }
      if (!r.key("entry_dist") || !r.integer(&iv)) { rc = DCP_EFDATA; break; }
      hdr->entry_dist = (int)iv;
      if (!r.key("epsilon") || !r.f32(&hdr->epsilon)) { rc = DCP_EFDATA; break; }
      if (!r.key("abc")) { rc = DCP_EFDATA; break; }
      {
        // alphabet: only the symbols matter here (imm_abc_unpack is third-party)
        size_t const start = r.p;
        if (!r.skip()) { rc = DCP_EFDATA; break; }
        std::string blob(reinterpret_cast<char const *>(&r.buf[start]), r.p - start);
        hdr->is_rna = blob.find("ACGU") != std::string::npos;
        if (blob.find("ACGT") == std::string::npos && !hdr->is_rna) { rc = DCP_ENUCLTNOSUPPORT; break; } // scan.c:107-108
      }
      if (!r.key("amino") || !r.skip()) { rc = DCP_EFDATA; break; }
      if (!r.key("has_ga") || !r.boolean(&hdr->has_ga)) { rc = DCP_EFDATA; break; }
      if (!r.key("protein_sizes")) { rc = DCP_EFDATA; break; }
      {
        hdr->protein_sizes.clear();
        int const b = r.peek();
        if (b >= 0 && ((b & 0xf0) == 0x90 || b == 0xdc || b == 0xdd))
        { // current writer: array of ints
          if (!r.array(&n)) { rc = DCP_EFDATA; break; }
          bool bad = false;
          for (uint32_t i = 0; i < n && !bad; ++i)
          {
            if (!r.integer(&iv) || iv < 0) bad = true;
            else hdr->protein_sizes.push_back((uint64_t)iv);
          }
          if (bad) { rc = DCP_EFDATA; break; }
        }
        else
        { // golden file: ext blob of big-endian u32
          bool be;
          size_t bytes;
          if (!r.blob(&be, &bytes) || bytes % 4) { rc = DCP_EFDATA; break; }
          for (size_t i = 0; i < bytes / 4; ++i)
          {
            unsigned char const *q = &r.buf[r.p + 4 * i];
            uint32_t const v = be ? ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]
                                  : ((uint32_t)q[3] << 24) | ((uint32_t)q[2] << 16) | ((uint32_t)q[1] << 8) | q[0];
            hdr->protein_sizes.push_back(v);
          }
          r.p += bytes;
        }
      }
      if (!r.key("proteins") || !r.array(&hdr->num_proteins)) { rc = DCP_EFDATA; break; }
    } while (false);
    if (rc == 0 && r.ok)
    {
      header_end = (long)r.p;
      break;
    }
    if (!r.ok && got == want)
    { // ran off the end of the prefix: take a larger one
      want *= 4;
      continue;
    }
    return rc ? rc : DCP_EFDATA;
  }
  if (hdr->protein_sizes.size() != hdr->num_proteins) return DCP_EINVALNUMPROTEINS;

  if (fseek(fp, header_end, SEEK_SET)) return DCP_EFREAD;
  ProteinRecord rec;
  rec.nul.resize(NCODES);
  rec.bg.resize(NCODES);
  for (uint32_t pi = 0; pi < hdr->num_proteins; ++pi)
  {
    size_t const bytes = (size_t)hdr->protein_sizes[pi];
    r.buf.resize(bytes);
    r.p = 0;
    r.ok = true;
    if (fread(r.buf.data(), 1, bytes, fp) != bytes) return DCP_EENDOFFILE;
    uint32_t n;
    int64_t iv;
    std::string consensus;
    ProfileInfo &m = rec.info;
    m = ProfileInfo{};
    if (!r.map(&n) || n != 10) return r.ok ? DCP_EFDATA : DCP_EENDOFFILE;
    if (!r.key("accession") || !r.str(&m.accession)) return DCP_EFDATA;
    if (m.accession.size() >= 32) return DCP_ELONGACCESSION;
    if (!r.key("gencode") || !r.integer(&iv)) return DCP_EFDATA;
    m.gencode = (int)iv;
    if (!dcpb::gencode_table(m.gencode)) return DCP_EFREAD; // protein.c:297-298: imm_gencode_get fails -> DCP_EFREAD
    if (!r.key("consensus") || !r.str(&consensus)) return DCP_EFDATA;
    if (!r.key("core_size") || !r.integer(&iv)) return DCP_EFDATA;
    if (iv < 1 || iv > DCPGPU_MAX_CORE_SIZE) return DCP_ELARGECORESIZE;
    int const K = m.K = (int)iv;
    rec.dists.resize((size_t)(K + 2) * 129);
    NucltDist tmp;
    auto put = [&](int row) { memcpy(&rec.dists[(size_t)row * 129], &tmp, sizeof tmp); };
    static_assert(sizeof(NucltDist) == 129 * sizeof(float), "packed");
    if (!r.key("null_nuclt_dist") || !read_nuclt_dist(r, &tmp)) return DCP_EFDATA;
    put(K);
    if (!r.key("null_emission") || !r.f32array(NCODES, rec.nul.data())) return DCP_EFDATA;
    if (!r.key("bg_nuclt_dist") || !read_nuclt_dist(r, &tmp)) return DCP_EFDATA;
    put(K + 1);
    if (!r.key("bg_emission") || !r.f32array(NCODES, rec.bg.data())) return DCP_EFDATA;
    if (!r.key("nodes") || !r.map(&n) || n != (uint32_t)(K + 1) * 3) return DCP_EFDATA;
    rec.emission.resize((size_t)(K + 1) * NCODES);
    rec.trans.resize((size_t)(K + 1) * 7);
    for (int i = 0; i <= K; ++i)
    {
      if (!r.key("nuclt_dist") || !read_nuclt_dist(r, &tmp)) return DCP_EFDATA;
      if (i < K) put(i);
      if (!r.key("trans") || !r.f32array(7, &rec.trans[(size_t)i * 7])) return DCP_EFDATA;
      if (want_tables)
      {
        if (!r.key("emission") || !r.f32array(NCODES, &rec.emission[(size_t)i * NCODES])) return DCP_EFDATA;
      }
      else if (!r.key("emission") || !r.skip())
        return DCP_EFDATA;
    }
    rec.bmk.resize((size_t)K);
    if (!r.key("BMk") || !r.f32array((size_t)K, rec.bmk.data())) return r.ok ? DCP_EFDATA : DCP_EENDOFFILE;
    int const rc = on_protein(*hdr, (int)pi, rec);
    if (rc) return rc;
  }
  return 0;
}

// contiguous ranges of `world` shards balanced by the records' byte sizes (5,456 bytes per node:
// proportional to core size, i.e. to DP cells)
std::vector<int> shard_cuts(std::vector<uint64_t> const &sizes, int world)
{
  std::vector<double> csum(sizes.size() + 1, 0.0);
  for (size_t i = 0; i < sizes.size(); ++i) csum[i + 1] = csum[i] + (double)sizes[i];
  std::vector<int> cuts{0};
  for (int r = 1; r < world; ++r)
  {
    double const target = csum.back() * r / world;
    int c = (int)(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin());
    c = std::min(std::max(c, cuts.back()), (int)sizes.size());
    cuts.push_back(c);
  }
  cuts.push_back((int)sizes.size());
  return cuts;
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

// ---- scan ---------------------------------------------------------------------------------------

struct dcp_scan *dcp_scan_new(void) { return new (std::nothrow) dcp_scan; }

void dcp_scan_del(struct dcp_scan const *scan)
{
  dcp_scan *x = const_cast<dcp_scan *>(scan);
  if (!x) return;
  close_shards(x);
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
  close_shards(x);
  if (char const *env = getenv("DCP_CHUNK_CELLS"))
    if (atof(env) > 0) x->chunk_cells = atof(env);
  if (char const *env = getenv("DCP_WRITE_AMINOS")) x->write_aminos = env[0] && env[0] != '0';

  // partitions -> GPUs: min(num_threads, visible devices, DCP_GPU_COUNT), starting at DCP_GPU_DEVICE;
  // every GPU takes DCP_SHARDS_PER_GPU (default 2) shards when the database is large enough: while
  // one shard's host thread decodes hits and formats rows, the other shard's kernels keep the GPU busy
  int first_device = 0, max_gpus = std::max(1, num_threads), shards_per_gpu = 2;
  if (char const *env = getenv("DCP_GPU_DEVICE")) first_device = atoi(env);
  if (char const *env = getenv("DCP_GPU_COUNT"))
    if (atoi(env) > 0) max_gpus = std::min(max_gpus, atoi(env));
  if (char const *env = getenv("DCP_SHARDS_PER_GPU"))
    if (atoi(env) > 0) shards_per_gpu = std::min(atoi(env), 8);
  // (the devices are looked at only once the file has proved to be a database: the reference's
  // error order, scan.c:102-108)
  auto world_max = [&](int *out) -> int {
    int const visible = (int)dcpgpu_device_count();
    if (visible <= 0) return DCP_EGPUNODEVICE;
    if (first_device < 0 || first_device >= visible) return DCP_EGPUINTERNAL;
    *out = std::min(max_gpus, visible - first_device);
    return 0;
  };

  DbHeader hdr;
  std::vector<int> cuts;
  int rc = parse_db(dbfile, &hdr, true, [&](DbHeader const &h, int pi, ProteinRecord &rec) -> int {
    if (pi == 0)
    { // the header is known: plan the shards and bring the devices up
      x->epsilon = h.epsilon;
      x->is_rna = h.is_rna;
      x->abc_name = h.is_rna ? "rna" : "dna";
      int wm = 1;
      if (int const e = world_max(&wm)) return e;
      int const gpus = std::max(1, std::min<int>(wm, (int)h.num_proteins)); // scan.c:105
      int const spg = (long)h.num_proteins >= 128L * gpus * shards_per_gpu ? shards_per_gpu : 1;
      int const world = gpus * spg;
      cuts = shard_cuts(h.protein_sizes, world);
      for (int r = 0; r < world; ++r)
      {
        Shard sh;
        sh.device = first_device + r / spg;
        sh.stagger = (r % spg) != 0;
        sh.p0 = cuts[(size_t)r];
        sh.p1 = cuts[(size_t)r + 1];
        int const e = dcpgpu_open(&sh.gpu, sh.device);
        if (e) return map_gpu_error(e);
        x->shards.push_back(sh);
      }
    }
    size_t si = 0;
    while (si + 1 < x->shards.size() && pi >= x->shards[si].p1) ++si;
    dcpgpu_ctx *gpu = x->shards[si].gpu;
    // the GPU analogue of work_setup/protein_setup_viterbi (work.c:24-46, protein.c:353-394)
    int64_t first = 0;
    int32_t index = -1;
    int e;
    int const K = rec.info.K;
    if ((e = dcpgpu_pool_add(gpu, K, rec.emission.data(), rec.trans.data(), &first))) return map_gpu_error(e);
    if ((e = dcpgpu_profile_add(gpu, K, nullptr, first, rec.bmk.data(), rec.nul.data(), rec.bg.data(), &index)))
      return map_gpu_error(e);
    if ((e = dcpgpu_pool_release(gpu))) return map_gpu_error(e);
    if ((e = dcpgpu_profile_set_decoder(gpu, index, rec.dists.data(), &rec.dists[(size_t)K * 129],
                                        &rec.dists[(size_t)(K + 1) * 129], dcpb::gencode_table(rec.info.gencode))))
      return map_gpu_error(e);
    x->profiles.push_back(rec.info);
    return 0;
  });
  if (rc)
  {
    close_shards(x);
    return rc;
  }
  if (x->shards.empty())
  { // empty database: still needs a device for dcp_scan_run
    x->epsilon = hdr.epsilon;
    x->is_rna = hdr.is_rna;
    x->abc_name = hdr.is_rna ? "rna" : "dna";
    int wm = 1;
    if ((rc = world_max(&wm))) return rc;
    Shard sh;
    sh.device = first_device;
    if ((rc = dcpgpu_open(&sh.gpu, sh.device))) return map_gpu_error(rc);
    x->shards.push_back(sh);
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
  int const rc = parse_db(dbfile, &hdr, false, [&](DbHeader const &, int, ProteinRecord &rec) -> int {
    total += rec.info.K;
    ++n;
    return 0;
  });
  if (rc) return rc;
  if (num_proteins) *num_proteins = n;
  if (total_core_size) *total_core_size = total;
  if (epsilon) *epsilon = hdr.epsilon;
  return 0;
}

int dcpb200_scan_num_gpus(struct dcp_scan const *x)
{
  if (!x || x->shards.empty()) return 0;
  return x->shards.back().device - x->shards.front().device + 1;
}

int dcpb200_scan_num_shards(struct dcp_scan const *x) { return x ? (int)x->shards.size() : 0; }

double dcpb200_scan_counter(struct dcp_scan const *x, int what)
{
  if (!x) return 0;
  if (what == 3) return (double)x->cells;
  if (what == 4) return (double)x->windows;
  if (what == 5) return (double)x->lrt_windows;
  if (what == 6) return (double)x->speculative_windows;
  double v = 0;
  for (auto const &sh : x->shards) v += dcpgpu_counter(sh.gpu, what);
  return v;
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

} // extern "C"

namespace {

// One shard's share of dcp_scan_run: every chunk of its profiles against the whole batch.
int run_shard(dcp_scan *x, size_t shard_index, dcp_batch const *batch, std::vector<uint8_t> const &symbols,
              std::vector<int64_t> const &offsets, std::vector<Row> *rows)
{
  Shard const &sh = x->shards[shard_index];
  dcpgpu_ctx *gpu = sh.gpu;
  int const S = (int)batch->seqs.size();
  uint32_t const flags = (x->multi_hits ? DCPGPU_MULTI_HITS : 0u) | (x->hmmer3_compat ? DCPGPU_HMMER3_COMPAT : 0u);
  int rc;
  PhaseTimer tm;
  bool const timing = getenv("DCP_TIMING") != nullptr;
  if ((rc = dcpgpu_reads_set(gpu, S, symbols.data(), offsets.data()))) return map_gpu_error(rc);
  if (sh.p1 <= sh.p0 || S == 0) return 0;
  tm.lap(0);

  // cells of the first windows of one profile against the whole batch
  auto profile_cells = [&](int p) {
    int const K = x->profiles[(size_t)p].K, w = std::min(K * 50, 100000);
    double c = 0;
    for (int s = 0; s < S; ++s) c += (double)std::min<int64_t>(w, offsets[(size_t)s + 1] - offsets[(size_t)s]) * K;
    return c;
  };

  // trace the hits of one wave, turn them into rows and next-window state
  auto process_hits = [&](std::vector<dcpgpu_pair> const &pairs, std::vector<int> const &win_idx,
                          std::vector<float> const &nulc, std::vector<float> const &altc,
                          std::vector<Active *> const &owner) -> int {
    size_t i0 = 0;
    while (i0 < pairs.size())
    { // in rounds bounded by the bytes of their value dumps
      size_t i1 = i0;
      double bytes = 0;
      while (i1 < pairs.size())
      {
        double const b = (double)(pairs[i1].len + 1) * (12.0 * x->profiles[(size_t)(sh.p0 + pairs[i1].profile)].K + 32.0);
        if (i1 > i0 && bytes + b > 16e9) break;
        bytes += b;
        ++i1;
      }
      size_t const n = i1 - i0;
      tm.lap(7);
      if ((rc = dcpgpu_trace_pairs(gpu, (int64_t)n, &pairs[i0], flags, nullptr, nullptr))) return map_gpu_error(rc);
      tm.lap(2);
      std::vector<int32_t> hit(n), hstart(n), hstop(n);
      std::vector<int64_t> toff(n + 1, 0);
      rc = dcpgpu_match_build(gpu, x->epsilon, x->is_rna ? 1 : 0, hit.data(), hstart.data(), hstop.data(), toff.data());
      if (rc == DCPGPU_EDECODE) return DCP_EDECODON; // decoder.c:52-56
      if (rc) return map_gpu_error(rc);
      std::vector<char> text((size_t)toff[n] + 1);
      if ((rc = dcpgpu_match_fetch(gpu, text.data()))) return map_gpu_error(rc);
      tm.lap(3);
      for (size_t i = 0; i < n; ++i)
      {
        if (!hit[i]) continue; // no B..E segment: thread.c:136,148
        dcpgpu_pair const &pr = pairs[i0 + i];
        int const gp = sh.p0 + pr.profile;
        ProfileInfo const &pm = x->profiles[(size_t)gp];
        Sequence const &sq = batch->seqs[(size_t)pr.seq];
        if (owner[i0 + i]) owner[i0 + i]->last_hit = hstop[i] - 1; // thread.c:162
        // row (product_thread.c:40-79); the match column comes from the device
        float const null_ll = -nulc[i0 + i], alt_ll = -altc[i0 + i];
        float const lrt = -2 * (null_ll - alt_ll); // lrt.h:6-9
        char head[256];
        snprintf(head, sizeof head, "%ld\t%d\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%.1f\t%.2g\t", sq.id, win_idx[i0 + i], pr.start,
                 pr.start + pr.len, 0, hstart[i], hstop[i], pm.accession.c_str(), x->abc_name.c_str(), (double)lrt, 0.0);
        std::string row(head);
        row.append(text.data() + toff[i], (size_t)(toff[i + 1] - toff[i]));
        row.push_back('\n');
        std::string fasta;
        if (x->write_aminos)
        {
          char tag[192];
          snprintf(tag, sizeof tag, ">%ld/%d/%s window=[%d,%d) hit=[%d,%d)\n", sq.id, win_idx[i0 + i], pm.accession.c_str(),
                   pr.start, pr.start + pr.len, hstart[i], hstop[i]);
          fasta = tag;
          fasta += amino_of_match(text.data() + toff[i], (size_t)(toff[i + 1] - toff[i]));
          fasta.push_back('\n');
        }
        rows->push_back(Row{gp, pr.seq, win_idx[i0 + i], std::move(row), std::move(fasta)});
      }
      tm.lap(4);
      i0 = i1;
    }
    return 0;
  };

  int c0 = sh.p0;
  while (c0 < sh.p1 && !x->interrupted)
  {
    // ---- the chunk [c0, c1): about chunk_cells DP cells of first windows ----
    // (the second shard of a GPU starts with half a chunk: its host phases then fall into the
    // other shard's score passes instead of coinciding with them)
    double const target = (c0 == sh.p0 && sh.stagger) ? x->chunk_cells / 2 : x->chunk_cells;
    int c1 = c0;
    double cells = 0;
    do
    {
      cells += profile_cells(c1);
      ++c1;
    } while (c1 < sh.p1 && cells < target);
    int const P = c1 - c0, l0 = c0 - sh.p0; // shard-local profile indices l0 .. l0 + P

    // wave 0: the first window of every (sequence, profile), generated on the device
    tm.lap(7);
    if ((rc = dcpgpu_score_grid(gpu, l0, l0 + P, 0, S, flags))) return map_gpu_error(rc);
    int64_t nhits = 0;
    if ((rc = dcpgpu_hits_fetch(gpu, 0, nullptr, &nhits))) return map_gpu_error(rc);
    tm.lap(1);
    std::vector<int64_t> hit((size_t)nhits);
    if (nhits && (rc = dcpgpu_hits_fetch(gpu, nhits, hit.data(), &nhits))) return map_gpu_error(rc);
    x->windows += (long)P * S;
    x->cells += (long long)cells; // first windows of the chunk (profile_cells above)
    x->lrt_windows += (long)nhits;
    std::vector<float> hn((size_t)nhits), ha((size_t)nhits); // costs of the hits only (not the whole P x S grid)
    if (nhits && (rc = dcpgpu_scores_gather(gpu, nhits, hit.data(), hn.data(), ha.data()))) return map_gpu_error(rc);

    // pairs whose sequence is longer than the first window keep iterating (window.c:27-31)
    std::vector<Active> active;
    for (int p = 0; p < P; ++p)
    {
      int const w = std::min(x->profiles[(size_t)(c0 + p)].K * 50, 100000);
      for (int s = 0; s < S; ++s)
      {
        int const len = (int)(offsets[(size_t)s + 1] - offsets[(size_t)s]);
        if (len > w) active.push_back(Active{l0 + p, s, 0, w, 0, -1});
      }
    }
    std::map<std::pair<int, int>, Active *> by_pair;
    for (auto &a : active) by_pair[{a.profile, a.seq}] = &a;

    std::vector<dcpgpu_pair> hp((size_t)nhits);
    std::vector<int> widx((size_t)nhits, 0);
    std::vector<Active *> owner((size_t)nhits, nullptr);
    for (int64_t i = 0; i < nhits; ++i)
    {
      int const p = l0 + (int)(hit[(size_t)i] / S), s = (int)(hit[(size_t)i] % S);
      int const len = (int)(offsets[(size_t)s + 1] - offsets[(size_t)s]);
      hp[(size_t)i] = dcpgpu_pair{p, s, 0, std::min(len, std::min(x->profiles[(size_t)(sh.p0 + p)].K * 50, 100000))};
      auto it = by_pair.find({p, s});
      if (it != by_pair.end()) owner[(size_t)i] = it->second;
    }
    tm.lap(5);
    if ((rc = process_hits(hp, widx, hn, ha, owner))) return rc;

    // Later windows.  A pair's next window depends on the hit found in the previous one only
    // through last_hit_pos (window.c:13-37, thread.c:162), and most windows hold no hit: every
    // round PLANS each active pair's whole remaining chain of windows as if none of them had a
    // hit, scores them all in one pass, commits every pair's windows up to and including its first
    // one that passes the lrt gate, traces those, and replans from there.  Rounds = 1 + the largest
    // number of gate-passing windows of a pair, instead of one GPU pass per window index.
    while (!active.empty() && !x->interrupted)
    {
      std::vector<dcpgpu_pair> wp;
      std::vector<int> pidx;                            // window index of every planned window
      std::vector<size_t> first(active.size() + 1, 0);  // planned windows of active[i]: [first[i], first[i+1])
      for (size_t i = 0; i < active.size(); ++i)
      {
        Active t = active[i];
        int const len = (int)(offsets[(size_t)t.seq + 1] - offsets[(size_t)t.seq]);
        int const K = x->profiles[(size_t)(sh.p0 + t.profile)].K;
        while (window_next(t, len, K))
        {
          wp.push_back(dcpgpu_pair{t.profile, t.seq, t.start, t.stop - t.start});
          pidx.push_back(t.idx);
        }
        first[i + 1] = wp.size();
      }
      if (wp.empty()) break;
      std::vector<float> nulc(wp.size()), altc(wp.size());
      tm.lap(7);
      if ((rc = dcpgpu_score_pairs(gpu, (int64_t)wp.size(), wp.data(), flags, nulc.data(), altc.data())))
        return map_gpu_error(rc);
      tm.lap(6);
      std::vector<Active> next;
      std::vector<size_t> gate; // planned index of the gate-passing window of next[j]
      next.reserve(active.size());
      long committed = 0;
      long long ccells = 0;
      for (size_t i = 0; i < active.size(); ++i)
      {
        int const Ki = x->profiles[(size_t)(sh.p0 + active[i].profile)].K;
        size_t j = first[i];
        for (; j < first[i + 1]; ++j)
        {
          float const lrt = -2 * ((-nulc[j]) - (-altc[j]));
          if (std::isfinite(lrt) && lrt >= 0) break; // thread.c:121
        }
        if (j == first[i + 1])
        { // no further candidate: the pair's chain ends here, every planned window stands
          committed += (long)(first[i + 1] - first[i]);
          for (size_t w = first[i]; w < first[i + 1]; ++w) ccells += (long long)wp[w].len * Ki;
          continue;
        }
        committed += (long)(j - first[i] + 1);
        for (size_t w = first[i]; w <= j; ++w) ccells += (long long)wp[w].len * Ki;
        Active t = active[i]; // the committed state: window j (last_hit_pos carried over, window.c never resets it)
        t.start = wp[j].start;
        t.stop = wp[j].start + wp[j].len;
        t.idx = pidx[j];
        next.push_back(t);
        gate.push_back(j);
      }
      std::vector<dcpgpu_pair> hp2(next.size());
      std::vector<int> widx2(next.size());
      std::vector<float> hn2(next.size()), ha2(next.size());
      std::vector<Active *> owner2(next.size());
      for (size_t j = 0; j < next.size(); ++j)
      {
        hp2[j] = wp[gate[j]];
        widx2[j] = pidx[gate[j]];
        hn2[j] = nulc[gate[j]];
        ha2[j] = altc[gate[j]];
        owner2[j] = &next[j];
      }
      x->windows += committed;
      x->cells += ccells;
      x->lrt_windows += (long)hp2.size();
      x->speculative_windows += (long)wp.size() - committed;
      if ((rc = process_hits(hp2, widx2, hn2, ha2, owner2))) return rc;
      active.swap(next);
      if (shard_index == 0 && x->callback) x->callback(x->userdata);
    }

    x->done_proteins += P; // thread.c:81-82
    if (shard_index == 0 && x->callback) x->callback(x->userdata); // rank 0 only, scan.c:196-198
    c0 = c1;
  }
  tm.lap(7);
  if (timing)
    fprintf(stderr,
            "[dcp_scan_run shard %zu] reads_set %.3f  grid+hit count %.3f  trace %.3f  match build+fetch %.3f  rows %.3f  "
            "scores fetch + hit lists %.3f  later-window score passes %.3f  other host %.3f s\n",
            shard_index, tm.t[0], tm.t[1], tm.t[2], tm.t[3], tm.t[4], tm.t[5], tm.t[6], tm.t[7]);
  return 0;
}

} // namespace

extern "C" {

int dcp_scan_run(struct dcp_scan *x, struct dcp_batch *batch, char const *product_dir)
{
  if (!x || !batch || !product_dir || x->shards.empty()) return DCP_EFUNCUSE;
  x->interrupted = false;
  x->done_proteins = 0;
  int const S = (int)batch->seqs.size();

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
  if (!x->profiles.empty())
    for (int s = 0; s < S; ++s)
      if (batch->seqs[(size_t)s].data.empty()) return DCP_EZEROSEQ;

  // product_open (product.c:15-32)
  std::string const dir(product_dir);
  int rc;
  if ((rc = mkdir_p(dir)) || (rc = mkdir_p(dir + "/hmmer"))) return rc;

  // one host thread per shard (the reference: one OpenMP thread per partition, scan.c:188-208)
  size_t const W = x->shards.size();
  std::vector<std::vector<Row>> rows(W);
  std::vector<int> rcs(W, 0);
  if (W == 1) rcs[0] = run_shard(x, 0, batch, symbols, offsets, &rows[0]);
  else
  {
    std::vector<std::thread> threads;
    for (size_t i = 0; i < W; ++i)
      threads.emplace_back([&, i] {
        rcs[i] = run_shard(x, i, batch, symbols, offsets, &rows[i]);
        if (rcs[i]) x->interrupted = true; // scan.c:199-203: an error stops every partition
      });
    for (auto &t : threads) t.join();
  }
  for (int r : rcs)
    if (r) return r;

  // product_close (product.c:34-87): header + the shards' rows in shard (= profile) order, each
  // shard's rows in (profile, batch order, window) order.  The reference concatenates per-thread
  // files; here every shard's thread sorts its rows and writes them at the shard's offset of the
  // one file (pwrite), so the N shards' text (hundreds of MB per batch at 8 GPUs) goes out in parallel.
  static char const header[] = "sequence\twindow\twindow_start\twindow_stop\thit\thit_start\thit_stop\tprofile\tabc\tlrt\tevalue\tmatch\n";
  std::vector<size_t> bytes(W, 0);
  auto sort_rows = [&](size_t i) {
    std::stable_sort(rows[i].begin(), rows[i].end(), [](Row const &a, Row const &b) {
      if (a.profile != b.profile) return a.profile < b.profile;
      if (a.seq_order != b.seq_order) return a.seq_order < b.seq_order;
      return a.window < b.window;
    });
    for (auto const &r : rows[i]) bytes[i] += r.text.size();
  };
  int const fd = open((dir + "/products.tsv").c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) return DCP_EWRITEPROD;
  std::vector<char> okv(W, 1);
  auto write_rows = [&](size_t i, size_t offset) {
    std::string buf;
    buf.reserve(std::min<size_t>(bytes[i], size_t(8) << 20) + 65536);
    auto flush = [&] {
      size_t done = 0;
      while (done < buf.size())
      {
        ssize_t const n = pwrite(fd, buf.data() + done, buf.size() - done, (off_t)(offset + done));
        if (n <= 0) { okv[i] = 0; return; }
        done += (size_t)n;
      }
      offset += buf.size();
      buf.clear();
    };
    for (auto const &r : rows[i])
    {
      buf += r.text;
      if (buf.size() >= (size_t(8) << 20)) flush();
    }
    flush();
  };
  bool ok = pwrite(fd, header, sizeof header - 1, 0) == (ssize_t)(sizeof header - 1);
  if (W == 1)
  {
    sort_rows(0);
    write_rows(0, sizeof header - 1);
  }
  else
  {
    std::vector<std::thread> ts;
    for (size_t i = 0; i < W; ++i) ts.emplace_back(sort_rows, i);
    for (auto &t : ts) t.join();
    ts.clear();
    size_t off = sizeof header - 1;
    for (size_t i = 0; i < W; ++i)
    {
      ts.emplace_back(write_rows, i, off);
      off += bytes[i];
    }
    for (auto &t : ts) t.join();
  }
  for (char c : okv) ok = ok && c;
  ok = (close(fd) == 0) && ok;
  if (ok && x->write_aminos)
  { // the sequences the reference sends to the HMMER daemon one hit at a time (hmmer.c:83-108), as ONE
    // FASTA file in row order: the input of a single batched hmmscan / hmmsearch --cut_ga run
    FILE *fa = fopen((dir + "/hmmer/aminos.fa").c_str(), "wb");
    if (!fa) return DCP_EWRITEPROD;
    for (auto const &rs : rows)
      for (auto const &r : rs) ok = ok && fwrite(r.amino.data(), 1, r.amino.size(), fa) == r.amino.size();
    ok = (fclose(fa) == 0) && ok;
  }
  return ok ? 0 : DCP_EWRITEPROD;
}

} // extern "C"
