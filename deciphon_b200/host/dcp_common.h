// dcp_common.h -- MessagePack-subset reader and writer of the .dcp database format, shared by
// the scan (database_reader.c:26-80, protein.c:283-351) and the press (database_writer.c:136-193,
// protein.c:234-281, write.c).  Layout in SURVEY App. A.7.
#pragma once
#include "gencode.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dcpb {

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

using dcpb::NucltDist;

inline bool read_nuclt_dist(Reader &r, NucltDist *d)
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

// ---- writer (lite-pack's choices: the smallest MessagePack form of every value, write.c) --------

struct Writer
{
  std::string out;
  void byte(unsigned v) { out.push_back((char)(unsigned char)v); }
  void be(uint64_t v, int n)
  {
    for (int i = n - 1; i >= 0; --i) byte((unsigned)((v >> (8 * i)) & 0xff));
  }
  void str(std::string const &s)
  { // write_cstring
    size_t const n = s.size();
    if (n < 32) byte(0xa0 | (unsigned)n);
    else if (n < 256) { byte(0xd9); be(n, 1); }
    else if (n < 65536) { byte(0xda); be(n, 2); }
    else { byte(0xdb); be(n, 4); }
    out.append(s);
  }
  void map(uint32_t n)
  {
    if (n < 16) byte(0x80 | n);
    else if (n < 65536) { byte(0xde); be(n, 2); }
    else { byte(0xdf); be(n, 4); }
  }
  void array(uint32_t n)
  {
    if (n < 16) byte(0x90 | n);
    else if (n < 65536) { byte(0xdc); be(n, 2); }
    else { byte(0xdd); be(n, 4); }
  }
  void uint(uint64_t v)
  { // write_int of a non-negative value
    if (v < 128) byte((unsigned)v);
    else if (v < 256) { byte(0xcc); be(v, 1); }
    else if (v < 65536) { byte(0xcd); be(v, 2); }
    else if (v < (uint64_t(1) << 32)) { byte(0xce); be(v, 4); }
    else { byte(0xcf); be(v, 8); }
  }
  void boolean(bool v) { byte(v ? 0xc3 : 0xc2); }
  void f32(float v)
  { // write_float
    uint32_t u;
    memcpy(&u, &v, 4);
    byte(0xca);
    be(u, 4);
  }
  void f32bin(float const *a, size_t count)
  { // write_f32array: bin + host-endian floats (write.c:59-66)
    size_t const n = count * 4;
    if (n < 256) { byte(0xc4); be(n, 1); }
    else if (n < 65536) { byte(0xc5); be(n, 2); }
    else { byte(0xc6); be(n, 4); }
    out.append(reinterpret_cast<char const *>(a), n);
  }
  void f32list(float const *a, size_t count)
  { // array of float32 values (how the current imm packs base log-probs and codon marginals:
    // 650 bytes per nuclt_dist, the size test_press.c:26 implies)
    array((uint32_t)count);
    for (size_t i = 0; i < count; ++i) f32(a[i]);
  }
  void bin8(unsigned char const *a, size_t n)
  {
    byte(0xc4);
    be(n, 1);
    out.append(reinterpret_cast<char const *>(a), n);
  }
};

} // namespace dcpb
