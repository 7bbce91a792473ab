// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of BioD's BGZF-inflate -> BAM record decode -> pileup path.
// Nothing in the product (biod_b200/, include/) links, imports or executes this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and there only as the checker / CPU baseline.
//
// Parity pin: the reference is D and cannot be compiled in this image (no dmd /
// ldc2 / gdc), so this restatement is pinned against the reference's OWN golden
// vectors instead (tests/test_oracle_golden.py): ex1_header.sam (3270 records,
// test/unittests.d:303-312), first-record fields (unittests.d:92-103), the
// pileup column counts {1470,1567} and first columns (unittests.d:334-367), the
// in-module 10-read pileup vector (bam/pileup.d:735-825), the zero-coverage
// vector (pileup.d:830-856), examples/make_pileup.d:20-30 and the five corrupted
// BAMs (unittests.d:132-142).  Inflate itself is system libz — the very library
// the reference links (bio/core/utils/zlib.d:6-245, Makefile:10 `-L-lz`).
//
// Every function cites the reference lines it follows.  Paths are relative to
// /root/reference.

#include <zlib.h>

#include <cfloat>
#include <cmath>
#include <cctype>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

enum {
  ORC_OK = 0,
  ORC_ERR_BGZF = -1,    // BgzfException  (bio/core/bgzf/inputstream.d:41-43)
  ORC_ERR_ZLIB = -2,    // ZlibException  (bio/core/utils/zlib.d:247-274)
  ORC_ERR_FORMAT = -3,  // plain Exception (bam/reader.d:113, bgzf/block.d:150)
  ORC_ERR_TRUNC = -4,   // ReadException from readExact (bam/readrange.d:169)
  ORC_ERR_CIGAR = -7,   // PileupRead.assertCigarIndexIsValid (bam/pileup.d:224-228)
  ORC_ERR_UNSORTED = -8, // not a reference error: input order the engine assumes
  ORC_ERR_ARG = -9      // enforce() on arguments (randomaccessmanager.d:206-208)
};

const uint32_t BGZF_MAX_BLOCK_SIZE = 65536;  // bio/core/bgzf/constants.d:60

struct Block {
  uint64_t coffset;   // start_offset          (bgzf/block.d:44)
  uint32_t bsize;     // total block size - 1  (block.d:47)
  uint32_t xlen;
  uint32_t cdata_size;
  uint32_t crc32;
  uint32_t isize;     // input_size
  uint64_t uoffset;   // prefix sum of isize over the blocks delivered so far
  uint64_t payload;   // file offset of the deflate payload
};

struct Error {
  int status = ORC_OK;
  int zerr = 0;
  uint64_t offset = 0;
  std::string msg;
};

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// fillBgzfBufferFromStream — bio/core/bgzf/inputstream.d:54-199.
// Returns 1 = block parsed, 0 = clean end of stream (:57-58, :75-76), <0 = error.
int parse_bgzf_header(const uint8_t* d, uint64_t len, uint64_t pos, Block* b, Error* e) {
  auto fail = [&](const std::string& m) {
    e->status = ORC_ERR_BGZF;
    e->offset = pos;
    e->msg = "Error reading BGZF block starting from offset " + std::to_string(pos) + ": " + m;  // :63-64
    return ORC_ERR_BGZF;
  };
  const char* kStream = "stream error: not enough data in stream";  // :194-196
  if (pos >= len) return 0;                       // stream.eof() (:57)
  // :72-79 — a short read of the 4 magic bytes is a clean EOF only if *zero*
  // bytes were available; a partial magic keeps looping until read returns 0.
  if (len - pos < 4) return 0;
  if (!(d[pos] == 0x1f && d[pos + 1] == 0x8b && d[pos + 2] == 0x08 && d[pos + 3] == 0x04))
    return fail("wrong BGZF magic");              // :82-84
  uint64_t p = pos + 4 + 4 + 1 + 1;               // skip MTIME, XFL, OS (:88-97)
  if (p + 2 > len) return fail(kStream);
  uint32_t xlen = le16(d + p);                    // :99
  p += 2;
  uint32_t bsize = 0;
  bool found = false;
  uint32_t n = 0;
  while (n < xlen) {                              // :106
    if (p + 4 > len) return fail(kStream);
    uint8_t si1 = d[p], si2 = d[p + 1];
    uint32_t slen = le16(d + p + 2);
    p += 4;
    if (si1 == 66 && si2 == 67) {                 // 'B','C' (:115)
      if (slen != 2)
        return fail("wrong BC subfield length: " + std::to_string(slen) + "; expected 2");  // :118-121
      if (found) return fail("duplicate field with block size");                            // :123-125
      if (p + 2 > len) return fail(kStream);
      bsize = le16(d + p);
      found = true;
    }
    p += slen;                                    // :131-142 (seekCur never fails by itself)
    n += 4 + slen;                                // :145-148
  }
  if (n != xlen)
    return fail("total length of subfields in bytes (" + std::to_string(n) +
                ") is not equal to gzip_extra_length (" + std::to_string(xlen) + ")");  // :152-157
  if (!found) return fail("block size was not found in any subfield");                  // :159-161
  int64_t cdata = (int64_t)bsize - (int64_t)xlen - 19;                                  // :164
  if (cdata > (int64_t)BGZF_MAX_BLOCK_SIZE)
    return fail("compressed data size is more than 65536 bytes, which is not allowed by current BAM specification");
  // :176 readExact(buffer, cdata_size): a negative size converts to a huge
  // size_t and the stream throws ReadException -> "stream error" (:194-196).
  if (cdata < 0 || p + (uint64_t)cdata + 8 > len) return fail(kStream);
  b->coffset = pos;
  b->bsize = bsize;
  b->xlen = xlen;
  b->cdata_size = (uint32_t)cdata;
  b->payload = p;
  b->crc32 = le32(d + p + cdata);                 // :184
  b->isize = le32(d + p + cdata + 4);             // :185
  return 1;
}

// decompressBgzfBlock — bio/core/bgzf/block.d:127-216.
int inflate_block(const uint8_t* d, const Block& b, uint8_t* out, Error* e) {
  if (b.isize > BGZF_MAX_BLOCK_SIZE) {            // :150-152
    e->status = ORC_ERR_FORMAT;
    e->offset = b.coffset;
    e->msg = "Uncompressed block size must be within 65536 bytes";
    return e->status;
  }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  zs.next_in = const_cast<Bytef*>(d + b.payload);
  zs.avail_in = b.cdata_size;
  int err = inflateInit2(&zs, -15);               // :162
  if (err) { e->status = ORC_ERR_ZLIB; e->zerr = err; e->offset = b.coffset; e->msg = "zlib init"; return e->status; }
  zs.next_out = out;
  zs.avail_out = b.isize;                         // :170
  err = inflate(&zs, Z_FINISH);                   // :172
  uint64_t total_out = zs.total_out;
  inflateEnd(&zs);
  if (err != Z_STREAM_END) {                      // :182-185
    e->status = ORC_ERR_ZLIB;
    e->zerr = err;
    e->offset = b.coffset;
    e->msg = "zlib error " + std::to_string(err);
    return e->status;
  }
  // :175 `assert(zs.total_out == block.input_size)` and :187 the CRC assert are
  // compiled out by -release (Makefile:33); a short stream would hand garbage
  // stack bytes to the caller.  Restatement-defined: report it as Z_DATA_ERROR.
  if (total_out != b.isize) {
    e->status = ORC_ERR_ZLIB;
    e->zerr = Z_DATA_ERROR;
    e->offset = b.coffset;
    e->msg = "inflated size differs from ISIZE";
    return e->status;
  }
  return ORC_OK;
}

struct Bam {
  const uint8_t* file = nullptr;
  uint64_t flen = 0;
  Error err;            // first error met (sticky)
  // BGZF stream state: BgzfInputStream — inputstream.d:349-541
  std::vector<Block> blocks;   // data blocks delivered, in order
  std::vector<uint8_t> u;      // concatenated uncompressed bytes
  uint64_t next_coffset = 0;
  bool stream_done = false;    // supplier exhausted / ISIZE==0 block met (:393-394)
  uint64_t end_coffset = 0;    // end_offset of the last delivered block
  // header
  std::string text;
  std::vector<std::string> ref_names;
  std::vector<int32_t> ref_lens;
  uint64_t reads_start_u = 0;  // offset into u of the first record
  uint64_t reads_start_vo = 0;
  // records
  bool decoded = false;
  std::vector<uint64_t> rec_off;   // offset in u of each record's block_size prefix
  std::vector<int32_t> block_size, ref_id, pos, end_pos, l_seq, next_ref, next_pos, tlen;
  std::vector<uint16_t> bin, flag, n_cigar;
  std::vector<uint8_t> mapq, l_read_name;
  std::vector<uint64_t> start_vo, end_vo;
  std::vector<uint64_t> cigar_off;
  std::vector<uint32_t> cigar;
};

// One step of fillNextBlock + decompress (inputstream.d:386-424, block.d:127).
// Returns 1 if a data block was appended, 0 at end of stream, <0 on error.
int next_block(Bam* s) {
  if (s->stream_done) return 0;
  if (s->err.status) return s->err.status;
  Block b;
  int r = parse_bgzf_header(s->file, s->flen, s->next_coffset, &b, &s->err);
  if (r < 0) return r;
  if (r == 0) { s->stream_done = true; return 0; }
  if (b.isize == 0) { s->stream_done = true; return 0; }   // BGZF EOF block (:393-394)
  b.uoffset = s->u.size();
  // enforce ISIZE <= 65536 before growing the buffer (block.d:150)
  if (b.isize > BGZF_MAX_BLOCK_SIZE) { inflate_block(s->file, b, nullptr, &s->err); return s->err.status; }
  s->u.resize(b.uoffset + b.isize);
  r = inflate_block(s->file, b, s->u.data() + b.uoffset, &s->err);
  if (r < 0) { s->u.resize(b.uoffset); return r; }
  s->blocks.push_back(b);
  s->next_coffset = b.coffset + b.bsize + 1;      // end_offset (block.d:53-55)
  s->end_coffset = s->next_coffset;
  return 1;
}

// Make at least `upto` uncompressed bytes available (or hit EOF / an error).
int ensure(Bam* s, uint64_t upto) {
  while (s->u.size() < upto) {
    int r = next_block(s);
    if (r <= 0) return r;
  }
  return 1;
}

// VirtualOffset of stream offset x — virtualoffset.d:43, with the normalisation
// of inputstream.d:443-445,516-524: an offset that is exactly a block end is
// reported as (next block's coffset, 0).
uint64_t voffset_of(const Bam* s, uint64_t x) {
  const auto& bl = s->blocks;
  size_t lo = 0, hi = bl.size();
  while (lo < hi) {            // last block with uoffset <= x
    size_t m = (lo + hi) / 2;
    if (bl[m].uoffset <= x) lo = m + 1; else hi = m;
  }
  if (lo == 0) return 0;
  const Block& b = bl[lo - 1];
  if (x - b.uoffset >= b.isize) return (uint64_t)(b.coffset + b.bsize + 1) << 16;
  return (b.coffset << 16) | (x - b.uoffset);
}

// BamReader ctor — bam/reader.d:100-124, 579-597; referenceinfo.d:57-62.
int parse_header(Bam* s) {
  auto need = [&](uint64_t upto) -> int {
    int r = ensure(s, upto);
    if (r < 0) return r;
    if (s->u.size() < upto) {
      s->err.status = ORC_ERR_TRUNC;
      s->err.msg = "not enough data in stream";
      return s->err.status;
    }
    return 1;
  };
  int r;
  // readString(4) on an empty/short stream yields a short string -> enforce fails.
  r = ensure(s, 4);
  if (r < 0) return r;
  if (s->u.size() < 4 || memcmp(s->u.data(), "BAM\1", 4) != 0) {   // reader.d:111-113
    s->err.status = ORC_ERR_FORMAT;
    s->err.msg = "Invalid file format: expected BAM\\1";
    return s->err.status;
  }
  uint64_t p = 4;
  if ((r = need(p + 4)) < 0) return r;
  int32_t l_text = (int32_t)le32(s->u.data() + p);                  // :580-581
  p += 4;
  if (l_text < 0) { s->err.status = ORC_ERR_FORMAT; s->err.msg = "negative l_text"; return s->err.status; }
  if ((r = need(p + l_text)) < 0) return r;
  s->text.assign((const char*)s->u.data() + p, l_text);             // :583
  p += l_text;
  if ((r = need(p + 4)) < 0) return r;
  int32_t n_ref = (int32_t)le32(s->u.data() + p);                   // :588-589
  p += 4;
  for (int32_t i = 0; i < n_ref; i++) {
    if ((r = need(p + 4)) < 0) return r;
    int32_t l_name = (int32_t)le32(s->u.data() + p);                // referenceinfo.d:58-59
    p += 4;
    if (l_name < 0) { s->err.status = ORC_ERR_FORMAT; s->err.msg = "negative l_name"; return s->err.status; }
    if ((r = need(p + l_name + 4)) < 0) return r;
    std::string nm((const char*)s->u.data() + p, l_name);           // keeps the NUL (:60)
    if (!nm.empty() && nm.back() == '\0') nm.pop_back();            // name() drops it (:41-43)
    s->ref_names.push_back(nm);
    p += l_name;
    s->ref_lens.push_back((int32_t)le32(s->u.data() + p));          // :61
    p += 4;
  }
  s->reads_start_u = p;
  // reader.d:121-123: virtualTell() right after the header.  When the header
  // ends exactly at a block end the stream has already stepped to the next
  // block (inputstream.d:516-524) — which needs that block to be known.
  ensure(s, p + 1);
  if (s->err.status) {
    // the error belongs to a later block; it must surface during iteration
    // (unittests.d:139), not here.  Keep it sticky but report open as OK.
  }
  s->reads_start_vo = voffset_of(s, p);
  return ORC_OK;
}

inline bool op_query(uint32_t raw) { return ((0x3C1A7u >> ((raw & 0xF) * 2)) & 1) != 0; }  // cigar.d:116-121
inline bool op_ref(uint32_t raw)   { return ((0x3C1A7u >> ((raw & 0xF) * 2)) & 2) != 0; }  // cigar.d:116,124-126
// 0x3C1A7 == 0b11_11_00_00_01_10_10_01_11 (cigar.d:116); op codes 9..15 shift past bit 17 -> 0.

// BamReadRange.readNext — bam/readrange.d:118-173, fields — bam/read.d:907-1003,
// basesCovered — read.d:255-262, end_position — read.d:1380-1383.
int decode_records(Bam* s) {
  if (s->decoded) return s->err.status;
  s->decoded = true;
  // inflate everything first (errors are sticky and re-raised below at the
  // point the sequential reader would have met them)
  while (next_block(s) > 0) {}
  Error blk_err = s->err;
  const uint64_t avail = s->u.size();
  uint64_t p = s->reads_start_u;
  s->cigar_off.push_back(0);
  while (true) {
    if (p >= avail) break;                      // stream.eof() / 0 bytes read (:125,:141-148)
    if (avail - p < 4) {
      // readBlock returned fewer than 4 bytes and then 0: range ends (:139-150),
      // unless the stream stopped because of an error in the next block.
      break;
    }
    int32_t bs = (int32_t)le32(s->u.data() + p);
    if (bs < 32 || (uint64_t)bs > avail - p - 4) {
      if (blk_err.status) break;                // the pending block error wins (raised below)
      s->err.status = ORC_ERR_TRUNC;            // readExact throws (:169); <32 would be a RangeError in BamRead
      s->err.offset = p;
      s->err.msg = bs < 32 ? "record shorter than the fixed 32-byte core" : "not enough data in stream";
      break;
    }
    const uint8_t* r = s->u.data() + p + 4;
    uint32_t bin_mq_nl = le32(r + 8), flag_nc = le32(r + 12);
    int32_t lseq = (int32_t)le32(r + 16);
    uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
    uint64_t need = 32ull + lname + 4ull * nc + (lseq > 0 ? (uint64_t)(lseq + 1) / 2 + lseq : 0);
    if (lseq < 0 || need > (uint64_t)bs) {
      s->err.status = ORC_ERR_TRUNC;
      s->err.offset = p;
      s->err.msg = "record fields exceed block_size";
      break;
    }
    s->rec_off.push_back(p);
    s->block_size.push_back(bs);
    s->ref_id.push_back((int32_t)le32(r));
    int32_t pos = (int32_t)le32(r + 4);
    s->pos.push_back(pos);
    s->bin.push_back((uint16_t)(bin_mq_nl >> 16));
    s->mapq.push_back((uint8_t)(bin_mq_nl >> 8));
    s->l_read_name.push_back((uint8_t)lname);
    uint16_t flag = (uint16_t)(flag_nc >> 16);
    s->flag.push_back(flag);
    s->n_cigar.push_back((uint16_t)nc);
    s->l_seq.push_back(lseq);
    s->next_ref.push_back((int32_t)le32(r + 20));
    s->next_pos.push_back((int32_t)le32(r + 24));
    s->tlen.push_back((int32_t)le32(r + 28));
    uint32_t covered = 0;
    for (uint32_t k = 0; k < nc; k++) {
      uint32_t raw = le32(r + 32 + lname + 4 * k);
      s->cigar.push_back(raw);
      if (op_ref(raw)) covered += raw >> 4;
    }
    if (flag & 0x4) covered = 0;                // read.d:257-259
    s->cigar_off.push_back(s->cigar.size());
    s->end_pos.push_back((int32_t)((uint32_t)pos + covered));
    s->start_vo.push_back(voffset_of(s, p));    // readrange.d:64-66
    p += 4 + (uint64_t)bs;
    s->end_vo.push_back(voffset_of(s, p));      // readrange.d:55-57
  }
  if (!s->err.status && blk_err.status) s->err = blk_err;
  return s->err.status;
}

// ---------------------------------------------------------------- pileup ----

// ---- reference bases from MD tags (oracle only so far: the CUDA path is the next step) ----
// Base16(char) -> char: bio/core/base.d:39-62 (_char2code) and :85 (_code2char)
inline char norm16(uint8_t ch) {
  static const uint8_t low[128] = {
      15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,15, 15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,
      15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,15,  1, 2, 4, 8,15,15,15,15,15,15,15,15,15, 0,15,15,
      15, 1,14, 2,13,15,15, 4,11,15,15,12,15, 3,15,15, 15,15, 5, 6, 8,15, 7, 9,15,10,15,15,15,15,15,15,
      15, 1,14, 2,13,15,15, 4,11,15,15,12,15, 3,15,15, 15,15, 5, 6, 8,15, 7, 9,15,10,15,15,15,15,15,15};
  return "=ACMGRSVTWYHKDBN"[ch < 128 ? low[ch] : 15];
}

// BamRead["MD"] as a string: TagStorage.opIndex / skipValue (read.d:1070-1087, 1219-1230), charToSizeof
// (tagvalue.d:106-119).  Restatement-defined: a value that is not of type Z, or a tag area that ends inside a value,
// gives "no MD tag" (BioD would throw or read out of bounds).
bool find_md(const Bam* s, uint32_t r, std::string* md) {
  const uint8_t* rec = s->u.data() + s->rec_off[r] + 4;
  const uint64_t bs = (uint64_t)s->block_size[r];
  const uint64_t lseq = (uint64_t)std::max(0, s->l_seq[r]);
  uint64_t off = 32ull + s->l_read_name[r] + 4ull * s->n_cigar[r] + (lseq + 1) / 2 + lseq;
  if (off > bs) return false;
  const uint8_t* t = rec + off;
  const uint64_t n = bs - off;
  if (n < 4) return false;
  uint64_t o = 0;
  auto size_of = [](char c) -> int {
    switch (c) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; default: return -1; }
  };
  while (o + 1 < n) {
    const bool hit = t[o] == 'M' && t[o + 1] == 'D';
    o += 2;
    if (o >= n) return false;
    const char type = (char)t[o++];
    if (type == 'Z' || type == 'H') {
      const uint64_t b = o;
      while (o < n && t[o] != 0) ++o;
      if (o >= n) return false;
      if (hit) { if (type != 'Z') return false; md->assign((const char*)t + b, (size_t)(o - b)); return true; }
      ++o;
    } else if (type == 'B') {
      if (o + 5 > n) return false;
      const int es = size_of((char)t[o]);
      if (es < 0) return false;
      const uint64_t cnt = le32(t + o + 1);
      o += 5 + (uint64_t)es * cnt;
      if (hit || o > n) return false;
    } else {
      const int es = size_of(type);
      if (es < 0 || hit) return false;
      o += (uint64_t)es;
    }
  }
  return false;
}

struct MdOp { int type; uint32_t match; char mismatch; std::string del; };   // md/operation.d: 0 Match, 1 Mismatch, 2 Deletion

// mdOperations (md/parse.d:13-143), literally: a bidirectional range that parses its first operation from the front of
// the string and caches its LAST operation from the back at construction; forward iteration therefore yields the
// operations parsed from the front of what is left, then the cached back one.  Zero-length matches are filtered the
// way std.algorithm.filterBidirectional does (from both ends at construction, then at every popFront).
struct MdRange {
  std::string md;
  MdOp f{}, b{};
  uint8_t rem = 255;
  static bool up(char c) { return c >= 'A' && c <= 'Z'; }
  static bool dig(char c) { return c >= '0' && c <= '9'; }
  static uint32_t to_uint(const std::string& x) {        // to!uint; restatement-defined: saturates instead of throwing
    uint64_t v = 0;
    for (char c : x) { v = v * 10 + (uint64_t)(c - '0'); if (v > 0xffffffffull) v = 0xffffffffull; }
    return (uint32_t)v;
  }
  bool cache_front() {
    if (md.empty()) return false;
    if (md[0] == '^') {
      md.erase(0, 1);
      size_t len = 0;
      while (len < md.size() && up(md[len])) ++len;
      f = MdOp{2, 0, 0, md.substr(0, len)};
      md.erase(0, len);
    } else if (dig(md[0])) {
      size_t len = 0;
      while (len < md.size() && dig(md[len])) ++len;
      f = MdOp{0, to_uint(md.substr(0, len)), 0, ""};
      md.erase(0, len);
    } else {
      f = MdOp{1, 0, md[0], ""};
      md.erase(0, 1);
    }
    return true;
  }
  bool cache_back() {
    if (md.empty()) return false;
    const size_t n = md.size();
    if (dig(md[n - 1])) {
      size_t len = 0;
      while (len < n && dig(md[n - 1 - len])) ++len;
      b = MdOp{0, to_uint(md.substr(n - len)), 0, ""};
      md.erase(n - len);
    } else if (n == 1 || dig(md[n - 2])) {
      b = MdOp{1, 0, md[n - 1], ""};
      md.erase(n - 1);
    } else {
      size_t len = 0;                                        // countUntil!"a == '^'"(retro(_md)); restatement-defined: no '^' = whole string
      while (len < n && md[n - 1 - len] != '^') ++len;
      b = MdOp{2, 0, 0, md.substr(n - len)};
      md.erase(len < n ? n - len - 1 : 0);
    }
    return true;
  }
  explicit MdRange(const std::string& m) : md(m) {
    if (!cache_front()) rem = 0;
    else if (!cache_back()) { b = f; rem = 1; }
  }
  bool empty() const { return rem == 0; }
  void pop_front() {
    if (md.empty()) { if (rem == 255) { f = b; rem = 1; } else rem = 0; }
    else if (!cache_front()) rem = 0;
  }
  void pop_back() {
    if (md.empty()) { if (rem == 255) { b = f; rem = 1; } else rem = 0; }
    else if (!cache_back()) rem = 0;
  }
};
inline bool zero_match(const MdOp& o) { return o.type == 0 && o.match == 0; }
std::vector<MdOp> md_operations(const std::string& md) {
  MdRange r(md);
  while (!r.empty() && zero_match(r.f)) r.pop_front();
  while (!r.empty() && zero_match(r.b)) r.pop_back();
  std::vector<MdOp> out;
  while (!r.empty()) {
    out.push_back(r.f);
    do { r.pop_front(); } while (!r.empty() && zero_match(r.f));
  }
  return out;
}

// dna(read) (md/reconstruct.d:38-214): reference bases over the read's M/=/X and deleted positions, from SEQ + MD.
// Restatement-defined: where BioD's lazy range gets stuck (a match / mismatch operation with the query bases used up:
// popFront returns without advancing, :143-156) or reads an empty range, the sequence simply ends.
std::string dna_of_read(const Bam* s, uint32_t r) {
  const uint8_t* rec = s->u.data() + s->rec_off[r] + 4;
  const uint32_t lname = s->l_read_name[r], nc = s->n_cigar[r];
  const int64_t lseq = s->l_seq[r];
  const uint8_t* seq = rec + 32 + lname + 4 * nc;
  std::string q;                                              // joiner of the SEQ chunks of M/=/X operations (:87-92,:196-198)
  int64_t qoff = 0;
  const uint32_t* cg = s->cigar.data() + s->cigar_off[r];
  for (uint32_t k = 0; k < nc; ++k) {
    const uint32_t raw = cg[k], len = raw >> 4;
    if (!op_query(raw)) continue;
    if (op_ref(raw))
      for (int64_t i = qoff; i < qoff + (int64_t)len && i < lseq; ++i) {
        const uint8_t byte = seq[i >> 1];
        q.push_back("=ACMGRSVTWYHKDBN"[(i & 1) ? (byte & 0xF) : (byte >> 4)]);
      }
    qoff += len;
  }
  std::string md;
  find_md(s, r, &md);
  const std::vector<MdOp> ops = md_operations(md);
  std::string out;
  size_t qi = 0, k = 0, di = 0;
  if (ops.empty()) return out;
  MdOp cur = ops[k++];
  while (true) {
    if (cur.type == 2) {
      if (di >= cur.del.size()) { if (k >= ops.size()) break; cur = ops[k++]; di = 0; continue; }   // (empty deletion)
      out.push_back(norm16((uint8_t)cur.del[di]));
      if (++di >= cur.del.size()) { if (k >= ops.size()) break; cur = ops[k++]; di = 0; }
    } else {
      if (qi >= q.size()) break;                             // _qseqIsSuddenlyEmpty
      out.push_back(cur.type == 0 ? q[qi] : norm16((uint8_t)cur.mismatch));
      ++qi;
      if (cur.type == 1 || --cur.match == 0) { if (k >= ops.size()) break; cur = ops[k++]; di = 0; }
    }
  }
  return out;
}

struct Cursor {              // PileupRead — bam/pileup.d:86-230
  uint32_t read;             // index into the record table
  int32_t end_position;      // EagerBamRead (read.d:1380-1383)
  uint32_t op_index, op_raw, op_offset, qoff;
};

struct Pileup {
  int status = 0;
  std::string msg;
  int ref_id = -1;
  std::vector<int32_t> col_ref;
  std::vector<uint64_t> col_pos, col_off;
  std::vector<uint32_t> n_start;
  std::vector<uint32_t> read_idx, qoff, op_index, op_offset;
  std::vector<uint8_t> base, qual;
  std::vector<uint8_t> ref_base;      // PileupColumn.reference_base, one per column ('N' unless use_md_tag)
  // lazy mode (orc_cpu_baseline): nothing above is filled; only what `foreach (column; pileup) n += column.coverage` sees
  uint64_t lazy_columns = 0, lazy_entries = 0;
};

struct PileupSim {
  const Bam* s;
  Pileup* out;
  std::vector<uint32_t> reads;   // filtered, in file order
  size_t next = 0;
  std::vector<Cursor> buf;
  uint64_t position = 0;
  int32_t ref = -1;
  size_t n_starting = 0;
  bool skip_zero;
  // lazy = what BioD's own sweep costs: popFront keeps and advances the live reads (pileup.d:345-397) and the column
  // hands out a slice of them; bases and qualities are computed only if the consumer asks (pileup.d:115-134)
  bool lazy = false;
  // PileupRangeUsingMdTag (pileup.d:522-654)
  bool use_md = false;
  std::string chunk;             // _chunk: reference bases reconstructed from the current provider
  size_t chunk_i = 0;
  uint32_t chunk_end = 0;        // _chunk_end_position
  uint32_t provider = 0;         // _next_chunk_provider
  bool has_provider = false;
  uint64_t prev_cov = 0;         // _prev_coverage
  int32_t curr_ref = -1;         // _curr_ref_id
  uint8_t ref_base = 'N';        // _column._reference_base (PileupColumn default 'N')
  bool chunk_empty() const { return chunk_i >= chunk.size(); }
  void take_front() {            // _column._reference_base = _chunk.front; _chunk.popFront();
    // restatement-defined: `front` of an exhausted chunk (BioD reads an empty range) gives 'N'
    if (chunk_empty()) { ref_base = 'N'; return; }
    ref_base = (uint8_t)chunk[chunk_i++];
  }
  void md_add(uint32_t r) {      // override add :563-596, after super.add(read)
    const bool had_zero = prev_cov == 0;
    if (s->ref_id[r] != curr_ref) {
      curr_ref = s->ref_id[r];
      has_provider = true;
      provider = r;
      return;
    }
    if ((uint32_t)s->pos[r] > chunk_end && !had_zero) return;
    if ((uint32_t)s->end_pos[r] > chunk_end) {
      if (!has_provider) { has_provider = true; provider = r; }
      else if (s->end_pos[r] > s->end_pos[provider]) provider = r;
    }
  }

  const uint32_t* cig(uint32_t r) const { return s->cigar.data() + s->cigar_off[r]; }
  uint32_t ncig(uint32_t r) const { return (uint32_t)(s->cigar_off[r + 1] - s->cigar_off[r]); }

  bool reads_empty() const { return next >= reads.size(); }

  void add(uint32_t r) {     // PileupRange.add :315-317, PileupRead ctor :175-192
    Cursor c{};
    c.read = r;
    c.end_position = s->end_pos[r];
    const uint32_t* cg = cig(r);
    uint32_t n = ncig(r);
    bool skipped_n = false;
    for (c.op_index = 0; c.op_index < n; ++c.op_index) {
      c.op_raw = cg[c.op_index];
      if (op_ref(c.op_raw)) {
        if ((c.op_raw & 0xF) != 3) break;            // type != 'N'
        if ((c.op_raw >> 4) > 0) skipped_n = true;
      } else if (op_query(c.op_raw)) {
        c.qoff += c.op_raw >> 4;
      }
    }
    // assertCigarIndexIsValid :224-228.  Restatement-defined: a leading N that
    // is skipped without advancing the column makes the cursor run off the end
    // of the CIGAR later (undefined in -release); both are reported as errors.
    if ((c.op_index >= n || skipped_n) && !out->status) {
      out->status = ORC_ERR_CIGAR;
      out->msg = "Invalid read - CIGAR has no usable reference-consuming operation (record " + std::to_string(r) + ")";
    }
    buf.push_back(c);
    if (use_md) md_add(r);
  }

  void increment(Cursor& c) {   // incrementPosition :195-222
    ++c.op_offset;
    if (op_query(c.op_raw)) ++c.qoff;
    if (c.op_offset >= (c.op_raw >> 4)) {
      c.op_offset = 0;
      const uint32_t* cg = cig(c.read);
      uint32_t n = ncig(c.read);
      for (++c.op_index; c.op_index < n; ++c.op_index) {
        c.op_raw = cg[c.op_index];
        if (op_ref(c.op_raw)) break;
        if (op_query(c.op_raw)) c.qoff += c.op_raw >> 4;
      }
      // running past the end only happens on the very last live position of a
      // read whose zero-length ops stole positions; the read dies before use.
    }
  }

  void init_new_reference() {   // PileupRangeUsingMdTag.initNewReference :598-611 (virtual: the base class calls this one)
    if (!use_md) { base_init_new_reference(); return; }
    prev_cov = 0;
    base_init_new_reference();
    if (has_provider) {
      chunk = dna_of_read(s, provider);
      chunk_i = 0;
      chunk_end = (uint32_t)s->end_pos[provider];
      has_provider = false;
      take_front();
    } else {
      ref_base = 'N';
    }
  }

  void base_init_new_reference() {   // :399-424
    uint32_t r = reads[next];
    position = (uint64_t)(int64_t)s->pos[r];
    ref = s->ref_id[r];
    size_t n = 1;
    add(r);
    ++next;
    while (!reads_empty()) {
      r = reads[next];
      if (s->ref_id[r] == ref && (uint64_t)(int64_t)s->pos[r] == position) { add(r); ++n; ++next; }
      else break;
    }
    n_starting = n;
  }

  bool empty() const { return reads_empty() && buf.empty(); }   // :340-342

  void pop_front() {            // PileupRangeUsingMdTag.popFront :614-653
    if (!use_md) { base_pop_front(); return; }
    if (!chunk_empty()) ref_base = (uint8_t)chunk[chunk_i++];
    else ref_base = 'N';
    prev_cov = buf.size();
    base_pop_front();           // may call init_new_reference() above
    if (chunk_empty() && has_provider) {
      chunk = dna_of_read(s, provider);
      chunk_i = 0;
      chunk_end = (uint32_t)s->end_pos[provider];
      has_provider = false;
      const uint64_t skip = position - (uint64_t)(int64_t)s->pos[provider];     // _chunk.popFrontN(...)
      chunk_i = (size_t)std::min<uint64_t>(skip, chunk.size());
      take_front();
    }
  }

  void base_pop_front() {       // :345-397
    uint64_t pos = ++position;
    size_t survived = 0;
    for (size_t i = 0; i < buf.size(); ++i) {
      if ((uint64_t)(int64_t)buf[i].end_position > pos) {
        if (survived < i) buf[survived] = buf[i];
        ++survived;
      }
    }
    for (size_t i = 0; i < survived; ++i) increment(buf[i]);
    buf.resize(survived);
    n_starting = 0;
    if (!reads_empty()) {
      if (s->ref_id[reads[next]] != ref && survived == 0) {
        init_new_reference();
      } else {
        size_t n = 0;
        while (!reads_empty() && (uint64_t)(int64_t)s->pos[reads[next]] == pos && s->ref_id[reads[next]] == ref) {
          add(reads[next]);
          ++next;
          ++n;
        }
        n_starting = n;
        if (survived == 0 && n == 0 && skip_zero) init_new_reference();
      }
    }
  }

  void emit() {
    if (lazy) { ++out->lazy_columns; out->lazy_entries += buf.size(); return; }
    out->col_ref.push_back(ref);
    out->col_pos.push_back(position);
    out->n_start.push_back((uint32_t)n_starting);
    out->ref_base.push_back(ref_base);
    for (const Cursor& c : buf) {
      const uint8_t* rec = s->u.data() + s->rec_off[c.read] + 4;
      uint32_t lname = s->l_read_name[c.read];
      uint32_t nc = s->n_cigar[c.read];
      int32_t lseq = s->l_seq[c.read];
      const uint8_t* seq = rec + 32 + lname + 4 * nc;
      const uint8_t* ql = seq + (lseq + 1) / 2;
      uint8_t base = '-', q = 255;                       // :115-134
      if (op_query(c.op_raw) && op_ref(c.op_raw)) {
        if ((int64_t)c.qoff < (int64_t)lseq) {
          uint8_t byte = seq[c.qoff >> 1];               // read.d:364-383
          uint8_t code = (c.qoff & 1) ? (byte & 0xF) : (byte >> 4);
          base = (uint8_t)"=ACMGRSVTWYHKDBN"[code];      // bio/core/base.d:85
          q = ql[c.qoff];                                // read.d:468-470
        } else {
          // sequence[query_offset] past l_seq (e.g. SEQ '*' with a CIGAR, as in bins.bam) is a RangeError
          // in D — but only if the consumer asks for the base; building the column never touches it.
          // Restatement-defined for an eager builder: base 0x00, quality 255, no error.
          base = 0;
          q = 255;
        }
      }
      out->read_idx.push_back(c.read);
      out->base.push_back(base);
      out->qual.push_back(q);
      out->qoff.push_back(c.qoff);
      out->op_index.push_back(c.op_index);
      out->op_offset.push_back(c.op_offset);
    }
    out->col_off.push_back(out->read_idx.size());
  }
};

// makePileup / pileupInstance (pileup.d:480-507, 683-694) when single_ref != 0,
// pileupColumns (pileup.d:509-519) otherwise.
Pileup* run_pileup(const Bam* s, int single_ref, uint64_t start_from, uint64_t end_at, int skip_zero,
                   int64_t rec_begin, int64_t rec_end, int use_md = 0, const int64_t* rec_list = nullptr,
                   uint64_t n_list = 0, bool lazy = false) {
  Pileup* out = new Pileup;
  out->col_off.push_back(0);
  PileupSim sim;
  sim.lazy = lazy;
  sim.s = s;
  sim.out = out;
  sim.skip_zero = skip_zero != 0;
  sim.use_md = use_md != 0;
  int64_t n = (int64_t)s->rec_off.size();
  if (rec_end < 0 || rec_end > n) rec_end = n;
  if (rec_begin < 0) rec_begin = 0;
  if (rec_list) {
    // the reads of an arbitrary range (e.g. bam[ref][beg .. end), examples/read_bam_file.d:22-25), given by index
    for (uint64_t k = 0; k < n_list; ++k) {
      const int64_t r = rec_list[k];
      if (r >= 0 && r < n && s->end_pos[r] - s->pos[r] > 0) sim.reads.push_back((uint32_t)r);
    }
  } else
  for (int64_t r = rec_begin; r < rec_end; ++r)
    if (s->end_pos[r] - s->pos[r] > 0) sim.reads.push_back((uint32_t)r);   // filter basesCovered()>0 (:481,:510)
  if (single_ref) {
    size_t k = 0;                                                           // :482-489
    while (k < sim.reads.size() && (uint64_t)(int64_t)s->end_pos[sim.reads[k]] < start_from) ++k;
    sim.reads.erase(sim.reads.begin(), sim.reads.begin() + k);
    if (!sim.reads.empty()) {
      out->ref_id = s->ref_id[sim.reads[0]];                                // :490-493
      size_t m = 0;
      while (m < sim.reads.size() && s->ref_id[sim.reads[m]] == out->ref_id) ++m;   // takeUntil :494
      sim.reads.resize(m);
    }
  } else {
    start_from = 0;
    end_at = ~0ull;
  }
  if (!sim.reads.empty()) sim.init_new_reference();                         // :328-331
  while (!sim.empty() && sim.position < start_from) sim.pop_front();        // :497-504
  while (!sim.empty()) {
    if (sim.position >= end_at) break;                                      // takeUntil :505
    sim.emit();
    sim.pop_front();
  }
  return out;
}

}  // namespace


// ----------------------------------------------------- BAI random access ----
// SURVEY.md §8f row N2 (oracle side).  BaiFile.parse (bam/baifile.d:126-169), Index.getMinimumOffset (:77-82),
// Bin.canOverlapWith (bai/bin.d:49-78), RandomAccessManager.getChunks / appendChunks (randomaccessmanager.d:211-244),
// nonOverlapping (core/utils/algo.d:95-162), StreamChunksSupplier (bgzf/inputstream.d:257-345) feeding the record
// framing of readrange.d:118-173, BamReadFilter (randomaccessmanager.d:366-462), getReadsBetween (:186-196).
namespace {

struct BaiBin { uint32_t id; std::vector<std::pair<uint64_t, uint64_t>> chunks; };
struct BaiIndex { std::vector<BaiBin> bins; std::vector<uint64_t> ioffsets; };
struct Bai {
  int status = 0;
  std::string msg;
  std::vector<BaiIndex> idx;
};

inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

Bai* bai_parse(const uint8_t* d, uint64_t n) {            // baifile.d:126-169
  Bai* b = new Bai;
  uint64_t o = 0;
  auto need = [&](uint64_t k) {
    if (b->status) return false;
    if (n - o < k || o > n) { b->status = ORC_ERR_TRUNC; b->msg = "not enough data in stream"; return false; }   // Stream.readExact
    return true;
  };
  if (!need(4)) return b;
  if (memcmp(d, "BAI\1", 4) != 0) { b->status = ORC_ERR_FORMAT; b->msg = "Invalid file format: expected BAI\\1"; return b; }
  o = 4;
  if (!need(4)) return b;
  const int32_t n_ref = (int32_t)le32(d + o);
  o += 4;
  for (int32_t i = 0; i < n_ref && !b->status; ++i) {
    BaiIndex ix;
    if (!need(4)) break;
    const int32_t n_bin = (int32_t)le32(d + o);
    o += 4;
    for (int32_t j = 0; j < n_bin; ++j) {
      if (!need(8)) break;
      BaiBin bin;
      bin.id = le32(d + o);
      const int32_t n_chunk = (int32_t)le32(d + o + 4);
      o += 8;
      for (int32_t k = 0; k < n_chunk; ++k) {
        if (!need(16)) break;
        bin.chunks.push_back({le64(d + o), le64(d + o + 8)});
        o += 16;
      }
      if (b->status) break;
      ix.bins.push_back(std::move(bin));
    }
    if (b->status || !need(4)) break;
    const int32_t n_intv = (int32_t)le32(d + o);
    o += 4;
    for (int32_t j = 0; j < n_intv; ++j) {
      if (!need(8)) break;
      ix.ioffsets.push_back(le64(d + o));
      o += 8;
    }
    if (b->status) break;
    b->idx.push_back(std::move(ix));
  }
  return b;
}

bool bin_can_overlap(uint32_t id, int32_t begin, int32_t end) {      // bai/bin.d:49-78
  if (id == 0) return true;
  if (id > 37449) return false;                                       // BAI_MAX_BIN_ID
  if (begin < 0) begin = 0;
  int32_t magic = 4681;
  int32_t b = begin >> 14, e = end >> 14;
  while (true) {
    const uint32_t delta = id - (uint32_t)magic;                      // uint - int in D: unsigned arithmetic and comparisons
    if ((uint32_t)b <= delta && delta <= (uint32_t)e) return true;
    magic >>= 3;
    if (magic == 0) return false;
    b >>= 3;
    e >>= 3;
  }
}

typedef std::pair<uint64_t, uint64_t> Chunk;    // (beg, end) virtual offsets; compares like Chunk.opCmp (bgzf/chunk.d:34-40)

std::vector<Chunk> region_chunks(const Bai* bai, uint32_t ref_id, uint32_t beg, uint32_t end, int* status) {   // :222-244
  std::vector<Chunk> out;
  if (ref_id >= bai->idx.size()) { *status = ORC_ERR_ARG; return out; }          // "Invalid reference sequence index"
  const BaiIndex& ix = bai->idx[ref_id];
  // getMinimumOffset(int position), baifile.d:77-82
  const int32_t pos = std::max(0, (int32_t)beg);
  const int32_t i = std::min(pos / 16384, (int32_t)ix.ioffsets.size() - 1);
  const uint64_t min_offset = i == -1 ? 0 : ix.ioffsets[(size_t)i];
  for (const BaiBin& b : ix.bins) {
    if (!bin_can_overlap(b.id, (int32_t)beg, (int32_t)end)) continue;
    for (const Chunk& c : b.chunks)                                              // appendChunks :211-220
      if (c.second > min_offset) out.push_back({std::max(c.first, min_offset), c.second});
  }
  std::sort(out.begin(), out.end());
  std::vector<Chunk> merged;                                                     // algo.d:95-140
  for (const Chunk& c : out) {
    if (!merged.empty() && merged.back().second >= c.first) merged.back().second = std::max(merged.back().second, c.second);
    else merged.push_back(c);
  }
  return merged;
}

struct RegionRead { int64_t index; uint64_t start_vo, end_vo; int32_t ref_id, pos, end_pos; };

// The byte stream StreamChunksSupplier + BgzfInputStream present for a chunk list, framed into records.
// `until_vo`: getReadsBetween's `until!offsetTooBig` (records whose end offset is beyond it are cut off), ~0 = none.
int chunk_stream_reads(Bam* s, std::vector<Chunk> chunks, uint64_t until_vo, std::vector<RegionRead>* out) {
  decode_records(s);            // the whole file: block table + the record table the indices refer to
  struct Piece { uint64_t cat, coffset, cend; uint32_t within, len; };
  std::vector<uint8_t> cat;
  std::vector<Piece> pieces;
  auto block_at = [&](uint64_t coffset) -> const Block* {
    size_t lo = 0, hi = s->blocks.size();
    while (lo < hi) { size_t m = (lo + hi) / 2; if (s->blocks[m].coffset < coffset) lo = m + 1; else hi = m; }
    return (lo < s->blocks.size() && s->blocks[lo].coffset == coffset) ? &s->blocks[lo] : nullptr;
  };
  size_t ci = 0;
  uint64_t cur_beg = 0, cur_end = 0, next_off = 0;
  auto move_to_next_chunk = [&]() {                                             // inputstream.d:262-277
    if (ci >= chunks.size()) return;
    const uint64_t beg = chunks[ci].first;
    size_t i = ci + 1;
    for (; i < chunks.size(); ++i)
      if ((chunks[i].first >> 16) > (chunks[ci].first >> 16)) break;
    ci = i - 1;
    chunks[ci].first = beg;
    cur_beg = chunks[ci].first;
    cur_end = chunks[ci].second;
    next_off = cur_beg >> 16;
  };
  move_to_next_chunk();
  bool eof = false;
  while (ci < chunks.size() && !eof) {
    // getNextBgzfBlock :285-338 (a block that is not in the table — EOF block, end of file, garbage — ends the stream)
    const Block* b = block_at(next_off);
    if (!b) break;
    const uint64_t offset = b->coffset;
    next_off = b->coffset + b->bsize + 1;
    const uint32_t skip_start = offset == (cur_beg >> 16) ? (uint32_t)(cur_beg & 0xFFFF) : 0;
    const int64_t skip_end = offset == (cur_end >> 16) ? (int64_t)b->isize - (int64_t)(cur_end & 0xFFFF) : 0;
    if (offset >= (cur_end >> 16)) { ++ci; move_to_next_chunk(); }
    if (b->isize > 0 && skip_end == (int64_t)b->isize) continue;                // the chunk ended on the edge of two blocks
    const int64_t from = skip_start, to = (int64_t)b->isize - (int64_t)(uint16_t)skip_end;
    if (to <= from) { eof = true; break; }                                      // setupReadBuffer :432-435 (restatement-defined for to < from)
    pieces.push_back(Piece{cat.size(), b->coffset, b->coffset + b->bsize + 1, (uint32_t)from, (uint32_t)(to - from)});
    cat.insert(cat.end(), s->u.begin() + b->uoffset + from, s->u.begin() + b->uoffset + to);
  }
  auto vo_of = [&](uint64_t x) -> uint64_t {                                     // _current_vo bookkeeping :438-445,516-524
    size_t lo = 0, hi = pieces.size();
    while (lo < hi) { size_t m = (lo + hi) / 2; if (pieces[m].cat <= x) lo = m + 1; else hi = m; }
    if (lo == 0) return 0;
    const Piece& p = pieces[lo - 1];
    if (x - p.cat >= p.len) {
      // end of a piece: the next block's start, or (skip_end > 0) the cut position inside this block
      const Block* b = block_at(p.coffset);
      if (b && p.within + p.len < b->isize) return (p.coffset << 16) | (p.within + p.len);
      return p.cend << 16;
    }
    return (p.coffset << 16) | (p.within + (x - p.cat));
  };
  uint64_t p = 0;
  const uint64_t avail = cat.size();
  while (p < avail && avail - p >= 4) {                                          // readrange.d:118-173
    const int32_t bs = (int32_t)le32(cat.data() + p);
    if (bs < 32 || (uint64_t)bs > avail - p - 4) return ORC_ERR_TRUNC;
    const uint8_t* r = cat.data() + p + 4;
    RegionRead rr;
    rr.start_vo = vo_of(p);
    rr.end_vo = vo_of(p + 4 + (uint64_t)bs);
    if (until_vo != ~0ull && rr.end_vo > until_vo) break;                        // randomaccessmanager.d:190-195
    rr.ref_id = (int32_t)le32(r);
    rr.pos = (int32_t)le32(r + 4);
    const uint32_t bin_mq_nl = le32(r + 8), flag_nc = le32(r + 12);
    const uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
    if (32ull + lname + 4ull * nc > (uint64_t)bs) return ORC_ERR_TRUNC;
    uint32_t covered = 0;
    for (uint32_t k = 0; k < nc; ++k) {
      const uint32_t raw = le32(r + 32 + lname + 4 * k);
      if (op_ref(raw)) covered += raw >> 4;
    }
    if ((flag_nc >> 16) & 0x4) covered = 0;
    rr.end_pos = (int32_t)((uint32_t)rr.pos + covered);
    const auto it = std::lower_bound(s->start_vo.begin(), s->start_vo.end(), rr.start_vo);
    rr.index = (it != s->start_vo.end() && *it == rr.start_vo) ? (int64_t)(it - s->start_vo.begin()) : -1;
    out->push_back(rr);
    p += 4 + (uint64_t)bs;
  }
  return 0;
}

// BamReadFilter.findNext for one region (randomaccessmanager.d:396-450)
void region_filter(const std::vector<RegionRead>& in, uint32_t ref_id, uint32_t start, uint32_t end, std::vector<RegionRead>* out) {
  for (const RegionRead& r : in) {
    const uint32_t cur = (uint32_t)r.ref_id;
    if (cur > ref_id) return;                                   // no more records for this reference (-1 compares as uint.max)
    if (cur < ref_id) continue;
    // _current_read.position (int) against region.end / region.start (uint): D compares them as unsigned
    if ((uint32_t)r.pos >= end) return;
    if ((uint32_t)r.pos > start) { out->push_back(r); continue; }
    if ((uint32_t)(r.pos + (r.end_pos - r.pos)) <= start) continue;     // position + basesCovered() <= region.start
    out->push_back(r);
  }
}

// getGroupChunks (randomaccessmanager.d:246-296): the chunks of every bin the bitset of ALL regions (of one reference,
// sorted, non-overlapping) marks, cut at the linear index's minimum offset for the FIRST region's start, merged.
// Restatement-defined: the reference's bitset has BAI_MAX_BIN_ID = 37449 entries and is indexed with every bin id of the
// index — the pseudo-bin 37450 that samtools (and BioD's own IndexBuilder) write is out of its range, an error in D.
// Pinned: bins with an id beyond the bitset are passed over, as Bin.canOverlapWith does for a single region (bin.d:58).
std::vector<Chunk> group_chunks(const Bai* bai, uint32_t ref_id, const std::vector<std::pair<uint32_t, uint32_t>>& regions, int* status) {
  std::vector<Chunk> out;
  if (regions.empty() || ref_id >= bai->idx.size()) { *status = ORC_ERR_ARG; return out; }
  std::vector<char> bitset(37449, 0);
  bitset[0] = 1;
  for (const auto& rg : regions) {
    const uint32_t beg = rg.first;
    uint32_t end = rg.second;
    if (!(beg < end)) { *status = ORC_ERR_ARG; return out; }        // enforce(beg < end)
    --end;
    uint32_t k;
    for (k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) bitset[k] = 1;
    for (k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) bitset[k] = 1;
    for (k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) bitset[k] = 1;
    for (k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) bitset[k] = 1;
    for (k = 4681 + (beg >> 14); k <= 4681 + (end >> 14) && k < 37449; ++k) bitset[k] = 1;
  }
  const BaiIndex& ix = bai->idx[ref_id];
  const int32_t pos = std::max(0, (int32_t)regions.front().first);
  const int32_t i = std::min(pos / 16384, (int32_t)ix.ioffsets.size() - 1);
  const uint64_t min_offset = i == -1 ? 0 : ix.ioffsets[(size_t)i];
  for (const BaiBin& b : ix.bins) {
    if (b.id >= bitset.size() || !bitset[b.id]) continue;
    for (const Chunk& c : b.chunks)
      if (c.second > min_offset) out.push_back({std::max(c.first, min_offset), c.second});
  }
  std::sort(out.begin(), out.end());
  std::vector<Chunk> merged;
  for (const Chunk& c : out) {
    if (!merged.empty() && merged.back().second >= c.first) merged.back().second = std::max(merged.back().second, c.second);
    else merged.push_back(c);
  }
  return merged;
}

// BamReadFilter.findNext for several regions of one reference (randomaccessmanager.d:396-450): the region pointer only
// moves forward, when a read starts at or beyond the current region's end
void regions_filter(const std::vector<RegionRead>& in, uint32_t ref_id, const std::vector<std::pair<uint32_t, uint32_t>>& regions,
                    std::vector<RegionRead>* out) {
  size_t k = 0;
  for (size_t i = 0; i < in.size();) {
    const RegionRead& r = in[i];
    const uint32_t cur = (uint32_t)r.ref_id;
    if (cur > ref_id) return;
    if (cur < ref_id) { ++i; continue; }
    if ((uint32_t)r.pos >= regions[k].second) {
      if (++k == regions.size()) return;
      continue;                                                 // the same read against the next region
    }
    if ((uint32_t)r.pos > regions[k].first) { out->push_back(r); ++i; continue; }
    if ((uint32_t)(r.pos + (r.end_pos - r.pos)) <= regions[k].first) { ++i; continue; }
    out->push_back(r);
    ++i;
  }
}

}  // namespace

// ------------------------------------------------------------------ C API ----
extern "C" {

typedef struct Bam orc_bam;
typedef struct Pileup orc_pileup;

orc_bam* orc_open(const uint8_t* data, uint64_t len) {
  Bam* s = new Bam;
  s->file = data;
  s->flen = len;
  parse_header(s);
  return s;
}
void orc_close(orc_bam* s) { delete s; }
// status of open: an error in a block beyond the header is deferred to decode.
int orc_open_status(const orc_bam* s) { return s->reads_start_u ? 0 : s->err.status; }
int orc_status(const orc_bam* s) { return s->err.status; }
int orc_zlib_errnum(const orc_bam* s) { return s->err.zerr; }
uint64_t orc_err_offset(const orc_bam* s) { return s->err.offset; }
const char* orc_errmsg(const orc_bam* s) { return s->err.msg.c_str(); }

const char* orc_header_text(const orc_bam* s, uint64_t* len) { *len = s->text.size(); return s->text.data(); }
int orc_n_refs(const orc_bam* s) { return (int)s->ref_names.size(); }
const char* orc_ref_name(const orc_bam* s, int i) { return s->ref_names[i].c_str(); }
int orc_ref_len(const orc_bam* s, int i) { return s->ref_lens[i]; }
uint64_t orc_reads_start_voffset(const orc_bam* s) { return s->reads_start_vo; }
uint64_t orc_reads_start_uoffset(const orc_bam* s) { return s->reads_start_u; }

int orc_decode(orc_bam* s) { return decode_records(s); }
uint64_t orc_n_blocks(const orc_bam* s) { return s->blocks.size(); }
// field: 0 coffset, 1 bsize, 2 cdata_size, 3 crc32, 4 isize, 5 uoffset, 6 payload offset
uint64_t orc_block_field(const orc_bam* s, uint64_t i, int f) {
  const Block& b = s->blocks[i];
  switch (f) {
    case 0: return b.coffset; case 1: return b.bsize; case 2: return b.cdata_size; case 3: return b.crc32;
    case 4: return b.isize; case 5: return b.uoffset; default: return b.payload;
  }
}
const uint8_t* orc_udata(const orc_bam* s, uint64_t* len) { *len = s->u.size(); return s->u.data(); }
uint64_t orc_n_records(const orc_bam* s) { return s->rec_off.size(); }
uint64_t orc_n_cigar_total(const orc_bam* s) { return s->cigar.size(); }
#define ORC_ARR(name, type, member) const type* orc_##name(const orc_bam* s) { return s->member.data(); }
ORC_ARR(rec_off, uint64_t, rec_off)
ORC_ARR(block_size, int32_t, block_size)
ORC_ARR(ref_id, int32_t, ref_id)
ORC_ARR(pos, int32_t, pos)
ORC_ARR(end_pos, int32_t, end_pos)
ORC_ARR(l_seq, int32_t, l_seq)
ORC_ARR(next_ref, int32_t, next_ref)
ORC_ARR(next_pos, int32_t, next_pos)
ORC_ARR(tlen, int32_t, tlen)
ORC_ARR(bin, uint16_t, bin)
ORC_ARR(flag, uint16_t, flag)
ORC_ARR(n_cigar, uint16_t, n_cigar)
ORC_ARR(mapq, uint8_t, mapq)
ORC_ARR(l_read_name, uint8_t, l_read_name)
ORC_ARR(start_vo, uint64_t, start_vo)
ORC_ARR(end_vo, uint64_t, end_vo)
ORC_ARR(cigar_off, uint64_t, cigar_off)
ORC_ARR(cigar, uint32_t, cigar)

orc_pileup* orc_pileup_run(orc_bam* s, int single_ref, uint64_t start_from, uint64_t end_at, int skip_zero) {
  decode_records(s);
  return run_pileup(s, single_ref, start_from, end_at, skip_zero, 0, -1);
}
// pileup over records [rec_begin, rec_end) only — used to restate pileupChunks
// (pileup.d:859-1015), whose front is makePileup(chain(prev_chunk, current_chunk), ...).
orc_pileup* orc_pileup_run_range(orc_bam* s, int single_ref, uint64_t start_from, uint64_t end_at, int skip_zero,
                                 int64_t rec_begin, int64_t rec_end) {
  decode_records(s);
  return run_pileup(s, single_ref, start_from, end_at, skip_zero, rec_begin, rec_end);
}
// pileup over the reads whose record indices are listed (any range of reads is acceptable to makePileup:
// examples/read_bam_file.d:22-25 feeds it bam["chr2"][150 .. 160])
orc_pileup* orc_pileup_run_list(orc_bam* s, const int64_t* rec_list, uint64_t n_list, int single_ref, uint64_t start_from,
                                uint64_t end_at, int skip_zero, int use_md) {
  decode_records(s);
  return run_pileup(s, single_ref, start_from, end_at, skip_zero, 0, -1, use_md, rec_list, n_list);
}
// the same with use_md_tag = true: every column also gets its reference_base (pileup.d:522-654)
orc_pileup* orc_pileup_run_md(orc_bam* s, int single_ref, uint64_t start_from, uint64_t end_at, int skip_zero) {
  decode_records(s);
  return run_pileup(s, single_ref, start_from, end_at, skip_zero, 0, -1, 1);
}
// dna(read) of one record (md/reconstruct.d:38-214); returns its length, copies at most cap bytes
uint64_t orc_dna_of_read(orc_bam* s, uint64_t r, char* out, uint64_t cap) {
  decode_records(s);
  const std::string d = dna_of_read(s, (uint32_t)r);
  memcpy(out, d.data(), (size_t)std::min<uint64_t>(cap, d.size()));
  return d.size();
}
void orc_pileup_free(orc_pileup* p) { delete p; }
int orc_pileup_status(const orc_pileup* p) { return p->status; }
const char* orc_pileup_errmsg(const orc_pileup* p) { return p->msg.c_str(); }
int orc_pileup_ref_id(const orc_pileup* p) { return p->ref_id; }
uint64_t orc_pileup_n_columns(const orc_pileup* p) { return p->col_pos.size(); }
uint64_t orc_pileup_n_entries(const orc_pileup* p) { return p->read_idx.size(); }
#define ORC_PARR(name, type, member) const type* orc_pileup_##name(const orc_pileup* p) { return p->member.data(); }
ORC_PARR(col_ref, int32_t, col_ref)
ORC_PARR(col_pos, uint64_t, col_pos)
ORC_PARR(col_off, uint64_t, col_off)
ORC_PARR(n_start, uint32_t, n_start)
ORC_PARR(read_idx, uint32_t, read_idx)
ORC_PARR(qoff, uint32_t, qoff)
ORC_PARR(op_index, uint32_t, op_index)
ORC_PARR(op_offset, uint32_t, op_offset)
ORC_PARR(base, uint8_t, base)
ORC_PARR(qual, uint8_t, qual)
ORC_PARR(ref_base, uint8_t, ref_base)


typedef struct Bai orc_bai;
orc_bai* orc_bai_open(const uint8_t* data, uint64_t len) { return bai_parse(data, len); }
void orc_bai_close(orc_bai* b) { delete b; }
int orc_bai_status(const orc_bai* b) { return b->status; }
const char* orc_bai_errmsg(const orc_bai* b) { return b->msg.c_str(); }
uint64_t orc_bai_n_refs(const orc_bai* b) { return b->idx.size(); }
uint64_t orc_bai_n_bins(const orc_bai* b, uint64_t ref) { return ref < b->idx.size() ? b->idx[ref].bins.size() : 0; }
uint64_t orc_bai_n_intervals(const orc_bai* b, uint64_t ref) { return ref < b->idx.size() ? b->idx[ref].ioffsets.size() : 0; }
// getChunks(BamRegion) — writes at most cap (beg, end) pairs, returns their number (-1: invalid reference index)
int64_t orc_region_chunks(const orc_bai* b, uint32_t ref_id, uint32_t beg, uint32_t end, uint64_t* out2, uint64_t cap) {
  int st = 0;
  const std::vector<Chunk> c = region_chunks(b, ref_id, beg, end, &st);
  if (st) return -1;
  for (uint64_t k = 0; k < c.size() && k < cap; ++k) { out2[2 * k] = c[k].first; out2[2 * k + 1] = c[k].second; }
  return (int64_t)c.size();
}
// bam[ref][beg .. end) = getReads(BamRegion) (:300-305): record indices (into the whole-file record table; -1 for a
// record the sequential walk does not know) and their virtual offsets.  Returns the count, or -(status) on error.
int64_t orc_region_reads(orc_bam* s, const orc_bai* b, uint32_t ref_id, uint32_t beg, uint32_t end, int64_t* index,
                         uint64_t* start_vo, uint64_t* end_vo, uint64_t cap) {
  int st = 0;
  const std::vector<Chunk> c = region_chunks(b, ref_id, beg, end, &st);
  if (st) return st;
  std::vector<RegionRead> all, keep;
  st = chunk_stream_reads(s, c, ~0ull, &all);
  if (st) return st;
  region_filter(all, ref_id, beg, end, &keep);
  for (uint64_t k = 0; k < keep.size() && k < cap; ++k) {
    index[k] = keep[k].index;
    if (start_vo) start_vo[k] = keep[k].start_vo;
    if (end_vo) end_vo[k] = keep[k].end_vo;
  }
  return (int64_t)keep.size();
}
// getGroupChunks for the (sorted, non-overlapping) regions of one reference: the chunk list, for the tests of the
// product's biodb_index_regions_chunks
int64_t orc_group_chunks(const orc_bai* b, uint32_t ref_id, uint64_t n, const uint32_t* begs, const uint32_t* ends, uint64_t* out2, uint64_t cap) {
  std::vector<std::pair<uint32_t, uint32_t>> group;
  for (uint64_t k = 0; k < n; ++k) group.push_back({begs[k], ends[k]});
  int st = 0;
  const std::vector<Chunk> c = group_chunks(b, ref_id, group, &st);
  if (st) return st;
  for (uint64_t k = 0; k < c.size() && k < cap; ++k) { out2[2 * k] = c[k].first; out2[2 * k + 1] = c[k].second; }
  return (int64_t)c.size();
}
// getReads(BamRegion[]) (:316-337): regions sorted, grouped by reference, overlapping ones joined (nonOverlapping:
// prev.end >= next.start); per group filteredReads (:286-296); the groups one after the other.  Regions: n triples
// (ref_id, start, end).  Output as orc_region_reads.
int64_t orc_regions_reads(orc_bam* s, const orc_bai* b, const uint32_t* regions3, uint64_t n, int64_t* index,
                          uint64_t* start_vo, uint64_t* end_vo, uint64_t cap) {
  struct Rg { uint32_t ref, beg, end; };
  std::vector<Rg> rg(n);
  for (uint64_t k = 0; k < n; ++k) rg[k] = Rg{regions3[3 * k], regions3[3 * k + 1], regions3[3 * k + 2]};
  std::sort(rg.begin(), rg.end(), [](const Rg& a, const Rg& c) {               // BamRegion.opCmp (region.d)
    if (a.ref != c.ref) return a.ref < c.ref;
    if (a.beg != c.beg) return a.beg < c.beg;
    return a.end < c.end;
  });
  uint64_t total = 0;
  for (size_t i = 0; i < rg.size();) {
    size_t j = i;
    std::vector<std::pair<uint32_t, uint32_t>> group;
    for (; j < rg.size() && rg[j].ref == rg[i].ref; ++j) {
      if (!group.empty() && group.back().second >= rg[j].beg) group.back().second = std::max(group.back().second, rg[j].end);
      else group.push_back({rg[j].beg, rg[j].end});
    }
    int st = 0;
    const std::vector<Chunk> c = group_chunks(b, rg[i].ref, group, &st);
    if (st) return st;
    std::vector<RegionRead> all, keep;
    st = chunk_stream_reads(s, c, ~0ull, &all);
    if (st) return st;
    regions_filter(all, rg[i].ref, group, &keep);
    for (const RegionRead& r : keep) {
      if (total < cap) {
        index[total] = r.index;
        if (start_vo) start_vo[total] = r.start_vo;
        if (end_vo) end_vo[total] = r.end_vo;
      }
      ++total;
    }
    i = j;
  }
  return (int64_t)total;
}
// getReadsBetween(from, to) (:186-196)
int64_t orc_reads_between(orc_bam* s, uint64_t from_vo, uint64_t to_vo, int64_t* index, uint64_t cap) {
  // createStreamStartingFrom(from): every block from from.coffset on, the first one cut at from.uoffset
  decode_records(s);
  std::vector<Chunk> c{{from_vo, s->end_coffset << 16}};
  std::vector<RegionRead> all;
  const int st = chunk_stream_reads(s, c, to_vo, &all);
  if (st) return st;
  for (uint64_t k = 0; k < all.size() && k < cap; ++k) index[k] = all[k].index;
  return (int64_t)all.size();
}

// ------------------------------------------------------- MAQ genotype likelihoods (row N3) ----
// bio/std/hts/snpcallers/maq.d: ErrorModelCoefficients (:66-132), computeLikelihoods (:138-248), GenotypeLikelihoodInfo
// (:252-310), MaqSnpCaller.genotypeLikelihoodInfo / makeCall / findSNPs (:388-540).  D's `real` is the x87 80-bit
// type on x86 = long double here; float / double roundings happen where D's types make them happen.
//
// Restatement-defined (the reference leaves them open, pinned here and in the CUDA kernel):
//  * computeLikelihoods sorts the read bases with Phobos' UNSTABLE sort!"a.quality < b.quality" (maq.d:151): the order
//    of bases of equal quality is whatever that Phobos version's shortSort does.  Pinned: STABLE — bases of equal quality
//    keep their order in the column.  (The order matters: it decides which strand's dependency weight fk(w) meets which
//    beta(q, n, k).)
//  * more than 255 valid bases are subsampled with randomSample (maq.d:142-147), which is seeded unpredictably.  Pinned:
//    the first 255 valid bases of the column.
struct MaqTables {
  std::vector<double> fk, beta, lhet;
  MaqTables(double depcorr, double eta) : fk(256), beta((size_t)256 * 256 * 64, 0.0), lhet((size_t)256 * 256, 0.0) {
    for (size_t n = 0; n < 256; ++n) fk[n] = pow(1.0 - depcorr, (double)n) * (1.0 - eta) + eta;      // maq.d:85-87
    static double lC[256][256];
    double lG[256];
    const long double LN2l = 0.693147180559945309417232121458176568L, LN10l = 2.302585092994045684017991454684364208L;
    for (size_t n = 0; n <= 255; ++n) {                                                              // :95-103
      lG[n] = lgamma((double)(n + 1));
      for (size_t k = 0; k <= n / 2; ++k) {
        lC[n][n - k] = lC[n][k] = lG[n] - lG[k] - lG[n - k];
        // `lC[n][k] - n * cast(double)LN2`
        lhet[n << 8 | (n - k)] = lhet[n << 8 | k] = lC[n][k] - (double)n * (double)LN2l;
      }
    }
    for (size_t q = 1; q < 64; ++q) {                                                                // :105-119
      const long double e = powl(10.0L, -((long double)q) / 10.0L);
      const long double le = logl(e), le1 = logl(1.0L - e);
      for (int n = 1; n <= 255; ++n) {
        long double sum = 0.0L, sum1 = 0.0L;
        for (int k = n; k >= 0; --k) {
          sum = sum1 + expl((long double)lC[n][k] + k * le + (n - k) * le1);
          beta[q << 16 | (size_t)n << 8 | (size_t)k] = (double)(-10.0L / LN10l * logl(sum1 / sum));
          sum1 = sum;
        }
      }
    }
  }
};

// Base16 internal codes of A, C, G, T (bio/core/base.d); Base5 codes are 0..3, N = 4
static inline int base16_code(uint8_t ch) {
  static const char* tab = "=ACMGRSVTWYHKDBN";
  for (int i = 0; i < 16; ++i) if (tab[i] == (char)toupper(ch)) return i;
  return 15;
}

// computeLikelihoods (maq.d:138-248): scores[25] indexed by DiploidGenotype!Base5 code (first * 5 + second), FLT_MIN =
// absent (TinyMap useDefaultValue with float.min).  base / qual / rev: the valid read bases of the column, in column order.
static void maq_likelihoods(const MaqTables& T, const uint8_t* base, const uint8_t* qual, const uint8_t* rev, size_t n_in,
                            float scores[25]) {
  const size_t n = std::min<size_t>(n_in, 255);
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return qual[a] < qual[b]; });
  uint32_t w[32] = {0}, c[16] = {0};
  double fsum[16] = {0}, bsum[16] = {0};
  for (size_t r = n; r-- > 0;) {                       // foreach_reverse: highest quality first
    const uint32_t i = order[r];
    uint32_t quality = qual[i];
    if (quality < 4) quality = 4;
    if (quality > 63) quality = 63;
    const int b = base16_code(base[i]);
    const int bws = b * 2 + (rev[i] ? 1 : 0);
    fsum[b] += T.fk[w[bws]];
    bsum[b] += T.fk[w[bws]] * T.beta[(size_t)quality << 16 | n << 8 | c[b]];
    c[b] += 1;
    w[bws] += 1;
  }
  static const int nuc[4] = {1, 2, 4, 8};                // Base16 codes of A, C, G, T
  const long double C = 10.0L / 2.302585092994045684017991454684364208L;   // `immutable C = 10.0 / LN10` is a real
  for (int g = 0; g < 25; ++g) scores[g] = FLT_MIN;
  for (int i = 0; i < 4; ++i) {
    float tmp1 = 0.0f, tmp3 = 0.0f;
    int tmp2 = 0;
    for (int k = 0; k < 4; ++k)
      if (k != i) { tmp1 = (float)((double)tmp1 + bsum[nuc[k]]); tmp2 += (int)c[nuc[k]]; tmp3 = (float)((double)tmp3 + fsum[nuc[k]]); }
    scores[i * 5 + i] = tmp2 > 0 ? tmp1 : 0.0f;          // homozygous
    for (int j = i + 1; j < 4; ++j) {                    // heterozygous: dG(b2, b1) with b2 the later nucleotide
      const int cij = (int)(c[nuc[i]] + c[nuc[j]]);
      tmp1 = tmp3 = 0.0f;
      tmp2 = 0;
      for (int k = 0; k < 4; ++k)
        if (k != i && k != j) { tmp1 = (float)((double)tmp1 + bsum[nuc[k]]); tmp2 += (int)c[nuc[k]]; tmp3 = (float)((double)tmp3 + fsum[nuc[k]]); }
      const double lh = T.lhet[(size_t)cij << 8 | c[nuc[j]]];
      scores[j * 5 + i] = tmp2 > 0 ? (float)((long double)tmp1 - C * (long double)lh) : (float)(-C * (long double)lh);
    }
    for (int k = 0; k < 4; ++k) {                        // :236-241
      const int g = i * 5 + k;
      if (scores[g] != FLT_MIN && scores[g] < 0.0f) scores[g] = 0.0f;
    }
  }
}

struct MaqResult {
  std::vector<uint32_t> n_valid;        // per column
  std::vector<uint8_t> gt0, gt1;        // best and second genotype (Base5 pair codes), 255 = none
  std::vector<float> s0, s1;            // their scores
  std::vector<float> scores;            // 25 per column (FLT_MIN = absent)
};

extern "C" {

typedef struct MaqTables orc_maq;
orc_maq* orc_maq_new(double depcorr, double eta) { return new MaqTables(depcorr, eta); }
void orc_maq_free(orc_maq* t) { delete t; }
const double* orc_maq_fk(const orc_maq* t) { return t->fk.data(); }
const double* orc_maq_beta(const orc_maq* t) { return t->beta.data(); }
const double* orc_maq_lhet(const orc_maq* t) { return t->lhet.data(); }
// computeLikelihoods over explicit read bases (already filtered): scores[25]
void orc_maq_compute(const orc_maq* t, const uint8_t* base, const uint8_t* qual, const uint8_t* rev, uint64_t n, float* scores) {
  maq_likelihoods(*t, base, qual, rev, (size_t)n, scores);
}

// MaqSnpCaller.genotypeLikelihoodInfo over every column of a pileup (maq.d:388-457) + the ordering of
// GenotypeLikelihoodInfo (:258-277: genotypes in code order, insertion-sorted by score, equal scores keep code order).
// Per column: n_valid, best / second genotype and their scores; makeCall's quality is s1 - s0 (:480-482).
typedef struct MaqResult orc_maq_result;
orc_maq_result* orc_maq_run(orc_bam* s, const orc_pileup* p, const orc_maq* t, int min_base_quality) {
  decode_records(s);
  MaqResult* r = new MaqResult;
  const size_t nc = p->col_pos.size();
  std::vector<uint8_t> b, q, rv;
  for (size_t c = 0; c < nc; ++c) {
    b.clear(); q.clear(); rv.clear();
    for (uint64_t e = p->col_off[c]; e < p->col_off[c + 1]; ++e) {
      const uint32_t rd = p->read_idx[e];
      if (p->qual[e] < min_base_quality) continue;        // current_base_quality < minimum_base_quality
      if (p->base[e] == '-') continue;
      b.push_back(p->base[e]);
      q.push_back(std::min<uint8_t>(p->qual[e], s->mapq[rd]));   // min(base quality, mapping quality)
      rv.push_back((s->flag[rd] & 0x10) ? 1 : 0);
    }
    float sc[25];
    for (int g = 0; g < 25; ++g) sc[g] = FLT_MIN;
    uint8_t g0 = 255, g1 = 255;
    float s0 = 0, s1 = 0;
    if (!b.empty()) {
      maq_likelihoods(*t, b.data(), q.data(), rv.data(), b.size(), sc);
      int buf[25], k = 0;
      for (int g = 0; g < 25; ++g) {
        if (sc[g] == FLT_MIN) continue;
        int j = k;
        while (j > 0 && sc[buf[j - 1]] > sc[g]) { buf[j] = buf[j - 1]; --j; }
        buf[j] = g;
        ++k;
      }
      if (k >= 2) { g0 = (uint8_t)buf[0]; g1 = (uint8_t)buf[1]; s0 = sc[buf[0]]; s1 = sc[buf[1]]; }
    }
    r->n_valid.push_back((uint32_t)b.size());
    r->gt0.push_back(g0); r->gt1.push_back(g1); r->s0.push_back(s0); r->s1.push_back(s1);
    r->scores.insert(r->scores.end(), sc, sc + 25);
  }
  return r;
}
void orc_maq_result_free(orc_maq_result* r) { delete r; }
const uint32_t* orc_maq_n_valid(const orc_maq_result* r) { return r->n_valid.data(); }
const uint8_t* orc_maq_gt0(const orc_maq_result* r) { return r->gt0.data(); }
const uint8_t* orc_maq_gt1(const orc_maq_result* r) { return r->gt1.data(); }
const float* orc_maq_s0(const orc_maq_result* r) { return r->s0.data(); }
const float* orc_maq_s1(const orc_maq_result* r) { return r->s1.data(); }
const float* orc_maq_scores(const orc_maq_result* r) { return r->scores.data(); }

}  // extern "C"

// ------------------------------------------------------- CPU baseline legs ----
// "Restated BioD CPU path (libz, g++ -O3), not the D binary" (BASELINE.md §2):
// inflate with `threads` worker threads the way BgzfInputStream hands whole
// blocks to a TaskPool (inputstream.d:414-417; default totalCPUs-1 workers),
// framing + field decode + pileup single-threaded by construction
// (readrange.d:118-173, pileup.d:345-397).  Returns seconds per leg.
static thread_local double g_last_lazy_pileup_s = 0;
int orc_cpu_baseline(const uint8_t* data, uint64_t len, int threads, int do_pileup, double* t_inflate,
                     double* t_decode, double* t_pileup, uint64_t* n_records, uint64_t* n_columns,
                     uint64_t* n_entries, uint64_t* checksum) {
  auto now = [] {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  };
  Bam* s = new Bam;
  s->file = data;
  s->flen = len;
  double t0 = now();
  // pass 1: header chain walk (serial, as the consumer thread does)
  std::vector<Block> blocks;
  uint64_t pos = 0, utotal = 0;
  while (true) {
    Block b;
    int r = parse_bgzf_header(data, len, pos, &b, &s->err);
    if (r < 0) { int st = s->err.status; delete s; return st; }
    if (r == 0 || b.isize == 0) break;
    if (b.isize > BGZF_MAX_BLOCK_SIZE) { delete s; return ORC_ERR_FORMAT; }
    b.uoffset = utotal;
    utotal += b.isize;
    blocks.push_back(b);
    pos = b.coffset + b.bsize + 1;
  }
  s->u.resize(utotal);
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  std::vector<int> bad(threads, 0);
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([&, t] {
      Error e;
      for (size_t i = t; i < blocks.size(); i += threads)
        if (inflate_block(data, blocks[i], s->u.data() + blocks[i].uoffset, &e) < 0) bad[t] = e.status;
    });
  }
  for (auto& th : pool) th.join();
  for (int t = 0; t < threads; t++) if (bad[t]) { int st = bad[t]; delete s; return st; }
  s->blocks = blocks;
  s->stream_done = true;
  double t1 = now();
  // header parse over the already inflated stream, then the record walk
  int r = parse_header(s);
  if (r < 0) { delete s; return r; }
  r = decode_records(s);
  if (r < 0) { delete s; return r; }
  double t2 = now();
  *t_inflate = t1 - t0;
  *t_decode = t2 - t1;
  *n_records = s->rec_off.size();
  *t_pileup = 0;
  *n_columns = *n_entries = 0;
  uint64_t cs = 0;
  for (size_t i = 0; i < s->pos.size(); i++) cs = cs * 1099511628211ull + (uint32_t)s->end_pos[i];
  if (do_pileup) {
    Pileup* p = run_pileup(s, 0, 0, ~0ull, 1, 0, -1);
    double t3 = now();
    *t_pileup = t3 - t2;
    *n_columns = p->col_pos.size();
    *n_entries = p->read_idx.size();
    for (size_t i = 0; i < p->base.size(); i++) cs = cs * 1099511628211ull + p->base[i] + 256u * p->qual[i];
    int st = p->status;
    delete p;
    if (st) { delete s; return st; }
  }
  *checksum = cs;
  // the same sweep without materialising the columns: BioD's `foreach (column; pileupColumns(reads)) n += column.coverage`
  g_last_lazy_pileup_s = 0;
  if (do_pileup) {
    double t4 = now();
    Pileup* p = run_pileup(s, 0, 0, ~0ull, 1, 0, -1, 0, nullptr, 0, true);
    g_last_lazy_pileup_s = now() - t4;
    const bool same = p->lazy_columns == *n_columns && p->lazy_entries == *n_entries;
    delete p;
    if (!same) { delete s; return ORC_ERR_FORMAT; }
  }
  delete s;
  return 0;
}

// seconds the lazy sweep of the last orc_cpu_baseline call took (same thread)
double orc_cpu_baseline_lazy_seconds() { return g_last_lazy_pileup_s; }

}  // extern "C"
