// See leveldb_reader.h.
#include "leveldb_reader.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>

namespace psd_host {
namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error("LevelDB: " + m); }

struct Slice {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

bool get_varint64(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift <= 63 && p < end; shift += 7) {
    const uint8_t b = *p++;
    v |= static_cast<uint64_t>(b & 0x7f) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}
bool get_varint32(const uint8_t*& p, const uint8_t* end, uint32_t& v) {
  uint64_t t;
  if (!get_varint64(p, end, t) || t > 0xffffffffull) return false;
  v = static_cast<uint32_t>(t);
  return true;
}
bool get_length_prefixed(const uint8_t*& p, const uint8_t* end, Slice& s) {
  uint32_t n;
  if (!get_varint32(p, end, n) || static_cast<size_t>(end - p) < n) return false;
  s.p = p;
  s.n = n;
  p += n;
  return true;
}
uint32_t rd32(const uint8_t* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}
uint64_t rd64(const uint8_t* p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;
}

// A whole file mapped read-only for the life of the index.
struct Mapped {
  int fd = -1;
  const uint8_t* base = nullptr;
  size_t size = 0;
  explicit Mapped(const std::string& path) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) fail("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); fail("cannot stat " + path); }
    size = static_cast<size_t>(st.st_size);
    if (size > 0) {
      void* m = mmap(nullptr, size, PROT_READ, MAP_SHARED, fd, 0);
      if (m == MAP_FAILED) { ::close(fd); fail("mmap failed for " + path); }
      base = static_cast<const uint8_t*>(m);
    }
  }
  ~Mapped() {
    if (base != nullptr) munmap(const_cast<uint8_t*>(base), size);
    if (fd >= 0) ::close(fd);
  }
  Mapped(const Mapped&) = delete;
  Mapped& operator=(const Mapped&) = delete;
};

// ---- log format (MANIFEST and write-ahead logs): 32 KiB blocks of [crc32c u32 | length u16 | type u8 | payload] ----------------
constexpr size_t kLogBlock = 32768;
enum { kFull = 1, kFirst = 2, kMiddle = 3, kLast = 4 };

std::vector<std::vector<uint8_t>> read_log_records(const Mapped& f) {
  std::vector<std::vector<uint8_t>> out;
  std::vector<uint8_t> cur;
  bool in_fragment = false;
  size_t off = 0;
  while (off + 7 <= f.size) {
    const size_t left = kLogBlock - (off % kLogBlock);
    if (left < 7) { off += left; continue; }                      // block trailer (zero padding)
    const uint32_t len = f.base[off + 4] | (static_cast<uint32_t>(f.base[off + 5]) << 8);
    const uint8_t type = f.base[off + 6];
    if (type == 0 && len == 0) { off += left; continue; }         // preallocated / zeroed tail
    if (7 + static_cast<size_t>(len) > left || off + 7 + len > f.size) break;   // truncated tail: stop (as leveldb does)
    const uint8_t* payload = f.base + off + 7;
    switch (type) {
      case kFull:
        out.emplace_back(payload, payload + len);
        in_fragment = false;
        break;
      case kFirst:
        cur.assign(payload, payload + len);
        in_fragment = true;
        break;
      case kMiddle:
        if (in_fragment) cur.insert(cur.end(), payload, payload + len);
        break;
      case kLast:
        if (in_fragment) {
          cur.insert(cur.end(), payload, payload + len);
          out.push_back(std::move(cur));
          cur.clear();
          in_fragment = false;
        }
        break;
      default:
        fail("unknown log record type");
    }
    off += 7 + len;
  }
  return out;
}

struct Entry {
  Slice user_key;
  uint64_t seq;
  uint8_t type;          // 1 value, 0 deletion
  Slice value;
};

}  // namespace

// ---- snappy (raw block format) -----------------------------------------------------------------------------------------------
std::vector<uint8_t> snappy_uncompress(const uint8_t* src, size_t n) {
  const uint8_t* p = src;
  const uint8_t* end = src + n;
  uint64_t ulen;
  if (!get_varint64(p, end, ulen) || ulen > (1ull << 32)) throw std::runtime_error("snappy: bad length preamble");
  std::vector<uint8_t> out;
  out.reserve(static_cast<size_t>(ulen));
  while (p < end) {
    const uint8_t tag = *p++;
    const int kind = tag & 3;
    if (kind == 0) {                                   // literal
      size_t len = (tag >> 2) + 1;
      if (len > 60) {
        const int extra = static_cast<int>(len) - 60;  // 1..4 bytes of length follow
        if (end - p < extra) throw std::runtime_error("snappy: truncated literal length");
        len = 0;
        for (int i = 0; i < extra; ++i) len |= static_cast<size_t>(p[i]) << (8 * i);
        len += 1;
        p += extra;
      }
      if (static_cast<size_t>(end - p) < len) throw std::runtime_error("snappy: truncated literal");
      out.insert(out.end(), p, p + len);
      p += len;
    } else {
      size_t len, offset;
      if (kind == 1) {
        if (end - p < 1) throw std::runtime_error("snappy: truncated copy");
        len = 4 + ((tag >> 2) & 7);
        offset = (static_cast<size_t>(tag >> 5) << 8) | *p++;
      } else if (kind == 2) {
        if (end - p < 2) throw std::runtime_error("snappy: truncated copy");
        len = (tag >> 2) + 1;
        offset = p[0] | (static_cast<size_t>(p[1]) << 8);
        p += 2;
      } else {
        if (end - p < 4) throw std::runtime_error("snappy: truncated copy");
        len = (tag >> 2) + 1;
        offset = rd32(p);
        p += 4;
      }
      if (offset == 0 || offset > out.size()) throw std::runtime_error("snappy: bad copy offset");
      const size_t start = out.size() - offset;
      for (size_t i = 0; i < len; ++i) out.push_back(out[start + i]);     // may overlap: byte by byte
    }
  }
  if (out.size() != ulen) throw std::runtime_error("snappy: length mismatch");
  return out;
}

struct LevelDBIndex::Impl {
  std::deque<std::unique_ptr<Mapped>> files;             // keep every mapping alive: records point into them
  std::deque<std::vector<uint8_t>> owned;                // decompressed blocks, reassembled keys, WAL payloads
  std::vector<Entry> entries;

  const uint8_t* keep(std::vector<uint8_t>&& v) {
    owned.push_back(std::move(v));
    return owned.back().data();
  }

  // a block (data or index): returns a pointer to its uncompressed contents
  Slice read_block(const Mapped& f, uint64_t off, uint64_t size) {
    if (off + size + 5 > f.size) fail("block handle past end of table");
    const uint8_t type = f.base[off + size];
    if (type == 0) return Slice{f.base + off, static_cast<size_t>(size)};
    if (type == 1) {
      std::vector<uint8_t> raw = snappy_uncompress(f.base + off, static_cast<size_t>(size));
      const size_t n = raw.size();
      return Slice{keep(std::move(raw)), n};
    }
    fail("unknown block compression type");
  }

  template <class Fn>
  void for_each_block_entry(Slice blk, Fn&& fn) {
    if (blk.n < 4) fail("block too small");
    const uint32_t nrestart = rd32(blk.p + blk.n - 4);
    if (static_cast<uint64_t>(nrestart) * 4 + 4 > blk.n) fail("bad restart array");
    const uint8_t* p = blk.p;
    const uint8_t* end = blk.p + blk.n - 4 - 4 * static_cast<size_t>(nrestart);
    std::vector<uint8_t> key;
    while (p < end) {
      uint32_t shared, non_shared, vlen;
      if (!get_varint32(p, end, shared) || !get_varint32(p, end, non_shared) || !get_varint32(p, end, vlen)) fail("bad block entry");
      if (shared > key.size() || static_cast<size_t>(end - p) < static_cast<size_t>(non_shared) + vlen) fail("bad block entry sizes");
      key.resize(shared);
      key.insert(key.end(), p, p + non_shared);
      p += non_shared;
      fn(key, Slice{p, vlen});
      p += vlen;
    }
  }

  void read_table(const std::string& path) {
    files.push_back(std::make_unique<Mapped>(path));
    const Mapped& f = *files.back();
    if (f.size < 48) fail(path + ": too small for a table");
    const uint8_t* footer = f.base + f.size - 48;
    if (rd64(footer + 40) != 0xdb4775248b80fb57ull) fail(path + ": bad table magic");
    const uint8_t* p = footer;
    const uint8_t* fend = footer + 40;
    uint64_t mo, ms, io, is;
    if (!get_varint64(p, fend, mo) || !get_varint64(p, fend, ms) || !get_varint64(p, fend, io) || !get_varint64(p, fend, is))
      fail(path + ": bad footer");
    const Slice index = read_block(f, io, is);
    std::vector<std::pair<uint64_t, uint64_t>> handles;
    for_each_block_entry(index, [&](const std::vector<uint8_t>&, Slice v) {
      const uint8_t* q = v.p;
      uint64_t bo, bs;
      if (!get_varint64(q, v.p + v.n, bo) || !get_varint64(q, v.p + v.n, bs)) fail("bad block handle in index");
      handles.emplace_back(bo, bs);
    });
    for (auto& h : handles) {
      const Slice data = read_block(f, h.first, h.second);
      for_each_block_entry(data, [&](const std::vector<uint8_t>& ikey, Slice v) {
        if (ikey.size() < 8) fail("internal key shorter than its trailer");
        const uint64_t tag = rd64(ikey.data() + ikey.size() - 8);
        std::vector<uint8_t> uk(ikey.begin(), ikey.end() - 8);      // keys are prefix-compressed: reassembled copies are owned
        const size_t n = uk.size();
        entries.push_back(Entry{Slice{keep(std::move(uk)), n}, tag >> 8, static_cast<uint8_t>(tag & 0xff), v});
      });
    }
  }

  void read_wal(const std::string& path) {
    files.push_back(std::make_unique<Mapped>(path));
    for (auto& rec : read_log_records(*files.back())) {
      if (rec.size() < 12) continue;
      const size_t n = rec.size();
      const uint8_t* base = keep(std::move(rec));
      uint64_t seq = rd64(base);
      const uint32_t count = rd32(base + 8);
      const uint8_t* p = base + 12;
      const uint8_t* end = base + n;
      for (uint32_t i = 0; i < count && p < end; ++i, ++seq) {
        const uint8_t type = *p++;
        Slice k, v;
        if (!get_length_prefixed(p, end, k)) fail("bad WriteBatch key");
        if (type == 1 && !get_length_prefixed(p, end, v)) fail("bad WriteBatch value");
        entries.push_back(Entry{k, seq, type, v});
      }
    }
  }
};

LevelDBIndex::LevelDBIndex(const std::string& dir) : impl_(new Impl) {
  // CURRENT -> MANIFEST
  std::ifstream cur(dir + "/CURRENT");
  std::string manifest;
  if (!cur || !std::getline(cur, manifest) || manifest.empty()) fail(dir + ": no CURRENT file");
  while (!manifest.empty() && (manifest.back() == '\n' || manifest.back() == '\r')) manifest.pop_back();
  Mapped mf(dir + "/" + manifest);
  std::set<uint64_t> live;
  uint64_t log_number = 0, prev_log = 0;
  for (auto& rec : read_log_records(mf)) {
    const uint8_t* p = rec.data();
    const uint8_t* end = p + rec.size();
    while (p < end) {
      uint32_t tag;
      if (!get_varint32(p, end, tag)) fail("bad VersionEdit tag");
      uint64_t v64;
      uint32_t level;
      Slice s;
      switch (tag) {
        case 1:                                   // comparator name
          if (!get_length_prefixed(p, end, s)) fail("bad comparator");
          if (std::string(reinterpret_cast<const char*>(s.p), s.n) != "leveldb.BytewiseComparator")
            fail("unsupported comparator (only leveldb.BytewiseComparator)");
          break;
        case 2: if (!get_varint64(p, end, log_number)) fail("bad log number"); break;
        case 9: if (!get_varint64(p, end, prev_log)) fail("bad prev log number"); break;
        case 3: case 4: if (!get_varint64(p, end, v64)) fail("bad VersionEdit number"); break;
        case 5:                                   // compact pointer
          if (!get_varint32(p, end, level) || !get_length_prefixed(p, end, s)) fail("bad compact pointer");
          break;
        case 6:                                   // deleted file
          if (!get_varint32(p, end, level) || !get_varint64(p, end, v64)) fail("bad deleted-file record");
          live.erase(v64);
          break;
        case 7: {                                 // new file
          uint64_t num, fsize;
          Slice a, b;
          if (!get_varint32(p, end, level) || !get_varint64(p, end, num) || !get_varint64(p, end, fsize) ||
              !get_length_prefixed(p, end, a) || !get_length_prefixed(p, end, b))
            fail("bad new-file record");
          live.insert(num);
          break;
        }
        default:
          fail("unknown VersionEdit tag " + std::to_string(tag));
      }
    }
  }
  auto name = [&](uint64_t num, const char* ext) {
    char buf[32];
    snprintf(buf, sizeof buf, "/%06llu.%s", static_cast<unsigned long long>(num), ext);
    return dir + buf;
  };
  auto exists = [](const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; };
  for (uint64_t num : live) {
    if (exists(name(num, "ldb"))) impl_->read_table(name(num, "ldb"));
    else if (exists(name(num, "sst"))) impl_->read_table(name(num, "sst"));
    else fail("live table " + std::to_string(num) + " is missing");
  }
  // write-ahead logs that were never flushed: every NNNNNN.log with number >= log_number (or == prev_log)
  std::vector<uint64_t> logs;
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string fn = e->d_name;
      if (fn.size() > 4 && fn.substr(fn.size() - 4) == ".log") {
        char* endp = nullptr;
        const uint64_t num = strtoull(fn.c_str(), &endp, 10);
        if (endp != fn.c_str() && (num >= log_number || num == prev_log)) logs.push_back(num);
      }
    }
    closedir(d);
  }
  std::sort(logs.begin(), logs.end());
  for (uint64_t num : logs) impl_->read_wal(name(num, "log"));

  // newest sequence per user key wins; deletions drop the key
  auto& es = impl_->entries;
  std::sort(es.begin(), es.end(), [](const Entry& a, const Entry& b) {
    const size_t n = std::min(a.user_key.n, b.user_key.n);
    const int c = n ? std::memcmp(a.user_key.p, b.user_key.p, n) : 0;
    if (c != 0) return c < 0;
    if (a.user_key.n != b.user_key.n) return a.user_key.n < b.user_key.n;
    return a.seq > b.seq;
  });
  for (size_t i = 0; i < es.size(); ++i) {
    if (i > 0 && es[i].user_key.n == es[i - 1].user_key.n &&
        (es[i].user_key.n == 0 || std::memcmp(es[i].user_key.p, es[i - 1].user_key.p, es[i].user_key.n) == 0))
      continue;                                                  // an older version of the same key
    if (es[i].type != 1) continue;                               // deleted
    recs_.push_back(LdbRecord{es[i].user_key.p, static_cast<uint32_t>(es[i].user_key.n), es[i].value.p,
                              static_cast<uint32_t>(es[i].value.n)});
  }
}

LevelDBIndex::~LevelDBIndex() = default;

}  // namespace psd_host
