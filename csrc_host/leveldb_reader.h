// Read-only LevelDB reader (no libleveldb / libsnappy needed).
//
// Caffe's default DataParameter backend is LEVELDB (reference: src/caffe/layers/data_layer.cpp:35-58, tools/convert_imageset.cpp,
// tools/partition_data.cpp) — existing Poseidon datasets are LevelDB directories.  This implements the read path from the
// on-disk format (leveldb/doc/{log_format,table_format,impl}.md):
//   CURRENT            -> name of the live MANIFEST
//   MANIFEST-nnnnnn    -> log-format file of VersionEdit records: which table files are live (tag 7 new file, tag 6 deleted
//                         file), the current write-ahead log number (tag 2) ...
//   nnnnnn.ldb / .sst  -> sorted tables: data blocks (prefix-compressed entries + restart array, optional snappy), index
//                         block, 48-byte footer ending in magic 0xdb4775248b80fb57
//   nnnnnn.log         -> write-ahead log of WriteBatches not yet flushed to a table
// Entries carry internal keys (user key + 8 bytes: sequence << 8 | type); the newest sequence of every user key wins and
// deletions are dropped.  Keys are ordered bytewise (leveldb.BytewiseComparator, what Caffe uses).
//
// Validated against a writer following the same documents (tests/test_leveldb_reader.py); the snappy decoder is additionally
// cross-checked against pyarrow's codec.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace psd_host {

struct LdbRecord {
  const uint8_t* key;
  uint32_t klen;
  const uint8_t* val;
  uint32_t vlen;
};

class LevelDBIndex {
 public:
  explicit LevelDBIndex(const std::string& dir);
  ~LevelDBIndex();
  const std::vector<LdbRecord>& records() const { return recs_; }

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
  std::vector<LdbRecord> recs_;
};

// raw snappy block format; throws std::runtime_error on malformed input
std::vector<uint8_t> snappy_uncompress(const uint8_t* src, size_t n);

}  // namespace psd_host
