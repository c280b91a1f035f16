// LibSVM text parser for the ml/ helper library (reference: ps/src/ml/parsers/libsvm_parser.hpp,
// ps/src/ml/util/data_loading.cpp — one sample per line, "label id:value id:value ...").
//
// The buffer is cut at line boundaries into one slice per hardware thread; every slice is parsed independently into
// CSR pieces that are stitched together afterwards.  Output: labels int32 [rows], indptr int64 [rows + 1],
// indices int32 [nnz], values float32 [nnz].
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace psd_host {

struct CsrPiece {
  std::vector<int32_t> labels, indices;
  std::vector<int64_t> row_nnz;
  std::vector<float> values;
  std::string error;
};

// std::from_chars: locale-free, bounds-checked (no NUL terminator needed) and several times faster than strtod
template <typename T>
static inline bool parse_num(const char*& q, const char* end, T* out) {
  if (q < end && *q == '+') ++q;                         // from_chars rejects an explicit plus sign
  const auto r = std::from_chars(q, end, *out);
  if (r.ec != std::errc() || r.ptr == q) return false;
  q = r.ptr;
  return true;
}

static void parse_slice(const char* p, const char* end, int id_shift, int label_shift, CsrPiece* out) {
  while (p < end) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', end - p));
    if (eol == nullptr) eol = end;
    const char* q = p;
    while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
    if (q < eol && *q != '#') {
      double lab = 0;
      if (!parse_num(q, eol, &lab)) { out->error = "libsvm: line does not start with a label"; return; }
      out->labels.push_back(static_cast<int32_t>(lab) - label_shift);
      int64_t nnz = 0;
      while (q < eol) {
        while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        if (q >= eol || *q == '#') break;
        long id = 0;
        if (!parse_num(q, eol, &id) || q >= eol || *q != ':') { out->error = "libsvm: expected id:value"; return; }
        ++q;
        float v = 0.f;
        if (!parse_num(q, eol, &v)) { out->error = "libsvm: expected a value after ':'"; return; }
        if (id - id_shift < 0) { out->error = "libsvm: negative feature id (is the file one-based?)"; return; }
        out->indices.push_back(static_cast<int32_t>(id - id_shift));
        out->values.push_back(v);
        ++nnz;
      }
      out->row_nnz.push_back(nnz);
    }
    p = eol + 1;
  }
}

static py::tuple parse_libsvm(py::bytes data, bool feature_one_based, bool label_one_based, int64_t max_rows, int threads) {
  char* raw = nullptr;
  Py_ssize_t len = 0;
  if (PyBytes_AsStringAndSize(data.ptr(), &raw, &len) != 0) throw py::error_already_set();
  const char* base = raw;                    // parsed in place: the bytes object outlives the call
  const size_t n = static_cast<size_t>(len);
  int T = threads > 0 ? threads : static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  T = static_cast<int>(std::min<size_t>(T, std::max<size_t>(1, n / (1 << 16))));
  std::vector<size_t> cut(T + 1, n);
  cut[0] = 0;
  for (int t = 1; t < T; ++t) {
    size_t c = n * t / T;
    while (c < n && base[c] != '\n') ++c;
    cut[t] = std::min(n, c + 1);
  }
  std::vector<CsrPiece> pieces(T);
  {
    py::gil_scoped_release nogil;
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
      pool.emplace_back(parse_slice, base + cut[t], base + cut[t + 1], feature_one_based ? 1 : 0, label_one_based ? 1 : 0,
                        &pieces[t]);
    for (auto& th : pool) th.join();
  }
  int64_t rows = 0, nnz = 0;
  for (auto& pc : pieces) {
    if (!pc.error.empty()) throw std::runtime_error(pc.error);
    rows += static_cast<int64_t>(pc.labels.size());
  }
  if (max_rows >= 0) rows = std::min(rows, max_rows);
  {
    int64_t left = rows;
    for (auto& pc : pieces) {
      const int64_t take = std::min<int64_t>(left, pc.labels.size());
      for (int64_t i = 0; i < take; ++i) nnz += pc.row_nnz[i];
      left -= take;
    }
  }
  py::array_t<int32_t> labels(rows), indices(nnz);
  py::array_t<int64_t> indptr(rows + 1);
  py::array_t<float> values(nnz);
  int32_t* lp = labels.mutable_data();
  int32_t* ip = indices.mutable_data();
  int64_t* pp = indptr.mutable_data();
  float* vp = values.mutable_data();
  int64_t r = 0, z = 0;
  pp[0] = 0;
  for (auto& pc : pieces) {
    int64_t off = 0;
    for (size_t i = 0; i < pc.labels.size() && r < rows; ++i, ++r) {
      lp[r] = pc.labels[i];
      const int64_t k = pc.row_nnz[i];
      std::copy(pc.indices.begin() + off, pc.indices.begin() + off + k, ip + z);
      std::copy(pc.values.begin() + off, pc.values.begin() + off + k, vp + z);
      off += k;
      z += k;
      pp[r + 1] = z;
    }
  }
  return py::make_tuple(labels, indptr, indices, values);
}

// CRC-32C (Castagnoli), slicing-by-4 — the checksum of LevelDB blocks and log records (data/leveldb_writer.py).
static uint32_t crc32c_update(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[4][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 4; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xFF];
    init = true;
  }
  crc = ~crc;
  while (n >= 4) {
    crc ^= static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
           (static_cast<uint32_t>(p[3]) << 24);
    crc = table[3][crc & 0xFF] ^ table[2][(crc >> 8) & 0xFF] ^ table[1][(crc >> 16) & 0xFF] ^ table[0][crc >> 24];
    p += 4;
    n -= 4;
  }
  while (n--) crc = table[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

void bind_ml(py::module_& m) {
  m.def("crc32c", [](py::bytes data, uint32_t seed) {
    const std::string s = data;
    return crc32c_update(seed, reinterpret_cast<const uint8_t*>(s.data()), s.size());
  }, py::arg("data"), py::arg("seed") = 0, "CRC-32C (Castagnoli) of a byte string, continuing from `seed`");
  m.def("parse_libsvm", &parse_libsvm, py::arg("data"), py::arg("feature_one_based") = false,
        py::arg("label_one_based") = false, py::arg("max_rows") = -1, py::arg("threads") = 0,
        "LibSVM text -> (labels int32, indptr int64, indices int32, values float32), parsed on a thread per slice");
}

}  // namespace psd_host
