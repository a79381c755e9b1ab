// Multithreaded mmap LIBSVM parser (C ABI, loaded through ctypes).
//
// Behavioural parity with MLUtils.loadLibSVMFile / parseLibSVMRecord
// (mllib/util/MLUtils.scala:97-148): lines are trimmed; empty lines and lines starting with
// '#' are skipped; tokens are split on single spaces with empty tokens dropped; the first
// token is the label (double); the rest are `index:value` with 1-based indices converted to
// 0-based; indices must be strictly increasing (so a 0 index in the file fails); a label-only
// line gives an all-zero row.  Output is CSR (labels, indptr, indices, values) plus the
// maximum index seen (for the auto feature count, MLUtils.scala:91-95).
//
// Design: the file is mmap'ed, cut into T byte ranges aligned to line starts, each parsed by
// a std::thread into thread-local vectors, then stitched.  No JVM, no per-record allocation.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Chunk {
  std::vector<double> labels;
  std::vector<int64_t> row_nnz;
  std::vector<int32_t> indices;
  std::vector<double> values;
  int32_t max_index = -1;
  std::string error;          // "{LINE}" is replaced by the absolute line number by afb_libsvm_parse
  int64_t error_line = 0;     // line of the error, relative to the chunk (1-based)
};

// A parsed file stays in its per-thread chunks; afb_libsvm_copy() scatters them straight into the caller's arrays
// (one parallel copy instead of a serial stitch followed by a second copy).
struct Parsed {
  std::vector<Chunk> chunks;
  std::vector<size_t> row_off;   // first row of chunk c
  std::vector<size_t> nnz_off;   // first non-zero of chunk c
  size_t rows = 0;
  size_t nnz = 0;
  int32_t max_index = -1;
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

// strtod on a bounded token (tokens are short; copy into a stack buffer to NUL-terminate).  Slow path only.
inline bool parse_double_slow(const char* b, const char* e, double* out) {
  char buf[64];
  size_t n = static_cast<size_t>(e - b);
  if (n == 0 || n >= sizeof(buf)) return false;
  memcpy(buf, b, n);
  buf[n] = 0;
  char* endp = nullptr;
  errno = 0;
  double v = strtod(buf, &endp);
  if (endp != buf + n) return false;
  *out = v;
  return true;
}

constexpr double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Plain decimals ([+-]digits[.digits], the only form LIBSVM writers emit for pixel / tf-idf data) are converted
// exactly without strtod: when the digit string fits in 2^53 and there are <= 22 fraction digits, mantissa and
// 10^frac are both exactly representable and one IEEE division gives the correctly rounded result (Clinger's
// fast path).  Anything else (exponents, inf/nan, hex, > 15-16 digits) goes to strtod, so acceptance and values
// are identical to the slow path.
inline bool parse_double(const char* b, const char* e, double* out) {
  const char* p = b;
  if (p == e) return false;
  bool neg = false;
  if (*p == '-' || *p == '+') { neg = (*p == '-'); ++p; }
  uint64_t mant = 0;
  int digits = 0, frac = 0;
  while (p < e && static_cast<unsigned>(*p - '0') <= 9u) { mant = mant * 10 + static_cast<unsigned>(*p - '0'); ++p; ++digits; }
  if (p < e && *p == '.') {
    ++p;
    while (p < e && static_cast<unsigned>(*p - '0') <= 9u) { mant = mant * 10 + static_cast<unsigned>(*p - '0'); ++p; ++digits; ++frac; }
  }
  if (p != e || digits == 0 || digits > 18 || mant > (1ULL << 53) || frac > 22) return parse_double_slow(b, e, out);
  double v = static_cast<double>(mant);
  if (frac) v /= kPow10[frac];
  *out = neg ? -v : v;
  return true;
}

inline bool parse_int(const char* b, const char* e, int64_t* out) {
  if (b == e) return false;
  bool neg = false;
  if (*b == '-' || *b == '+') { neg = (*b == '-'); ++b; if (b == e) return false; }
  int64_t v = 0;
  for (; b < e; ++b) {
    if (*b < '0' || *b > '9') return false;
    v = v * 10 + (*b - '0');
    if (v > (1LL << 40)) return false;
  }
  *out = neg ? -v : v;
  return true;
}

void parse_range(const char* data, size_t begin, size_t end, Chunk* out) {
  size_t pos = begin;
  int64_t line_no = 0;
  // size the output once: >= 4 bytes per "i:v " token, >= 2 bytes per line
  const size_t span = end - begin;
  out->indices.reserve(span / 6 + 16);
  out->values.reserve(span / 6 + 16);
  out->labels.reserve(span / 512 + 16);
  out->row_nnz.reserve(span / 512 + 16);
  while (pos < end) {
    const void* nl = memchr(data + pos, '\n', end - pos);
    size_t eol = nl ? static_cast<size_t>(static_cast<const char*>(nl) - data) : end;
    const char* b = data + pos;
    const char* e = data + eol;
    pos = eol + 1;
    ++line_no;
    while (b < e && is_ws(*b)) ++b;       // trim
    while (e > b && is_ws(e[-1])) --e;
    if (b == e || *b == '#') continue;
    // label
    const char* t = b;
    while (t < e && *t != ' ') ++t;
    double label;
    if (!parse_double(b, t, &label)) {
      out->error = "bad label at line {LINE}: '" + std::string(b, t) + "'";
      out->error_line = line_no;
      return;
    }
    out->labels.push_back(label);
    int64_t nnz = 0;
    int64_t prev = -1;
    b = t;
    while (b < e) {
      while (b < e && *b == ' ') ++b;      // empty tokens dropped
      if (b >= e) break;
      // fused scan of `index:value`: digits up to ':' are accumulated while scanning (the common case); any
      // other shape falls back to the generic token parsers below
      const char* q = b;
      int64_t idx1 = 0;
      int nd = 0;
      while (q < e && static_cast<unsigned>(*q - '0') <= 9u && nd < 12) { idx1 = idx1 * 10 + (*q - '0'); ++q; ++nd; }
      const char* colon;
      bool ok;
      if (nd > 0 && q < e && *q == ':') {
        colon = q;
        ok = true;
      } else {
        t = b;
        while (t < e && *t != ' ') ++t;
        colon = b;
        while (colon < t && *colon != ':') ++colon;
        ok = (colon != t) && parse_int(b, colon, &idx1);
      }
      t = colon;
      while (t < e && *t != ' ') ++t;
      double val;
      if (!ok || !parse_double(colon + 1, t, &val)) {
        out->error = "bad index:value token '" + std::string(b, t) + "' at line {LINE}";
        out->error_line = line_no;
        return;
      }
      int64_t idx0 = idx1 - 1;  // 1-based -> 0-based
      if (!(idx0 > prev)) {
        out->error = "indices should be one-based and in ascending order; found current=" +
                     std::to_string(idx0) + ", previous=" + std::to_string(prev) + " at line {LINE}";
        out->error_line = line_no;
        return;
      }
      if (idx0 > static_cast<int64_t>(INT32_MAX)) {
        // MLlib's `toInt` fails on such an index; narrowing it silently would wrap to a small or negative column and let the
        // gather / scatter kernels go out of bounds
        out->error = "feature index " + std::to_string(idx1) + " does not fit a 32-bit column id at line {LINE}";
        out->error_line = line_no;
        return;
      }
      prev = idx0;
      out->indices.push_back(static_cast<int32_t>(idx0));
      out->values.push_back(val);
      ++nnz;
      b = t;
    }
    out->row_nnz.push_back(nnz);
    if (prev > out->max_index) out->max_index = static_cast<int32_t>(prev);
  }
}

void set_err(char* err, int errlen, const std::string& msg) {
  if (err && errlen > 0) {
    snprintf(err, static_cast<size_t>(errlen), "%s", msg.c_str());
  }
}

}  // namespace

extern "C" {

void* afb_libsvm_parse(const char* path, int num_threads, char* err, int errlen) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) { set_err(err, errlen, std::string("cannot open ") + path + ": " + strerror(errno)); return nullptr; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); set_err(err, errlen, "fstat failed"); return nullptr; }
  size_t size = static_cast<size_t>(st.st_size);
  auto* result = new Parsed();
  if (size == 0) { close(fd); return result; }
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) { delete result; set_err(err, errlen, "mmap failed"); return nullptr; }
  madvise(map, size, MADV_SEQUENTIAL);
  const char* data = static_cast<const char*>(map);

  int T = std::max(1, num_threads);
  if (size < (1u << 16)) T = 1;
  std::vector<size_t> cuts(T + 1);
  cuts[0] = 0;
  cuts[T] = size;
  for (int t = 1; t < T; ++t) {
    size_t p = size / T * t;
    if (p < cuts[t - 1]) p = cuts[t - 1];
    while (p < size && data[p] != '\n') ++p;
    cuts[t] = std::min(size, p + 1);
  }
  std::vector<Chunk>& chunks = result->chunks;
  chunks.resize(T);
  std::vector<std::thread> threads;
  for (int t = 0; t < T; ++t)
    threads.emplace_back(parse_range, data, cuts[t], cuts[t + 1], &chunks[t]);
  for (auto& th : threads) th.join();
  // errors carry ABSOLUTE line numbers: lines before the failing chunk are counted here (error path only)
  for (int t = 0; t < T; ++t) {
    if (chunks[t].error.empty()) continue;
    int64_t before = 0;
    for (const char* q = data; q < data + cuts[t]; ) {
      const void* nl = memchr(q, '\n', static_cast<size_t>(data + cuts[t] - q));
      if (!nl) break;
      ++before;
      q = static_cast<const char*>(nl) + 1;
    }
    std::string& msg = chunks[t].error;
    const size_t at = msg.find("{LINE}");
    if (at != std::string::npos) msg.replace(at, 6, std::to_string(before + chunks[t].error_line));
    break;                                       // the first failing chunk in file order is the one reported
  }
  munmap(map, size);

  for (auto& c : chunks) {
    if (!c.error.empty()) { set_err(err, errlen, c.error); delete result; return nullptr; }
    result->row_off.push_back(result->rows);
    result->nnz_off.push_back(result->nnz);
    result->rows += c.labels.size();
    result->nnz += c.values.size();
    result->max_index = std::max(result->max_index, c.max_index);
  }
  return result;
}

int64_t afb_libsvm_rows(void* h) { return static_cast<int64_t>(static_cast<Parsed*>(h)->rows); }
int64_t afb_libsvm_nnz(void* h) { return static_cast<int64_t>(static_cast<Parsed*>(h)->nnz); }
int32_t afb_libsvm_max_index(void* h) { return static_cast<Parsed*>(h)->max_index; }

void afb_libsvm_copy(void* h, double* labels, int64_t* indptr, int32_t* indices, double* values) {
  auto* p = static_cast<Parsed*>(h);
  indptr[0] = 0;
  auto scatter = [&](size_t ci) {
    const Chunk& c = p->chunks[ci];
    const size_t r0 = p->row_off[ci], z0 = p->nnz_off[ci];
    if (!c.labels.empty()) memcpy(labels + r0, c.labels.data(), c.labels.size() * sizeof(double));
    int64_t acc = static_cast<int64_t>(z0);
    for (size_t i = 0; i < c.row_nnz.size(); ++i) { acc += c.row_nnz[i]; indptr[r0 + i + 1] = acc; }
    if (!c.values.empty()) {
      memcpy(indices + z0, c.indices.data(), c.indices.size() * sizeof(int32_t));
      memcpy(values + z0, c.values.data(), c.values.size() * sizeof(double));
    }
  };
  if (p->chunks.size() <= 1) {
    if (!p->chunks.empty()) scatter(0);
    return;
  }
  std::vector<std::thread> threads;
  for (size_t ci = 0; ci < p->chunks.size(); ++ci) threads.emplace_back(scatter, ci);
  for (auto& th : threads) th.join();
}

void afb_libsvm_free(void* h) { delete static_cast<Parsed*>(h); }

// Fast writer: one row per line, `label idx:val ...` with 1-based indices (MLUtils.scala:176-187).
int afb_libsvm_write(const char* path, int64_t rows, const double* labels, const int64_t* indptr,
                     const int32_t* indices, const double* values, char* err, int errlen) {
  FILE* f = fopen(path, "w");
  if (!f) { set_err(err, errlen, std::string("cannot open ") + path); return -1; }
  std::vector<char> buf(1 << 20);
  setvbuf(f, buf.data(), _IOFBF, buf.size());
  for (int64_t r = 0; r < rows; ++r) {
    double l = labels[r];
    if (l == std::floor(l) && std::fabs(l) < 1e15) fprintf(f, "%.1f", l); else fprintf(f, "%.17g", l);
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
      double v = values[k];
      if (v == std::floor(v) && std::fabs(v) < 1e15) fprintf(f, " %d:%.1f", indices[k] + 1, v);
      else fprintf(f, " %d:%.17g", indices[k] + 1, v);
    }
    fputc('\n', f);
  }
  fclose(f);
  return 0;
}

}  // extern "C"
