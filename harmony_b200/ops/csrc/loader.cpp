// Native data-loader kernels of the host runtime: multi-threaded parsers for the sparse text formats of the
// apps and the counting-sort "group the non-zeros by feature" step that turns a mini-batch (CSR) into the
// feature-major layout the push kernels consume.
//
// Reference behaviour (formats only): MLRETDataParser / LassoETParser / GBTETDataParser (`y idx:val ...`),
// NMFETDataParser.java (`row: col,val ...`, 1-based, non-negative values), LDAETDataParser (`w w w ...`),
// dataloader HdfsSplitFetcher line-record semantics (a line belongs to the split holding its first byte --
// handled by the Python caller, this file parses one in-memory byte range).
#include <torch/extension.h>

#include <charconv>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

namespace hb {
namespace {

struct Chunk {
  std::vector<double> head;      // label / row id per record
  std::vector<int64_t> len;      // entries per record
  std::vector<int32_t> idx;
  std::vector<float> val;
  std::string error;
};

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
  return p;
}

inline const char* parse_double(const char* p, const char* e, double& out) {
  auto r = std::from_chars(p, e, out);
  if (r.ec != std::errc()) {
    if (p < e && *p == '+') return parse_double(p + 1, e, out);
    return nullptr;
  }
  return r.ptr;
}

inline const char* parse_int(const char* p, const char* e, long long& out) {
  if (p < e && *p == '+') ++p;
  auto r = std::from_chars(p, e, out);
  return r.ec == std::errc() ? r.ptr : nullptr;
}

// mode 0: `head idx:val idx:val ...`      (MLR / Lasso / GBT; head is a real number)
// mode 1: `head: idx,val idx,val ...`     (NMF; indices are 1-based in the file, returned 0-based)
// mode 2: `idx idx idx ...`               (LDA documents; head = 0)
void parse_range(const char* b, const char* e, int mode, Chunk& out) {
  const char pair_sep = mode == 1 ? ',' : ':';
  const char* p = b;
  while (p < e) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', e - p));
    if (!eol) eol = e;
    const char* q = skip_ws(p, eol);
    if (q < eol && *q != '#') {
      double head = 0.0;
      if (mode != 2) {
        q = parse_double(q, eol, head);
        if (!q) { out.error = "bad record head: " + std::string(p, std::min<size_t>(eol - p, 60)); return; }
        if (mode == 1) {
          q = skip_ws(q, eol);
          if (q < eol && *q == ':') ++q;
          if (head <= 0) { out.error = "row index must be positive"; return; }
        }
      }
      int64_t n = 0;
      for (;;) {
        q = skip_ws(q, eol);
        if (q >= eol) break;
        long long i;
        const char* r = parse_int(q, eol, i);
        if (!r) { out.error = "bad index in: " + std::string(p, std::min<size_t>(eol - p, 60)); return; }
        double v = 1.0;
        if (mode != 2) {
          if (r >= eol || *r != pair_sep) {
            out.error = "missing value in: " + std::string(p, std::min<size_t>(eol - p, 60)); return;
          }
          r = parse_double(r + 1, eol, v);
          if (!r) { out.error = "bad value in: " + std::string(p, std::min<size_t>(eol - p, 60)); return; }
          if (mode == 1) {
            if (i <= 0) { out.error = "column index must be positive"; return; }
            if (v < 0) { out.error = "matrix values must be non-negative"; return; }
            i -= 1;
          }
        }
        out.idx.push_back(static_cast<int32_t>(i));
        if (mode != 2) out.val.push_back(static_cast<float>(v));
        ++n;
        q = r;
      }
      out.head.push_back(head);
      out.len.push_back(n);
    }
    p = eol + 1;
  }
}

}  // namespace

// -> (head f64[n], ptr i64[n+1], idx i32[nnz], val f32[nnz] (empty for mode 2))
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> parse_sparse_text(const std::string& text, int64_t mode,
                                                                              int64_t threads) {
  TORCH_CHECK(mode >= 0 && mode <= 2, "unknown text format");
  const char* b = text.data();
  const char* e = b + text.size();
  int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, 64)));
  if (text.size() < (1u << 16)) T = 1;
  std::vector<const char*> cut(T + 1, e);
  cut[0] = b;
  for (int t = 1; t < T; ++t) {
    const char* p = b + text.size() * t / T;
    if (p < cut[t - 1]) p = cut[t - 1];
    const char* nl = static_cast<const char*>(memchr(p, '\n', e - p));
    cut[t] = nl ? nl + 1 : e;
  }
  std::vector<Chunk> chunks(T);
  {
    pybind11::gil_scoped_release nogil;
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t)
      pool.emplace_back([&, t] { parse_range(cut[t], cut[t + 1], (int)mode, chunks[t]); });
    parse_range(cut[0], cut[1], (int)mode, chunks[0]);
    for (auto& th : pool) th.join();
  }
  int64_t n = 0, nnz = 0;
  for (auto& c : chunks) {
    if (!c.error.empty()) throw std::invalid_argument(c.error);
    n += (int64_t)c.head.size();
    nnz += (int64_t)c.idx.size();
  }
  at::Tensor head = at::empty({n}, at::kDouble), ptr = at::empty({n + 1}, at::kLong);
  at::Tensor idx = at::empty({nnz}, at::kInt), val = at::empty({mode == 2 ? 0 : nnz}, at::kFloat);
  double* hp = head.data_ptr<double>();
  int64_t* pp = ptr.data_ptr<int64_t>();
  int32_t* ip = idx.data_ptr<int32_t>();
  float* vp = val.data_ptr<float>();
  int64_t r = 0, z = 0;
  pp[0] = 0;
  for (auto& c : chunks) {
    if (!c.head.empty()) memcpy(hp + r, c.head.data(), c.head.size() * sizeof(double));
    for (size_t i = 0; i < c.len.size(); ++i) pp[r + i + 1] = pp[r + i] + c.len[i];
    if (!c.idx.empty()) memcpy(ip + z, c.idx.data(), c.idx.size() * sizeof(int32_t));
    if (mode != 2 && !c.val.empty()) memcpy(vp + z, c.val.data(), c.val.size() * sizeof(float));
    r += (int64_t)c.head.size();
    z += (int64_t)c.idx.size();
  }
  return {head, ptr, idx, val};
}

// CSR mini-batch -> feature-major grouping (stable: instances stay in ascending order inside a feature).
// -> (uniq_feat i32[U], feat_ptr i32[U+1], samp i32[nnz], xval f32[nnz])
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> group_by_feature(at::Tensor csr_ptr, at::Tensor csr_col,
                                                                             at::Tensor csr_val) {
  TORCH_CHECK(!csr_ptr.is_cuda() && csr_ptr.scalar_type() == at::kInt && csr_col.scalar_type() == at::kInt &&
              csr_val.scalar_type() == at::kFloat, "group_by_feature expects CPU int32 / int32 / float32");
  csr_ptr = csr_ptr.contiguous(); csr_col = csr_col.contiguous(); csr_val = csr_val.contiguous();
  const int64_t B = csr_ptr.numel() - 1, nnz = csr_col.numel();
  TORCH_CHECK(B >= 0 && csr_val.numel() == nnz);
  const int32_t* ptr = csr_ptr.data_ptr<int32_t>();
  const int32_t* col = csr_col.data_ptr<int32_t>();
  const float* val = csr_val.data_ptr<float>();
  TORCH_CHECK(B == 0 || ptr[B] == nnz, "csr_ptr does not cover csr_col");
  at::Tensor samp = at::empty({nnz}, at::kInt), xval = at::empty({nnz}, at::kFloat);
  if (nnz == 0)
    return {at::empty({0}, at::kInt), at::zeros({1}, at::kInt), samp, xval};
  int32_t maxc = 0;
  for (int64_t i = 0; i < nnz; ++i) {
    TORCH_CHECK(col[i] >= 0, "negative feature index");
    maxc = std::max(maxc, col[i]);
  }
  // two-pass LSD radix sort on (feature) keeps the work O(nnz) even for a 1M+ feature space with few non-zeros
  const int bits = 32 - __builtin_clz((unsigned)std::max(maxc, 1));
  const int lo_bits = std::min(bits, 16), hi_bits = bits - lo_bits;
  std::vector<int32_t> row_of(nnz), perm(nnz), tmp(nnz);
  for (int64_t b = 0; b < B; ++b)
    for (int32_t k = ptr[b]; k < ptr[b + 1]; ++k) row_of[k] = (int32_t)b;
  auto pass = [&](const int32_t* in, int32_t* out, int shift, int nbits, bool identity_in) {
    const size_t buckets = (size_t)1 << nbits;
    std::vector<int64_t> cnt(buckets + 1, 0);
    const int32_t mask = (int32_t)(buckets - 1);
    for (int64_t i = 0; i < nnz; ++i) cnt[((col[identity_in ? i : in[i]] >> shift) & mask) + 1]++;
    for (size_t k = 0; k < buckets; ++k) cnt[k + 1] += cnt[k];
    for (int64_t i = 0; i < nnz; ++i) {
      const int32_t src = identity_in ? (int32_t)i : in[i];
      out[cnt[(col[src] >> shift) & mask]++] = src;
    }
  };
  pass(nullptr, hi_bits > 0 ? tmp.data() : perm.data(), 0, lo_bits, true);
  if (hi_bits > 0) pass(tmp.data(), perm.data(), lo_bits, hi_bits, false);
  int32_t* sp = samp.data_ptr<int32_t>();
  float* xp = xval.data_ptr<float>();
  std::vector<int32_t> uniq, starts;
  int32_t prev = -1;
  for (int64_t i = 0; i < nnz; ++i) {
    const int32_t src = perm[i], c = col[src];
    if (c != prev) { uniq.push_back(c); starts.push_back((int32_t)i); prev = c; }
    sp[i] = row_of[src];
    xp[i] = val[src];
  }
  starts.push_back((int32_t)nnz);
  at::Tensor u = at::empty({(int64_t)uniq.size()}, at::kInt), fp = at::empty({(int64_t)starts.size()}, at::kInt);
  memcpy(u.data_ptr<int32_t>(), uniq.data(), uniq.size() * sizeof(int32_t));
  memcpy(fp.data_ptr<int32_t>(), starts.data(), starts.size() * sizeof(int32_t));
  return {u, fp, samp, xval};
}

}  // namespace hb
