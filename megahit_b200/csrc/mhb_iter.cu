// mhb_iter.cu -- `iterate` on the device: host-level entry point mhb_iterate_host and the host mirror
// used by the CPU tests.  Kernels and building blocks: mhb_iter.cuh.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mhb_common.cuh"
#include "mhb_iter.cuh"

using namespace mhb;

int scan32(cudaStream_t st, const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *total_dev, uint64_t *bsum);

namespace {

struct IBuf {
  void *p = nullptr;
  ~IBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t b, const char *what) {
    if (p) cudaFree(p);
    p = nullptr;
    b = ((b ? b : 1) + 255) & ~(size_t)255;
    if (cudaMalloc(&p, b) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return mhb_set_error(MHB_ERR_NOMEM, "iterate: cudaMalloc of %zu bytes for %s failed", b, what);
    }
    return MHB_OK;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

#define CKR(call)        \
  do {                   \
    int rc_ = (call);    \
    if (rc_) return rc_; \
  } while (0)

unsigned igrid(uint64_t n, unsigned threads, unsigned per_sm = 16) {
  uint64_t g = (n + threads - 1) / threads;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

// ascending byte positions of a record of `words` words that hold the first `bits` bits (from the top)
uint32_t top_bytes(uint32_t words, uint32_t bits, uint8_t *out) {
  const uint32_t lo = (32 * words - bits) / 8;
  uint32_t n = 0;
  for (uint32_t b = lo; b < 4 * words; ++b) out[n++] = (uint8_t)b;
  return n;
}

// register words for the kernels: capacity classes instead of one instantiation per width
int cap_class(uint32_t words) { return words <= 2 ? 2 : words <= 4 ? 4 : words <= 8 ? 8 : 17; }
#define IT_FOR_WC(M) M(2) M(4) M(8) M(17)

struct ReadIndex {
  uint32_t fixed_len = 0;
  std::vector<uint64_t> rec_off, base_off;
  uint64_t n_bases = 0;
};
int index_reads(const uint32_t *bin, uint64_t bin_words, uint64_t n_reads, ReadIndex *ix) {
  if (n_reads == 0) return MHB_OK;
  if (bin_words == 0) return mhb_set_error(MHB_ERR_ARG, "empty .bin image for %llu reads", (unsigned long long)n_reads);
  const uint32_t L0 = bin[0];
  const uint64_t stride = 1 + div_ceil(L0, 16);
  bool fixed = L0 > 0 && bin_words == n_reads * stride;
  for (uint64_t r = 0; r < n_reads && fixed; ++r) fixed = bin[r * stride] == L0;
  if (fixed) {
    ix->fixed_len = L0;
    ix->n_bases = n_reads * (uint64_t)L0;
    return MHB_OK;
  }
  ix->rec_off.resize(n_reads + 1);
  ix->base_off.resize(n_reads + 1);
  uint64_t pos = 0, b = 0;
  for (uint64_t r = 0; r < n_reads; ++r) {
    if (pos >= bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated at read %llu", (unsigned long long)r);
    ix->rec_off[r] = pos;
    ix->base_off[r] = b;
    b += bin[pos];
    pos += 1 + div_ceil(bin[pos], 16);
  }
  if (pos > bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated");
  ix->rec_off[n_reads] = pos;
  ix->base_off[n_reads] = b;
  ix->n_bases = b;
  return MHB_OK;
}

}  // namespace

extern "C" int mhb_iterate_host(const mhb_iterate_args *a, mhb_iterate_result *res) {
  if (!a || !res) return mhb_set_error(MHB_ERR_ARG, "null argument");
  memset(res, 0, sizeof(*res));
  const uint32_t k = a->k, step = a->step, K1 = k + 1, KN = k + step + 1;
  // main_iterate.cpp:73-93: step even, 0 < step <= 28
  if (k < 9 || step == 0 || step > 28 || (step & 1)) return mhb_set_error(MHB_ERR_ARG, "iterate: invalid k / step");
  const uint32_t wk = div_ceil(K1, 16), w2 = words_per_edge(k + step), wn = div_ceil(KN, 16);
  if (wk + 2 > 17 || w2 > 17) return mhb_set_error(MHB_ERR_ARG, "iterate: k + step + 1 = %u is beyond the 17-word records of the device sort", KN);
  if (mhb_device_count() <= 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  res->words_per_edge = w2;
  cudaStream_t st = 0;
  struct Events {
    cudaEvent_t a, b;
    Events() {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
    }
    ~Events() {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } ev;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  cudaEventRecord(e0, st);
  const int WCc = cap_class(std::max(wn, wk));

  // ---- flank index (FeedBatchContigs) ----
  IBuf d_cw, d_co, d_cl, d_fl, d_fl2, d_ws, d_flag, d_off, d_bsum, d_cnt, d_tab, d_lut;
  CKR(d_cnt.alloc(64, "counters"));
  CK(cudaMemsetAsync(d_cnt.p, 0, 64, st));
  unsigned long long *cnt = d_cnt.as<unsigned long long>();
  const uint32_t frw = wk + 2;
  uint64_t n_tab = 0;
  if (a->n_contigs) {
    const uint64_t cw = a->contig_word_off[a->n_contigs];
    CKR(d_cw.alloc(cw * 4 + 64, "contigs"));
    CKR(d_co.alloc((a->n_contigs + 1) * 8, "contig offsets"));
    CKR(d_cl.alloc(a->n_contigs * 4, "contig lengths"));
    CK(cudaMemcpyAsync(d_cw.p, a->contig_words, cw * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_co.p, a->contig_word_off, (a->n_contigs + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cl.p, a->contig_len, a->n_contigs * 4, cudaMemcpyHostToDevice, st));
    const uint64_t cap = 2 * a->n_contigs;
    CKR(d_fl.alloc(cap * frw * 4 + 16, "flank records"));
    CKR(d_fl2.alloc(cap * frw * 4 + 16, "flank records (sort buffer)"));
    IterContigs cs{d_cw.as<u32>(), d_co.as<u64>(), d_cl.as<u32>(), a->n_contigs};
#define M(WW) \
  if (WCc == WW) k_iter_flanks<WW><<<igrid(cap, 256), 256, 0, st>>>(cs, k, step, wk, d_fl.as<u32>(), cnt);
    IT_FOR_WC(M)
#undef M
    CK_LAUNCH();
    unsigned long long nf = 0;
    CK(cudaMemcpyAsync(&nf, cnt, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nf) {
      // ascending on key, then on ~val: within a key the largest (ext_len, ext_seq) comes first and survives
      uint8_t bytes[72];
      uint32_t nb = 0;
      for (uint32_t b = 0; b < 8; ++b) bytes[nb++] = (uint8_t)b;  // the two ~val words
      uint8_t kb[72];
      const uint32_t nkb = top_bytes(wk, 2 * K1, kb);             // key bytes inside the key words ...
      for (uint32_t i = 0; i < nkb; ++i) bytes[nb++] = (uint8_t)(kb[i] + 8);  // ... sit above the 8 val bytes
      const size_t wsb = mhb_sort_workspace_bytes(nf, frw);
      CKR(d_ws.alloc(wsb, "sort workspace"));
      int in_b = 0;
      CKR(mhb_sort_records(st, d_fl.as<u32>(), d_fl2.as<u32>(), nf, frw, bytes, nb, nullptr, d_ws.p, wsb, &in_b));
      const u32 *sorted = in_b ? d_fl2.as<u32>() : d_fl.as<u32>();
      CKR(d_flag.alloc(nf * 4 + 16, "flags"));
      CKR(d_off.alloc(nf * 8 + 16, "offsets"));
      CKR(d_bsum.alloc((nf / 4096 + 4) * 8, "scan sums"));
      k_iter_heads<<<igrid(nf, 256), 256, 0, st>>>(sorted, nf, frw, wk, d_flag.as<u32>());
      CK_LAUNCH();
      CKR(scan32(st, d_flag.as<u32>(), nf, d_off.as<u64>(), (uint64_t *)(cnt + 2), d_bsum.as<u64>()));
      unsigned long long nu = 0;
      CK(cudaMemcpyAsync(&nu, cnt + 2, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      CKR(d_tab.alloc((size_t)nu * frw * 4 + 16, "flank table"));
      k_iter_compact<<<igrid(nf, 256), 256, 0, st>>>(sorted, nf, frw, d_flag.as<u32>(), d_off.as<u64>(), d_tab.as<u32>());
      CK_LAUNCH();
      n_tab = nu;
    }
  }
  res->n_flanks = n_tab;
  CKR(d_lut.alloc(65537 * 4, "flank prefix table"));
  CK(cudaMemsetAsync(d_lut.p, 0, 65537 * 4, st));
  if (n_tab) {
    k_iter_lut<<<(65537 + 255) / 256, 256, 0, st>>>(d_tab.as<u32>(), n_tab, frw, d_lut.as<u32>());
    CK_LAUNCH();
  }
  FlankTable tab{d_tab.as<u32>(), n_tab, wk, d_lut.as<u32>()};

  // ---- reads (FindNextKmersFromReads) ----
  ReadIndex ix;
  CKR(index_reads(a->bin, a->bin_words, a->n_reads, &ix));
  IBuf d_bin, d_ro, d_bo, d_exist, d_out, d_out2, d_uniq;
  uint64_t n_cand = 0, n_edges = 0;
  if (a->n_reads && n_tab) {
    CKR(d_bin.alloc(a->bin_words * 4 + 64, ".bin image"));
    CK(cudaMemcpyAsync(d_bin.p, a->bin, a->bin_words * 4, cudaMemcpyHostToDevice, st));
    IterReads rd{d_bin.as<u32>(), a->n_reads, ix.fixed_len, nullptr, nullptr};
    if (!ix.fixed_len) {
      CKR(d_ro.alloc(ix.rec_off.size() * 8, "record offsets"));
      CKR(d_bo.alloc(ix.base_off.size() * 8, "base offsets"));
      CK(cudaMemcpyAsync(d_ro.p, ix.rec_off.data(), ix.rec_off.size() * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(d_bo.p, ix.base_off.data(), ix.base_off.size() * 8, cudaMemcpyHostToDevice, st));
      rd.rec_off = d_ro.as<u64>();
      rd.base_off = d_bo.as<u64>();
    }
    const uint64_t bw = ix.n_bases / 32 + 2;
    CKR(d_exist.alloc(bw * 4, "position marks"));
    CK(cudaMemsetAsync(d_exist.p, 0, bw * 4, st));
    CK(cudaMemsetAsync(cnt + 4, 0, 16, st));
#define M(WW)                                                                                                         \
  if (WCc == WW) {                                                                                                    \
    k_iter_mark<WW><<<igrid(a->n_reads, 128, 32), 128, 0, st>>>(rd, k, step, tab, d_exist.as<u32>());                  \
    k_iter_emit<WW, false><<<igrid(a->n_reads, 128, 32), 128, 0, st>>>(rd, k, step, d_exist.as<u32>(), w2, nullptr,    \
                                                                       cnt + 4, 0);                                   \
  }
    IT_FOR_WC(M)
#undef M
    CK_LAUNCH();
    unsigned long long hc[2] = {0, 0};
    CK(cudaMemcpyAsync(hc, cnt + 4, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    n_cand = hc[0];
    res->n_aligned_reads = hc[1];
    if (n_cand) {
      CKR(d_out.alloc((size_t)n_cand * w2 * 4 + 16, "edges"));
      CKR(d_out2.alloc((size_t)n_cand * w2 * 4 + 16, "edges (sort buffer)"));
      CK(cudaMemsetAsync(cnt + 6, 0, 8, st));
#define M(WW)                                                                                                        \
  if (WCc == WW)                                                                                                     \
    k_iter_emit<WW, true><<<igrid(a->n_reads, 128, 32), 128, 0, st>>>(rd, k, step, d_exist.as<u32>(), w2, d_out.as<u32>(), \
                                                                      cnt + 6, n_cand);
      IT_FOR_WC(M)
#undef M
      CK_LAUNCH();
      // KmerCollector is a set (kmer_collector.h:37-48): sort + unique; the multiplicity bits are all zero
      uint8_t bytes[72];
      const uint32_t nb = top_bytes(w2, 2 * KN, bytes);
      const size_t wsb = mhb_sort_workspace_bytes(n_cand, w2);
      CKR(d_ws.alloc(wsb, "sort workspace"));
      int in_b = 0;
      CKR(mhb_sort_records_relaxed(st, d_out.as<u32>(), d_out2.as<u32>(), n_cand, w2, bytes, nb, nullptr, d_ws.p, wsb, &in_b));
      const u32 *sorted = in_b ? d_out2.as<u32>() : d_out.as<u32>();
      u32 *uniq = in_b ? d_out.as<u32>() : d_out2.as<u32>();
      CKR(d_flag.alloc(n_cand * 4 + 16, "flags"));
      CKR(d_off.alloc(n_cand * 8 + 16, "offsets"));
      CKR(d_bsum.alloc((n_cand / 4096 + 4) * 8, "scan sums"));
      k_iter_heads<<<igrid(n_cand, 256), 256, 0, st>>>(sorted, n_cand, w2, w2, d_flag.as<u32>());
      CK_LAUNCH();
      CKR(scan32(st, d_flag.as<u32>(), n_cand, d_off.as<u64>(), (uint64_t *)(cnt + 7), d_bsum.as<u64>()));
      unsigned long long nu = 0;
      CK(cudaMemcpyAsync(&nu, cnt + 7, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      k_iter_compact<<<igrid(n_cand, 256), 256, 0, st>>>(sorted, n_cand, w2, d_flag.as<u32>(), d_off.as<u64>(), uniq);
      CK_LAUNCH();
      n_edges = nu;
      res->edges = (uint32_t *)malloc(std::max<size_t>(1, (size_t)n_edges * w2 * 4));
      if (!res->edges) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
      CK(cudaMemcpyAsync(res->edges, uniq, (size_t)n_edges * w2 * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  if (!res->edges) res->edges = (uint32_t *)malloc(4);
  res->n_candidates = n_cand;
  res->n_edges = n_edges;
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  res->t_total_ms = ms;
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// Host mirror for the CPU tests: the same __host__ __device__ building blocks (flank records, flank search, read
// marking, edge emission) driven serially; std::sort stands in for the device radix sort.  Not a compute path of the
// library (nothing calls it but tests/test_iter_cpu.py).
// ------------------------------------------------------------------------------------------------
extern "C" int mhb_selftest_iterate(const mhb_iterate_args *a, mhb_iterate_result *res) {
  if (!a || !res) return mhb_set_error(MHB_ERR_ARG, "null argument");
  memset(res, 0, sizeof(*res));
  const uint32_t k = a->k, step = a->step, K1 = k + 1, KN = k + step + 1;
  if (k < 9 || step == 0 || step > 28 || (step & 1)) return mhb_set_error(MHB_ERR_ARG, "iterate: invalid k / step");
  const uint32_t wk = div_ceil(K1, 16), w2 = words_per_edge(k + step), wn = div_ceil(KN, 16), frw = wk + 2;
  if (frw > 17 || w2 > 17) return mhb_set_error(MHB_ERR_ARG, "iterate: record too wide");
  res->words_per_edge = w2;
  const int WCc = cap_class(std::max(wn, wk));
  std::vector<std::vector<u32>> fl;
  for (uint64_t c = 0; c < a->n_contigs; ++c)
    for (u32 strand = 0; strand < 2; ++strand) {
      u32 rec[20];
      bool ok = false;
#define M(WW) \
  if (WCc == WW) ok = iter_flank_record<WW>(a->contig_words + a->contig_word_off[c], a->contig_len[c], k, step, strand, wk, rec);
      IT_FOR_WC(M)
#undef M
      if (ok) fl.emplace_back(rec, rec + frw);
    }
  std::sort(fl.begin(), fl.end());  // key words, then ~val: the largest extension first within a key
  std::vector<u32> tab;
  uint64_t nt = 0;
  for (size_t i = 0; i < fl.size(); ++i)
    if (i == 0 || !std::equal(fl[i].begin(), fl[i].begin() + wk, fl[i - 1].begin())) {
      tab.insert(tab.end(), fl[i].begin(), fl[i].end());
      ++nt;
    }
  std::vector<u32> lut(65537);
  for (u32 p = 0; p <= 65536; ++p) {
    uint64_t lo = 0, hi = nt;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if ((tab[mid * frw] >> 16) < p) lo = mid + 1; else hi = mid;
    }
    lut[p] = (u32)lo;
  }
  res->n_flanks = nt;
  FlankTable t{tab.data(), nt, wk, lut.data()};
  ReadIndex ix;
  CKR(index_reads(a->bin, a->bin_words, a->n_reads, &ix));
  IterReads rd{a->bin, a->n_reads, ix.fixed_len, ix.rec_off.data(), ix.base_off.data()};
  std::vector<u32> exist(ix.n_bases / 32 + 2, 0u);
  std::vector<std::vector<u32>> out;
  std::vector<u32> tmp;
  for (uint64_t r = 0; r < a->n_reads && nt; ++r) {
#define M(WW)                                                                          \
  if (WCc == WW) {                                                                     \
    iter_mark_read<WW>(rd, r, k, step, t, exist.data());                               \
    const u32 n = iter_emit_read<WW>(rd, r, k, step, exist.data(), w2, nullptr);       \
    if (n) {                                                                           \
      tmp.assign((size_t)n * w2, 0u);                                                  \
      iter_emit_read<WW>(rd, r, k, step, exist.data(), w2, tmp.data());                \
      for (u32 i = 0; i < n; ++i) out.emplace_back(tmp.begin() + (size_t)i * w2, tmp.begin() + (size_t)(i + 1) * w2); \
      ++res->n_aligned_reads;                                                          \
    }                                                                                  \
  }
    IT_FOR_WC(M)
#undef M
  }
  res->n_candidates = out.size();
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
  res->n_edges = out.size();
  res->edges = (uint32_t *)malloc(std::max<size_t>(4, out.size() * w2 * 4));
  for (size_t i = 0; i < out.size(); ++i) memcpy(res->edges + i * w2, out[i].data(), w2 * 4);
  return MHB_OK;
}
