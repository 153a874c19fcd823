// mhb_r2s.cu -- `read2sdbg` on the device: host-level entry point mhb_read2sdbg_host and the
// self-test hooks of its building blocks.  Kernels: mhb_r2s.cuh.  Reference: main_sdbg_build.cpp:88-156,
// sorting/read_to_sdbg_s1.cpp, sorting/read_to_sdbg_s2.cpp.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mhb_common.cuh"
#include "mhb_r2s.cuh"

using namespace mhb;

// three-phase scan of mhb_device.cu
int scan32(cudaStream_t st, const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *total_dev, uint64_t *bsum);
#define MHB_FOR_RW(M) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16) M(17)

namespace {

struct DevBuf {  // cudaMalloc'ed scratch released on scope exit
  void *p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int alloc(size_t b, const char *what) {
    release();
    b = (b + 255) & ~(size_t)255;
    if (b == 0) b = 256;
    cudaError_t e = cudaMalloc(&p, b);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return mhb_set_error(MHB_ERR_NOMEM, "read2sdbg: cudaMalloc of %zu bytes for %s failed: %s", b, what,
                           cudaGetErrorString(e));
    }
    bytes = b;
    return MHB_OK;
  }
  // keep the allocation when it is already large enough (the two stages share their big buffers)
  int ensure(size_t b, const char *what) { return bytes >= b ? MHB_OK : alloc(b, what); }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

// the record / item ping-pong buffers and the sort workspace, shared by stage 1 and stage 2 (cudaMalloc + cudaFree of
// 20-GB buffers between the stages cost ~0.2 s at 10 M reads)
struct BigBufs {
  DevBuf a, b, ws;
};

#define CKR(call)        \
  do {                   \
    int rc_ = (call);    \
    if (rc_) return rc_; \
  } while (0)

// phase marks on the stream: deltas include host-side gaps (allocations, synchronisations) between the marks
struct PhaseTrace {
  std::vector<std::pair<const char *, cudaEvent_t>> ev;
  cudaStream_t st = 0;
  void mark(const char *name) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.emplace_back(name, e);
  }
  // ms spent in the phases whose name starts with `prefix` (all when empty); call after a stream synchronise
  double sum(const char *prefix) const {
    double t = 0;
    for (size_t i = 1; i < ev.size(); ++i)
      if (strncmp(ev[i].first, prefix, strlen(prefix)) == 0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
        t += ms;
      }
    return t;
  }
  void report() const {
    if (!getenv("MHB_R2S_TRACE")) return;
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      fprintf(stderr, "[r2s] %-28s %9.3f ms\n", ev[i].first, ms);
    }
  }
  ~PhaseTrace() {
    for (auto &e : ev) cudaEventDestroy(e.second);
  }
};

unsigned grid_for(uint64_t n, unsigned threads, unsigned per_sm = 16) {
  uint64_t g = (n + threads - 1) / threads;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// host index of the `.bin` image: package geometry of every read
struct PkgIndex {
  uint32_t fixed_len = 0, fixed_words = 0, max_len = 0;
  uint64_t n_reads = 0, n_words = 0, n_bases = 0, n_s1 = 0, n_edges = 0;
  std::vector<uint64_t> rec_off, word_off, base_off, s1_off, edge_off;
  std::vector<uint32_t> len;
};

int index_pkg(const uint32_t *bin, uint64_t bin_words, uint64_t n_reads, uint32_t k, PkgIndex *ix) {
  ix->n_reads = n_reads;
  if (n_reads == 0) return MHB_OK;
  if (bin_words == 0) return mhb_set_error(MHB_ERR_ARG, "empty .bin image for %llu reads", (unsigned long long)n_reads);
  const uint32_t L0 = bin[0];
  const uint64_t stride = 1 + div_ceil(L0, 16);
  bool fixed = L0 > 0 && bin_words == n_reads * stride;
  if (fixed) {
    int bad = 0;
#pragma omp parallel for reduction(| : bad) schedule(static)
    for (long long r = 0; r < (long long)n_reads; ++r) bad |= bin[(uint64_t)r * stride] != L0;
    fixed = !bad;
  }
  if (fixed) {
    ix->fixed_len = ix->max_len = L0;
    ix->fixed_words = div_ceil(L0, 16);
    ix->n_words = n_reads * ix->fixed_words;
    ix->n_bases = n_reads * (uint64_t)L0;
    if (L0 >= k + 1) {
      ix->n_s1 = n_reads * (uint64_t)(L0 - k + 4);
      ix->n_edges = n_reads * (uint64_t)(L0 - k);
    }
    return MHB_OK;
  }
  ix->rec_off.resize(n_reads + 1);
  ix->word_off.resize(n_reads + 1);
  ix->base_off.resize(n_reads + 1);
  ix->s1_off.resize(n_reads + 1);
  ix->edge_off.resize(n_reads + 1);
  ix->len.resize(n_reads);
  uint64_t pos = 0, w = 0, b = 0, s1 = 0, e = 0;
  for (uint64_t r = 0; r < n_reads; ++r) {
    if (pos >= bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated at read %llu", (unsigned long long)r);
    const uint32_t L = bin[pos];
    const uint32_t eff = L == 0 ? 1 : L;  // sequence_package.h:276-281
    ix->rec_off[r] = pos;
    ix->word_off[r] = w;
    ix->base_off[r] = b;
    ix->s1_off[r] = s1;
    ix->edge_off[r] = e;
    ix->len[r] = eff;
    ix->max_len = std::max(ix->max_len, eff);
    pos += 1 + div_ceil(L, 16);
    w += div_ceil(eff, 16);
    b += eff;
    if (eff >= k + 1) {
      s1 += eff - k + 4;
      e += eff - k;
    }
  }
  if (pos > bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated");
  ix->rec_off[n_reads] = pos;
  ix->word_off[n_reads] = w;
  ix->base_off[n_reads] = b;
  ix->s1_off[n_reads] = s1;
  ix->edge_off[n_reads] = e;
  ix->n_words = w;
  ix->n_bases = b;
  ix->n_s1 = s1;
  ix->n_edges = e;
  return MHB_OK;
}

template <class T>
int upload(DevBuf &d, const std::vector<T> &v, const char *what) {
  CKR(d.alloc(v.size() * sizeof(T), what));
  if (!v.empty()) CK(cudaMemcpy(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MHB_OK;
}

// ---- stage 1 on the device: is_solid bits, mercy planes, multiplicity histogram ----
int run_stage1(cudaStream_t st, const PkgView &pv, const PkgIndex &ix, uint32_t k, int32_t m, bool need_mercy, const S1Out &out,
               unsigned long long *d_mul_hist, PhaseTrace &tr, BigBufs &big, size_t min_rec_bytes, size_t min_ws_bytes) {
  const uint32_t NW = r2s_s1_key_words(k), RW = NW + 2;
  const uint64_t n = ix.n_s1;
  if (n == 0) return MHB_OK;
  if (RW > 17)
    return mhb_set_error(MHB_ERR_ARG, "read2sdbg: stage 1 supports k <= 237 (record of %u words > 17)", RW);
  if (n >= (1ull << 40)) return mhb_set_error(MHB_ERR_ARG, "read2sdbg: too many stage-1 records for one round");
  DevBuf bstart, segs0, segs1, counter, bnd;
  DevBuf &a = big.a, &b = big.b, &ws = big.ws;
  const size_t rec_bytes = (size_t)n * RW * 4 + 16, ws_bytes = mhb_sort_workspace_bytes(n, RW);
  CKR(a.ensure(std::max(rec_bytes, min_rec_bytes), "records"));
  CKR(b.ensure(std::max(rec_bytes, min_rec_bytes), "records (sort buffer)"));
  CKR(ws.ensure(std::max(ws_bytes, min_ws_bytes), "sort workspace"));
  CKR(bstart.alloc((MHB_NUM_BUCKETS + 1) * 8, "bucket bounds"));
  const uint64_t seg_cap = n / (kKmInsertThreshold + 1) + 2;
  CKR(segs0.alloc(seg_cap * sizeof(KmSeg), "kmsort ranges"));
  CKR(segs1.alloc(seg_cap * sizeof(KmSeg), "kmsort ranges"));
  CKR(counter.alloc(8, "counter"));
  CKR(bnd.alloc((n / 32 + 2) * 4, "range marks"));
  CK(cudaMemsetAsync(bnd.p, 0, (n / 32 + 2) * 4, st));
#define M(WW)                                                                                         \
  if (NW == WW) k_r2s_s1_extract<WW><<<grid_for(n, 256), 256, 0, st>>>(pv, k, a.as<u32>(), n);
  MHB_FOR_W(M)
#undef M
  CK_LAUNCH();
  tr.mark("s1.extract");
  // the reference's bucket input order: records of one 16-bit bucket in global read order = a STABLE sort on the two
  // leading key bytes (base_engine.cpp:323-348 fills every bucket thread by thread, i.e. in read order)
  const uint8_t bytes[2] = {(uint8_t)(4 * RW - 2), (uint8_t)(4 * RW - 1)};
  int in_b = 0;
  CKR(mhb_sort_records(st, a.as<u32>(), b.as<u32>(), n, RW, bytes, 2, nullptr, ws.p, ws_bytes, &in_b));
  u32 *recs = in_b ? b.as<u32>() : a.as<u32>();
  tr.mark("s1.partition");
  k_r2s_bucket_bounds<<<(MHB_NUM_BUCKETS + 1 + 255) / 256, 256, 0, st>>>(recs, n, RW, bstart.as<u64>());
  CK_LAUNCH();
  // kmsort (kmsort.h:103-117 entry, :43-101 per range), level by level
  int kb = 4 * (int)NW - 2 - 1;
  KmSeg *cur = segs0.as<KmSeg>(), *nxt = segs1.as<KmSeg>();
  unsigned long long *d_cnt = counter.as<unsigned long long>();
  CK(cudaMemsetAsync(d_cnt, 0, 8, st));
  const bool km_global = getenv("MHB_R2S_KMSORT_GLOBAL") != nullptr;  // the in-place form of every level (A/B, tests)
  static char level_names[72][24];
  int level = 0;
  DevBuf todo, src16;
  u32 *d_todo = nullptr;
  if (!km_global) {
    // level 0 on shared-memory tags: bucket sizes decide the tag capacity of a CTA
    std::vector<uint64_t> h_b(MHB_NUM_BUCKETS + 1);
    CK(cudaMemcpyAsync(h_b.data(), bstart.p, h_b.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t max_bucket = 0;
    for (int i = 0; i < MHB_NUM_BUCKETS; ++i) max_bucket = std::max(max_bucket, h_b[i + 1] - h_b[i]);
    uint32_t cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(max_bucket, 1024), 65535);
    cap = (cap + 1023) & ~1023u;
    if (const char *e = getenv("MHB_R2S_KM_CAP")) cap = std::max(1024u, (uint32_t)atoi(e) & ~1023u);  // tests: force the fall-back
    CKR(todo.alloc((n / 32 + 2) * 4, "unsorted-range marks"));
    CK(cudaMemsetAsync(todo.p, 0, (n / 32 + 2) * 4, st));
    CKR(src16.alloc((size_t)n * 2 + 64, "kmsort source indices"));
    d_todo = todo.as<u32>();
    u32 *other = in_b ? a.as<u32>() : b.as<u32>();
#define M(WW)                                                                                                           \
  if (RW == WW) {                                                                                                       \
    CK(cudaFuncSetAttribute(k_r2s_km_bucket<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));               \
    k_r2s_km_bucket<WW><<<MHB_NUM_BUCKETS, 128, cap, st>>>(recs, other, bstart.as<u64>(), NW, kb, cap,                   \
                                                           src16.as<uint16_t>(), bnd.as<u32>(), d_todo, nxt, d_cnt, seg_cap); \
  }
    MHB_FOR_RW(M)
#undef M
    CK_LAUNCH();
    recs = other;
  } else {
#define M(WW)                                                                                                      \
  if (RW == WW)                                                                                                    \
    k_r2s_kmsort_level<WW><<<MHB_NUM_BUCKETS / 128, 128, 0, st>>>(recs, NW, kb, nullptr, bstart.as<u64>(),          \
                                                                  MHB_NUM_BUCKETS, nxt, d_cnt, seg_cap, bnd.as<u32>());
    MHB_FOR_RW(M)
#undef M
    CK_LAUNCH();
  }
  for (;;) {
    snprintf(level_names[level], sizeof(level_names[level]), "s1.kmsort.L%d", level);
    tr.mark(level_names[level]);
    if (level < 71) ++level;
    unsigned long long n_next = 0;
    CK(cudaMemcpyAsync(&n_next, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_next == 0 || kb == 0) break;
    if (n_next > seg_cap) return mhb_set_error(MHB_ERR_CUDA, "read2sdbg: internal: kmsort range list overflow");
    std::swap(cur, nxt);
    --kb;
    CK(cudaMemsetAsync(d_cnt, 0, 8, st));
    if (!km_global) {
#define M(WW)                                                                                                            \
  if (RW == WW) {                                                                                                        \
    const size_t smem = (size_t)kKmWarps * km_warp_smem<WW>();                                                           \
    static bool attr = false;                                                                                            \
    if (!attr) {                                                                                                         \
      CK(cudaFuncSetAttribute(k_r2s_km_warp<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
      attr = true;                                                                                                       \
    }                                                                                                                    \
    uint64_t g = (n_next + kKmWarps - 1) / kKmWarps;                                                                     \
    if (g > (uint64_t)sm_count() * 8) g = (uint64_t)sm_count() * 8;                                                      \
    k_r2s_km_warp<WW><<<(unsigned)g, kKmWarps * 32, smem, st>>>(recs, NW, kb, cur, n_next, nxt, d_cnt, seg_cap,           \
                                                              bnd.as<u32>(), d_todo);                                    \
  }
      MHB_FOR_RW(M)
#undef M
    } else {
#define M(WW)                                                                                                         \
  if (RW == WW)                                                                                                       \
    k_r2s_kmsort_level<WW><<<(unsigned)((n_next + 127) / 128), 128, 0, st>>>(recs, NW, kb, cur, nullptr, n_next, nxt, \
                                                                            d_cnt, seg_cap, bnd.as<u32>());
      MHB_FOR_RW(M)
#undef M
    }
    CK_LAUNCH();
  }
#define M(WW) \
  if (RW == WW) k_r2s_kmsort_finish<WW><<<grid_for(n, 256, 32), 256, 0, st>>>(recs, n, NW, bnd.as<u32>(), d_todo);
  MHB_FOR_RW(M)
#undef M
  CK_LAUNCH();
  tr.mark("s1.kmsort.finish");
#define M(WW)                                                                                                          \
  if (RW == WW)                                                                                                        \
    k_r2s_s1_post<WW><<<grid_for(n, 256, 32), 256, 0, st>>>(recs, n, NW, k, m, pv, out, need_mercy ? 1 : 0, d_mul_hist);
  MHB_FOR_RW(M)
#undef M
  CK_LAUNCH();
  tr.mark("s1.post");
  CK(cudaStreamSynchronize(st));
  return MHB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// main_read2sdbg (main_sdbg_build.cpp:88-156) from a host `.bin` image: SdBG item stream + bucket table out.
// ------------------------------------------------------------------------------------------------
extern "C" int mhb_read2sdbg_host(const mhb_build_args *args, mhb_build_result *res) {
  if (!args || !res) return mhb_set_error(MHB_ERR_ARG, "null argument");
  memset(res, 0, sizeof(*res));
  const uint32_t k = args->k;
  const int32_t m = args->m;
  if (k < 9 || k > MHB_MAX_K || m < 1) return mhb_set_error(MHB_ERR_ARG, "read2sdbg: need 9 <= k <= 255 and m >= 1");
  if (mhb_device_count() <= 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  cudaStream_t st = 0;
  PhaseTrace tr;
  tr.mark("start");
  PkgIndex ix;
  CKR(index_pkg(args->bin, args->bin_words, args->n_reads, k, &ix));
  const uint32_t W = s2s_record_words(k), WPT = words_per_tip_label(k);
  res->words_per_tip_label = WPT;
  res->n_edge_records = ix.n_edges;
  res->bucket_table = (uint64_t *)calloc((size_t)MHB_NUM_BUCKETS * 4, 8);
  res->counting = (int64_t *)calloc(65536, 8);
  struct ResGuard {  // a failing call hands nothing back
    mhb_build_result *r;
    const uint8_t *caller_buf;
    bool ok = false;
    ~ResGuard() {
      if (ok) return;
      free(r->bucket_table);
      free(r->counting);
      if (r->bytes != caller_buf) free(r->bytes);
      r->bucket_table = nullptr;
      r->counting = nullptr;
      r->bytes = nullptr;
    }
  } guard{res, args->sdbg_out};
  if (!res->bucket_table || !res->counting) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");

  // ---- the package: reversed reads on the device ----
  DevBuf d_pkg, d_word_off, d_len, d_base_off, d_s1_off, d_edge_off;
  PkgView pv;
  memset(&pv, 0, sizeof(pv));
  pv.n_reads = ix.n_reads;
  pv.fixed_len = ix.fixed_len;
  pv.fixed_words = ix.fixed_words;
  CKR(d_pkg.alloc((size_t)ix.n_words * 4 + 64, "package"));
  if (ix.n_reads) {
    DevBuf d_bin, d_rec_off;
    CKR(d_bin.alloc((size_t)args->bin_words * 4 + 64, ".bin image"));
    CK(cudaMemcpyAsync(d_bin.p, args->bin, (size_t)args->bin_words * 4, cudaMemcpyHostToDevice, st));
    if (!ix.fixed_len) {
      CKR(upload(d_rec_off, ix.rec_off, "record offsets"));
      CKR(upload(d_word_off, ix.word_off, "word offsets"));
      CKR(upload(d_len, ix.len, "lengths"));
      CKR(upload(d_base_off, ix.base_off, "base offsets"));
      CKR(upload(d_s1_off, ix.s1_off, "stage-1 offsets"));
      CKR(upload(d_edge_off, ix.edge_off, "edge offsets"));
      pv.word_off = d_word_off.as<u64>();
      pv.len = d_len.as<u32>();
      pv.base_off = d_base_off.as<u64>();
      pv.s1_off = d_s1_off.as<u64>();
      pv.edge_off = d_edge_off.as<u64>();
    }
    if (ix.n_words) {
      k_r2s_reverse<<<grid_for(ix.n_words, 256), 256, 0, st>>>(d_bin.as<u32>(), ix.n_reads, ix.fixed_len, d_rec_off.as<u64>(), pv,
                                                                d_pkg.as<u32>(), ix.n_words);
      CK_LAUNCH();
    }
    CK(cudaStreamSynchronize(st));  // d_bin / d_rec_off go out of scope
  }
  pv.words = d_pkg.as<u32>();
  tr.mark("h2d.upload+reverse");

  // ---- stage 1 (only when the threshold can reject anything, main_sdbg_build.cpp:141-147) ----
  const uint64_t bit_words = ix.n_bases / 32 + 2;
  DevBuf d_solid, d_planes, d_hist, d_cnt;
  CKR(d_solid.alloc(bit_words * 4, "solid marker"));
  CK(cudaMemsetAsync(d_solid.p, 0, bit_words * 4, st));
  CKR(d_hist.alloc(65536 * 8, "multiplicity histogram"));
  CK(cudaMemsetAsync(d_hist.p, 0, 65536 * 8, st));
  CKR(d_cnt.alloc(64, "counters"));
  CK(cudaMemsetAsync(d_cnt.p, 0, 64, st));
  unsigned long long *d_counter = d_cnt.as<unsigned long long>();
  const bool mercy = args->need_mercy && m > 1;
  BigBufs big;
  if (m > 1 && ix.n_s1) {
    S1Out so;
    memset(&so, 0, sizeof(so));
    so.is_solid = d_solid.as<u32>();
    if (mercy) {
      CKR(d_planes.alloc(bit_words * 4 * 4, "mercy candidate planes"));
      CK(cudaMemsetAsync(d_planes.p, 0, bit_words * 4 * 4, st));
      so.no_in = d_planes.as<u32>();
      so.no_out = so.no_in + bit_words;
      so.any = so.no_out + bit_words;
    }
    // size the shared buffers for stage 2 as well: ~1.6 items per edge position on a 30x library, 2.2 to be safe
    // (stage 2 reallocates when its count launch says more)
    const uint64_t est_items = (uint64_t)(2.2 * (double)ix.n_edges) + 1024;
    CKR(run_stage1(st, pv, ix, k, m, mercy, so, d_hist.as<unsigned long long>(), tr, big, (size_t)est_items * W * 4 + 16,
                   mhb_sort_workspace_bytes(est_items, W)));
    if (mercy) {  // Read2SdbgS2::Initialize, read_to_sdbg_s2.cpp:117-263
      u32 *d_mercy = so.any + bit_words;
      k_r2s_mercy<<<grid_for(ix.n_reads, 256), 256, 0, st>>>(pv, k, so, d_mercy, d_counter + 1);
      CK_LAUNCH();
      k_r2s_or_words<<<grid_for(bit_words, 256), 256, 0, st>>>(d_solid.as<u32>(), d_mercy, bit_words);
      CK_LAUNCH();
      unsigned long long nm = 0;
      CK(cudaMemcpyAsync(&nm, d_counter + 1, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      res->n_mercy = nm;
      tr.mark("s1.mercy");
    }
    CK(cudaMemcpyAsync(res->counting, d_hist.p, 65536 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    d_planes.release();
  }

  // ---- stage 2 ----
  uint64_t n_items = 0;
  if (ix.n_edges) {
    CK(cudaMemsetAsync(d_counter, 0, 8, st));
#define M(WW)                                                                                                          \
  if (W == WW)                                                                                                         \
    k_r2s_s2_extract<WW, false><<<grid_for(ix.n_edges, 256), 256, 0, st>>>(pv, k, d_solid.as<u32>(), m == 1, ix.n_edges, \
                                                                          nullptr, d_counter, 0);
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, d_counter, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    n_items = c;
    tr.mark("s2.count");
  }
  res->n_sort_items = n_items;
  DevBuf d_table, d_totals;
  CKR(d_table.alloc((size_t)MHB_NUM_BUCKETS * 32, "bucket table"));
  CKR(d_totals.alloc(16 * 8, "totals"));
  if (n_items == 0) {
    res->bytes = args->sdbg_out ? args->sdbg_out : (uint8_t *)malloc(1);
  } else {
    DevBuf heads, tile_heads, tile_off, bsum, scr, d_bytes;
    DevBuf &a = big.a, &b = big.b, &ws = big.ws;
    const size_t rec_bytes = (size_t)n_items * W * 4 + 16, ws_bytes = mhb_sort_workspace_bytes(n_items, W);
    CKR(a.ensure(rec_bytes, "stage-2 items"));
    CKR(b.ensure(rec_bytes, "stage-2 items (sort buffer)"));
    CKR(ws.ensure(ws_bytes, "sort workspace"));
    CK(cudaMemsetAsync(d_counter, 0, 8, st));
#define M(WW)                                                                                                          \
  if (W == WW)                                                                                                         \
    k_r2s_s2_extract<WW, true><<<grid_for(ix.n_edges, 256), 256, 0, st>>>(pv, k, d_solid.as<u32>(), m == 1, ix.n_edges,  \
                                                                         a.as<u32>(), d_counter, n_items);
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    tr.mark("s2.extract");
    uint8_t sbytes[80];
    const uint32_t n_sb = mhb_s2s_sort_bytes(k, sbytes);
    int in_b = 0;
    CKR(mhb_sort_records_relaxed(st, a.as<u32>(), b.as<u32>(), n_items, W, sbytes, n_sb, nullptr, ws.p, ws_bytes, &in_b));
    tr.mark("s2.sort");
    const u32 *sorted = in_b ? b.as<u32>() : a.as<u32>();
    u32 *uniq = in_b ? a.as<u32>() : b.as<u32>();
    // equal items -> one item carrying the run length (read_to_sdbg_s2.cpp:560-572)
    const uint64_t n_tiles = (n_items + kDdTile - 1) / kDdTile;
    CKR(tile_heads.alloc(n_tiles * 4, "tile counts"));
    CKR(tile_off.alloc(n_tiles * 8, "tile offsets"));
    CKR(bsum.alloc((n_tiles / 4096 + 4) * 8, "scan sums"));
#define M(WW) \
  if (W == WW) k_r2s_dd_count<WW><<<(unsigned)n_tiles, kDdThreads, 0, st>>>(sorted, n_items, tile_heads.as<u32>());
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    CKR(scan32(st, tile_heads.as<u32>(), n_tiles, tile_off.as<u64>(), (uint64_t *)(d_counter + 2), bsum.as<u64>()));
    unsigned long long n_u = 0;
    CK(cudaMemcpyAsync(&n_u, d_counter + 2, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CKR(heads.alloc((size_t)n_u * 8, "run heads"));
#define M(WW) \
  if (W == WW) k_r2s_dd_heads<WW><<<(unsigned)n_tiles, kDdThreads, 0, st>>>(sorted, n_items, tile_off.as<u64>(), heads.as<u64>());
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
#define M(WW) \
  if (W == WW) k_r2s_dd_build<WW><<<grid_for(n_u, 256), 256, 0, st>>>(sorted, n_items, heads.as<u64>(), n_u, uniq);
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    tr.mark("s2.collapse");
    const size_t scr_bytes = mhb_s2s_emit_scratch_bytes(n_u, k);
    const uint64_t cap_bytes = (uint64_t)n_u * (4ull + 4ull * WPT) + 16;
    CKR(scr.alloc(scr_bytes, "emit scratch"));
    CKR(d_bytes.alloc(cap_bytes, "SdBG stream"));
    CKR(mhb_s2s_emit_fmt(st, uniq, n_u, k, d_bytes.as<uint8_t>(), cap_bytes, d_table.as<u64>(), d_totals.as<u64>(), scr.p,
                         scr_bytes, 1));
    uint64_t totals[16];
    CK(cudaMemcpyAsync(totals, d_totals.p, sizeof(totals), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    tr.mark("s2.emit");
    res->n_bytes = totals[0];
    res->n_items = totals[1];
    res->n_tips = totals[2];
    res->n_large_mul = totals[3];
    for (int i = 0; i < 9; ++i) res->w_count[i] = totals[4 + i];
    res->ones_in_last = totals[13];
    if (res->n_bytes > cap_bytes) return mhb_set_error(MHB_ERR_NOMEM, "internal: SdBG byte stream exceeds capacity");
    if (args->sdbg_out && args->sdbg_out_capacity >= res->n_bytes) res->bytes = args->sdbg_out;
    else res->bytes = (uint8_t *)malloc(std::max<size_t>(1, res->n_bytes));
    if (!res->bytes) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
    CK(cudaMemcpyAsync(res->bucket_table, d_table.p, (size_t)MHB_NUM_BUCKETS * 32, cudaMemcpyDeviceToHost, st));
    if (res->n_bytes) CK(cudaMemcpyAsync(res->bytes, d_bytes.p, res->n_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    tr.mark("d2h");
    res->n_solid = n_u;  // distinct stage-2 items
  }
  tr.mark("release");
  CK(cudaStreamSynchronize(st));
  res->t_h2d_ms = tr.sum("h2d");
  res->t_count_ms = tr.sum("s1.extract") + tr.sum("s1.partition");  // records + stable bucket partition
  res->t_mercy_ms = tr.sum("s1.kmsort");                            // kmsort emulation (levels + finish)
  res->t_s2s_ms = tr.sum("s2");
  res->t_d2h_ms = tr.sum("d2h");
  res->t_total_ms = tr.sum("");
  tr.report();
  guard.ok = true;
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// Self-test hooks (host): the same __host__ __device__ code the kernels run, for the CPU-only tests (tests/test_r2s_cpu.py).
// ------------------------------------------------------------------------------------------------
extern "C" int mhb_selftest_r2s_s1_record(const uint32_t *pkg_words, uint32_t nwords, uint32_t L, uint32_t k, uint32_t e,
                                          uint64_t base_off, uint32_t *rec_out) {
  if (k < 9 || k > MHB_MAX_K || L < k + 1 || e >= L - k + 4) return mhb_set_error(MHB_ERR_ARG, "bad args");
  const uint32_t NW = r2s_s1_key_words(k);
  u32 p, want;
  s1_emission(L, k, e, p, want);
#define M(WW)                                                       \
  if (NW == WW) {                                                   \
    u32 rec[WW + 2];                                                \
    make_s1_record<WW>(pkg_words, nwords, L, k, p, want, base_off, rec); \
    memcpy(rec_out, rec, sizeof(rec));                              \
  }
  MHB_FOR_W(M)
#undef M
  return MHB_OK;
}

extern "C" int mhb_selftest_r2s_item(const uint32_t *pkg_words, uint32_t nwords, uint32_t L, uint32_t k, uint32_t i,
                                     uint32_t strand, uint32_t type, uint32_t *rec_out, uint32_t *palindrome_out) {
  if (k < 9 || k > MHB_MAX_K || i + k >= L || strand > 1 || type > 2) return mhb_set_error(MHB_ERR_ARG, "bad args");
  const uint32_t W = s2s_record_words(k);
#define M(WW)                                                              \
  if (W == WW) {                                                           \
    u32 rec[WW];                                                           \
    make_r2s_item<WW>(pkg_words, nwords, k, i, strand, type, rec);         \
    memcpy(rec_out, rec, sizeof(rec));                                     \
    *palindrome_out = edge_is_palindrome<WW>(pkg_words, nwords, k, i) ? 1u : 0u; \
  }
  MHB_FOR_WR(M)
#undef M
  return MHB_OK;
}

// kmsort emulation of ONE bucket on the host, exactly as the kernels do it (records of nw + 2 words): radix levels that
// mark the range starts, then the insertion sorts of all marked small ranges
extern "C" int mhb_selftest_kmsort(uint32_t *recs, uint64_t n, uint32_t nw) {
  const uint32_t RW = nw + 2;
  if (nw < 1 || RW > 17 || n >= (1ull << 32)) return mhb_set_error(MHB_ERR_ARG, "bad args");
  if (n <= 1) return MHB_OK;
  std::vector<KmSeg> cur, nxt;
  std::vector<u32> count(256), last(256), bnd(n / 32 + 2, 0u);
  int kb = 4 * (int)nw - 2 - 1;
  bit_or(bnd.data(), 0);
#define M(WW)                                                                                           \
  if (RW == WW) {                                                                                       \
    if (n > (uint64_t)kKmInsertThreshold) cur.push_back(KmSeg{0, n});                                   \
    while (!cur.empty()) {                                                                              \
      nxt.clear();                                                                                      \
      for (const KmSeg &s : cur) {                                                                      \
        km_radix_range<WW>(recs + s.start * WW, (u32)s.len, nw, kb, count.data(), last.data());         \
        u32 b0 = 0;                                                                                     \
        for (int i = 0; i < 256; ++i) {                                                                 \
          const u32 c = count[i];                                                                       \
          if (c) bit_or(bnd.data(), s.start + b0);                                                      \
          if (c > (u32)kKmInsertThreshold && kb > 0) nxt.push_back(KmSeg{s.start + b0, c});             \
          b0 += c;                                                                                      \
        }                                                                                               \
      }                                                                                                 \
      cur.swap(nxt);                                                                                    \
      --kb;                                                                                             \
    }                                                                                                   \
    for (u64 i = 0; i < n; ++i) {                                                                       \
      if (!bit_at(bnd.data(), i)) continue;                                                             \
      const u32 len = km_small_range(bnd.data(), n, i, (u32)kKmInsertThreshold);                        \
      if (len >= 2) km_insertion<WW>(recs + i * WW, len, nw);                                           \
    }                                                                                                   \
  }
  MHB_FOR_RW(M)
#undef M
  return MHB_OK;
}

// The shared-memory form of the same sort, mirrored on the host with the same building blocks (km_walk_src on the tags,
// km_insertion_idx on the index array of a staged range): level 0 gathers the bucket into a second buffer, later levels
// stage ranges of at most `wcap` records (0 = km_wcap of the record width, as on the device), larger ones and buckets
// above `cap` tags take the in-place walk; what is left unsorted is marked and finished by insertion.
extern "C" int mhb_selftest_kmsort_smem(uint32_t *recs, uint64_t n, uint32_t nw, uint32_t cap, uint32_t wcap) {
  const uint32_t RW = nw + 2;
  if (nw < 1 || RW > 17 || n >= (1ull << 32)) return mhb_set_error(MHB_ERR_ARG, "bad args");
  if (n <= 1) return MHB_OK;
  std::vector<KmSeg> cur, nxt;
  std::vector<u32> cnt(256), last(256), beg(256), bnd(n / 32 + 2, 0u), todo(n / 32 + 2, 0u);
  std::vector<uint8_t> tags(n);
  std::vector<uint16_t> src(n);
  std::vector<u32> other((size_t)n * RW), staged;
  int kb = 4 * (int)nw - 2 - 1;
#define M(WW)                                                                                                         \
  if (RW == WW) {                                                                                                     \
    if (!wcap) wcap = km_wcap(WW);                                                                                    \
    bit_or(bnd.data(), 0);                                                                                            \
    auto children = [&](u64 start, bool sorted_small) {                                                               \
      u32 acc = 0;                                                                                                    \
      for (int i = 0; i < 256; ++i) {                                                                                 \
        const u32 c = cnt[i];                                                                                         \
        if (c) {                                                                                                      \
          bit_or(bnd.data(), start + acc);                                                                            \
          if (kb > 0) {                                                                                               \
            if (c > (u32)kKmInsertThreshold) nxt.push_back(KmSeg{start + acc, c});                                    \
            else if (c >= 2 && !sorted_small) bit_or(todo.data(), start + acc);                                       \
          }                                                                                                           \
        }                                                                                                             \
        acc += c;                                                                                                     \
      }                                                                                                               \
    };                                                                                                                \
    /* level 0 (k_r2s_km_bucket) */                                                                                   \
    if (n <= (uint64_t)kKmInsertThreshold) {                                                                          \
      bit_or(todo.data(), 0);                                                                                         \
    } else if (n > cap || n > 65535) {                                                                                \
      km_radix_range<WW>(recs, (u32)n, nw, kb, cnt.data(), last.data());                                              \
      children(0, false);                                                                                             \
    } else {                                                                                                          \
      std::fill(cnt.begin(), cnt.end(), 0u);                                                                          \
      for (u32 i = 0; i < n; ++i) ++cnt[tags[i] = (uint8_t)km_byte_mem(recs + (size_t)i * WW, nw, kb)];               \
      for (u32 i = 0, acc = 0; i < 256; ++i) {                                                                        \
        last[i] = acc;                                                                                                \
        acc += cnt[i];                                                                                                \
      }                                                                                                               \
      km_walk_src<uint16_t>(tags.data(), (u32)n, cnt.data(), last.data(), src.data());                                \
      for (u32 q = 0; q < n; ++q) memcpy(&other[(size_t)q * WW], recs + (size_t)src[q] * WW, 4 * WW);                 \
      memcpy(recs, other.data(), (size_t)n * WW * 4);                                                                 \
      children(0, false);                                                                                             \
    }                                                                                                                 \
    /* levels >= 1 (k_r2s_km_warp) */                                                                                 \
    while (!nxt.empty() && kb > 0) {                                                                                  \
      cur.swap(nxt);                                                                                                  \
      nxt.clear();                                                                                                    \
      --kb;                                                                                                           \
      for (const KmSeg &sg : cur) {                                                                                   \
        u32 *a = recs + sg.start * WW;                                                                                \
        const u32 len = (u32)sg.len;                                                                                  \
        if (len > wcap) {                                                                                             \
          km_radix_range<WW>(a, len, nw, kb, cnt.data(), last.data());                                                \
          children(sg.start, false);                                                                                  \
          continue;                                                                                                   \
        }                                                                                                             \
        staged.assign(a, a + (size_t)len * WW);                                                                       \
        std::fill(cnt.begin(), cnt.end(), 0u);                                                                        \
        for (u32 i = 0; i < len; ++i) ++cnt[tags[i] = (uint8_t)km_byte_mem(&staged[(size_t)i * WW], nw, kb)];         \
        for (u32 i = 0, acc = 0; i < 256; ++i) {                                                                      \
          beg[i] = last[i] = acc;                                                                                     \
          acc += cnt[i];                                                                                              \
        }                                                                                                             \
        km_walk_src<uint16_t>(tags.data(), len, cnt.data(), last.data(), src.data());                                 \
        if (kb > 0)                                                                                                   \
          for (int b = 0; b < 256; ++b)                                                                               \
            if (cnt[b] >= 2 && cnt[b] <= (u32)kKmInsertThreshold)                                                     \
              km_insertion_idx<uint16_t>(staged.data(), WW, src.data() + beg[b], cnt[b], nw);                         \
        for (u32 q = 0; q < len; ++q) memcpy(a + (size_t)q * WW, &staged[(size_t)src[q] * WW], 4 * WW);               \
        children(sg.start, true);                                                                                     \
      }                                                                                                               \
    }                                                                                                                 \
    /* k_r2s_kmsort_finish */                                                                                         \
    for (u64 i = 0; i < n; ++i) {                                                                                     \
      if (!bit_at(todo.data(), i)) continue;                                                                          \
      const u32 len = km_small_range(bnd.data(), n, i, (u32)kKmInsertThreshold);                                      \
      if (len >= 2) km_insertion<WW>(recs + i * WW, len, nw);                                                         \
    }                                                                                                                 \
  }
  MHB_FOR_RW(M)
#undef M
  return MHB_OK;
}

// stage-1 Lv2Postprocess + the mercy step on host arrays (one sorted bucket / one read), same code as the kernels
extern "C" int mhb_selftest_r2s_s1_group(const uint32_t *recs, uint64_t n, uint32_t k, int32_t m, uint32_t fixed_len,
                                         uint64_t n_reads, int need_mercy, uint32_t *is_solid, uint32_t *no_in,
                                         uint32_t *no_out, uint32_t *any, int64_t *counting) {
  const uint32_t nw = r2s_s1_key_words(k), rw = nw + 2;
  PkgView pv;
  memset(&pv, 0, sizeof(pv));
  pv.fixed_len = fixed_len;
  pv.n_reads = n_reads;
  S1Out o{is_solid, no_in, no_out, any};
  for (u64 g = 0; g < n;) {
    u32 hv[16], nh;
    g = s1_group(recs, n, g, rw, nw, k, m, pv, o, need_mercy != 0, hv, nh);
    for (u32 q = 0; q < nh; ++q) counting[hv[q] > MHB_MAX_MUL ? MHB_MAX_MUL : hv[q]]++;
  }
  return MHB_OK;
}

extern "C" int mhb_selftest_r2s_mercy_read(uint32_t fixed_len, uint64_t n_reads, uint64_t r, uint32_t k,
                                           const uint32_t *is_solid, const uint32_t *no_in, const uint32_t *no_out,
                                           const uint32_t *any, uint32_t *mercy, uint32_t *added_out) {
  PkgView pv;
  memset(&pv, 0, sizeof(pv));
  pv.fixed_len = fixed_len;
  pv.n_reads = n_reads;
  S1Out o{const_cast<u32 *>(is_solid), const_cast<u32 *>(no_in), const_cast<u32 *>(no_out), const_cast<u32 *>(any)};
  *added_out = r2s_mercy_read(pv, r, k, o, mercy);
  return MHB_OK;
}
