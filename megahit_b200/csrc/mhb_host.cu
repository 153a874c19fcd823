// mhb_host.cu -- host-level C ABI (include/mhb.h, layer 2): host buffers in, host buffers out.
// Orchestrates the device-level entry points on one GPU with a grow-only device arena that is kept
// between calls (so repeated steps do not pay cudaMalloc).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mhb.h"
#include "mhb_bits.cuh"
#include "mhb_internal.h"

using namespace mhb;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return mhb_set_error(MHB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                           cudaGetErrorString(e_));                                                \
  } while (0)
#define CKR(call)           \
  do {                      \
    int rc_ = (call);       \
    if (rc_) return rc_;    \
  } while (0)

namespace {

struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t bytes) {
    used = 0;
    if (bytes <= cap) return MHB_OK;
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void **)&base, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return mhb_set_error(MHB_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = bytes;
    return MHB_OK;
  }
  template <class T>
  T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    char *p = base + used;
    used += bytes;
    return reinterpret_cast<T *>(p);
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};
Arena g_arena;

struct Timer {
  cudaEvent_t a, b;
  cudaStream_t st;
  explicit Timer(cudaStream_t s) : st(s) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~Timer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start() { cudaEventRecord(a, st); }
  double stop() {
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};

// Index a `.bin` image (sequence_package.h:224-240).  Fixed-length libraries need no side arrays.
struct BinIndex {
  uint32_t fixed_len = 0;
  std::vector<uint64_t> rec_off, edge_off;
  uint64_t n_edges = 0;
};

// sampled = true: a library whose size matches n_reads x (1 + ceil(L0/16)) is taken as fixed-length after looking at
// ~2048 of its length words only; the caller must then verify ALL of them on the device (mhb_check_fixed_len) and come
// back with sampled = false when that fails.  (The full host scan touches every cache line of the image: ~10 ms for
// 10 M reads, all of it inside the end-to-end time of the fused build.)
int index_bin(const uint32_t *bin, uint64_t bin_words, uint64_t n_reads, uint32_t k, BinIndex *ix, bool sampled = false) {
  ix->fixed_len = 0;
  ix->n_edges = 0;
  if (n_reads == 0) return MHB_OK;
  if (bin_words == 0) return mhb_set_error(MHB_ERR_ARG, "empty .bin image for %llu reads", (unsigned long long)n_reads);
  const uint32_t L0 = bin[0];
  const uint64_t stride = 1 + div_ceil(L0, 16);
  bool fixed = L0 > 0 && bin_words == n_reads * stride;
  if (fixed && sampled) {
    const uint64_t step = std::max<uint64_t>(1, n_reads / 1024);
    for (uint64_t r = 0; r < n_reads && fixed; r += step) fixed = bin[r * stride] == L0;
    for (uint64_t r = 0; r < std::min<uint64_t>(n_reads, 1024) && fixed; ++r) fixed = bin[r * stride] == L0;
    fixed = fixed && bin[(n_reads - 1) * stride] == L0;
  } else if (fixed) {
    int bad = 0;
#pragma omp parallel for reduction(| : bad) schedule(static)
    for (long long r = 0; r < (long long)n_reads; ++r) bad |= bin[(uint64_t)r * stride] != L0;
    fixed = !bad;
  }
  if (fixed) {
    ix->fixed_len = L0;
    ix->n_edges = L0 >= k + 1 ? n_reads * (uint64_t)(L0 - k) : 0;
    if (L0 < k + 1) ix->fixed_len = L0;  // kernels skip short reads themselves
    return MHB_OK;
  }
  ix->rec_off.resize(n_reads + 1);
  ix->edge_off.resize(n_reads + 1);
  uint64_t pos = 0, e = 0;
  for (uint64_t r = 0; r < n_reads; ++r) {
    if (pos >= bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated at read %llu", (unsigned long long)r);
    const uint32_t L = bin[pos];
    ix->rec_off[r] = pos;
    ix->edge_off[r] = e;
    if (L >= k + 1) e += L - k;
    pos += 1 + div_ceil(L, 16);
  }
  if (pos > bin_words) return mhb_set_error(MHB_ERR_ARG, ".bin image truncated");
  ix->rec_off[n_reads] = pos;
  ix->edge_off[n_reads] = e;
  ix->n_edges = e;
  return MHB_OK;
}

}  // namespace

extern "C" int mhb_release(void) {
  if (g_arena.base) cudaFree(g_arena.base);
  g_arena = Arena();
  return MHB_OK;
}


// ------------------------------------------------------------------------------------------------
// The count stage on extracted records resident in d_a: either the LSD sort on every key byte followed by the
// run-length count, or - 8-byte records, the default where available - two partition passes + per-bucket hash
// aggregation (mhb_count_solid_hashed).  MHB_COUNT_MODE=sort forces the former.  Both leave the same edges / aux /
// histogram; both clobber d_a and d_b.
// ------------------------------------------------------------------------------------------------
namespace {
struct CountWork {
  bool hashed;
  size_t bytes;      // one work area: sort workspace + count scratch, or the hashed path's workspace
  size_t ws_bytes;   // sort path: size of the leading sort workspace
  int hist_byte;     // record byte whose histogram the extraction must deliver
};
CountWork count_work_plan(uint64_t n, uint32_t k, int32_t m) {
  static const bool force_sort = getenv("MHB_COUNT_MODE") && !strcmp(getenv("MHB_COUNT_MODE"), "sort");
  CountWork cw;
  cw.hashed = !force_sort && mhb_count_hashed_supported(k, m);
  if (cw.hashed) {
    cw.bytes = mhb_count_hashed_workspace_bytes(n, k, m);
    cw.ws_bytes = 0;
    cw.hist_byte = 5;
  } else {
    uint8_t sb[72];
    mhb_count_sort_bytes(k, sb);
    cw.ws_bytes = Arena::pad(mhb_sort_workspace_bytes(n, count_record_words(k)));
    cw.bytes = cw.ws_bytes + Arena::pad(mhb_count_solid_scratch_bytes(n));
    cw.hist_byte = sb[0];
  }
  return cw;
}
int run_count_stage(cudaStream_t st, const CountWork &cw, uint32_t *d_a, uint32_t *d_b, uint64_t n, uint32_t k, int32_t m,
                    const uint64_t *d_hist0, uint32_t *d_edges, uint8_t *d_aux, uint64_t cap_edges, uint64_t *d_mul_hist,
                    uint64_t *d_nsolid, char *work, double *pass_ms, uint32_t *n_passes) {
  const uint32_t WR = count_record_words(k);
  if (cw.hashed) {
    if (n_passes) *n_passes = n ? 2 : 0;
    CKR(mhb_count_solid_hashed(st, d_a, d_b, n, k, m, d_hist0, d_edges, d_aux, cap_edges, d_mul_hist, d_nsolid, work, cw.bytes));
    if (pass_ms && n) {
      uint32_t np = 0, w = 0;
      uint64_t nr = 0;
      CKR(mhb_sort_pass_ms(0, pass_ms, 64, &np, &nr, &w));
    }
    return MHB_OK;
  }
  uint8_t sort_bytes[72];
  const uint32_t n_sort = mhb_count_sort_bytes(k, sort_bytes);
  if (n_passes) *n_passes = n ? n_sort : 0;
  int in_b = 0;
  CKR(mhb_sort_records_impl(st, d_a, d_b, n, WR, sort_bytes, n_sort, d_hist0, work, cw.ws_bytes, &in_b, pass_ms));
  return mhb_count_solid(st, in_b ? d_b : d_a, n, k, m, d_edges, d_aux, cap_edges, d_mul_hist, d_nsolid, work + cw.ws_bytes,
                         cw.bytes - cw.ws_bytes);
}
}  // namespace

namespace {
size_t round_bytes(uint64_t n, uint32_t WR, uint32_t WE, int32_t m, uint32_t k) {
  const uint64_t cap_edges = n / (uint64_t)std::max(1, m) + 1;
  return 2 * Arena::pad((size_t)n * WR * 4 + 16) + Arena::pad(count_work_plan(n, k, m).bytes) +
         Arena::pad((size_t)cap_edges * WE * 4) + Arena::pad(cap_edges);
}
}  // namespace

// ================================================================================================
// count, out of core (A13): rounds over ranges of the leading record byte
// ================================================================================================
namespace {
uint64_t g_round_limit = 0;      // count records per round; 0 = derive from free device memory
uint64_t g_s2s_round_limit = 0;  // seq2sdbg sort items per round; 0 = derive from free device memory

// bytes of device memory one round of `n` records needs besides the resident read library

}  // namespace

// Greedy cut of the 256 leading-byte values into contiguous ranges of at most max_records records each (pure host
// logic, callable without a GPU).  Returns the number of ranges (>= 1), or -1 when a single byte value alone exceeds
// the cap (poly-A like skew: reported, never mis-sorted).
extern "C" int mhb_plan_rounds(const uint64_t *hist256, uint64_t max_records, uint32_t *lo_out, uint32_t *hi_out) {
  if (!hist256 || !lo_out || !hi_out || max_records == 0) {
    mhb_set_error(MHB_ERR_ARG, "bad round plan arguments");
    return -1;
  }
  int n = 0;
  uint32_t lo = 0;
  uint64_t acc = 0;
  for (uint32_t b = 0; b < 256; ++b) {
    if (hist256[b] > max_records) {
      mhb_set_error(MHB_ERR_NOMEM, "leading byte 0x%02x alone holds %llu records, more than one round can take (%llu)", b,
                    (unsigned long long)hist256[b], (unsigned long long)max_records);
      return -1;
    }
    if (acc + hist256[b] > max_records) {
      lo_out[n] = lo;
      hi_out[n] = b - 1;
      ++n;
      lo = b;
      acc = 0;
    }
    acc += hist256[b];
  }
  lo_out[n] = lo;
  hi_out[n] = 255;
  return n + 1;
}

// Two-level planner (cf. Lv1FindEndBuckets, base_engine.cpp:254-281, which cuts on the 65 536 8-base buckets): the
// unit is the leading byte (256 bucket ids) unless that byte alone exceeds the cap - canonical (k+1)-mers are skewed
// towards A-prefixes, poly-A / low-complexity data more so - in which case the byte is cut on its second byte, i.e. on
// bucket ids.  sub_hist = 256 x 256 counts (row b = second-byte histogram of leading byte b); only rows of oversized
// bytes are read, and it may be NULL when there are none.  Output: ranges of 16-bit bucket ids.  Returns the number of
// ranges, -1 when a single bucket exceeds the cap or more than cap_out ranges are needed.
extern "C" int mhb_plan_rounds16(const uint64_t *hist256, const uint64_t *sub_hist, uint64_t max_records, uint32_t *lo_out,
                                 uint32_t *hi_out, uint32_t cap_out) {
  if (!hist256 || !lo_out || !hi_out || max_records == 0 || cap_out == 0) {
    mhb_set_error(MHB_ERR_ARG, "bad round plan arguments");
    return -1;
  }
  uint32_t n = 0, lo = 0;
  uint64_t acc = 0;
  bool open = false;
  auto close_at = [&](uint32_t hi) -> bool {  // ends the open range at bucket id hi
    if (n >= cap_out) return false;
    lo_out[n] = lo;
    hi_out[n] = hi;
    ++n;
    return true;
  };
  auto add = [&](uint32_t a_lo, uint32_t a_hi, uint64_t cnt) -> int {  // next atom [a_lo, a_hi] with cnt records
    if (cnt > max_records) {
      mhb_set_error(MHB_ERR_NOMEM, "bucket 0x%04x alone holds %llu records, more than one round can take (%llu)", a_lo,
                    (unsigned long long)cnt, (unsigned long long)max_records);
      return -1;
    }
    if (open && acc + cnt > max_records) {
      if (!close_at(a_lo - 1)) return -2;
      open = false;
    }
    if (!open) {
      lo = a_lo;
      acc = 0;
      open = true;
    }
    acc += cnt;
    (void)a_hi;
    return 0;
  };
  for (uint32_t b = 0; b < 256; ++b) {
    int rc = 0;
    if (hist256[b] > max_records) {
      if (!sub_hist) {
        mhb_set_error(MHB_ERR_NOMEM, "leading byte 0x%02x alone holds %llu records, more than one round can take (%llu)", b,
                      (unsigned long long)hist256[b], (unsigned long long)max_records);
        return -1;
      }
      for (uint32_t c = 0; c < 256 && !rc; ++c) rc = add((b << 8) | c, (b << 8) | c, sub_hist[(size_t)b * 256 + c]);
    } else {
      rc = add(b << 8, (b << 8) | 255u, hist256[b]);
    }
    if (rc == -1) return -1;
    if (rc == -2) break;
  }
  if (!open) {  // no records at all: one empty range covering everything
    lo = 0;
    open = true;
  }
  if (lo_out && n < cap_out) {
    lo_out[n] = n ? lo : 0;
    hi_out[n] = 65535;
    return (int)n + 1;
  }
  mhb_set_error(MHB_ERR_NOMEM, "round plan needs more than %u ranges", cap_out);
  return -1;
}

extern "C" int mhb_set_round_limit(uint64_t max_records_per_round) {
  g_round_limit = max_records_per_round;
  return MHB_OK;
}
extern "C" int mhb_set_s2s_round_limit(uint64_t max_items_per_round) {
  g_s2s_round_limit = max_items_per_round;
  return MHB_OK;
}

// The reference plans Lv1 passes over bucket ranges so that every pass fits the memory it was given
// (base_engine.cpp:54-141 AdjustMemory, :254-281 Lv1FindEndBuckets); the output does not depend on where the pass
// boundaries fall.  Here a round = a contiguous range of leading record bytes (four bases) whose records fit in HBM
// next to the resident read library: extract that range -> sort -> solid edges -> append to the host result.  Rounds
// ascend, so the concatenated edges are sorted.  Tip edges (aux != 0) of all rounds are collected on the host and the
// mercy bookkeeping runs once at the end over the whole library.
static int count_host_rounds(const mhb_count_args *args, mhb_count_result *res, const BinIndex &ix, uint64_t max_records) {
  const uint32_t k = args->k;
  const int32_t m = args->m;
  const uint64_t n = ix.n_edges, n_reads = args->n_reads;
  const uint32_t WR = count_record_words(k), WE = words_per_edge(k);
  uint8_t sort_bytes[72];
  const uint32_t n_sort = mhb_count_sort_bytes(k, sort_bytes);
  const int top_byte = (int)(4 * WR - 1);
  cudaStream_t st = 0;
  Timer t_all(st), t(st);
  t_all.start();

  const size_t bin_bytes = (args->bin_words * 4 + 15) & ~(size_t)15;
  size_t fixed = Arena::pad(bin_bytes + 16) + Arena::pad((n_reads + 2) * 8) + Arena::pad(65536 * 8) + 2 * Arena::pad(256 * 8) +
                 Arena::pad(64) + 8192;
  if (!ix.fixed_len) fixed += 2 * Arena::pad((n_reads + 1) * 8);
  if (args->want_mercy) fixed += 2 * Arena::pad((size_t)(n_reads + 1) * 4);
  if (!max_records) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)((double)(free_b + g_arena.cap) * 0.92);
    if (avail <= fixed) return mhb_set_error(MHB_ERR_NOMEM, "the read library alone (%zu bytes) does not fit the device", fixed);
    uint64_t lo = 1, hi = n;  // largest round that fits (round_bytes is monotone)
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo + 1) / 2;
      if (fixed + round_bytes(mid, WR, WE, m, k) <= avail) lo = mid;
      else hi = mid - 1;
    }
    max_records = lo;
  }
  max_records = std::min<uint64_t>(std::max<uint64_t>(max_records, 1), std::max<uint64_t>(n, 1));
  const uint64_t cap_edges = max_records / (uint64_t)std::max(1, m) + 1;
  CKR(g_arena.reserve(fixed + round_bytes(max_records, WR, WE, m, k)));

  uint32_t *d_bin = g_arena.take<uint32_t>(bin_bytes / 4 + 4);
  uint64_t *d_per_read = g_arena.take<uint64_t>(n_reads + 2);
  uint64_t *d_mul_hist = g_arena.take<uint64_t>(65536);
  uint64_t *d_hist0 = g_arena.take<uint64_t>(256);
  uint64_t *d_hist_top = g_arena.take<uint64_t>(256);
  uint64_t *d_scalars = g_arena.take<uint64_t>(8);  // [0] n_solid, [1] round total
  uint64_t *d_rec_off = nullptr, *d_edge_off = nullptr;
  uint32_t *d_first = nullptr, *d_last = nullptr;
  if (!ix.fixed_len) {
    d_rec_off = g_arena.take<uint64_t>(n_reads + 1);
    d_edge_off = g_arena.take<uint64_t>(n_reads + 1);
  }
  if (args->want_mercy) {
    d_first = g_arena.take<uint32_t>(n_reads + 1);
    d_last = g_arena.take<uint32_t>(n_reads + 1);
  }
  uint32_t *d_a = g_arena.take<uint32_t>((size_t)max_records * WR + 4);
  uint32_t *d_b = g_arena.take<uint32_t>((size_t)max_records * WR + 4);
  const CountWork cw = count_work_plan(max_records, k, m);
  char *d_work = g_arena.take<char>(cw.bytes);
  uint32_t *d_edges = g_arena.take<uint32_t>((size_t)cap_edges * WE);
  uint8_t *d_aux = g_arena.take<uint8_t>(cap_edges);

  t.start();
  if (args->bin_words) CK(cudaMemcpyAsync(d_bin, args->bin, args->bin_words * 4, cudaMemcpyHostToDevice, st));
  if (!ix.fixed_len && n_reads) {
    CK(cudaMemcpyAsync(d_rec_off, ix.rec_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_edge_off, ix.edge_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(d_mul_hist, 0, 65536 * 8, st));
  CK(cudaMemsetAsync(d_hist_top, 0, 256 * 8, st));
  res->t_h2d_ms = t.stop();

  mhb_dev_reads reads;
  reads.bin = d_bin;
  reads.bin_words = args->bin_words;
  reads.n_reads = n_reads;
  reads.fixed_len = ix.fixed_len;
  reads.rec_off = d_rec_off;
  reads.edge_off = d_edge_off;

  // ---- plan: histogram of the leading byte over the whole library (+ of the second byte inside every leading byte
  // that alone exceeds a round), then greedy contiguous ranges of bucket ids ----
  t.start();
  uint64_t h_top[256];
  CKR(mhb_count_extract_range(st, &reads, k, 0, 65535, 0, d_per_read, nullptr, d_hist_top, top_byte, d_scalars + 1));
  CK(cudaMemcpyAsync(h_top, d_hist_top, sizeof(h_top), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<uint64_t> h_sub;
  for (uint32_t b = 0; b < 256; ++b) {
    if (h_top[b] <= max_records) continue;
    if (WR * 4 < 2) return mhb_set_error(MHB_ERR_NOMEM, "leading byte 0x%02x exceeds a round and the record has no second byte", b);
    if (h_sub.empty()) h_sub.assign(256 * 256, 0);
    CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
    CKR(mhb_count_extract_range(st, &reads, k, b << 8, (b << 8) | 255u, 0, d_per_read, nullptr, d_hist0, top_byte - 1, d_scalars + 1));
    CK(cudaMemcpyAsync(h_sub.data() + (size_t)b * 256, d_hist0, 256 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  std::vector<uint32_t> r_lo(65536), r_hi(65536);
  const int n_ranges = mhb_plan_rounds16(h_top, h_sub.empty() ? nullptr : h_sub.data(), max_records, r_lo.data(), r_hi.data(), 65536);
  if (n_ranges < 0) return MHB_ERR_NOMEM;  // message set by the planner
  std::vector<std::pair<uint32_t, uint32_t>> ranges;
  for (int i = 0; i < n_ranges; ++i) ranges.push_back({r_lo[i], r_hi[i]});
  res->t_extract_ms = t.stop();

  std::vector<uint32_t> h_edges;      // all solid edges, ascending
  std::vector<uint32_t> h_tip_edges;  // the ones with aux != 0
  std::vector<uint8_t> h_tip_aux, h_aux;
  uint64_t n_solid_total = 0;
  for (const auto &rg : ranges) {
    // ---- extract the range ----
    t.start();
    uint64_t n_round = 0;
    CKR(mhb_count_extract_range(st, &reads, k, rg.first, rg.second, 0, d_per_read, nullptr, nullptr, 0, d_scalars + 1));
    CK(cudaMemcpyAsync(&n_round, d_scalars + 1, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
    CK(cudaMemsetAsync(d_scalars, 0, 8, st));
    CK(cudaStreamSynchronize(st));
    if (n_round > max_records) return mhb_set_error(MHB_ERR_NOMEM, "internal: round of %llu records exceeds its plan", (unsigned long long)n_round);
    if (n_round == 0) continue;
    CKR(mhb_count_extract_range(st, &reads, k, rg.first, rg.second, 1, d_per_read, d_a, d_hist0, cw.hist_byte, nullptr));
    res->t_extract_ms += t.stop();
    // ---- sort / partition + solid edges ----
    t.start();
    CKR(run_count_stage(st, cw, d_a, d_b, n_round, k, m, d_hist0, d_edges, d_aux, cap_edges, d_mul_hist, d_scalars, d_work,
                        nullptr, nullptr));
    uint64_t n_solid = 0;
    CK(cudaMemcpyAsync(&n_solid, d_scalars, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    res->t_count_ms += t.stop();
    if (n_solid > cap_edges) return mhb_set_error(MHB_ERR_NOMEM, "internal: solid edges exceed capacity");
    // ---- append to the host result ----
    t.start();
    const size_t e0 = h_edges.size();
    h_edges.resize(e0 + (size_t)n_solid * WE);
    h_aux.resize(n_solid);
    if (n_solid) {
      CK(cudaMemcpyAsync(h_edges.data() + e0, d_edges, (size_t)n_solid * WE * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(h_aux.data(), d_aux, n_solid, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    if (args->want_mercy)
      for (uint64_t i = 0; i < n_solid; ++i)
        if (h_aux[i]) {
          h_tip_edges.insert(h_tip_edges.end(), h_edges.begin() + e0 + i * WE, h_edges.begin() + e0 + (i + 1) * WE);
          h_tip_aux.push_back(h_aux[i]);
        }
    n_solid_total += n_solid;
    res->t_d2h_ms += t.stop();
    ++res->n_rounds;
  }
  res->n_solid = n_solid_total;

  // ---- mercy bookkeeping over the whole library with the tip edges of all rounds ----
  std::vector<uint32_t> h_first, h_last;
  if (args->want_mercy && n_reads) {
    t.start();
    const uint64_t n_tip = h_tip_aux.size();
    const size_t ts_bytes = mhb_tipset_bytes(n_tip, k);
    const size_t list_bytes = Arena::pad(h_tip_edges.size() * 4 + 16) + Arena::pad(n_tip + 16);
    // the per-round buffers are free now
    char *d_tmp = nullptr;
    bool own = false;
    const size_t have = (size_t)max_records * WR * 4 * 2;
    if (ts_bytes + list_bytes + 512 <= have) d_tmp = (char *)d_a;
    else {
      CK(cudaMalloc((void **)&d_tmp, ts_bytes + list_bytes + 512));
      own = true;
    }
    uint32_t *d_tip_edges = (uint32_t *)d_tmp;
    uint8_t *d_tip_aux = (uint8_t *)(d_tmp + Arena::pad(h_tip_edges.size() * 4 + 16));
    char *d_tips = d_tmp + list_bytes;
    int rc = MHB_OK;
    if (n_tip) {
      if (cudaMemcpyAsync(d_tip_edges, h_tip_edges.data(), h_tip_edges.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(d_tip_aux, h_tip_aux.data(), n_tip, cudaMemcpyHostToDevice, st) != cudaSuccess)
        rc = mhb_set_error(MHB_ERR_CUDA, "tip edge upload failed");
    }
    if (!rc) rc = mhb_tipset_build(st, d_tip_edges, d_tip_aux, n_tip, k, d_tips, ts_bytes, n_tip);
    if (!rc) rc = mhb_count_mark_mercy(st, &reads, k, d_tips, ts_bytes, n_tip, d_first, d_last);
    h_first.resize(n_reads);
    h_last.resize(n_reads);
    if (!rc) {
      cudaMemcpyAsync(h_first.data(), d_first, n_reads * 4, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(h_last.data(), d_last, n_reads * 4, cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) rc = mhb_set_error(MHB_ERR_CUDA, "mercy marking failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (own) cudaFree(d_tmp);
    if (rc) return rc;
    res->t_mercy_ms = t.stop();
  }

  res->edges = (uint32_t *)malloc(std::max<size_t>(1, h_edges.size() * 4));
  if (!res->edges) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  if (!h_edges.empty()) memcpy(res->edges, h_edges.data(), h_edges.size() * 4);
  std::vector<uint64_t> h_hist(65536);
  CK(cudaMemcpy(h_hist.data(), d_mul_hist, 65536 * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i <= MHB_MAX_MUL; ++i) res->counting[i] = (int64_t)h_hist[i];
  if (args->want_mercy) {  // kmer_counter.cpp:390-401
    std::vector<uint64_t> ids;
    for (uint64_t r = 0; r < n_reads; ++r) {
      const uint32_t f = h_first[r], l = h_last[r];
      if (f != MHB_SENTINEL_OFFSET && l != MHB_SENTINEL_OFFSET) {
        ++res->n_has_tips;
        if (l > f) ids.push_back(r);
      }
    }
    res->n_cand = ids.size();
    res->cand_ids = (uint64_t *)malloc(std::max<size_t>(1, ids.size() * 8));
    if (!ids.empty()) memcpy(res->cand_ids, ids.data(), ids.size() * 8);
  }
  res->t_total_ms = t_all.stop();
  return MHB_OK;
}

// ================================================================================================
// count
// ================================================================================================
extern "C" int mhb_count_host(const mhb_count_args *args, mhb_count_result *res) {
  if (!args || !res) return mhb_set_error(MHB_ERR_ARG, "null args");
  memset(res, 0, sizeof(*res));
  const uint32_t k = args->k;
  if (k < 1 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "kmer size %u out of range", k);
  if (mhb_device_count() == 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  res->words_per_edge = words_per_edge(k);

  BinIndex ix;
  CKR(index_bin(args->bin, args->bin_words, args->n_reads, k, &ix));
  const uint64_t n = ix.n_edges, n_reads = args->n_reads;
  res->n_edge_records = n;

  cudaStream_t st = 0;
  Timer t_all(st), t(st);
  t_all.start();

  const uint32_t WR = count_record_words(k), WE = words_per_edge(k);
  uint8_t sort_bytes[72];
  const uint32_t n_sort = mhb_count_sort_bytes(k, sort_bytes);
  const int32_t m = args->m;
  const uint64_t cap_edges = n / (uint64_t)std::max(1, m) + 1;

  const size_t bin_bytes = (args->bin_words * 4 + 15) & ~(size_t)15;
  const CountWork cw = count_work_plan(n, k, m);
  size_t need = Arena::pad(bin_bytes + 16) + 2 * Arena::pad((size_t)n * WR * 4 + 16) + Arena::pad(cw.bytes) +
                Arena::pad((size_t)cap_edges * WE * 4) + Arena::pad(cap_edges) +
                Arena::pad(65536 * 8) + Arena::pad(256 * 8) + 4096;
  if (!ix.fixed_len) need += 2 * Arena::pad((n_reads + 1) * 8);
  if (args->want_mercy) need += 2 * Arena::pad((size_t)(n_reads + 1) * 4);
  {
    // A13: when one pass over all records does not fit the device (or the caller capped the round size), run the
    // stage in rounds over ranges of the leading record byte
    bool rounds = g_round_limit && n > g_round_limit;
    if (!rounds && need > g_arena.cap) {
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      rounds = (double)need > 0.92 * (double)(free_b + g_arena.cap);
    }
    if (rounds) return count_host_rounds(args, res, ix, g_round_limit);
  }
  CKR(g_arena.reserve(need));

  uint32_t *d_bin = g_arena.take<uint32_t>(bin_bytes / 4 + 4);
  uint32_t *d_a = g_arena.take<uint32_t>((size_t)n * WR + 4);
  uint32_t *d_b = g_arena.take<uint32_t>((size_t)n * WR + 4);
  char *d_work = g_arena.take<char>(cw.bytes);
  uint32_t *d_edges = g_arena.take<uint32_t>((size_t)cap_edges * WE);
  uint8_t *d_aux = g_arena.take<uint8_t>(cap_edges);
  uint64_t *d_mul_hist = g_arena.take<uint64_t>(65536);
  uint64_t *d_hist0 = g_arena.take<uint64_t>(256);
  uint64_t *d_nsolid = g_arena.take<uint64_t>(8);
  uint64_t *d_rec_off = nullptr, *d_edge_off = nullptr;
  uint32_t *d_first = nullptr, *d_last = nullptr;
  if (!ix.fixed_len) {
    d_rec_off = g_arena.take<uint64_t>(n_reads + 1);
    d_edge_off = g_arena.take<uint64_t>(n_reads + 1);
  }
  if (args->want_mercy) {
    d_first = g_arena.take<uint32_t>(n_reads + 1);
    d_last = g_arena.take<uint32_t>(n_reads + 1);
  }

  // ---- H2D ----
  t.start();
  if (args->bin_words) CK(cudaMemcpyAsync(d_bin, args->bin, args->bin_words * 4, cudaMemcpyHostToDevice, st));
  if (!ix.fixed_len && n_reads) {
    CK(cudaMemcpyAsync(d_rec_off, ix.rec_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_edge_off, ix.edge_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(d_mul_hist, 0, 65536 * 8, st));
  CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
  CK(cudaMemsetAsync(d_nsolid, 0, 64, st));
  res->t_h2d_ms = t.stop();

  mhb_dev_reads reads;
  reads.bin = d_bin;
  reads.bin_words = args->bin_words;
  reads.n_reads = n_reads;
  reads.fixed_len = ix.fixed_len;
  reads.rec_off = d_rec_off;
  reads.edge_off = d_edge_off;

  // ---- extract ----
  t.start();
  CKR(mhb_count_extract(st, &reads, k, d_a, n, d_hist0, cw.hist_byte));
  res->t_extract_ms = t.stop();

  // ---- sort (or partition) + count ----
  t.start();
  res->n_rounds = 1;
  CKR(run_count_stage(st, cw, d_a, d_b, n, k, m, d_hist0, d_edges, d_aux, cap_edges, d_mul_hist, d_nsolid, d_work,
                      res->sort_pass_ms, &res->n_sort_passes));
  for (uint32_t p = 0; p < res->n_sort_passes; ++p) res->t_sort_ms += res->sort_pass_ms[p];
  uint64_t n_solid = 0;
  CK(cudaMemcpyAsync(&n_solid, d_nsolid, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  res->t_count_ms = t.stop();
  if (n_solid > cap_edges) return mhb_set_error(MHB_ERR_NOMEM, "internal: solid edges exceed capacity");
  res->n_solid = n_solid;

  // ---- mercy bookkeeping ----
  std::vector<uint32_t> h_first, h_last;
  if (args->want_mercy && n_reads) {
    t.start();
    uint64_t n_tip = 0;
    CKR(mhb_count_tip_edges(st, d_aux, n_solid, &n_tip));
    const size_t ts_bytes = mhb_tipset_bytes(n_tip, k);
    // the ping-pong buffers are free now: put the tip set there if it fits
    char *d_tips = nullptr;
    uint32_t *d_free = d_a;
    bool own = false;
    if (ts_bytes <= (size_t)n * WR * 4) d_tips = (char *)d_free;
    else {
      CK(cudaMalloc((void **)&d_tips, ts_bytes));
      own = true;
    }
    int rc = mhb_tipset_build(st, d_edges, d_aux, n_solid, k, d_tips, ts_bytes, n_tip);
    if (!rc) rc = mhb_count_mark_mercy(st, &reads, k, d_tips, ts_bytes, n_tip, d_first, d_last);
    h_first.resize(n_reads);
    h_last.resize(n_reads);
    if (!rc) {
      cudaMemcpyAsync(h_first.data(), d_first, n_reads * 4, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(h_last.data(), d_last, n_reads * 4, cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) rc = mhb_set_error(MHB_ERR_CUDA, "mercy marking failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (own) cudaFree(d_tips);
    if (rc) return rc;
    res->t_mercy_ms = t.stop();
  }

  // ---- D2H ----
  t.start();
  res->edges = (uint32_t *)malloc(std::max<size_t>(1, (size_t)n_solid * WE * 4));
  if (!res->edges) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  if (n_solid) CK(cudaMemcpyAsync(res->edges, d_edges, (size_t)n_solid * WE * 4, cudaMemcpyDeviceToHost, st));
  std::vector<uint64_t> h_hist(65536);
  CK(cudaMemcpyAsync(h_hist.data(), d_mul_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  res->t_d2h_ms = t.stop();
  for (int i = 0; i <= MHB_MAX_MUL; ++i) res->counting[i] = (int64_t)h_hist[i];

  if (args->want_mercy) {  // kmer_counter.cpp:390-401
    std::vector<uint64_t> ids;
    for (uint64_t r = 0; r < n_reads; ++r) {
      const uint32_t f = h_first[r], l = h_last[r];
      if (f != MHB_SENTINEL_OFFSET && l != MHB_SENTINEL_OFFSET) {
        ++res->n_has_tips;
        if (l > f) ids.push_back(r);
      }
    }
    res->n_cand = ids.size();
    res->cand_ids = (uint64_t *)malloc(std::max<size_t>(1, ids.size() * 8));
    if (!ids.empty()) memcpy(res->cand_ids, ids.data(), ids.size() * 8);
  }
  res->t_total_ms = t_all.stop();
  return MHB_OK;
}

// ================================================================================================
// seq2sdbg, out of core (A13): rounds over ranges of the leading record byte
// ================================================================================================
namespace {
size_t s2s_round_bytes(uint64_t n, uint32_t W, uint32_t k) {
  return 2 * Arena::pad((size_t)n * W * 4 + 16) + Arena::pad(mhb_sort_workspace_bytes(n, W)) +
         Arena::pad(mhb_s2s_emit_scratch_bytes(n, k)) + Arena::pad((size_t)n * (4ull + 4ull * words_per_tip_label(k)) + 16);
}
}  // namespace

// Same idea as count_host_rounds: the sort items of all sequences do not fit in HBM next to the sequences, so the
// stage runs once per contiguous range of leading record bytes (a (k-1)-mer group, and a bucket, never spans two
// ranges): extract the range -> sort -> emit -> append the item bytes and that range's rows of the bucket table to the
// host result.  Ranges ascend, so the concatenated stream is in bucket order.
static int s2s_host_rounds(const mhb_s2s_args *args, mhb_s2s_result *res, const std::vector<uint64_t> &item_off,
                           uint64_t n_items, bool fixed, uint32_t L0, uint64_t n_words, uint64_t max_items) {
  const uint32_t k = args->k;
  const uint64_t ns = args->n_seqs;
  const uint32_t W = s2s_record_words(k), WPT = words_per_tip_label(k);
  uint8_t sort_bytes[72];
  const uint32_t n_sort = mhb_s2s_sort_bytes(k, sort_bytes);
  const int top_byte = (int)(4 * W - 1);
  cudaStream_t st = 0;
  Timer t_all(st), t(st);
  t_all.start();

  const size_t fixed_b = Arena::pad(n_words * 4 + 64) + Arena::pad((ns + 1) * 8) * 2 + Arena::pad((ns + 1) * 4) +
                         Arena::pad((ns + 1) * 2) + Arena::pad((size_t)MHB_NUM_BUCKETS * 4 * 8) + 2 * Arena::pad(256 * 8) +
                         Arena::pad(16 * 8) + Arena::pad(64) + 8192;
  if (!max_items) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)((double)(free_b + g_arena.cap) * 0.92);
    if (avail <= fixed_b) return mhb_set_error(MHB_ERR_NOMEM, "the sequences alone (%zu bytes) do not fit the device", fixed_b);
    uint64_t lo = 1, hi = n_items;
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo + 1) / 2;
      if (fixed_b + s2s_round_bytes(mid, W, k) <= avail) lo = mid;
      else hi = mid - 1;
    }
    max_items = lo;
  }
  max_items = std::min<uint64_t>(std::max<uint64_t>(max_items, 1), std::max<uint64_t>(n_items, 1));
  CKR(g_arena.reserve(fixed_b + s2s_round_bytes(max_items, W, k)));
  uint32_t *d_words = g_arena.take<uint32_t>(n_words + 16);
  uint64_t *d_word_off = g_arena.take<uint64_t>(ns + 1);
  uint64_t *d_item_off = g_arena.take<uint64_t>(ns + 1);
  uint32_t *d_len = g_arena.take<uint32_t>(ns + 1);
  uint16_t *d_mult = g_arena.take<uint16_t>(ns + 1);
  uint64_t *d_table = g_arena.take<uint64_t>((size_t)MHB_NUM_BUCKETS * 4);
  uint64_t *d_hist0 = g_arena.take<uint64_t>(256);
  uint64_t *d_hist_top = g_arena.take<uint64_t>(256);
  uint64_t *d_totals = g_arena.take<uint64_t>(16);
  uint64_t *d_cursor = g_arena.take<uint64_t>(8);
  uint32_t *d_a = g_arena.take<uint32_t>((size_t)max_items * W + 4);
  uint32_t *d_b = g_arena.take<uint32_t>((size_t)max_items * W + 4);
  const size_t ws_bytes = mhb_sort_workspace_bytes(max_items, W);
  const size_t scratch_bytes = mhb_s2s_emit_scratch_bytes(max_items, k);
  const uint64_t cap_bytes = max_items * (4ull + 4ull * WPT) + 16;
  char *d_ws = g_arena.take<char>(ws_bytes);
  char *d_scratch = g_arena.take<char>(scratch_bytes);
  uint8_t *d_bytes = g_arena.take<uint8_t>(cap_bytes);

  if (ns) {
    CK(cudaMemcpyAsync(d_words, args->words, n_words * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_word_off, args->word_off, (ns + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_item_off, item_off.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_len, args->len, ns * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_mult, args->mult, ns * 2, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(d_hist_top, 0, 256 * 8, st));
  mhb_dev_seqs seqs;
  seqs.words = d_words;
  seqs.n_words = n_words;
  seqs.n_seqs = ns;
  seqs.fixed_len = fixed ? L0 : 0;
  seqs.word_off = d_word_off;
  seqs.len = d_len;
  seqs.item_off = d_item_off;
  seqs.mult = d_mult;
  seqs.fixed_stride = 0;

  // ---- plan (two-level, as in count_host_rounds) ----
  t.start();
  uint64_t h_top[256];
  CKR(mhb_s2s_extract_range(st, &seqs, k, nullptr, n_items, 0, 65535, nullptr, 0, d_hist_top, top_byte));
  CK(cudaMemcpyAsync(h_top, d_hist_top, sizeof(h_top), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<uint64_t> h_sub;
  for (uint32_t b = 0; b < 256; ++b) {
    if (h_top[b] <= max_items) continue;
    if (h_sub.empty()) h_sub.assign(256 * 256, 0);
    CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
    CKR(mhb_s2s_extract_range(st, &seqs, k, nullptr, n_items, b << 8, (b << 8) | 255u, nullptr, 0, d_hist0, top_byte - 1));
    CK(cudaMemcpyAsync(h_sub.data() + (size_t)b * 256, d_hist0, 256 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  std::vector<uint32_t> r_lo(65536), r_hi(65536);
  const int n_ranges = mhb_plan_rounds16(h_top, h_sub.empty() ? nullptr : h_sub.data(), max_items, r_lo.data(), r_hi.data(), 65536);
  if (n_ranges < 0) return MHB_ERR_NOMEM;
  res->t_extract_ms = t.stop();

  std::vector<uint8_t> h_bytes;
  std::vector<uint64_t> h_table((size_t)MHB_NUM_BUCKETS * 4), h_round_table((size_t)MHB_NUM_BUCKETS * 4);
  uint64_t tot[16] = {0};
  for (int ri = 0; ri < n_ranges; ++ri) {
    t.start();
    CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
    CK(cudaMemsetAsync(d_cursor, 0, 64, st));
    CKR(mhb_s2s_extract_range(st, &seqs, k, d_a, n_items, r_lo[ri], r_hi[ri], d_cursor, max_items, d_hist0, sort_bytes[0]));
    uint64_t n_round = 0;
    CK(cudaMemcpyAsync(&n_round, d_cursor, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    res->t_extract_ms += t.stop();
    if (n_round > max_items) return mhb_set_error(MHB_ERR_NOMEM, "internal: round of %llu items exceeds its plan", (unsigned long long)n_round);
    if (n_round == 0) continue;
    t.start();
    int in_b = 0;
    CKR(mhb_sort_records_impl(st, d_a, d_b, n_round, W, sort_bytes, n_sort, d_hist0, d_ws, ws_bytes, &in_b, nullptr));
    res->t_sort_ms += t.stop();
    t.start();
    CKR(mhb_s2s_emit(st, in_b ? d_b : d_a, n_round, k, d_bytes, cap_bytes, d_table, d_totals, d_scratch, scratch_bytes));
    uint64_t rt[16];
    CK(cudaMemcpyAsync(rt, d_totals, sizeof(rt), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_round_table.data(), d_table, h_round_table.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (rt[0] > cap_bytes) return mhb_set_error(MHB_ERR_NOMEM, "internal: SdBG byte stream exceeds capacity");
    const size_t base = h_bytes.size();
    h_bytes.resize(base + rt[0]);
    if (rt[0]) CK(cudaMemcpy(h_bytes.data() + base, d_bytes, rt[0], cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < (size_t)MHB_NUM_BUCKETS; ++b)
      if (h_round_table[4 * b + 1]) {
        h_table[4 * b + 0] = h_round_table[4 * b + 0] + base;
        h_table[4 * b + 1] = h_round_table[4 * b + 1];
        h_table[4 * b + 2] = h_round_table[4 * b + 2];
        h_table[4 * b + 3] = h_round_table[4 * b + 3];
      }
    for (int i = 0; i < 16; ++i) tot[i] += rt[i];
    res->t_emit_ms += t.stop();
  }
  res->n_bytes = tot[0];
  res->n_items = tot[1];
  res->n_tips = tot[2];
  res->n_large_mul = tot[3];
  for (int i = 0; i < 9; ++i) res->w_count[i] = tot[4 + i];
  res->ones_in_last = tot[13];
  memcpy(res->bucket_table, h_table.data(), sizeof(res->bucket_table));
  res->bytes = (uint8_t *)malloc(std::max<size_t>(1, h_bytes.size()));
  if (!res->bytes) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  if (!h_bytes.empty()) memcpy(res->bytes, h_bytes.data(), h_bytes.size());
  res->t_total_ms = t_all.stop();
  return MHB_OK;
}

// ================================================================================================
// seq2sdbg
// ================================================================================================
extern "C" int mhb_s2s_host(const mhb_s2s_args *args, mhb_s2s_result *res) {
  if (!args || !res) return mhb_set_error(MHB_ERR_ARG, "null args");
  memset(res, 0, sizeof(*res));
  const uint32_t k = args->k;
  if (k < 9 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "kmer size must be >= 9!");
  if (mhb_device_count() == 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  res->words_per_tip_label = words_per_tip_label(k);
  const uint64_t ns = args->n_seqs;

  // item offsets; detect the fixed-length, gap-free layout (edges only)
  std::vector<uint64_t> item_off(ns + 1);
  uint64_t n_items = 0;
  bool fixed = ns > 0;
  const uint32_t L0 = ns ? args->len[0] : 0;
  for (uint64_t s = 0; s < ns; ++s) {
    item_off[s] = n_items;
    const uint32_t L = args->len[s];
    if (L >= k + 1) n_items += 2ull * (L - k + 2);
    if (L != L0 || args->word_off[s] != s * (uint64_t)div_ceil(L0, 16)) fixed = false;
  }
  item_off[ns] = n_items;
  if (L0 < k + 1) fixed = false;
  const uint64_t n_words = ns ? args->word_off[ns] : 0;
  res->n_records = n_items;

  cudaStream_t st = 0;
  Timer t_all(st), t(st);
  t_all.start();
  const uint32_t W = s2s_record_words(k);
  uint8_t sort_bytes[72];
  const uint32_t n_sort = mhb_s2s_sort_bytes(k, sort_bytes);
  const size_t ws_bytes = mhb_sort_workspace_bytes(n_items, W);
  const size_t scratch_bytes = mhb_s2s_emit_scratch_bytes(n_items, k);
  // worst case bytes per sort item: 2 + 2 + 4*WPT (every item a large-multiplicity tip)
  const uint64_t cap_bytes = n_items * (4ull + 4ull * res->words_per_tip_label) + 16;
  size_t need = Arena::pad(n_words * 4 + 64) + Arena::pad((ns + 1) * 8) * 2 + Arena::pad((ns + 1) * 4) +
                Arena::pad((ns + 1) * 2) + 2 * Arena::pad((size_t)n_items * W * 4 + 16) + Arena::pad(ws_bytes) +
                Arena::pad(scratch_bytes) + Arena::pad(cap_bytes) + Arena::pad((size_t)MHB_NUM_BUCKETS * 4 * 8) +
                Arena::pad(256 * 8) + Arena::pad(16 * 8) + 4096;
  {
    // A13: items that do not fit the device at once (or a caller-imposed cap) -> rounds over leading-byte ranges
    bool rounds = g_s2s_round_limit && n_items > g_s2s_round_limit;
    if (!rounds && need > g_arena.cap) {
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      rounds = (double)need > 0.92 * (double)(free_b + g_arena.cap);
    }
    if (rounds) return s2s_host_rounds(args, res, item_off, n_items, fixed, L0, n_words, g_s2s_round_limit);
  }
  CKR(g_arena.reserve(need));
  uint32_t *d_words = g_arena.take<uint32_t>(n_words + 16);
  uint64_t *d_word_off = g_arena.take<uint64_t>(ns + 1);
  uint64_t *d_item_off = g_arena.take<uint64_t>(ns + 1);
  uint32_t *d_len = g_arena.take<uint32_t>(ns + 1);
  uint16_t *d_mult = g_arena.take<uint16_t>(ns + 1);
  uint32_t *d_a = g_arena.take<uint32_t>((size_t)n_items * W + 4);
  uint32_t *d_b = g_arena.take<uint32_t>((size_t)n_items * W + 4);
  char *d_ws = g_arena.take<char>(ws_bytes);
  char *d_scratch = g_arena.take<char>(scratch_bytes);
  uint8_t *d_bytes = g_arena.take<uint8_t>(cap_bytes);
  uint64_t *d_table = g_arena.take<uint64_t>((size_t)MHB_NUM_BUCKETS * 4);
  uint64_t *d_hist0 = g_arena.take<uint64_t>(256);
  uint64_t *d_totals = g_arena.take<uint64_t>(16);

  if (ns) {
    CK(cudaMemcpyAsync(d_words, args->words, n_words * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_word_off, args->word_off, (ns + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_item_off, item_off.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_len, args->len, ns * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_mult, args->mult, ns * 2, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));

  mhb_dev_seqs seqs;
  seqs.words = d_words;
  seqs.n_words = n_words;
  seqs.n_seqs = ns;
  seqs.fixed_len = fixed ? L0 : 0;
  seqs.word_off = d_word_off;
  seqs.len = d_len;
  seqs.item_off = d_item_off;
  seqs.mult = d_mult;
  seqs.fixed_stride = 0;

  t.start();
  CKR(mhb_s2s_extract(st, &seqs, k, d_a, n_items, d_hist0, sort_bytes[0]));
  res->t_extract_ms = t.stop();
  t.start();
  int in_b = 0;
  CKR(mhb_sort_records_impl(st, d_a, d_b, n_items, W, sort_bytes, n_sort, d_hist0, d_ws, ws_bytes, &in_b, nullptr));
  res->t_sort_ms = t.stop();
  t.start();
  CKR(mhb_s2s_emit(st, in_b ? d_b : d_a, n_items, k, d_bytes, cap_bytes, d_table, d_totals, d_scratch, scratch_bytes));
  uint64_t totals[16];
  CK(cudaMemcpyAsync(totals, d_totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(res->bucket_table, d_table, sizeof(res->bucket_table), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  res->t_emit_ms = t.stop();
  res->n_bytes = totals[0];
  res->n_items = totals[1];
  res->n_tips = totals[2];
  res->n_large_mul = totals[3];
  for (int i = 0; i < 9; ++i) res->w_count[i] = totals[4 + i];
  res->ones_in_last = totals[13];
  if (res->n_bytes > cap_bytes) return mhb_set_error(MHB_ERR_NOMEM, "internal: SdBG byte stream exceeds capacity");
  res->bytes = (uint8_t *)malloc(std::max<size_t>(1, res->n_bytes));
  if (!res->bytes) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  if (res->n_bytes) CK(cudaMemcpy(res->bytes, d_bytes, res->n_bytes, cudaMemcpyDeviceToHost));
  res->t_total_ms = t_all.stop();
  return MHB_OK;
}

// ================================================================================================
// fused k_min build: count -> mercy edges -> seq2sdbg, device resident
// ================================================================================================
static int build_host_impl(const mhb_build_args *args, mhb_build_result *res, bool full_index);

// A13 for the fused build: when the records of the whole library do not fit next to it in HBM (or a round cap is set),
// the same graph is built stage by stage through the host-level calls that already work in rounds - count (rounds over
// bucket ranges) -> mercy edges -> seq2sdbg (rounds over bucket ranges) - with the solid edges passing through host
// memory once, as they do between the reference's two sub-commands (base_engine.cpp:54-141 plans its passes the same
// way: the output does not depend on where the boundaries fall).
static int build_host_rounds(const mhb_build_args *args, mhb_build_result *res) {
  memset(res, 0, sizeof(*res));
  const uint32_t k = args->k, WE = words_per_edge(k), NWE = div_ceil(k + 1, 16);
  mhb_count_args ca;
  memset(&ca, 0, sizeof(ca));
  ca.k = k;
  ca.m = args->m;
  ca.bin = args->bin;
  ca.bin_words = args->bin_words;
  ca.n_reads = args->n_reads;
  ca.want_mercy = args->need_mercy;
  std::vector<char> cbuf(sizeof(mhb_count_result));
  mhb_count_result *cr = reinterpret_cast<mhb_count_result *>(cbuf.data());
  CKR(mhb_count_host(&ca, cr));
  struct Owned {  // the count result's buffers, unless handed on to the caller
    mhb_count_result *r;
    ~Owned() {
      free(r->edges);
      free(r->cand_ids);
    }
  } owned{cr};
  res->n_edge_records = cr->n_edge_records;
  res->n_solid = cr->n_solid;
  res->n_cand = cr->n_cand;
  res->words_per_edge = WE;
  res->words_per_tip_label = words_per_tip_label(k);
  res->t_count_ms = cr->t_total_ms;
  uint32_t *mercy = nullptr;
  uint64_t n_mercy = 0;
  if (args->need_mercy && cr->n_cand) {  // the `.cand` image: candidate reads reversed (kmer_counter.cpp:387-401)
    std::vector<uint32_t> cand;
    uint64_t r = 0;
    size_t pos = 0;
    for (uint64_t c = 0; c < cr->n_cand; ++c) {
      while (r < cr->cand_ids[c]) {
        pos += 1 + div_ceil(args->bin[pos], 16);
        ++r;
      }
      const uint32_t L = args->bin[pos], nw = div_ceil(L, 16);
      const size_t at = cand.size();
      cand.resize(at + 1 + nw, 0u);
      cand[at] = L;
      for (uint32_t i = 0; i < L; ++i) cand[at + 1 + (i >> 4)] |= base_at(&args->bin[pos + 1], L - 1 - i) << (30 - 2 * (i & 15));
    }
    uint64_t n_cr = 0;
    CKR(mhb_mercy_host(k, cr->edges, cr->n_solid, cand.data(), cand.size(), &mercy, &n_mercy, &n_cr));
  }
  struct FreeMercy {
    uint32_t *&p;
    ~FreeMercy() { free(p); }
  } fm{mercy};
  res->n_mercy = n_mercy;
  // the edge (k+1)-mers as seq2sdbg loads them (seq_to_sdbg.cpp:424-450)
  const uint64_t n_seqs = cr->n_solid + n_mercy;
  std::vector<uint32_t> words((size_t)n_seqs * NWE + 1, 0u), len(n_seqs, k + 1);
  std::vector<uint64_t> word_off(n_seqs + 1);
  std::vector<uint16_t> mult(n_seqs);
  const uint32_t tail = (k + 1) % 16;
  for (uint64_t i = 0; i < n_seqs; ++i) {
    const uint32_t *e = i < cr->n_solid ? cr->edges + i * WE : mercy + (i - cr->n_solid) * WE;
    for (uint32_t j = 0; j < NWE; ++j) words[i * NWE + j] = e[j];
    if (tail) words[i * NWE + NWE - 1] &= top_mask(2 * tail);
    word_off[i] = i * NWE;
    mult[i] = (uint16_t)(e[WE - 1] & 0xFFFFu);
  }
  word_off[n_seqs] = n_seqs * NWE;
  mhb_s2s_args sa;
  memset(&sa, 0, sizeof(sa));
  sa.k = k;
  sa.words = words.data();
  sa.word_off = word_off.data();
  sa.len = len.data();
  sa.mult = mult.data();
  sa.n_seqs = n_seqs;
  std::vector<char> sbuf(sizeof(mhb_s2s_result));
  mhb_s2s_result *sr = reinterpret_cast<mhb_s2s_result *>(sbuf.data());
  CKR(mhb_s2s_host(&sa, sr));
  res->n_sort_items = sr->n_records;
  res->n_items = sr->n_items;
  res->n_tips = sr->n_tips;
  res->n_large_mul = sr->n_large_mul;
  res->n_bytes = sr->n_bytes;
  for (int i = 0; i < 9; ++i) res->w_count[i] = sr->w_count[i];
  res->ones_in_last = sr->ones_in_last;
  res->t_s2s_ms = sr->t_total_ms;
  res->bucket_table = (uint64_t *)malloc((size_t)MHB_NUM_BUCKETS * 32);
  if (!res->bucket_table) {
    free(sr->bytes);
    return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  }
  memcpy(res->bucket_table, sr->bucket_table, (size_t)MHB_NUM_BUCKETS * 32);
  if (args->sdbg_out && args->sdbg_out_capacity >= sr->n_bytes) {
    if (sr->n_bytes) memcpy(args->sdbg_out, sr->bytes, sr->n_bytes);
    free(sr->bytes);
    res->bytes = args->sdbg_out;
  } else {
    res->bytes = sr->bytes;
  }
  if (args->want_edges) {
    res->counting = (int64_t *)malloc(65536 * 8);
    if (!res->counting) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
    memcpy(res->counting, cr->counting, 65536 * 8);
    res->edges = cr->edges;
    res->cand_ids = cr->cand_ids;
    cr->edges = nullptr;
    cr->cand_ids = nullptr;
  }
  res->t_total_ms = res->t_count_ms + res->t_s2s_ms;
  return MHB_OK;
}

extern "C" int mhb_build_host(const mhb_build_args *args, mhb_build_result *res) {
  if (!args || !res) return mhb_set_error(MHB_ERR_ARG, "null args");
  if (g_round_limit || g_s2s_round_limit) return build_host_rounds(args, res);  // explicit caps (tests, small devices)
  const int rc = build_host_impl(args, res, false);
  if (rc == MHB_ERR_NOMEM) {  // everything resident does not fit: the staged build, whose stages plan their own rounds
    mhb_free(res->bytes == args->sdbg_out ? nullptr : res->bytes);
    mhb_free(res->bucket_table);
    mhb_free(res->edges);
    mhb_free(res->cand_ids);
    mhb_free(res->counting);
    mhb_release();
    return build_host_rounds(args, res);
  }
  return rc;
}
static int build_host_impl(const mhb_build_args *args, mhb_build_result *res, bool full_index) {
  memset(res, 0, sizeof(*res));
  const uint32_t k = args->k;
  if (k < 9 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "kmer size must be >= 9 and <= 255");
  if (args->need_mercy && k < 12) return mhb_set_error(MHB_ERR_ARG, "mercy edges need k >= 12");
  if (mhb_device_count() == 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  const uint32_t WR = count_record_words(k), WE = words_per_edge(k), W2 = s2s_record_words(k), WPT = words_per_tip_label(k);
  res->words_per_edge = WE;
  res->words_per_tip_label = WPT;

  BinIndex ix;
  CKR(index_bin(args->bin, args->bin_words, args->n_reads, k, &ix, !full_index));
  const bool verify_fixed = !full_index && ix.fixed_len != 0;  // the device checks every length word (below)
  const uint64_t n = ix.n_edges, n_reads = args->n_reads;
  res->n_edge_records = n;
  uint32_t max_len = ix.fixed_len;
  if (!ix.fixed_len)
    for (uint64_t r = 0; r < n_reads; ++r) max_len = std::max(max_len, args->bin[ix.rec_off[r]]);

  cudaStream_t st = 0;
  Timer t_all(st), t(st);
  t_all.start();

  uint8_t cbytes[72], sbytes[72];
  const uint32_t n_csort = mhb_count_sort_bytes(k, cbytes), n_ssort = mhb_s2s_sort_bytes(k, sbytes);
  const int32_t m = args->m;
  const uint64_t cap_edges = n / (uint64_t)std::max(1, m) + 1;
  const size_t bin_bytes = (args->bin_words * 4 + 15) & ~(size_t)15;
  const CountWork cw = count_work_plan(n, k, m);
  const size_t count_work = 2 * Arena::pad((size_t)n * WR * 4 + 16) + Arena::pad(cw.bytes);
  // fixed part
  size_t fixed = Arena::pad(bin_bytes + 16) + Arena::pad((size_t)cap_edges * WE * 4) + Arena::pad(cap_edges) +
                 Arena::pad(65536 * 8) + 2 * Arena::pad(256 * 8) + Arena::pad(64) + Arena::pad((size_t)MHB_NUM_BUCKETS * 32) +
                 Arena::pad(128) + 8192;
  if (!ix.fixed_len) fixed += 2 * Arena::pad((n_reads + 1) * 8);
  if (args->need_mercy) fixed += 2 * Arena::pad((n_reads + 1) * 4) + Arena::pad((n_reads + 1) * 8);
  CKR(g_arena.reserve(fixed + count_work));

  uint32_t *d_bin = g_arena.take<uint32_t>(bin_bytes / 4 + 4);
  uint32_t *d_edges = g_arena.take<uint32_t>((size_t)cap_edges * WE);
  uint8_t *d_aux = g_arena.take<uint8_t>(cap_edges);
  uint64_t *d_mul_hist = g_arena.take<uint64_t>(65536);
  uint64_t *d_hist0 = g_arena.take<uint64_t>(256);
  uint64_t *d_hist1 = g_arena.take<uint64_t>(256);
  uint64_t *d_nsolid = g_arena.take<uint64_t>(8);
  uint64_t *d_table = g_arena.take<uint64_t>((size_t)MHB_NUM_BUCKETS * 4);
  uint64_t *d_totals = g_arena.take<uint64_t>(16);
  uint64_t *d_rec_off = nullptr, *d_edge_off = nullptr, *d_cand = nullptr;
  uint32_t *d_first = nullptr, *d_last = nullptr;
  if (!ix.fixed_len) {
    d_rec_off = g_arena.take<uint64_t>(n_reads + 1);
    d_edge_off = g_arena.take<uint64_t>(n_reads + 1);
  }
  if (args->need_mercy) {
    d_first = g_arena.take<uint32_t>(n_reads + 1);
    d_last = g_arena.take<uint32_t>(n_reads + 1);
    d_cand = g_arena.take<uint64_t>(n_reads + 1);
  }
  char *work = g_arena.take<char>(count_work);
  size_t work_bytes = count_work;
  char *extra = nullptr;  // separately allocated work area when the SdBG stage outgrows the count stage's
  char *big_edges = nullptr;  // solid + mercy edges when the mercy edges do not fit behind the solid ones in d_edges
  struct ExtraGuard {
    char *&p;
    ~ExtraGuard() {
      if (p) cudaFree(p);
    }
  } guard{extra}, guard2{big_edges};
  uint32_t *d_all_edges = d_edges;

  // ---- H2D ----
  // Fixed-length libraries are uploaded in C pieces on a copy stream and the edges of piece i are extracted while piece
  // i+1 is still crossing PCIe, instead of upload-then-extract (C = 4; MHB_H2D_CHUNKS=1 restores the single copy;
  // measured e2e 115.7 -> 112.5 ms on the bench workload, profiles/r2a_bench_chunks.json).
  static const int h2d_chunks_env = getenv("MHB_H2D_CHUNKS") ? atoi(getenv("MHB_H2D_CHUNKS")) : 4;
  const bool chunked = h2d_chunks_env > 1 && ix.fixed_len >= k + 1 && n_reads >= (uint64_t)h2d_chunks_env * 64;
  t.start();
  CK(cudaMemsetAsync(d_mul_hist, 0, 65536 * 8, st));
  CK(cudaMemsetAsync(d_hist0, 0, 256 * 8, st));
  CK(cudaMemsetAsync(d_hist1, 0, 256 * 8, st));
  CK(cudaMemsetAsync(d_nsolid, 0, 64, st));
  if (!chunked) {
    if (args->bin_words) CK(cudaMemcpyAsync(d_bin, args->bin, args->bin_words * 4, cudaMemcpyHostToDevice, st));
    if (!ix.fixed_len && n_reads) {
      CK(cudaMemcpyAsync(d_rec_off, ix.rec_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(d_edge_off, ix.edge_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    }
  }
  res->t_h2d_ms = t.stop();

  mhb_dev_reads reads;
  reads.bin = d_bin;
  reads.bin_words = args->bin_words;
  reads.n_reads = n_reads;
  reads.fixed_len = ix.fixed_len;
  reads.rec_off = d_rec_off;
  reads.edge_off = d_edge_off;

  // ---- count stage ----
  t.start();
  uint32_t *c_a = (uint32_t *)work;
  uint32_t *c_b = (uint32_t *)(work + Arena::pad((size_t)n * WR * 4 + 16));
  char *c_wsp = work + 2 * Arena::pad((size_t)n * WR * 4 + 16);
  if (!chunked) {
    CKR(mhb_count_extract(st, &reads, k, c_a, n, d_hist0, cw.hist_byte));
  } else {
    static cudaStream_t copy_st = nullptr;
    static cudaEvent_t ev[64];
    static bool ev_ready = false;
    if (!ev_ready) {
      CK(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
      for (int i = 0; i < 64; ++i) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
      ev_ready = true;
    }
    const int C = std::min(h2d_chunks_env, 60);  // ev[63] is the start marker
    const uint64_t stride = 1 + div_ceil(ix.fixed_len, 16);          // words per read
    const uint64_t per = ((n_reads + C - 1) / C + 3) & ~(uint64_t)3;  // reads per piece, multiple of 4 -> 16-byte aligned
    const uint64_t e_per_read = ix.fixed_len - k;
    cudaEvent_t start_ev = ev[63];
    CK(cudaEventRecord(start_ev, st));  // the copy stream must not run ahead of this call's place in `st`
    CK(cudaStreamWaitEvent(copy_st, start_ev, 0));
    int c = 0;
    for (uint64_t r0 = 0; r0 < n_reads; r0 += per, ++c) {
      const uint64_t r1 = std::min(n_reads, r0 + per);
      const uint64_t w0 = r0 * stride, w1 = r1 * stride;
      CK(cudaMemcpyAsync(d_bin + w0, args->bin + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, copy_st));
      CK(cudaEventRecord(ev[c], copy_st));
      CK(cudaStreamWaitEvent(st, ev[c], 0));
      mhb_dev_reads piece = reads;
      piece.bin = d_bin + w0;
      piece.bin_words = w1 - w0;
      piece.n_reads = r1 - r0;
      CKR(mhb_count_extract(st, &piece, k, c_a + (size_t)r0 * e_per_read * WR, (r1 - r0) * e_per_read, d_hist0, cw.hist_byte));
    }
  }
  if (verify_fixed) CKR(mhb_check_fixed_len(st, d_bin, n_reads, ix.fixed_len, d_nsolid + 1));
  CKR(run_count_stage(st, cw, c_a, c_b, n, k, m, d_hist0, d_edges, d_aux, cap_edges, d_mul_hist, d_nsolid, c_wsp, nullptr, nullptr));
  uint64_t h_scal[2] = {0, 0};
  CK(cudaMemcpyAsync(h_scal, d_nsolid, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (verify_fixed && h_scal[1]) return build_host_impl(args, res, true);  // not fixed-length after all: indexed path
  const uint64_t n_solid = h_scal[0];
  if (n_solid > cap_edges) return mhb_set_error(MHB_ERR_NOMEM, "internal: solid edges exceed capacity");
  res->n_solid = n_solid;
  res->t_count_ms = t.stop();

  // ---- mercy: per-read marks -> candidates -> mercy edges appended to the solid edges ----
  uint64_t n_cand = 0, n_mercy = 0;
  if (args->need_mercy && n_reads) {
    t.start();
    uint64_t n_tip = 0;
    CKR(mhb_count_tip_edges(st, d_aux, n_solid, &n_tip));
    const size_t ts_bytes = mhb_tipset_bytes(n_tip, k);
    const size_t cs_bytes = mhb_mercy_candidates_scratch_bytes(n_reads);
    // the count stage's work area is free now; small inputs may need more than it offers (12-mer look-up table)
    auto work_area = [&](size_t bytes) -> char * {
      if (bytes <= work_bytes) return work;
      if (extra) cudaFree(extra);
      extra = nullptr;
      if (cudaMalloc((void **)&extra, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return extra;
    };
    char *d_tips = work_area(Arena::pad(ts_bytes) + Arena::pad(cs_bytes));
    if (!d_tips) return mhb_set_error(MHB_ERR_NOMEM, "cudaMalloc for the mercy stage failed");
    char *d_cs = d_tips + Arena::pad(ts_bytes);
    CKR(mhb_tipset_build(st, d_edges, d_aux, n_solid, k, d_tips, ts_bytes, n_tip));
    CKR(mhb_count_mark_mercy(st, &reads, k, d_tips, ts_bytes, n_tip, d_first, d_last));
    CKR(mhb_mercy_candidates(st, d_first, d_last, n_reads, d_cand, &n_cand, d_cs, cs_bytes));
    if (n_cand) {
      const size_t ms_bytes = mhb_mercy_edges_scratch_bytes(n_cand, max_len);
      CK(cudaStreamSynchronize(st));  // the tip set / candidate scratch may be released by work_area()
      char *d_ms = work_area(ms_bytes);
      if (!d_ms) return mhb_set_error(MHB_ERR_NOMEM, "cudaMalloc for the mercy stage failed");
      // count first (probe + per-read counts + scan), then size the destination: reads that overlap only at their ends
      // can put more mercy edges between two tips than n/m + 1 - n_solid (the reference reserves +25 % and grows,
      // seq_to_sdbg.cpp:371-379; here the exact number is known before anything is written)
      const size_t core = ms_bytes - mhb_edge_lut_bytes();
      void *lut = d_ms + core;
      CKR(mhb_edge_lut_build(st, d_edges, n_solid, k, lut));
      const uint32_t *seg_e[1] = {d_edges};
      const uint64_t seg_n[1] = {n_solid};
      const void *seg_l[1] = {lut};
      CKR(mhb_mercy_edges_count(st, &reads, d_cand, n_cand, max_len, k, 1, seg_e, seg_n, seg_l, nullptr, &n_mercy, d_ms, core));
      if (n_mercy > cap_edges - n_solid) {
        const size_t eb = Arena::pad((size_t)(n_solid + n_mercy) * WE * 4 + 16);
        if (cudaMalloc((void **)&big_edges, eb) != cudaSuccess) {
          cudaGetLastError();
          return mhb_set_error(MHB_ERR_NOMEM, "cudaMalloc of %zu bytes for solid + mercy edges failed", eb);
        }
        CK(cudaMemcpyAsync(big_edges, d_edges, (size_t)n_solid * WE * 4, cudaMemcpyDeviceToDevice, st));
        d_all_edges = (uint32_t *)big_edges;
      }
      CKR(mhb_mercy_edges_write(st, &reads, d_cand, n_cand, max_len, k, d_all_edges + (size_t)n_solid * WE, n_mercy, n_mercy,
                                d_ms, core));
    }
    res->t_mercy_ms = t.stop();
  }
  res->n_cand = n_cand;
  res->n_mercy = n_mercy;

  // ---- SdBG stage over solid + mercy edges, straight from the device-resident edge records ----
  t.start();
  const uint64_t n_seqs = n_solid + n_mercy;
  const uint64_t n_items = n_seqs * 6;  // 2 strands x (k+1 - k + 2)
  res->n_sort_items = n_items;
  const size_t s_ws = mhb_sort_workspace_bytes(n_items, W2), s_scr = mhb_s2s_emit_scratch_bytes(n_items, k);
  const uint64_t cap_bytes = n_items * (4ull + 4ull * WPT) + 16;
  const size_t s2s_work = 2 * Arena::pad((size_t)n_items * W2 * 4 + 16) + Arena::pad(s_ws) + Arena::pad(s_scr) + Arena::pad(cap_bytes);
  char *sw = work;
  if (s2s_work > work_bytes) {
    CK(cudaStreamSynchronize(st));
    if (extra) cudaFree(extra);
    extra = nullptr;
    cudaError_t e = cudaMalloc((void **)&extra, s2s_work);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return mhb_set_error(MHB_ERR_NOMEM, "cudaMalloc of %zu bytes for the SdBG stage failed", s2s_work);
    }
    sw = extra;
  }
  uint32_t *s_a = (uint32_t *)sw;
  uint32_t *s_b = (uint32_t *)(sw + Arena::pad((size_t)n_items * W2 * 4 + 16));
  char *s_wsp = sw + 2 * Arena::pad((size_t)n_items * W2 * 4 + 16);
  char *s_scrp = s_wsp + Arena::pad(s_ws);
  uint8_t *d_bytes = (uint8_t *)(s_scrp + Arena::pad(s_scr));
  mhb_dev_seqs seqs;
  memset(&seqs, 0, sizeof(seqs));
  seqs.words = d_all_edges;
  seqs.n_words = n_seqs * WE;
  seqs.n_seqs = n_seqs;
  seqs.fixed_len = k + 1;
  seqs.fixed_stride = WE;
  // the solid edges still have the count stage's in/out flags: the $-items the emitter would discard for certain are
  // not generated (mhb_s2s_extract_edges_pruned; 2.02 instead of 6 items per edge on a 30x genome), same bytes out
  static const bool no_prune = getenv("MHB_S2S_NO_PRUNE") != nullptr;
  uint64_t n_sorted = n_items;
  if (!no_prune) {
    CK(cudaMemsetAsync(d_nsolid + 4, 0, 8, st));
    CKR(mhb_s2s_extract_edges_pruned(st, d_all_edges, d_aux, n_seqs, n_solid, k, s_a, n_items, d_nsolid + 4, d_hist1, sbytes[0]));
    CK(cudaMemcpyAsync(&n_sorted, d_nsolid + 4, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_sorted > n_items) return mhb_set_error(MHB_ERR_CUDA, "internal: pruned item count exceeds 6 per edge");
    res->n_sort_items = n_sorted;
  } else {
    CKR(mhb_s2s_extract(st, &seqs, k, s_a, n_items, d_hist1, sbytes[0]));
  }
  int s_in_b = 0;
  CKR(mhb_sort_records_impl(st, s_a, s_b, n_sorted, W2, sbytes, n_ssort, d_hist1, s_wsp, s_ws, &s_in_b, nullptr));
  CKR(mhb_s2s_emit(st, s_in_b ? s_b : s_a, n_sorted, k, d_bytes, cap_bytes, d_table, d_totals, s_scrp, s_scr));
  uint64_t totals[16];
  CK(cudaMemcpyAsync(totals, d_totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  res->t_s2s_ms = t.stop();
  res->n_bytes = totals[0];
  res->n_items = totals[1];
  res->n_tips = totals[2];
  res->n_large_mul = totals[3];
  for (int i = 0; i < 9; ++i) res->w_count[i] = totals[4 + i];
  res->ones_in_last = totals[13];
  if (res->n_bytes > cap_bytes) return mhb_set_error(MHB_ERR_NOMEM, "internal: SdBG byte stream exceeds capacity");

  // ---- D2H ----
  t.start();
  res->bucket_table = (uint64_t *)malloc((size_t)MHB_NUM_BUCKETS * 32);
  if (args->sdbg_out && args->sdbg_out_capacity >= res->n_bytes) res->bytes = args->sdbg_out;
  else res->bytes = (uint8_t *)malloc(std::max<size_t>(1, res->n_bytes));
  if (!res->bucket_table || !res->bytes) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  CK(cudaMemcpyAsync(res->bucket_table, d_table, (size_t)MHB_NUM_BUCKETS * 32, cudaMemcpyDeviceToHost, st));
  if (res->n_bytes) CK(cudaMemcpyAsync(res->bytes, d_bytes, res->n_bytes, cudaMemcpyDeviceToHost, st));
  if (args->want_edges) {
    res->edges = (uint32_t *)malloc(std::max<size_t>(1, (size_t)n_solid * WE * 4));
    res->cand_ids = (uint64_t *)malloc(std::max<size_t>(1, n_cand * 8));
    res->counting = (int64_t *)malloc(65536 * 8);
    if (!res->edges || !res->cand_ids || !res->counting) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
    if (n_solid) CK(cudaMemcpyAsync(res->edges, d_edges, (size_t)n_solid * WE * 4, cudaMemcpyDeviceToHost, st));
    if (n_cand) CK(cudaMemcpyAsync(res->cand_ids, d_cand, n_cand * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(res->counting, d_mul_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  res->t_d2h_ms = t.stop();
  res->t_total_ms = t_all.stop();
  return MHB_OK;
}

// ================================================================================================
// mercy edges from host buffers (what `seq2sdbg --need_mercy` needs between reading `.edges`/`.cand` and SeqToSdbg::Run)
// ================================================================================================
extern "C" int mhb_mercy_host(uint32_t k, const uint32_t *edges, uint64_t n_edges, const uint32_t *cand_bin,
                              uint64_t cand_words, uint32_t **mercy_out, uint64_t *n_mercy_out, uint64_t *n_cand_reads_out) {
  if (!mercy_out || !n_mercy_out) return mhb_set_error(MHB_ERR_ARG, "null output");
  *mercy_out = nullptr;
  *n_mercy_out = 0;
  if (n_cand_reads_out) *n_cand_reads_out = 0;
  if (k < 12 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "mercy edges need 12 <= k <= 255");
  if (mhb_device_count() == 0) return mhb_set_error(MHB_ERR_CUDA, "no CUDA device: libmhb has no CPU path");
  const uint32_t WE = words_per_edge(k);
  // `.cand` holds the reads as KmerCounter held them: REVERSED (kmer_counter.cpp:387-401; read back with reverse=false,
  // seq_to_sdbg.cpp:175-176).  The device kernels take a library in file orientation and apply the reversal themselves,
  // so every candidate read is turned around once here (they are ~0.2 % of a library).
  std::vector<uint32_t> bin;
  std::vector<uint64_t> rec_off, edge_off;
  bin.reserve(cand_words + 8);
  uint32_t max_len = 0;
  uint64_t pos = 0, e = 0;
  while (pos < cand_words) {
    const uint32_t L = cand_bin[pos], nw = div_ceil(L, 16);
    if (pos + 1 + nw > cand_words) return mhb_set_error(MHB_ERR_IO, "candidate read image is truncated");
    rec_off.push_back(bin.size());
    edge_off.push_back(e);
    if (L >= k + 1) e += L - k;
    const size_t at = bin.size();
    bin.resize(at + 1 + nw, 0);
    bin[at] = L;
    for (uint32_t i = 0; i < L; ++i)
      bin[at + 1 + (i >> 4)] |= base_at(cand_bin + pos + 1, L - 1 - i) << (30 - 2 * (i & 15));
    max_len = std::max(max_len, L);
    pos += 1 + nw;
  }
  const uint64_t n_reads = rec_off.size();
  rec_off.push_back(bin.size());
  edge_off.push_back(e);
  if (n_cand_reads_out) *n_cand_reads_out = n_reads;
  if (n_reads == 0 || n_edges == 0) {
    *mercy_out = (uint32_t *)malloc(4);
    return MHB_OK;
  }
  const size_t bin_bytes = (bin.size() * 4 + 15) & ~(size_t)15;
  const size_t ms_bytes = mhb_mercy_edges_scratch_bytes(n_reads, max_len);
  const size_t need = Arena::pad(bin_bytes + 16) + 3 * Arena::pad((n_reads + 1) * 8) + Arena::pad((size_t)n_edges * WE * 4 + 16) +
                      Arena::pad(ms_bytes) + 4096;
  CKR(g_arena.reserve(need));
  cudaStream_t st = 0;
  uint32_t *d_bin = g_arena.take<uint32_t>(bin_bytes / 4 + 4);
  uint64_t *d_rec_off = g_arena.take<uint64_t>(n_reads + 1);
  uint64_t *d_edge_off = g_arena.take<uint64_t>(n_reads + 1);
  uint64_t *d_ids = g_arena.take<uint64_t>(n_reads + 1);
  uint32_t *d_edges = g_arena.take<uint32_t>((size_t)n_edges * WE + 4);
  char *d_ms = g_arena.take<char>(ms_bytes);
  std::vector<uint64_t> ids(n_reads);
  for (uint64_t r = 0; r < n_reads; ++r) ids[r] = r;
  CK(cudaMemcpyAsync(d_bin, bin.data(), bin.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rec_off, rec_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_edge_off, edge_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_ids, ids.data(), n_reads * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_edges, edges, (size_t)n_edges * WE * 4, cudaMemcpyHostToDevice, st));
  mhb_dev_reads reads;
  reads.bin = d_bin;
  reads.bin_words = bin.size();
  reads.n_reads = n_reads;
  reads.fixed_len = 0;
  reads.rec_off = d_rec_off;
  reads.edge_off = d_edge_off;
  const size_t core = ms_bytes - mhb_edge_lut_bytes();
  void *lut = d_ms + core;
  CKR(mhb_edge_lut_build(st, d_edges, n_edges, k, lut));
  const uint32_t *seg_e[1] = {d_edges};
  const uint64_t seg_n[1] = {n_edges};
  const void *seg_l[1] = {lut};
  uint64_t n_mercy = 0;
  CKR(mhb_mercy_edges_count(st, &reads, d_ids, n_reads, max_len, k, 1, seg_e, seg_n, seg_l, nullptr, &n_mercy, d_ms, core));
  *mercy_out = (uint32_t *)malloc(std::max<size_t>(4, (size_t)n_mercy * WE * 4));
  if (!*mercy_out) return mhb_set_error(MHB_ERR_NOMEM, "host malloc failed");
  if (n_mercy) {
    uint32_t *d_out = nullptr;
    CK(cudaMalloc((void **)&d_out, (size_t)n_mercy * WE * 4));
    int rc = mhb_mercy_edges_write(st, &reads, d_ids, n_reads, max_len, k, d_out, n_mercy, n_mercy, d_ms, core);
    if (!rc && cudaMemcpyAsync(*mercy_out, d_out, (size_t)n_mercy * WE * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)
      rc = mhb_set_error(MHB_ERR_CUDA, "mercy edge download failed");
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess)
      rc = mhb_set_error(MHB_ERR_CUDA, "mercy edge kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_out);
    if (rc) {
      free(*mercy_out);
      *mercy_out = nullptr;
      return rc;
    }
  }
  *n_mercy_out = n_mercy;
  return MHB_OK;
}

// ================================================================================================
// Self-test hooks: run the SAME record builders the kernels use (mhb_kernels.cuh, __host__ __device__)
// on the host, so that `pytest -m "not gpu"` can check the bit arithmetic against the oracle without a
// GPU.  They build one record at a time and are not a compute path.
// ================================================================================================
#include "mhb_kernels.cuh"

extern "C" int mhb_selftest_count_record(const uint32_t *read_words, uint32_t nwords, uint32_t L, uint32_t k, uint32_t q,
                                         uint32_t *rec_out, uint32_t *strand_out) {
  const uint32_t W = count_key_words(k), WR = count_record_words(k);
  if (k < 1 || k > MHB_MAX_K || L < k + 1 || q + k + 1 > L) return mhb_set_error(MHB_ERR_ARG, "bad selftest args");
#define M(WW)                                                                       \
  if (W == WW && WR == WW) {                                                        \
    uint32_t r[WW];                                                                 \
    make_count_record<WW, WW>(read_words, nwords, L, k, q, r, *strand_out);         \
    memcpy(rec_out, r, sizeof(r));                                                  \
    return MHB_OK;                                                                  \
  }                                                                                 \
  if (W == WW && WR == WW + 1) {                                                    \
    uint32_t r[WW + 1];                                                             \
    make_count_record<WW, WW + 1>(read_words, nwords, L, k, q, r, *strand_out);     \
    memcpy(rec_out, r, sizeof(r));                                                  \
    return MHB_OK;                                                                  \
  }
  M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16)
#undef M
  return mhb_set_error(MHB_ERR_ARG, "unsupported k");
}

// the rolling builder (mhb_kernels.cuh make_count_records_roll) run on the host: 4 records from position q on
extern "C" int mhb_selftest_count_records_roll(const uint32_t *read_words, uint32_t nwords, uint32_t L, uint32_t k, uint32_t q,
                                               uint64_t *rec4_out, uint32_t *strand4_out) {
  if (k + 1 < 17 || k + 1 > 32 || L < k + 1 || q + k + 1 > L) return mhb_set_error(MHB_ERR_ARG, "bad selftest args");
  u64 r[4];
  u32 st[4];
  make_count_records_roll<4>(read_words, nwords, L, k, q, r, st);
  for (int j = 0; j < 4; ++j) {
    rec4_out[j] = r[j];
    strand4_out[j] = st[j];
  }
  return MHB_OK;
}

extern "C" int mhb_selftest_s2s_record(const uint32_t *seq_words, uint32_t nwords, uint32_t L, uint32_t k, uint32_t strand,
                                       uint32_t offset, uint32_t mult, uint32_t *rec_out) {
  const uint32_t W = s2s_record_words(k);
  if (k < 9 || k > MHB_MAX_K || L < k + 1 || offset > L - k + 1) return mhb_set_error(MHB_ERR_ARG, "bad selftest args");
#define M(WW)                                                                \
  if (W == WW) {                                                             \
    uint32_t r[WW];                                                          \
    make_s2s_record<WW>(seq_words, nwords, L, k, strand, offset, mult, r);   \
    memcpy(rec_out, r, sizeof(r));                                           \
    return MHB_OK;                                                           \
  }
  M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16) M(17)
#undef M
  return mhb_set_error(MHB_ERR_ARG, "unsupported k");
}
