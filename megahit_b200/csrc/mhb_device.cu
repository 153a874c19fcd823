// mhb_device.cu -- device-level C ABI (see include/mhb.h, layer 1): kernel launches on caller-owned
// device memory.  Built for sm_100a only; there is no host fallback.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mhb.h"
#include "mhb_count.cuh"
#include "mhb_internal.h"
#include "mhb_mercy.cuh"
#include "mhb_s2s.cuh"
#include "mhb_sort.cuh"
#include "mhb_sort3.cuh"

using namespace mhb;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
int mhb_set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char *mhb_last_error(void) { return g_err; }
extern "C" const char *mhb_version(void) { return "megahit_b200 0.1 (sm_100a; formats of megahit v1.2.9)"; }
extern "C" int mhb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" void mhb_free(void *p) { free(p); }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return mhb_set_error(MHB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                           cudaGetErrorString(e_));                                                \
  } while (0)
#define CK_LAUNCH() CK(cudaGetLastError())

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
extern "C" uint32_t mhb_count_record_words(uint32_t k) { return count_record_words(k); }
extern "C" uint32_t mhb_words_per_edge(uint32_t k) { return words_per_edge(k); }
extern "C" uint32_t mhb_s2s_record_words(uint32_t k) { return s2s_record_words(k); }

extern "C" uint32_t mhb_count_sort_bytes(uint32_t k, uint8_t *bytes) {
  // key = top 2(k+1) bits of the record: every byte that holds at least one key bit
  const u32 wr = count_record_words(k), total_bits = 32 * wr, key_bits = 2 * (k + 1);
  const u32 lo = (total_bits - key_bits) / 8;
  u32 n = 0;
  for (u32 b = lo; b < 4 * wr; ++b) bytes[n++] = (uint8_t)b;
  return n;
}
extern "C" uint32_t mhb_s2s_sort_bytes(uint32_t k, uint8_t *bytes) {
  // The reference sorts the whole record (seq_to_sdbg.cpp: no payload words).  The low 16 bits
  // (65535 - multiplicity) only decide which of several records with identical bases and flags comes first, and
  // the only use of that is "the first record of an (a,b) run carries the largest multiplicity" (:778-785); the
  // emit kernel takes the run's minimum instead, so those two bytes are not sorted.  Byte 2 holds the flag bits
  // 16..19 (prev char, non-dollar); all-zero bytes between the flags and the k-mer are skipped.
  const u32 w = s2s_record_words(k), total_bits = 32 * w, key_bits = 2 * k;
  const u32 lo = (total_bits - key_bits) / 8;
  u32 n = 0;
  for (u32 b = 2; b < 4 * w; ++b)
    if (b == 2 || b >= lo) bytes[n++] = (uint8_t)b;
  return n;
}

// ------------------------------------------------------------------------------------------------
// dispatch helpers
// ------------------------------------------------------------------------------------------------
#define MHB_FOR_W(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16)
#define MHB_FOR_WR(M) MHB_FOR_W(M) M(17)

static int g_sm_count = 0;
static int g_bound_device = -1;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
    if (g_bound_device < 0) g_bound_device = dev;  // first compute call: the process stays on this device
  }
  return g_sm_count;
}

static ReadsView make_reads_view(const mhb_dev_reads *r) {
  ReadsView v;
  v.bin = r->bin;
  v.bin_words = r->bin_words;
  v.n_reads = r->n_reads;
  v.fixed_len = r->fixed_len;
  v.fixed_stride = r->fixed_len ? 1 + div_ceil(r->fixed_len, 16) : 0;
  v.rec_off = r->rec_off;
  v.edge_off = r->edge_off;
  return v;
}

static int check_reads(const mhb_dev_reads *r, uint32_t k) {
  if (!r || k < 1 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "bad reads/k (k=%u)", k);
  if (r->n_reads && !r->bin) return mhb_set_error(MHB_ERR_ARG, "reads->bin is NULL");
  if (((uintptr_t)r->bin & 15) != 0) return mhb_set_error(MHB_ERR_ARG, "reads->bin must be 16-byte aligned");
  if (!r->fixed_len && r->n_reads && (!r->rec_off || !r->edge_off))
    return mhb_set_error(MHB_ERR_ARG, "variable-length reads need rec_off and edge_off");
  return MHB_OK;
}

// grow-only per-device scratch for block sums of scans issued from entry points that take no scratch argument
static int grow_scratch(size_t bytes, void **out) {
  static void *buf[64] = {nullptr};
  static size_t cap[64] = {0};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return mhb_set_error(MHB_ERR_CUDA, "device index %d out of range", dev);
  if (cap[dev] < bytes) {
    CK(cudaDeviceSynchronize());
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    cap[dev] = 0;
    const size_t want = (bytes + 4095) & ~(size_t)4095;
    CK(cudaMalloc(&buf[dev], want));
    cap[dev] = want;
  }
  *out = buf[dev];
  return MHB_OK;
}
static int scan64(cudaStream_t st, u64 *v, u64 n, u64 *total_dev, u64 *bsum);

// ------------------------------------------------------------------------------------------------
// count: extract
// ------------------------------------------------------------------------------------------------
extern "C" int mhb_count_extract(void *stream, const mhb_dev_reads *reads, uint32_t k, uint32_t *records,
                                 uint64_t n_edges, uint64_t *hist256, int hist_byte) {
  if (int rc = check_reads(reads, k)) return rc;
  if (reads->n_reads == 0 || n_edges == 0) return MHB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const ReadsView rv = make_reads_view(reads);
  const u32 W = count_key_words(k), WR = count_record_words(k);
  const u64 n_batches = (rv.n_reads + kReadsPerBatch - 1) / kReadsPerBatch;
  const int grid = (int)(n_batches < (u64)(sm_count() * 8) ? n_batches : (u64)(sm_count() * 8));
  static const bool roll = getenv("MHB_EXTRACT_ROLL") != nullptr;  // opt-in rolling builder (8-byte records)
  if (roll && W == 2 && WR == 2 && k + 1 >= 17) {
    k_count_extract_roll<<<grid, kExtractThreads, 0, st>>>(rv, k, records, hist256, hist_byte);
    CK_LAUNCH();
    return MHB_OK;
  }
#define M(WW)                                                                                              \
  if (W == WW && WR == WW)                                                                                 \
    k_count_extract<WW, WW><<<grid, kExtractThreads, 0, st>>>(rv, k, records, hist256, hist_byte);         \
  else if (W == WW && WR == WW + 1)                                                                        \
    k_count_extract<WW, WW + 1><<<grid, kExtractThreads, 0, st>>>(rv, k, records, hist256, hist_byte);     \
  else
  MHB_FOR_W(M) return mhb_set_error(MHB_ERR_ARG, "unsupported k=%u", k);
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

// A13: one round of the out-of-core count stage.  write == 0: per_read[r] <- in-range edge count of read r, then an
// in-place exclusive scan (per_read[n_reads] and *total_dev <- total); hist256 (optional) += histogram of byte
// hist_byte over the in-range records.  write != 0: per_read holds the scanned offsets; the in-range records go to
// records[per_read[r] ...) in read order and hist256 += histogram of hist_byte (the first sort digit).
extern "C" int mhb_count_extract_range(void *stream, const mhb_dev_reads *reads, uint32_t k, uint32_t lo, uint32_t hi,
                                       int write, uint64_t *per_read, uint32_t *records, uint64_t *hist256, int hist_byte,
                                       uint64_t *total_dev) {
  if (int rc = check_reads(reads, k)) return rc;
  if (lo > hi || hi > 65535) return mhb_set_error(MHB_ERR_ARG, "bad bucket range [%u, %u]", lo, hi);
  if (!per_read || (write && !records) || (!write && !total_dev)) return mhb_set_error(MHB_ERR_ARG, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (reads->n_reads == 0) {
    if (!write) CK(cudaMemsetAsync(total_dev, 0, 8, st));
    return MHB_OK;
  }
  const ReadsView rv = make_reads_view(reads);
  const u32 W = count_key_words(k), WR = count_record_words(k);
  const u64 n_batches = (rv.n_reads + kReadsPerBatch - 1) / kReadsPerBatch;
  const int grid = (int)(n_batches < (u64)(sm_count() * 8) ? n_batches : (u64)(sm_count() * 8));
#define M2(WW, WRR)                                                                                                   \
  if (W == WW && WR == WRR) {                                                                                         \
    if (write)                                                                                                        \
      k_count_extract_range<WW, WRR, true><<<grid, kExtractThreads, 0, st>>>(rv, k, lo, hi, per_read, records, hist256, hist_byte); \
    else                                                                                                              \
      k_count_extract_range<WW, WRR, false><<<grid, kExtractThreads, 0, st>>>(rv, k, lo, hi, per_read, records, hist256, hist_byte); \
  } else
#define M(WW) M2(WW, WW) M2(WW, WW + 1)
  MHB_FOR_W(M) return mhb_set_error(MHB_ERR_ARG, "unsupported k=%u", k);
#undef M
#undef M2
  CK_LAUNCH();
  if (!write) {
    // three-phase scan (the block sums live in this device's small scratch): a single-CTA scan over all reads would
    // dominate a library that needs many rounds
    u64 *bsum = nullptr;
    if (int rc = grow_scratch(((rv.n_reads / kScanTile) + 2) * 8, (void **)&bsum)) return rc;
    if (int rc = scan64(st, per_read, rv.n_reads, total_dev, bsum)) return rc;
    CK(cudaMemcpyAsync(per_read + rv.n_reads, total_dev, 8, cudaMemcpyDeviceToDevice, st));
  }
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// sort
// ------------------------------------------------------------------------------------------------
// Radix-pass variants.  0..3 = v2 geometries (mhb_sort.cuh); 256 + bits = v3 (mhb_sort3.cuh, see SortCfg3 for the
// bit field).  Only the listed v3 combinations are instantiated (all for 8- and 12-byte records, the first one for
// every record width).
#define MHB_V3_DEFAULT 0x080
#define MHB_V3_LIST(X)                                                                                               \
  X(0x080) X(0x000) X(0x009) X(0x082) X(0x180) X(0x480) X(0x084) X(0x1080) X(0x0888) X(0x8080) X(0x10080) X(0x18080)  \
  X(0x9080) X(0x8082) X(0x10082)
static bool v3_listed(int bits) {
#define X(B) \
  if (bits == B) return true;
  MHB_V3_LIST(X)
#undef X
  return false;
}
static int g_sort_cfg = -1;
extern "C" int mhb_set_sort_cfg(int cfg) {
  if (!((cfg >= 0 && cfg <= 3) || (cfg >= 256 && v3_listed(cfg - 256))))
    return mhb_set_error(MHB_ERR_ARG, "unknown sort configuration %d", cfg);
  g_sort_cfg = cfg;
  return MHB_OK;
}
static int sort_cfg() {
  int &cfg = g_sort_cfg;
  if (cfg < 0) {
    const char *e = getenv("MHB_SORT_CFG");
    cfg = e ? atoi(e) : 256 + MHB_V3_DEFAULT;
    if (!((cfg >= 0 && cfg <= 3) || (cfg >= 256 && v3_listed(cfg - 256)))) cfg = 256 + MHB_V3_DEFAULT;
  }
  return cfg;
}
template <int WR, int CFG>
static u64 sort_tiles_cfg(u64 n) {
  return (n + SortCfg<WR, CFG>::TILE - 1) / SortCfg<WR, CFG>::TILE;
}
template <int WR, int CFG>
static u64 sort_tiles_cfg3(u64 n) {
  return (n + SortCfg3<WR, CFG>::TILE - 1) / SortCfg3<WR, CFG>::TILE;
}
template <int WR>
static u64 sort_tiles(u64 n) {
  const int cfg = sort_cfg();
  if (cfg >= 256) {
    if constexpr (WR == 2 || WR == 3) {
#define X(B) \
  if (cfg - 256 == B) return sort_tiles_cfg3<WR, B>(n);
      MHB_V3_LIST(X)
#undef X
    }
    return sort_tiles_cfg3<WR, MHB_V3_DEFAULT>(n);
  }
  if constexpr (WR <= 3) {
    switch (cfg) {
      case 1: return sort_tiles_cfg<WR, 1>(n);
      case 2: return sort_tiles_cfg<WR, 2>(n);
      case 3: return sort_tiles_cfg<WR, 3>(n);
      default: break;
    }
  }
  return sort_tiles_cfg<WR, 0>(n);
}
static u64 sort_num_tiles(u64 n, u32 words) {
#define M(WW) \
  if (words == WW) return sort_tiles<WW>(n);
  MHB_FOR_WR(M)
#undef M
  return 0;
}
static constexpr size_t kSortHeadBytes = (size_t)(72 + 1) * 256 * 8 /*hist*/ + 256 * 8 /*bin_base*/ + 128 * 4;
// look-back storage for `tiles` tiles: 256 64-bit descriptors per tile + (compact-descriptor variants) one 16-byte
// word per digit and group of four tiles behind them
static size_t lb_bytes(u64 tiles) { return (size_t)tiles * 256 * 8 + (size_t)((tiles + 3) / 4) * 256 * 16; }

extern "C" size_t mhb_sort_workspace_bytes(uint64_t n, uint32_t words) {
  // sized for the smallest tile of any configuration so that a workspace stays valid across MHB_SORT_CFG values
  u64 tiles = sort_num_tiles(n, words);
  if (words <= 3) tiles = (n + 256 * 8 - 1) / (256 * 8) > tiles ? (n + 256 * 8 - 1) / (256 * 8) : tiles;
#define M(WW) \
  if (words == WW && sort_tiles_cfg<WW, 0>(n) > tiles) tiles = sort_tiles_cfg<WW, 0>(n);  // partition pass geometry
  MHB_FOR_WR(M)
#undef M
  return kSortHeadBytes + lb_bytes(tiles) + 256;
}

template <int WR, int CFG>
static int launch_radix_pass_cfg(cudaStream_t st, const u32 *in, u64 n, int byte_idx, const u64 *bin_base,
                                 u64 *lookback, u32 *tile_counter, u64 *next_hist, int next_byte, u32 epoch) {
  using C = SortCfg<WR, CFG>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CK(cudaFuncSetAttribute(k_radix_pass<WR, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_radix_pass<WR, CFG>, C::THREADS, C::SMEM));
    if (blocks_per_sm < 1) return mhb_set_error(MHB_ERR_CUDA, "radix pass kernel (WR=%d) does not fit an SM", WR);
    if (getenv("MHB_VERBOSE")) fprintf(stderr, "[mhb] radix pass WR=%d cfg=%d: %d threads x %d rec, %zu B smem, %d CTA/SM\n", WR, CFG, C::THREADS, C::IPT, C::SMEM, blocks_per_sm);
  }
  const u64 tiles = sort_tiles_cfg<WR, CFG>(n);
  u64 grid = (u64)blocks_per_sm * sm_count();
  if (grid > tiles) grid = tiles;
  k_radix_pass<WR, CFG><<<(int)grid, C::THREADS, C::SMEM, st>>>(in, n, (u32)tiles, byte_idx, bin_base, lookback,
                                                                 tile_counter, next_hist, next_byte, epoch);
  CK_LAUNCH();
  return MHB_OK;
}

template <int WR, int CFG>
static int launch_radix_pass_cfg3(cudaStream_t st, const u32 *in, u64 n, int byte_idx, const u64 *bin_base,
                                  u64 *lookback, u32 *tile_counter, u64 *next_hist, int next_byte, u32 epoch) {
  using C = SortCfg3<WR, CFG>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CK(cudaFuncSetAttribute(k_radix_pass3<WR, CFG, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    CK(cudaFuncSetAttribute(k_radix_pass3<WR, CFG, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_radix_pass3<WR, CFG, false, true>, C::THREADS, C::SMEM));
    if (blocks_per_sm < 1) return mhb_set_error(MHB_ERR_CUDA, "radix pass v3 kernel (WR=%d) does not fit an SM", WR);
    if (getenv("MHB_VERBOSE")) fprintf(stderr, "[mhb] radix pass v3 WR=%d bits=0x%03x: %d threads x %d rec, rank %d, prefetch %d, look-back %d/%d, batch %d, early %d, %zu B smem, %d CTA/SM\n", WR, CFG, C::THREADS, C::IPT, C::RANK, (int)C::PREFETCH, C::LB1, C::LBW, (int)C::BATCH, (int)C::EARLY, C::SMEM, blocks_per_sm);
  }
  const u64 tiles = sort_tiles_cfg3<WR, CFG>(n);
  u64 grid = (u64)blocks_per_sm * sm_count();
  if (grid > tiles) grid = tiles;
  if (next_hist)
    k_radix_pass3<WR, CFG, false, true><<<(int)grid, C::THREADS, C::SMEM, st>>>(in, n, (u32)tiles, byte_idx, bin_base, lookback,
                                                                                 tile_counter, next_hist, next_byte, epoch);
  else
    k_radix_pass3<WR, CFG, false, false><<<(int)grid, C::THREADS, C::SMEM, st>>>(in, n, (u32)tiles, byte_idx, bin_base, lookback,
                                                                                  tile_counter, nullptr, 0, epoch);
  CK_LAUNCH();
  return MHB_OK;
}

template <int WR>
static int launch_radix_pass(cudaStream_t st, const u32 *in, u64 n, int byte_idx, const u64 *bin_base,
                             u64 *lookback, u32 *tile_counter, u64 *next_hist, int next_byte, u32 epoch) {
  const int cfg = sort_cfg();
  if (cfg >= 256) {
    if constexpr (WR == 2 || WR == 3) {
#define X(B) \
  if (cfg - 256 == B) return launch_radix_pass_cfg3<WR, B>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
      MHB_V3_LIST(X)
#undef X
    }
    return launch_radix_pass_cfg3<WR, MHB_V3_DEFAULT>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
  }
  if constexpr (WR <= 3) {
    switch (cfg) {
      case 1: return launch_radix_pass_cfg<WR, 1>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
      case 2: return launch_radix_pass_cfg<WR, 2>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
      case 3: return launch_radix_pass_cfg<WR, 3>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
      default: break;
    }
  }
  return launch_radix_pass_cfg<WR, 0>(st, in, n, byte_idx, bin_base, lookback, tile_counter, next_hist, next_byte, epoch);
}

#ifdef MHB_SORT_TIMELINE
// diagnostic build only: point the v3 radix pass at a device buffer of rows x 16 uint64 (see mhb_sort3.cuh)
extern "C" int mhb_debug_set_sort_timeline(unsigned long long *dev_buf, unsigned long long rows) {
  CK(cudaMemcpyToSymbol(g_sort_timeline, &dev_buf, sizeof(dev_buf)));
  CK(cudaMemcpyToSymbol(g_sort_timeline_rows, &rows, sizeof(rows)));
  return MHB_OK;
}
#endif

// Per-pass timing: every sort records one event before and after each pass into a small ring, so a
// caller can ask afterwards (mhb_sort_pass_ms) how long each pass of a recent sort took without putting a
// synchronisation inside its timed region.
namespace {
struct SortTrace {
  cudaEvent_t ev[74];
  bool created = false;
  uint32_t n_passes = 0, words = 0;
  uint64_t n = 0;
};
SortTrace g_trace[4];
uint64_t g_trace_seq = 0;
}  // namespace

int mhb_sort_records_impl(void *stream, uint32_t *a, uint32_t *b, uint64_t n, uint32_t words, const uint8_t *bytes,
                          uint32_t n_bytes, const uint64_t *first_hist, void *ws, size_t ws_bytes, int *result_in_b,
                          double *pass_ms_host) {
  if (words < 1 || words > 17 || n_bytes > 72 || !result_in_b)
    return mhb_set_error(MHB_ERR_ARG, "bad sort geometry (words=%u n_bytes=%u)", words, n_bytes);
  *result_in_b = 0;
  if (n == 0 || n_bytes == 0) return MHB_OK;
  if (ws_bytes < mhb_sort_workspace_bytes(n, words)) return mhb_set_error(MHB_ERR_ARG, "sort workspace too small");
  if (n >= (1ull << 53)) return mhb_set_error(MHB_ERR_ARG, "too many records");
  cudaStream_t st = (cudaStream_t)stream;
  u64 *hist = (u64 *)ws;                        // [n_bytes+1][256]
  u64 *bin_base = hist + (72 + 1) * 256;        // [256]
  u32 *tile_counter = (u32 *)(bin_base + 256);  // [128]
  u64 *lookback = (u64 *)((char *)ws + kSortHeadBytes);
  // only what this sort's tile geometry touches (the workspace itself is sized for the smallest tile of any variant)
  CK(cudaMemsetAsync(ws, 0, kSortHeadBytes + lb_bytes(sort_num_tiles(n, words)) + 256, st));
  if (first_hist) {
    CK(cudaMemcpyAsync(hist, first_hist, 256 * 8, cudaMemcpyDeviceToDevice, st));
  } else {
#define M(WW) \
  if (words == WW) k_hist_byte<WW><<<sm_count() * 4, 256, 0, st>>>(a, n, bytes[0], hist);
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
  }
  SortTrace &tr = g_trace[g_trace_seq++ & 3];
  if (!tr.created) {
    for (int i = 0; i < 74; ++i) CK(cudaEventCreate(&tr.ev[i]));
    tr.created = true;
  }
  tr.n_passes = n_bytes;
  tr.words = words;
  tr.n = n;
  CK(cudaEventRecord(tr.ev[0], st));
  u32 *in = a, *out = b;
  for (u32 p = 0; p < n_bytes; ++p) {
    k_hist_scan256<<<1, 256, 0, st>>>(hist + (u64)p * 256, bin_base, (u64)(uintptr_t)out, words * 4);
    CK_LAUNCH();
    u64 *next_hist = p + 1 < n_bytes ? hist + (u64)(p + 1) * 256 : nullptr;
    const int next_byte = p + 1 < n_bytes ? bytes[p + 1] : 0;
    int rc = MHB_ERR_ARG;
#define M(WW) \
  if (words == WW) rc = launch_radix_pass<WW>(st, in, n, bytes[p], bin_base, lookback, tile_counter + p, next_hist, next_byte, p + 1);
    MHB_FOR_WR(M)
#undef M
    if (rc) return rc;
    CK(cudaEventRecord(tr.ev[p + 1], st));
    u32 *t = in;
    in = out;
    out = t;
  }
  *result_in_b = (in == b) ? 1 : 0;
  if (pass_ms_host) {
    CK(cudaEventSynchronize(tr.ev[n_bytes]));
    for (u32 p = 0; p < n_bytes; ++p) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, tr.ev[p], tr.ev[p + 1]));
      pass_ms_host[p] = ms;
    }
  }
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// fused partition + exchange: one radix pass whose per-digit destinations are arbitrary device addresses, e.g.
// slots inside OTHER GPUs' receive buffers opened through CUDA IPC.  The scatter stores travel over NVLink while
// the rest of the tile is still being ranked - no separate all-to-all.
// ------------------------------------------------------------------------------------------------
template <int WR>
static int launch_partition_pass(cudaStream_t st, const u32 *in, u64 n, int byte_idx, const u64 *bin_addr, u64 *lookback,
                                 u32 *tile_counter, const uint8_t *lut) {
  using C = SortCfg<WR, 0>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CK(cudaFuncSetAttribute(k_radix_pass<WR, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_radix_pass<WR, 0, true>, C::THREADS, C::SMEM));
    if (blocks_per_sm < 1) return mhb_set_error(MHB_ERR_CUDA, "partition pass kernel (WR=%d) does not fit an SM", WR);
  }
  const u64 tiles = sort_tiles_cfg<WR, 0>(n);
  u64 grid = (u64)blocks_per_sm * sm_count();
  if (grid > tiles) grid = tiles;
  k_radix_pass<WR, 0, true><<<(int)grid, C::THREADS, C::SMEM, st>>>(in, n, (u32)tiles, byte_idx, bin_addr, lookback,
                                                                     tile_counter, nullptr, 0, 1, lut);
  CK_LAUNCH();
  return MHB_OK;
}

extern "C" int mhb_partition_scatter(void *stream, const uint32_t *recs, uint64_t n, uint32_t words, int byte,
                                     const uint8_t *owner_of_byte_dev, const uint64_t *bin_addr_dev, void *ws,
                                     size_t ws_bytes) {
  if (words < 1 || words > 17 || byte < 0 || byte >= (int)(4 * words)) return mhb_set_error(MHB_ERR_ARG, "bad geometry");
  if (n == 0) return MHB_OK;
  // the partition pass always runs the v2 kernel in geometry 0, whatever variant the sorts use
  u64 tiles = 0;
#define M(WW) \
  if (words == WW) tiles = sort_tiles_cfg<WW, 0>(n);
  MHB_FOR_WR(M)
#undef M
  const size_t need = kSortHeadBytes + (size_t)tiles * 256 * 8 + 256;
  if (ws_bytes < need) return mhb_set_error(MHB_ERR_ARG, "sort workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  u64 *hist = (u64 *)ws;
  u32 *tile_counter = (u32 *)(hist + (72 + 1) * 256 + 256);
  u64 *lookback = (u64 *)((char *)ws + kSortHeadBytes);
  CK(cudaMemsetAsync(ws, 0, need, st));
  int rc = MHB_ERR_ARG;
  if (owner_of_byte_dev) {
#define M(WW) \
  if (words == WW) rc = launch_partition_pass<WW>(st, recs, n, byte, bin_addr_dev, lookback, tile_counter, owner_of_byte_dev);
    MHB_FOR_WR(M)
#undef M
  } else {
#define M(WW) \
  if (words == WW) rc = launch_radix_pass_cfg<WW, 0>(st, recs, n, byte, bin_addr_dev, lookback, tile_counter, nullptr, 0, 1);
    MHB_FOR_WR(M)
#undef M
  }
  return rc;
}

extern "C" int mhb_dev_malloc(void **ptr, size_t bytes) {
  CK(cudaMalloc(ptr, bytes));
  return MHB_OK;
}
extern "C" int mhb_dev_free(void *ptr) {
  CK(cudaFree(ptr));
  return MHB_OK;
}
extern "C" int mhb_ipc_export(const void *dev_ptr, uint8_t *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
  memcpy(handle64, &h, 64);
  return MHB_OK;
}
extern "C" int mhb_ipc_open(const uint8_t *handle64, void **peer_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MHB_OK;
}
extern "C" int mhb_ipc_close(void *peer_ptr) {
  CK(cudaIpcCloseMemHandle(peer_ptr));
  return MHB_OK;
}

extern "C" int mhb_sort_pass_ms(int back, double *pass_ms, uint32_t max_passes, uint32_t *n_passes, uint64_t *n_records,
                                uint32_t *words) {
  if (back < 0 || back > 3 || (uint64_t)back >= g_trace_seq) return mhb_set_error(MHB_ERR_ARG, "no such sort in the trace ring");
  SortTrace &tr = g_trace[(g_trace_seq - 1 - back) & 3];
  CK(cudaEventSynchronize(tr.ev[tr.n_passes]));
  for (u32 p = 0; p < tr.n_passes && p < max_passes; ++p) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, tr.ev[p], tr.ev[p + 1]));
    pass_ms[p] = ms;
  }
  if (n_passes) *n_passes = tr.n_passes;
  if (n_records) *n_records = tr.n;
  if (words) *words = tr.words;
  return MHB_OK;
}

extern "C" int mhb_sort_records(void *stream, uint32_t *a, uint32_t *b, uint64_t n, uint32_t words,
                                const uint8_t *bytes, uint32_t n_bytes, const uint64_t *first_hist, void *ws,
                                size_t ws_bytes, int *result_in_b) {
  return mhb_sort_records_impl(stream, a, b, n, words, bytes, n_bytes, first_hist, ws, ws_bytes, result_in_b, nullptr);
}

// ------------------------------------------------------------------------------------------------
// count: solid edges
// ------------------------------------------------------------------------------------------------
extern "C" int mhb_count_solid(void *stream, const uint32_t *sorted_records, uint64_t n, uint32_t k, int32_t m,
                               uint32_t *edges_out, uint8_t *aux_out, uint64_t capacity_edges, uint64_t *mul_hist,
                               uint64_t *n_solid_out, void *scratch, size_t scratch_bytes) {
  if (k < 1 || k > MHB_MAX_K || !mul_hist || !n_solid_out) return mhb_set_error(MHB_ERR_ARG, "bad args");
  if (n == 0) return MHB_OK;
  const u64 nblk = (n + kCompactTile - 1) / kCompactTile;
  const size_t need = mhb_count_solid_scratch_bytes(n);
  if (scratch_bytes < need) return mhb_set_error(MHB_ERR_ARG, "count scratch too small (%zu < %zu)", scratch_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const u32 WR = count_record_words(k);
  static const bool use_v1 = getenv("MHB_COUNT_V1") != nullptr, use_v2 = getenv("MHB_COUNT_V2") != nullptr;
  if (!use_v1 && !use_v2) {
    // v3: lane-blocked judge -> scan of chunk totals -> gather/pack (mhb_count.cuh)
#define M(WW)                                                                                                          \
  if (WR == WW) {                                                                                                      \
    constexpr int CH = 32 * count3_ipl(WW);                                                                            \
    const u64 n_chunks = (n + CH - 1) / CH;                                                                            \
    if (n_chunks >= (1ull << 32)) return mhb_set_error(MHB_ERR_ARG, "too many records for one count launch");          \
    const u64 n_sblk = (n_chunks + kScanTile - 1) / kScanTile;                                                         \
    char *p = (char *)scratch;                                                                                         \
    u32 *ticket = (u32 *)p;                                                                                            \
    p += 256;                                                                                                          \
    uint2 *solid_list = (uint2 *)p;                                                                                    \
    p += ((size_t)n_chunks * CH * 8 + 255) & ~(size_t)255;                                                             \
    u32 *chunk_count = (u32 *)p;                                                                                       \
    p += ((size_t)n_chunks * 4 + 255) & ~(size_t)255;                                                                  \
    u64 *chunk_off = (u64 *)p;                                                                                         \
    p += ((size_t)n_chunks * 8 + 255) & ~(size_t)255;                                                                  \
    u64 *bsum = (u64 *)p;                                                                                              \
    CK(cudaMemsetAsync(ticket, 0, 256, st));                                                                           \
    const size_t smem = (size_t)kCount3Warps * count3_warp_words(WW) * 4;                                              \
    static int bps = 0;                                                                                                \
    if (!bps) {                                                                                                        \
      CK(cudaFuncSetAttribute(k_count_lanes<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_count_lanes<WW>, kCount3Warps * 32, smem));             \
      if (bps < 1) bps = 1;                                                                                            \
    }                                                                                                                  \
    u64 grid = (u64)sm_count() * bps;                                                                                  \
    if (grid > (n_chunks + kCount3Warps - 1) / kCount3Warps) grid = (n_chunks + kCount3Warps - 1) / kCount3Warps;      \
    k_count_lanes<WW><<<(unsigned)grid, kCount3Warps * 32, smem, st>>>(sorted_records, n, k, m, (u32)n_chunks, ticket,  \
                                                                      solid_list, chunk_count, mul_hist);             \
    CK_LAUNCH();                                                                                                       \
    k_scan32_sums<<<(unsigned)n_sblk, kScanThreads, 0, st>>>(chunk_count, n_chunks, bsum);                             \
    CK_LAUNCH();                                                                                                       \
    k_scan_u64<<<1, 1024, 0, st>>>(bsum, n_sblk, n_solid_out);                                                         \
    CK_LAUNCH();                                                                                                       \
    k_scan32_apply<<<(unsigned)n_sblk, kScanThreads, 0, st>>>(chunk_count, n_chunks, bsum, chunk_off);                 \
    CK_LAUNCH();                                                                                                       \
    u64 gw = (n_chunks + 7) / 8;                                                                                       \
    if (gw > (u64)sm_count() * 16) gw = (u64)sm_count() * 16;                                                          \
    k_count_write<WW><<<(unsigned)gw, 256, 0, st>>>(sorted_records, k, (u32)n_chunks, solid_list, chunk_count,         \
                                                    chunk_off, edges_out, aux_out, capacity_edges);                   \
  }
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    return MHB_OK;
  }
  if (use_v2) {
    // v2: single pass, warp-cooperative ballots (mhb_count.cuh k_count_warp); kept for A/B checks
    const u32 WE = words_per_edge(k);
    const int CH = count_chunk((int)WE);
    const u64 n_chunks = (n + CH - 1) / CH;
    if (n_chunks >= (1ull << 32)) return mhb_set_error(MHB_ERR_ARG, "too many records for one count launch");
    u32 *ticket = (u32 *)scratch;
    u64 *lookback = (u64 *)((char *)scratch + 64);
    CK(cudaMemsetAsync(scratch, 0, 64 + n_chunks * 8, st));
    const size_t smem = (size_t)kCountWarps * CH * (WE * 4 + 1);
    u64 grid = (u64)sm_count() * 4;
    if (grid > (n_chunks + kCountWarps - 1) / kCountWarps) grid = (n_chunks + kCountWarps - 1) / kCountWarps;
#define M(WW)                                                                                                        \
  if (WR == WW) {                                                                                                    \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      CK(cudaFuncSetAttribute(k_count_warp<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));           \
      attr = true;                                                                                                   \
    }                                                                                                                \
    k_count_warp<WW><<<(unsigned)grid, kCountWarps * 32, smem, st>>>(sorted_records, n, k, m, (u32)n_chunks, ticket,  \
                                                                    lookback, edges_out, aux_out, capacity_edges,    \
                                                                    mul_hist, n_solid_out);                         \
  }
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    return MHB_OK;
  }
  u32 *info = (u32 *)scratch;
  u64 *btot = (u64 *)((char *)scratch + (((size_t)n * 4 + 63) & ~(size_t)63));
  const u64 gmark = (n + 255) / 256;
#define M(WW) \
  if (WR == WW) k_count_mark<WW><<<(unsigned)gmark, 256, 0, st>>>(sorted_records, n, m, info, mul_hist);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  k_solid_block_totals<<<(unsigned)nblk, kCompactThreads, 0, st>>>(info, n, btot);
  CK_LAUNCH();
  k_scan_u64<<<1, 1024, 0, st>>>(btot, nblk, n_solid_out);
  CK_LAUNCH();
#define M(WW)                                                                                                       \
  if (WR == WW)                                                                                                     \
    k_count_emit<WW><<<(unsigned)nblk, kCompactThreads, 0, st>>>(sorted_records, info, n, k, btot, edges_out, aux_out, \
                                                                 capacity_edges);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

extern "C" size_t mhb_count_solid_scratch_bytes(uint64_t n) {
  // solid list (8 B per record slot, chunk-rounded) + per-chunk count/offset (chunks of >= 128 records) + slack
  return (size_t)(n + 1024) * 8 + (size_t)(n / 128 + 2) * 12 + (size_t)(n / 128 / kScanTile + 2) * 8 + 4096;
}

// ------------------------------------------------------------------------------------------------
// count: mercy bookkeeping
// ------------------------------------------------------------------------------------------------
extern "C" size_t mhb_tipset_bytes(uint64_t n_tip_edges, uint32_t k) {
  return 16 + tipset_filter_words(n_tip_edges) * 4 + tipset_capacity(n_tip_edges) * (size_t)(count_key_words(k) + 1) * 4;
}

// a few device words for scalar results, allocated once per device (cudaMallocAsync/cudaFreeAsync per call
// was measured to cost tens to hundreds of ms when most of HBM is already reserved)
static int small_scratch(unsigned long long **out) {
  static unsigned long long *buf[64] = {nullptr};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return mhb_set_error(MHB_ERR_CUDA, "device index %d out of range", dev);
  if (!buf[dev]) CK(cudaMalloc((void **)&buf[dev], 256));
  *out = buf[dev];
  return MHB_OK;
}

extern "C" int mhb_count_tip_edges(void *stream, const uint8_t *aux, uint64_t n_solid, uint64_t *n_tip_host) {
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long *d = nullptr;
  *n_tip_host = 0;
  if (n_solid == 0) return MHB_OK;
  if (int rc = small_scratch(&d)) return rc;
  CK(cudaMemsetAsync(d, 0, 8, st));
  k_count_tips<<<sm_count() * 4, 256, 0, st>>>(aux, n_solid, d);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(n_tip_host, d, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MHB_OK;
}

extern "C" int mhb_tipset_build(void *stream, const uint32_t *edges, const uint8_t *aux, uint64_t n_solid, uint32_t k,
                                void *tipset, size_t tipset_bytes, uint64_t n_tip_edges) {
  if (tipset_bytes < mhb_tipset_bytes(n_tip_edges, k)) return mhb_set_error(MHB_ERR_ARG, "tipset too small");
  cudaStream_t st = (cudaStream_t)stream;
  const u64 hdr[2] = {tipset_capacity(n_tip_edges), tipset_filter_words(n_tip_edges)};
  const u64 cap = hdr[0], fwords = hdr[1];
  CK(cudaMemsetAsync(tipset, 0, mhb_tipset_bytes(n_tip_edges, k), st));
  CK(cudaMemcpyAsync(tipset, hdr, 16, cudaMemcpyHostToDevice, st));
  if (n_solid == 0) return MHB_OK;
  u32 *filter = (u32 *)((char *)tipset + 16);
  u32 *table = filter + fwords;
  const u32 W = count_key_words(k);
  const u64 g = (n_solid + 255) / 256;
#define M(WW) \
  if (W == WW) k_tipset_insert<WW><<<(unsigned)g, 256, 0, st>>>(edges, aux, n_solid, k, filter, fwords, table, cap);
  MHB_FOR_W(M)
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

extern "C" int mhb_count_mark_mercy(void *stream, const mhb_dev_reads *reads, uint32_t k, const void *tipset,
                                    size_t tipset_bytes, uint64_t n_tip_edges, uint32_t *first_0_out,
                                    uint32_t *last_0_in) {
  if (int rc = check_reads(reads, k)) return rc;
  if (reads->n_reads == 0) return MHB_OK;
  if (tipset_bytes < mhb_tipset_bytes(n_tip_edges, k)) return mhb_set_error(MHB_ERR_ARG, "bad tipset");
  cudaStream_t st = (cudaStream_t)stream;
  const ReadsView rv = make_reads_view(reads);
  const u32 W = count_key_words(k), WR = count_record_words(k);
  const u64 cap = tipset_capacity(n_tip_edges), fwords = tipset_filter_words(n_tip_edges);
  const u32 *filter = (const u32 *)((const char *)tipset + 16);
  const u32 *table = filter + fwords;
  u64 g64 = (rv.n_reads + 7) / 8;
  if (g64 > (u64)sm_count() * 16) g64 = (u64)sm_count() * 16;
  const int grid = (int)g64;
  static const bool roll = getenv("MHB_EXTRACT_ROLL") != nullptr;
  if (roll && W == 2 && WR == 2 && k + 1 >= 17) {
    k_mark_mercy_roll<<<grid, 256, 0, st>>>(rv, k, filter, fwords, table, cap, first_0_out, last_0_in);
    CK_LAUNCH();
    return MHB_OK;
  }
#define M(WW)                                                                                                          \
  if (W == WW && WR == WW)                                                                                             \
    k_mark_mercy<WW, WW><<<grid, 256, 0, st>>>(rv, k, filter, fwords, table, cap, first_0_out, last_0_in);              \
  else if (W == WW && WR == WW + 1)                                                                                    \
    k_mark_mercy<WW, WW + 1><<<grid, 256, 0, st>>>(rv, k, filter, fwords, table, cap, first_0_out, last_0_in);          \
  else
  MHB_FOR_W(M) return mhb_set_error(MHB_ERR_ARG, "unsupported k=%u", k);
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// seq2sdbg
// ------------------------------------------------------------------------------------------------
static SeqsView make_seqs_view(const mhb_dev_seqs *s) {
  SeqsView v;
  v.words = s->words;
  v.n_words = s->n_words;
  v.n_seqs = s->n_seqs;
  v.fixed_len = s->fixed_len;
  v.word_off = s->word_off;
  v.len = s->len;
  v.item_off = s->item_off;
  v.mult = s->mult;
  v.fixed_stride = s->fixed_stride;
  return v;
}

extern "C" int mhb_s2s_extract(void *stream, const mhb_dev_seqs *seqs, uint32_t k, uint32_t *records,
                               uint64_t n_items, uint64_t *hist256, int hist_byte) {
  if (!seqs || k < 9 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "kmer size must be >= 9 and <= 255");
  if (n_items == 0) return MHB_OK;
  if (!seqs->mult && !(seqs->fixed_len && seqs->fixed_stride))
    return mhb_set_error(MHB_ERR_ARG, "seqs->mult is NULL (only allowed for fixed-stride edge records)");
  if (!seqs->fixed_len && (!seqs->word_off || !seqs->len || !seqs->item_off))
    return mhb_set_error(MHB_ERR_ARG, "variable-length sequences need word_off, len and item_off");
  if (seqs->fixed_len && seqs->fixed_len < k + 1) return mhb_set_error(MHB_ERR_ARG, "fixed_len < k+1");
  cudaStream_t st = (cudaStream_t)stream;
  const SeqsView sv = make_seqs_view(seqs);
  const u32 W = s2s_record_words(k);
  if (seqs->fixed_len == k + 1 && seqs->fixed_stride && !seqs->mult && k + 1 <= 32 && k + 1 > 16 &&
      n_items == seqs->n_seqs * 6) {
    // `.edges` records of short (k+1)-mers: one thread per edge, 64-bit arithmetic
    u64 ge = (seqs->n_seqs + 255) / 256;
    if (ge > (u64)sm_count() * 32) ge = (u64)sm_count() * 32;
    if (W == 2) k_s2s_extract_edges<2><<<(unsigned)ge, 256, 0, st>>>(seqs->words, seqs->n_seqs, seqs->fixed_stride, k, records, hist256, hist_byte);
    else if (W == 3) k_s2s_extract_edges<3><<<(unsigned)ge, 256, 0, st>>>(seqs->words, seqs->n_seqs, seqs->fixed_stride, k, records, hist256, hist_byte);
    else return mhb_set_error(MHB_ERR_ARG, "internal: unexpected record width %u", W);
    CK_LAUNCH();
    return MHB_OK;
  }
  u64 g = (n_items + 255) / 256;
  if (g > (u64)sm_count() * 32) g = (u64)sm_count() * 32;
#define M(WW) \
  if (W == WW) k_s2s_extract<WW><<<(unsigned)g, 256, 0, st>>>(sv, k, records, n_items, hist256, hist_byte);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

// A13 for seq2sdbg: the items whose leading record byte lies in [lo, hi].  records == NULL counts only (hist256 +=
// histogram of record byte hist_byte over the in-range items); otherwise the in-range records are appended at
// records[*cursor_dev ...) (cursor_dev: device uint64, caller-zeroed; ends at the number of in-range items even when
// that exceeds `capacity`, in which case the surplus was not stored).
extern "C" int mhb_s2s_extract_range(void *stream, const mhb_dev_seqs *seqs, uint32_t k, uint32_t *records, uint64_t n_items,
                                     uint32_t lo, uint32_t hi, uint64_t *cursor_dev, uint64_t capacity, uint64_t *hist256,
                                     int hist_byte) {
  if (!seqs || k < 9 || k > MHB_MAX_K) return mhb_set_error(MHB_ERR_ARG, "kmer size must be >= 9 and <= 255");
  if (lo > hi || hi > 65535) return mhb_set_error(MHB_ERR_ARG, "bad bucket range [%u, %u]", lo, hi);
  if (records && !cursor_dev) return mhb_set_error(MHB_ERR_ARG, "cursor is NULL");
  if (n_items == 0) return MHB_OK;
  if (!seqs->mult && !(seqs->fixed_len && seqs->fixed_stride))
    return mhb_set_error(MHB_ERR_ARG, "seqs->mult is NULL (only allowed for fixed-stride edge records)");
  if (!seqs->fixed_len && (!seqs->word_off || !seqs->len || !seqs->item_off))
    return mhb_set_error(MHB_ERR_ARG, "variable-length sequences need word_off, len and item_off");
  if (seqs->fixed_len && seqs->fixed_len < k + 1) return mhb_set_error(MHB_ERR_ARG, "fixed_len < k+1");
  cudaStream_t st = (cudaStream_t)stream;
  const SeqsView sv = make_seqs_view(seqs);
  const u32 W = s2s_record_words(k);
  u64 g = (n_items + 255) / 256;
  if (g > (u64)sm_count() * 32) g = (u64)sm_count() * 32;
#define M(WW)                                                                                                        \
  if (W == WW)                                                                                                       \
    k_s2s_extract_range<WW><<<(unsigned)g, 256, 0, st>>>(sv, k, records, n_items, lo, hi, (unsigned long long *)cursor_dev, \
                                                         capacity, hist256, hist_byte);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  return MHB_OK;
}

static int scan32(cudaStream_t st, const u32 *in, u64 n, u64 *out, u64 *total_dev, u64 *bsum);
// in-place exclusive scan of n u64 values (three phases, no serial chain); total -> *total_dev
static int scan64(cudaStream_t st, u64 *v, u64 n, u64 *total_dev, u64 *bsum) {
  const u64 nb = (n + kScanTile - 1) / kScanTile;
  k_scan64_sums<<<(unsigned)nb, kScanThreads, 0, st>>>(v, n, bsum);
  CK_LAUNCH();
  k_scan_u64<<<1, 1024, 0, st>>>(bsum, nb, total_dev);
  CK_LAUNCH();
  k_scan64_apply<<<(unsigned)nb, kScanThreads, 0, st>>>(v, n, bsum);
  CK_LAUNCH();
  return MHB_OK;
}

static u64 emit2_chunks(u64 n, u32 W) {
#define M(WW) \
  if (W == WW) return (n + emit2_chunk(WW) - 1) / emit2_chunk(WW);
  MHB_FOR_WR(M)
#undef M
  return 0;
}
static u32 emit2_chunk_records(u32 W) {
#define M(WW) \
  if (W == WW) return (u32)emit2_chunk(WW);
  MHB_FOR_WR(M)
#undef M
  return 0;
}

extern "C" size_t mhb_s2s_emit_scratch_bytes(uint64_t n, uint32_t k) {
  const u32 W = s2s_record_words(k);
  const u64 nc = emit2_chunks(n, W);
  const u64 nblk = (n + kEmitThreads - 1) / kEmitThreads;  // v1 layout (MHB_EMIT_V1)
  const size_t v1 = (size_t)nblk * 4 * 8 + (size_t)MHB_NUM_BUCKETS * 4 * 8 + (size_t)(nblk / kScanTile + 2) * 8 + 512;
  const size_t v2 = (size_t)MHB_NUM_BUCKETS * (4 * 8 + 5 * 4) + (size_t)nc * 4 * (4 + 8) + (size_t)(nc / kScanTile + 2) * 8 +
                    (size_t)nc * emit2_chunk_records(W) * emit2_max_item_bytes(k) + 4096;
  return v1 > v2 ? v1 : v2;
}

extern "C" int mhb_s2s_emit(void *stream, const uint32_t *sorted_records, uint64_t n, uint32_t k, uint8_t *bytes_out,
                            uint64_t capacity_bytes, uint64_t *bucket_table, uint64_t *totals, void *scratch,
                            size_t scratch_bytes) {
  if (k < 9 || k > MHB_MAX_K || !bucket_table || !totals) return mhb_set_error(MHB_ERR_ARG, "bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(totals, 0, 16 * 8, st));
  CK(cudaMemsetAsync(bucket_table, 0, (size_t)MHB_NUM_BUCKETS * 4 * 8, st));
  if (n == 0) return MHB_OK;
  if (scratch_bytes < mhb_s2s_emit_scratch_bytes(n, k)) return mhb_set_error(MHB_ERR_ARG, "emit scratch too small");
  const u32 W = s2s_record_words(k);
  static const bool use_v1 = getenv("MHB_EMIT_V1") != nullptr;
  if (!use_v1) {
    const u64 nc = emit2_chunks(n, W);
    if (nc >= (1ull << 32)) return mhb_set_error(MHB_ERR_ARG, "too many records for one emit launch");
    const u32 chrec = emit2_chunk_records(W), maxb = emit2_max_item_bytes(k);
    char *p = (char *)scratch;
    u64 *bucket_start = (u64 *)p;
    p += (size_t)MHB_NUM_BUCKETS * 4 * 8;
    u32 *bucket_local = (u32 *)p;
    p += (size_t)MHB_NUM_BUCKETS * 5 * 4;
    u64 *chunk_off = (u64 *)p;
    p += (size_t)nc * 4 * 8;
    u64 *bsum = (u64 *)p;
    p += (size_t)(nc / kScanTile + 2) * 8;
    u32 *chunk_tot = (u32 *)p;
    p += ((size_t)nc * 4 * 4 + 255) & ~(size_t)255;
    uint8_t *tmp = (uint8_t *)p;
    CK(cudaMemsetAsync(bucket_local, 0xFF, (size_t)MHB_NUM_BUCKETS * 5 * 4, st));
#define M(WW)                                                                                                       \
  if (W == WW) {                                                                                                    \
    const size_t smem = (size_t)kEmit2Warps * emit2_slots(WW) * WW * 4;                                             \
    static int bps = 0;                                                                                             \
    if (!bps) {                                                                                                     \
      CK(cudaFuncSetAttribute(k_s2s_judge<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_s2s_judge<WW>, kEmit2Warps * 32, smem));             \
      if (bps < 1) bps = 1;                                                                                         \
    }                                                                                                               \
    u64 grid = (u64)sm_count() * bps;                                                                               \
    if (grid > (nc + kEmit2Warps - 1) / kEmit2Warps) grid = (nc + kEmit2Warps - 1) / kEmit2Warps;                   \
    k_s2s_judge<WW><<<(unsigned)grid, kEmit2Warps * 32, smem, st>>>(sorted_records, n, k, (u32)nc, tmp, chunk_tot,   \
                                                                   bucket_local, totals);                          \
  }
    MHB_FOR_WR(M)
#undef M
    CK_LAUNCH();
    for (int q = 0; q < 4; ++q)
      if (int rc = scan32(st, chunk_tot + (u64)q * nc, nc, chunk_off + (u64)q * nc, totals + q, bsum)) return rc;
    u64 gg = (nc + 7) / 8;
    if (gg > (u64)sm_count() * 16) gg = (u64)sm_count() * 16;
    k_s2s_gather<<<(unsigned)gg, 256, 0, st>>>(tmp, chrec, maxb, (u32)nc, chunk_tot, chunk_off, bytes_out, capacity_bytes);
    CK_LAUNCH();
    k_bucket_starts<<<64, 256, 0, st>>>(bucket_local, chunk_off, nc, bucket_start);
    CK_LAUNCH();
    k_bucket_finalize<<<64, 256, 0, st>>>(bucket_start, totals, bucket_table);
    CK_LAUNCH();
    return MHB_OK;
  }
  const u64 nblk = (n + kEmitThreads - 1) / kEmitThreads;
  u64 *bucket_start = (u64 *)scratch;
  u64 *btot = bucket_start + (size_t)MHB_NUM_BUCKETS * 4;
  CK(cudaMemsetAsync(bucket_start, 0xFF, (size_t)MHB_NUM_BUCKETS * 4 * 8, st));
#define M(WW) \
  if (W == WW) k_s2s_size<WW><<<(unsigned)nblk, kEmitThreads, 0, st>>>(sorted_records, n, k, btot);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  u64 *bsum = btot + 4 * nblk;
  for (int q = 0; q < 4; ++q)
    if (int rc = scan64(st, btot + (u64)q * nblk, nblk, totals + q, bsum)) return rc;
#define M(WW)                                                                                                    \
  if (W == WW)                                                                                                   \
    k_s2s_write<WW><<<(unsigned)nblk, kEmitThreads, 0, st>>>(sorted_records, n, k, btot, bytes_out, capacity_bytes, \
                                                            bucket_start, totals);
  MHB_FOR_WR(M)
#undef M
  CK_LAUNCH();
  k_bucket_finalize<<<64, 256, 0, st>>>(bucket_start, totals, bucket_table);
  CK_LAUNCH();
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// mercy edges on the device (A11)
// ------------------------------------------------------------------------------------------------
static int scan32(cudaStream_t st, const u32 *in, u64 n, u64 *out, u64 *total_dev, u64 *bsum) {
  const u64 nb = (n + kScanTile - 1) / kScanTile;
  k_scan32_sums<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, bsum);
  CK_LAUNCH();
  k_scan_u64<<<1, 1024, 0, st>>>(bsum, nb, total_dev);
  CK_LAUNCH();
  k_scan32_apply<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, bsum, out);
  CK_LAUNCH();
  return MHB_OK;
}

extern "C" size_t mhb_mercy_candidates_scratch_bytes(uint64_t n_reads) {
  return (size_t)(n_reads + 64) * 12 + (size_t)(n_reads / kScanTile + 2) * 8 + 1024;
}

extern "C" int mhb_mercy_candidates(void *stream, const uint32_t *first_0_out, const uint32_t *last_0_in,
                                    uint64_t n_reads, uint64_t *cand_ids, uint64_t *n_cand_host, void *scratch,
                                    size_t scratch_bytes) {
  *n_cand_host = 0;
  if (n_reads == 0) return MHB_OK;
  if (scratch_bytes < mhb_mercy_candidates_scratch_bytes(n_reads)) return mhb_set_error(MHB_ERR_ARG, "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  char *p = (char *)scratch;
  u64 *total = (u64 *)p;
  p += 256;
  u32 *flag = (u32 *)p;
  p += ((size_t)n_reads * 4 + 255) & ~(size_t)255;
  u64 *off = (u64 *)p;
  p += ((size_t)n_reads * 8 + 255) & ~(size_t)255;
  u64 *bsum = (u64 *)p;
  const unsigned g = (unsigned)((n_reads + 255) / 256);
  k_cand_flags<<<g, 256, 0, st>>>(first_0_out, last_0_in, n_reads, flag);
  CK_LAUNCH();
  if (int rc = scan32(st, flag, n_reads, off, total, bsum)) return rc;
  k_cand_compact<<<g, 256, 0, st>>>(flag, off, n_reads, cand_ids);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(n_cand_host, total, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MHB_OK;
}

extern "C" size_t mhb_edge_lut_bytes(void);
static size_t mercy_core_scratch(uint64_t n_cand, uint32_t max_read_len) {
  const size_t wpr = (max_read_len + 31) / 32 + 1;
  return (size_t)n_cand * 3 * wpr * 4 + (size_t)(n_cand + 64) * 12 + (size_t)(n_cand / kScanTile + 2) * 8 + 2048;
}
// scratch for mhb_mercy_edges (single segment: includes room for its look-up table); the segmented call needs
// this minus mhb_edge_lut_bytes()
extern "C" size_t mhb_mercy_edges_scratch_bytes(uint64_t n_cand, uint32_t max_read_len) {
  return ((mercy_core_scratch(n_cand, max_read_len) + 255) & ~(size_t)255) + 512 + mhb_edge_lut_bytes();
}

extern "C" size_t mhb_edge_lut_bytes(void) { return (size_t)kLutEntries * sizeof(uint2); }

extern "C" int mhb_edge_lut_build(void *stream, const uint32_t *edges, uint64_t n_edges, uint32_t k, void *lut) {
  if (n_edges >= 0xFFFFFFFFull) return mhb_set_error(MHB_ERR_ARG, "too many edges for the 32-bit look-up table");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(lut, 0xFF, mhb_edge_lut_bytes(), st));
  if (n_edges == 0) return MHB_OK;
  u64 g = (n_edges + 255) / 256;
  if (g > (u64)sm_count() * 16) g = (u64)sm_count() * 16;
  k_edge_lut<<<(unsigned)g, 256, 0, st>>>(edges, n_edges, words_per_edge(k), (uint2 *)lut);
  CK_LAUNCH();
  return MHB_OK;
}

// Layout of the mercy scratch shared by the count and the write half
struct MercyScratch {
  u64 *total;
  u32 *bits, *count;
  u64 *off, *bsum;
  u32 wpr;
};
static MercyScratch mercy_scratch_layout(void *scratch, uint64_t n_cand, uint32_t max_read_len) {
  MercyScratch m;
  m.wpr = (max_read_len + 31) / 32 + 1;
  char *p = (char *)scratch;
  m.total = (u64 *)p;
  p += 256;
  m.bits = (u32 *)p;
  p += ((size_t)n_cand * 3 * m.wpr * 4 + 255) & ~(size_t)255;
  m.count = (u32 *)p;
  p += ((size_t)n_cand * 4 + 255) & ~(size_t)255;
  m.off = (u64 *)p;
  p += ((size_t)n_cand * 8 + 255) & ~(size_t)255;
  m.bsum = (u64 *)p;
  return m;
}

extern "C" int mhb_mercy_edges_count(void *stream, const mhb_dev_reads *reads, const uint64_t *cand_ids, uint64_t n_cand,
                                     uint32_t max_read_len, uint32_t k, uint32_t n_segs, const uint32_t *const *seg_edges,
                                     const uint64_t *seg_counts, const void *const *seg_luts, const uint8_t *owner_of_byte,
                                     uint64_t *n_mercy_host, void *scratch, size_t scratch_bytes) {
  *n_mercy_host = 0;
  if (int rc = check_reads(reads, k)) return rc;
  if (n_cand == 0) return MHB_OK;
  if (k < 12) return mhb_set_error(MHB_ERR_ARG, "mercy edges need k >= 12 (12-mer look-up prefix)");
  if (n_segs < 1 || n_segs > 16) return mhb_set_error(MHB_ERR_ARG, "1..16 edge segments supported");
  if (scratch_bytes < mercy_core_scratch(n_cand, max_read_len)) return mhb_set_error(MHB_ERR_ARG, "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const ReadsView rv = make_reads_view(reads);
  EdgeSegs sg;
  memset(&sg, 0, sizeof(sg));
  for (u32 i = 0; i < n_segs; ++i) {
    sg.ptr[i] = seg_edges[i];
    sg.n[i] = (long long)seg_counts[i];
    sg.lut[i] = (const uint2 *)seg_luts[i];
    if (!sg.lut[i]) return mhb_set_error(MHB_ERR_ARG, "segment %u has no look-up table (mhb_edge_lut_build)", i);
  }
  for (int b = 0; b < 256; ++b) {
    sg.owner[b] = owner_of_byte ? owner_of_byte[b] : 0;
    if (sg.owner[b] >= n_segs) return mhb_set_error(MHB_ERR_ARG, "owner_of_byte[%d] = %u out of range", b, sg.owner[b]);
  }
  const u32 WE = words_per_edge(k), WM = div_ceil(k + 1, 16);
  const MercyScratch ms = mercy_scratch_layout(scratch, n_cand, max_read_len);
  u64 g64 = (n_cand + 7) / 8;
  if (g64 > (u64)sm_count() * 16) g64 = (u64)sm_count() * 16;
#define M(WW) \
  if (WM == WW) k_mercy_probe<WW><<<(unsigned)g64, 256, 0, st>>>(rv, cand_ids, n_cand, k, sg, WE, ms.bits, ms.wpr);
  MHB_FOR_W(M)
#undef M
  CK_LAUNCH();
  const unsigned g = (unsigned)((n_cand + 127) / 128);
  k_mercy_emit<false><<<g, 128, 0, st>>>(rv, cand_ids, n_cand, k, ms.bits, ms.wpr, ms.count, nullptr, nullptr, WE);
  CK_LAUNCH();
  if (int rc = scan32(st, ms.count, n_cand, ms.off, ms.total, ms.bsum)) return rc;
  CK(cudaMemcpyAsync(n_mercy_host, ms.total, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MHB_OK;
}

extern "C" int mhb_mercy_edges_write(void *stream, const mhb_dev_reads *reads, const uint64_t *cand_ids, uint64_t n_cand,
                                     uint32_t max_read_len, uint32_t k, uint32_t *mercy_out, uint64_t capacity,
                                     uint64_t n_mercy, void *scratch, size_t scratch_bytes) {
  if (int rc = check_reads(reads, k)) return rc;
  if (n_cand == 0 || n_mercy == 0) return MHB_OK;
  if (scratch_bytes < mercy_core_scratch(n_cand, max_read_len)) return mhb_set_error(MHB_ERR_ARG, "scratch too small");
  if (n_mercy > capacity) return mhb_set_error(MHB_ERR_NOMEM, "mercy edges (%llu) exceed capacity (%llu)",
                                               (unsigned long long)n_mercy, (unsigned long long)capacity);
  cudaStream_t st = (cudaStream_t)stream;
  const ReadsView rv = make_reads_view(reads);
  const MercyScratch ms = mercy_scratch_layout(scratch, n_cand, max_read_len);
  const unsigned g = (unsigned)((n_cand + 127) / 128);
  k_mercy_emit<true><<<g, 128, 0, st>>>(rv, cand_ids, n_cand, k, ms.bits, ms.wpr, ms.count, ms.off, mercy_out, words_per_edge(k));
  CK_LAUNCH();
  return MHB_OK;
}

extern "C" int mhb_mercy_edges_segs(void *stream, const mhb_dev_reads *reads, const uint64_t *cand_ids, uint64_t n_cand,
                                    uint32_t max_read_len, uint32_t k, uint32_t n_segs, const uint32_t *const *seg_edges,
                                    const uint64_t *seg_counts, const void *const *seg_luts, const uint8_t *owner_of_byte,
                                    uint32_t *mercy_out, uint64_t capacity, uint64_t *n_mercy_host, void *scratch,
                                    size_t scratch_bytes) {
  if (int rc = mhb_mercy_edges_count(stream, reads, cand_ids, n_cand, max_read_len, k, n_segs, seg_edges, seg_counts, seg_luts,
                                     owner_of_byte, n_mercy_host, scratch, scratch_bytes))
    return rc;
  return mhb_mercy_edges_write(stream, reads, cand_ids, n_cand, max_read_len, k, mercy_out, capacity, *n_mercy_host, scratch,
                               scratch_bytes);
}

extern "C" int mhb_mercy_edges(void *stream, const mhb_dev_reads *reads, const uint64_t *cand_ids, uint64_t n_cand,
                               uint32_t max_read_len, uint32_t k, const uint32_t *edges, uint64_t n_edges,
                               uint32_t *mercy_out, uint64_t capacity, uint64_t *n_mercy_host, void *scratch,
                               size_t scratch_bytes) {
  *n_mercy_host = 0;
  if (n_cand == 0) return MHB_OK;
  // single segment: the look-up table lives behind the core scratch in the caller's buffer
  const size_t core = (mercy_core_scratch(n_cand, max_read_len) + 255) & ~(size_t)255;
  if (scratch_bytes < core + mhb_edge_lut_bytes()) return mhb_set_error(MHB_ERR_ARG, "scratch too small");
  void *lut = (char *)scratch + core;
  if (int rc = mhb_edge_lut_build(stream, edges, n_edges, k, lut)) return rc;
  const void *luts[1] = {lut};
  return mhb_mercy_edges_segs(stream, reads, cand_ids, n_cand, max_read_len, k, 1, &edges, &n_edges, luts, nullptr, mercy_out,
                              capacity, n_mercy_host, scratch, core);
}

// One process drives one GPU (the multi-GPU build is one process per GPU): kernel attributes, occupancy caches and the
// host-level arena are per process, so the device can be chosen once, before the first compute call; choosing the same
// device again is a no-op, switching afterwards is refused instead of silently dereferencing the other GPU's memory.
extern "C" int mhb_set_device(int device) {
  if (g_bound_device >= 0 && device != g_bound_device)
    return mhb_set_error(MHB_ERR_ARG, "this process is bound to CUDA device %d: mhb_set_device(%d) must be the first libmhb call "
                         "(one process per GPU)", g_bound_device, device);
  CK(cudaSetDevice(device));
  g_bound_device = device;
  g_sm_count = 0;
  return MHB_OK;
}
