// mhb_mgpu.cpp -- `count` on several GPUs of one node behind the file-level C ABI (mhb_count_run_multi, include/mhb.h).
//
// One worker PROCESS per GPU (forked before CUDA is touched; libmhb keeps per-process state: arena, kernel attributes),
// the same pipeline as megahit_b200/multigpu.py but with no Python, torch or NCCL underneath:
//   * records travel GPU -> GPU inside the fused partition + exchange kernel (mhb_partition_scatter storing into the
//     owners' receive buffers, opened through CUDA IPC);
//   * the small collectives (256-bin histograms, counters, IPC handles) go through one MAP_SHARED control block with a
//     process-shared barrier; the medium ones (tip edges, candidate reads, answer planes of the mercy searches, the
//     per-bucket tables) through files in /dev/shm written by one rank and read by the others;
//   * the plan of a stage (owner ranges from the all-gathered histograms) is computed by every rank from the same data
//     (plan_partition_host = the rule of k_plan_partition / multigpu.plan_ranges).
// The mercy searches are answered by the owners of the searched prefixes (mhb_mercy_probe_owned); the k_min SdBG is
// built in the same run because the solid edges are already on the devices.
//
// Reference behaviour reproduced: KmerCounter::Run (sorting/kmer_counter.cpp) + SeqToSdbg::Run with need_mercy
// (sorting/seq_to_sdbg.cpp) at k_min; files as edge_io_meta.h:25-44 / sdbg_meta.cpp:44-61 with num_files = n_gpus.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <string>
#include <vector>

#include "mhb.h"
#include "mhb_bits.cuh"
#include "mhb_internal.h"

using namespace mhb;

namespace {

constexpr int kMaxRanks = 16;

#define XINFO(...)                                                         \
  do {                                                                     \
    fprintf(stderr, "INFO  %-30s: %4d - ", "megahit_b200", __LINE__);      \
    fprintf(stderr, __VA_ARGS__);                                          \
  } while (0)

double now_s() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

// control block shared by the workers (anonymous MAP_SHARED mapping created before the fork)
struct Control {
  pthread_barrier_t bar;
  int world;
  char err[kMaxRanks][512];
  uint64_t hist[2][kMaxRanks][256];
  uint8_t ipc[2][kMaxRanks][64];
  uint64_t n_solid[kMaxRanks], n_tip[kMaxRanks], n_cand[kMaxRanks], n_mercy[kMaxRanks], n_records[kMaxRanks];
  uint64_t sdbg_totals[kMaxRanks][16];
  uint64_t has_tips[kMaxRanks];
};

struct Fail {
  std::string msg;
};
[[noreturn]] void fail(const char *fmt, ...) {
  char buf[480];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Fail{buf};
}
#define CKC(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)
#define CKM(call)                                     \
  do {                                                \
    if ((call) != MHB_OK) fail("%s", mhb_last_error()); \
  } while (0)

// device memory owned by a worker, released at its end
struct DevPool {
  std::vector<void *> ptrs;
  template <class T>
  T *get(size_t count) {
    void *p = nullptr;
    CKC(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256)));
    ptrs.push_back(p);
    return (T *)p;
  }
  void drop(void *p) {
    for (auto &q : ptrs)
      if (q == p) {
        cudaFree(q);
        q = nullptr;
      }
  }
  ~DevPool() {
    for (void *p : ptrs)
      if (p) cudaFree(p);
  }
};

// ---- exchange of medium-sized host data through /dev/shm files: rank r writes <base>.<tag>.<r>, the others read it ----
struct Exchange {
  std::string base;
  int rank, world;
  Control *C;
  void barrier() { pthread_barrier_wait(&C->bar); }
  std::string path(const char *tag, int r) const { return base + "." + tag + "." + std::to_string(r); }
  void publish(const char *tag, const void *data, size_t bytes) {
    FILE *f = fopen(path(tag, rank).c_str(), "wb");
    if (!f) fail("cannot create %s", path(tag, rank).c_str());
    if (bytes && fwrite(data, 1, bytes, f) != bytes) {
      fclose(f);
      fail("short write on %s", path(tag, rank).c_str());
    }
    fclose(f);
  }
  // whole file of rank r, or the byte range [off, off + bytes)
  std::vector<char> fetch(const char *tag, int r, size_t off = 0, size_t bytes = (size_t)-1) {
    FILE *f = fopen(path(tag, r).c_str(), "rb");
    if (!f) fail("cannot open %s", path(tag, r).c_str());
    if (bytes == (size_t)-1) {
      fseek(f, 0, SEEK_END);
      bytes = (size_t)ftell(f) - off;
    }
    fseek(f, (long)off, SEEK_SET);
    std::vector<char> v(bytes);
    if (bytes && fread(v.data(), 1, bytes, f) != bytes) {
      fclose(f);
      fail("short read on %s", path(tag, r).c_str());
    }
    fclose(f);
    return v;
  }
  void cleanup(const char *tag) { unlink(path(tag, rank).c_str()); }
};

// owner ranges from the all-gathered top-byte histograms: bound r = the byte value whose cumulative count is closest
// to r/world of the total, leaving at least one value for every later rank (k_plan_partition, multigpu.plan_ranges)
struct Plan {
  uint32_t bounds[kMaxRanks + 1];
  uint8_t owner[256];
  uint64_t recv_tot[kMaxRanks], send[kMaxRanks], my_off[kMaxRanks];
};
Plan plan_partition_host(const uint64_t (*hist)[256], int world, int rank) {
  Plan p;
  uint64_t cum[257];
  cum[0] = 0;
  for (int b = 0; b < 256; ++b) {
    uint64_t a = 0;
    for (int r = 0; r < world; ++r) a += hist[r][b];
    cum[b + 1] = cum[b] + a;
  }
  const uint64_t total = cum[256];
  p.bounds[0] = 0;
  for (int r = 1; r < world; ++r) {
    const uint32_t lo = p.bounds[r - 1] + 1, hi = 256 - (world - r);
    const uint64_t target = total * r / world;
    uint32_t best = lo;
    uint64_t bestd = ~0ull;
    for (uint32_t c = lo; c <= hi; ++c) {
      const uint64_t d = cum[c] > target ? cum[c] - target : target - cum[c];
      if (d < bestd) {
        bestd = d;
        best = c;
      }
    }
    p.bounds[r] = best;
  }
  p.bounds[world] = 256;
  for (int o = 0; o < world; ++o) {
    for (uint32_t b = p.bounds[o]; b < p.bounds[o + 1]; ++b) p.owner[b] = (uint8_t)o;
    uint64_t before = 0, tot = 0, mine = 0;
    for (int r = 0; r < world; ++r) {
      uint64_t s = 0;
      for (uint32_t b = p.bounds[o]; b < p.bounds[o + 1]; ++b) s += hist[r][b];
      if (r < rank) before += s;
      if (r == rank) mine = s;
      tot += s;
    }
    p.recv_tot[o] = tot;
    p.send[o] = mine;
    p.my_off[o] = before;
  }
  return p;
}

struct PeerBuf {
  void *mine = nullptr;
  void *peer[kMaxRanks] = {nullptr};
  size_t bytes = 0;
};

// extract-side records -> the rank owning their leading byte; returns the local receive buffer and the plan
Plan partition_and_exchange(Exchange &X, int stage, uint32_t *recs, uint64_t n, uint32_t words, int top_byte, uint64_t *d_hist,
                            void *d_ws, size_t ws_bytes, DevPool &pool, PeerBuf *pb) {
  Control *C = X.C;
  const int W = X.world, r = X.rank;
  CKC(cudaMemcpy(C->hist[stage][r], d_hist, 256 * 8, cudaMemcpyDeviceToHost));
  X.barrier();
  const Plan P = plan_partition_host(C->hist[stage], W, r);
  uint64_t mx = 0;
  for (int o = 0; o < W; ++o) mx = std::max(mx, P.recv_tot[o]);
  pb->bytes = (size_t)mx * words * 4 + 256;  // the same size on every rank
  CKM(mhb_dev_malloc(&pb->mine, pb->bytes));
  CKM(mhb_ipc_export(pb->mine, C->ipc[stage][r]));
  X.barrier();
  for (int o = 0; o < W; ++o) {
    if (o == r) pb->peer[o] = pb->mine;
    else CKM(mhb_ipc_open(C->ipc[stage][o], &pb->peer[o]));
  }
  uint64_t addr[256] = {0};
  for (int o = 0; o < W; ++o) addr[o] = (uint64_t)(uintptr_t)pb->peer[o] + P.my_off[o] * (uint64_t)words * 4;
  uint64_t *d_addr = pool.get<uint64_t>(256);
  uint8_t *d_lut = pool.get<uint8_t>(256);
  CKC(cudaMemcpy(d_addr, addr, sizeof(addr), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_lut, P.owner, 256, cudaMemcpyHostToDevice));
  CKM(mhb_partition_scatter(nullptr, recs, n, words, top_byte, d_lut, d_addr, d_ws, ws_bytes));
  CKC(cudaDeviceSynchronize());
  X.barrier();  // every rank's scatter has completed: my receive buffer is complete
  return P;
}

void close_peers(Exchange &X, PeerBuf *pb) {
  X.barrier();  // nobody reads or writes the buffers any more
  for (int o = 0; o < X.world; ++o)
    if (o != X.rank && pb->peer[o]) mhb_ipc_close(pb->peer[o]);
  X.barrier();
  if (pb->mine) mhb_dev_free(pb->mine);
  pb->mine = nullptr;
}

struct Job {
  uint32_t k;
  int32_t m;
  const uint32_t *bin;  // whole library (host, inherited by the workers)
  uint64_t n_reads;
  uint32_t read_len;
  std::string prefix;
};

void write_file(const std::string &path, const void *data, size_t bytes) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) fail("cannot open %s for writing", path.c_str());
  if (bytes && fwrite(data, 1, bytes, f) != bytes) {
    fclose(f);
    fail("write to %s failed", path.c_str());
  }
  fclose(f);
}

// ================================================================================================
// one worker = one GPU
// ================================================================================================
void worker(const Job &J, Exchange &X) {
  Control *C = X.C;
  const int W = X.world, r = X.rank;
  const uint32_t k = J.k, L = J.read_len;
  const int32_t m = J.m;
  CKC(cudaSetDevice(r));
  CKM(mhb_set_device(r));
  DevPool pool;
  const uint32_t WR = mhb_count_record_words(k), WE = mhb_words_per_edge(k), W2 = mhb_s2s_record_words(k);
  const uint32_t stride = 1 + div_ceil(L, 16), wpt = div_ceil(k, 16);
  const uint64_t per = (J.n_reads + W - 1) / W;
  const uint64_t r0 = std::min<uint64_t>(J.n_reads, (uint64_t)r * per), r1 = std::min<uint64_t>(J.n_reads, r0 + per);
  const uint64_t nr = r1 - r0;                                    // my block of reads
  const uint64_t n = L >= k + 1 ? nr * (uint64_t)(L - k) : 0;     // my edge records
  const uint32_t *my_bin = J.bin + r0 * stride;
  uint8_t cbytes[72], sbytes[72];
  const uint32_t n_csort = mhb_count_sort_bytes(k, cbytes), n_ssort = mhb_s2s_sort_bytes(k, sbytes);
  const int top = (int)(4 * WR - 1), top2 = (int)(4 * W2 - 1);

  // ---- reads to the device, extraction ----
  uint32_t *d_bin = pool.get<uint32_t>(nr * stride + 16);
  if (nr) CKC(cudaMemcpy(d_bin, my_bin, nr * stride * 4, cudaMemcpyHostToDevice));
  mhb_dev_reads reads;
  reads.bin = d_bin;
  reads.bin_words = nr * stride;
  reads.n_reads = nr;
  reads.fixed_len = L;
  reads.rec_off = nullptr;
  reads.edge_off = nullptr;
  uint32_t *d_a = pool.get<uint32_t>(n * WR + 16);
  uint64_t *d_hist = pool.get<uint64_t>(256);
  CKC(cudaMemset(d_hist, 0, 256 * 8));
  CKM(mhb_count_extract(nullptr, &reads, k, d_a, n, d_hist, top));
  size_t ws_bytes = mhb_sort_workspace_bytes(std::max<uint64_t>(n, 1), WR);
  void *d_ws = pool.get<char>(ws_bytes);
  PeerBuf pc;
  const Plan P = partition_and_exchange(X, 0, d_a, n, WR, top, d_hist, d_ws, ws_bytes, pool, &pc);
  pool.drop(d_a);
  pool.drop(d_ws);
  const uint64_t n_own = P.recv_tot[r];
  C->n_records[r] = n_own;

  // ---- count stage on the owned records ----
  const uint64_t cap = n_own / (uint64_t)std::max(1, m) + 1;
  uint32_t *d_edges = pool.get<uint32_t>(cap * WE + 16);
  uint8_t *d_aux = pool.get<uint8_t>(cap + 16);
  uint64_t *d_mul = pool.get<uint64_t>(65536);
  uint64_t *d_ns = pool.get<uint64_t>(8);
  CKC(cudaMemset(d_mul, 0, 65536 * 8));
  CKC(cudaMemset(d_ns, 0, 64));
  {
    uint32_t *d_tmp = pool.get<uint32_t>(n_own * WR + 16);
    if (mhb_count_hashed_supported(k, m) && !(getenv("MHB_COUNT_MODE") && !strcmp(getenv("MHB_COUNT_MODE"), "sort"))) {
      const size_t hb = mhb_count_hashed_workspace_bytes(std::max<uint64_t>(n_own, 1), k, m);
      void *d_h = pool.get<char>(hb);
      CKM(mhb_count_solid_hashed(nullptr, (uint32_t *)pc.mine, d_tmp, n_own, k, m, nullptr, d_edges, d_aux, cap, d_mul, d_ns, d_h, hb));
      CKC(cudaDeviceSynchronize());
      pool.drop(d_h);
    } else {
      const size_t sb = mhb_sort_workspace_bytes(std::max<uint64_t>(n_own, 1), WR), cb = mhb_count_solid_scratch_bytes(n_own);
      void *d_s = pool.get<char>(sb), *d_c = pool.get<char>(cb);
      int in_b = 0;
      CKM(mhb_sort_records_relaxed(nullptr, (uint32_t *)pc.mine, d_tmp, n_own, WR, cbytes, n_csort, nullptr, d_s, sb, &in_b));
      CKM(mhb_count_solid(nullptr, in_b ? d_tmp : (uint32_t *)pc.mine, n_own, k, m, d_edges, d_aux, cap, d_mul, d_ns, d_c, cb));
      CKC(cudaDeviceSynchronize());
      pool.drop(d_s);
      pool.drop(d_c);
    }
    pool.drop(d_tmp);
  }
  uint64_t n_solid = 0;
  CKC(cudaMemcpy(&n_solid, d_ns, 8, cudaMemcpyDeviceToHost));
  if (n_solid > cap) fail("internal: solid edges exceed capacity");
  C->n_solid[r] = n_solid;
  close_peers(X, &pc);  // the count records are gone: give the memory back before the SdBG stage
  {
    std::vector<uint64_t> h(65536);
    CKC(cudaMemcpy(h.data(), d_mul, 65536 * 8, cudaMemcpyDeviceToHost));
    X.publish("mul", h.data(), 65536 * 8);
  }

  // ---- mercy bookkeeping: tip edges of every rank -> per-read marks -> candidate reads ----
  uint64_t n_tip = 0;
  CKM(mhb_count_tip_edges(nullptr, d_aux, n_solid, &n_tip));
  {
    uint32_t *d_tips = pool.get<uint32_t>(n_tip * WE + 16);
    uint8_t *d_taux = pool.get<uint8_t>(n_tip + 16);
    CKC(cudaMemset(d_ns, 0, 8));
    CKM(mhb_compact_tip_edges(nullptr, d_edges, d_aux, n_solid, k, d_tips, d_taux, n_tip, d_ns));
    std::vector<char> buf(n_tip * (WE * 4 + 1));
    if (n_tip) {
      CKC(cudaMemcpy(buf.data(), d_tips, n_tip * WE * 4, cudaMemcpyDeviceToHost));
      CKC(cudaMemcpy(buf.data() + n_tip * WE * 4, d_taux, n_tip, cudaMemcpyDeviceToHost));
    }
    C->n_tip[r] = n_tip;
    X.publish("tips", buf.data(), buf.size());
    pool.drop(d_tips);
    pool.drop(d_taux);
  }
  X.barrier();
  uint64_t n_tip_all = 0;
  for (int o = 0; o < W; ++o) n_tip_all += C->n_tip[o];
  uint32_t *d_first = pool.get<uint32_t>(nr + 1), *d_last = pool.get<uint32_t>(nr + 1);
  uint64_t *d_cand = pool.get<uint64_t>(nr + 1);
  uint64_t n_cand = 0;
  {
    std::vector<uint32_t> te(n_tip_all * WE + 4);
    std::vector<uint8_t> ta(n_tip_all + 4);
    uint64_t at = 0;
    for (int o = 0; o < W; ++o) {
      const uint64_t c = C->n_tip[o];
      if (!c) continue;
      const std::vector<char> v = X.fetch("tips", o);
      memcpy(te.data() + at * WE, v.data(), c * WE * 4);
      memcpy(ta.data() + at, v.data() + c * WE * 4, c);
      at += c;
    }
    uint32_t *d_te = pool.get<uint32_t>(n_tip_all * WE + 16);
    uint8_t *d_ta = pool.get<uint8_t>(n_tip_all + 16);
    if (n_tip_all) {
      CKC(cudaMemcpy(d_te, te.data(), n_tip_all * WE * 4, cudaMemcpyHostToDevice));
      CKC(cudaMemcpy(d_ta, ta.data(), n_tip_all, cudaMemcpyHostToDevice));
    }
    const size_t tb = mhb_tipset_bytes(n_tip_all, k);
    void *d_tipset = pool.get<char>(tb);
    CKM(mhb_tipset_build(nullptr, d_te, d_ta, n_tip_all, k, d_tipset, tb, n_tip_all));
    CKM(mhb_count_mark_mercy(nullptr, &reads, k, d_tipset, tb, n_tip_all, d_first, d_last));
    const size_t cs = mhb_mercy_candidates_scratch_bytes(nr);
    void *d_cs = pool.get<char>(cs);
    CKM(mhb_mercy_candidates(nullptr, d_first, d_last, nr, d_cand, &n_cand, d_cs, cs));
    pool.drop(d_cs);
    pool.drop(d_tipset);
    pool.drop(d_te);
    pool.drop(d_ta);
  }
  C->n_cand[r] = n_cand;
  std::vector<uint64_t> cand_ids(n_cand);
  if (n_cand) CKC(cudaMemcpy(cand_ids.data(), d_cand, n_cand * 8, cudaMemcpyDeviceToHost));
  {  // number of reads with both marks set (the "(%d)" of the reference's log line) + my candidate reads, file orientation
    std::vector<uint32_t> f(nr), l(nr);
    if (nr) {
      CKC(cudaMemcpy(f.data(), d_first, nr * 4, cudaMemcpyDeviceToHost));
      CKC(cudaMemcpy(l.data(), d_last, nr * 4, cudaMemcpyDeviceToHost));
    }
    uint64_t ht = 0;
    for (uint64_t i = 0; i < nr; ++i) ht += f[i] != MHB_SENTINEL_OFFSET && l[i] != MHB_SENTINEL_OFFSET;
    C->has_tips[r] = ht;
    std::vector<uint32_t> cr(n_cand * stride);
    for (uint64_t c = 0; c < n_cand; ++c) memcpy(cr.data() + c * stride, my_bin + cand_ids[c] * stride, stride * 4);
    X.publish("cand", cr.data(), cr.size() * 4);
  }
  X.barrier();
  // ---- mercy edges: every rank answers, for the candidates of ALL ranks, the searches that land in its bucket range ----
  uint64_t n_cand_all = 0, cand_off[kMaxRanks + 1];
  for (int o = 0; o < W; ++o) {
    cand_off[o] = n_cand_all;
    n_cand_all += C->n_cand[o];
  }
  cand_off[W] = n_cand_all;
  uint64_t n_mercy = 0;
  uint32_t *d_all_edges = d_edges;  // solid + mercy edges, the sequences of the SdBG stage
  if (n_cand_all) {
    std::vector<uint32_t> all(n_cand_all * stride + 4);
    for (int o = 0; o < W; ++o)
      if (C->n_cand[o]) {
        const std::vector<char> v = X.fetch("cand", o);
        memcpy(all.data() + cand_off[o] * stride, v.data(), v.size());
      }
    uint32_t *d_call = pool.get<uint32_t>(n_cand_all * stride + 16);
    CKC(cudaMemcpy(d_call, all.data(), n_cand_all * stride * 4, cudaMemcpyHostToDevice));
    mhb_dev_reads greads = reads;
    greads.bin = d_call;
    greads.bin_words = n_cand_all * stride;
    greads.n_reads = n_cand_all;
    void *d_lut = pool.get<char>(mhb_edge_lut_bytes());
    CKM(mhb_edge_lut_build(nullptr, d_edges, n_solid, k, d_lut));
    const size_t pw_all = mhb_mercy_planes_words(n_cand_all, L);
    uint32_t *d_planes = pool.get<uint32_t>(pw_all);
    CKM(mhb_mercy_probe_owned(nullptr, &greads, nullptr, n_cand_all, L, k, d_edges, n_solid, d_lut, P.owner, (uint32_t)r, d_planes));
    std::vector<uint32_t> hp(pw_all);
    CKC(cudaMemcpy(hp.data(), d_planes, pw_all * 4, cudaMemcpyDeviceToHost));
    X.publish("planes", hp.data(), pw_all * 4);
    pool.drop(d_planes);
    pool.drop(d_lut);
    pool.drop(d_call);
    X.barrier();
    if (n_cand) {
      // the answers of every rank about MY candidates: rank s's file holds them at [cand_off[r], cand_off[r] + n_cand)
      const size_t pw1 = mhb_mercy_planes_words(1, L), pw_mine = pw1 * n_cand;
      std::vector<uint32_t> mine((size_t)W * pw_mine);
      for (int s = 0; s < W; ++s) {
        const std::vector<char> v = X.fetch("planes", s, cand_off[r] * pw1 * 4, pw_mine * 4);
        memcpy(mine.data() + (size_t)s * pw_mine, v.data(), pw_mine * 4);
      }
      uint32_t *d_mine = pool.get<uint32_t>(mine.size());
      CKC(cudaMemcpy(d_mine, mine.data(), mine.size() * 4, cudaMemcpyHostToDevice));
      const size_t ms = mhb_mercy_edges_scratch_bytes(n_cand, L) - mhb_edge_lut_bytes();
      void *d_ms = pool.get<char>(ms);
      CKM(mhb_mercy_count_planes(nullptr, &reads, d_cand, n_cand, L, k, d_mine, (uint32_t)W, pw_mine, &n_mercy, d_ms, ms));
      if (n_mercy) {
        if (n_solid + n_mercy > cap) {  // reads overlapping only at their ends: more mercy than solid edges
          uint32_t *big = pool.get<uint32_t>((n_solid + n_mercy) * WE + 16);
          CKC(cudaMemcpy(big, d_edges, n_solid * WE * 4, cudaMemcpyDeviceToDevice));
          d_all_edges = big;
        }
        CKM(mhb_mercy_edges_write(nullptr, &reads, d_cand, n_cand, L, k, d_all_edges + n_solid * WE, n_mercy, n_mercy, d_ms, ms));
        CKC(cudaDeviceSynchronize());
      }
      pool.drop(d_ms);
      pool.drop(d_mine);
    }
  }
  C->n_mercy[r] = n_mercy;
  X.barrier();

  // ---- SdBG stage over solid + mercy edges ----
  const uint64_t n_seqs = n_solid + n_mercy;
  uint64_t n_items = n_seqs * 6;
  mhb_dev_seqs seqs;
  memset(&seqs, 0, sizeof(seqs));
  seqs.words = d_all_edges;
  seqs.n_words = n_seqs * WE;
  seqs.n_seqs = n_seqs;
  seqs.fixed_len = k + 1;
  seqs.fixed_stride = WE;
  uint32_t *d_sa = pool.get<uint32_t>(n_items * W2 + 16);
  CKC(cudaMemset(d_hist, 0, 256 * 8));
  if (getenv("MHB_S2S_NO_PRUNE")) {
    CKM(mhb_s2s_extract(nullptr, &seqs, k, d_sa, n_items, d_hist, top2));
  } else {
    // the owned solid edges still carry the count stage's in/out flags: the $-items the emitter is certain to discard are
    // neither generated nor exchanged (mhb_s2s_extract_edges_pruned)
    CKC(cudaMemset(d_ns + 4, 0, 8));
    CKM(mhb_s2s_extract_edges_pruned(nullptr, d_all_edges, d_aux, n_seqs, n_solid, k, d_sa, n_items, d_ns + 4, d_hist, top2));
    uint64_t kept = 0;
    CKC(cudaMemcpy(&kept, d_ns + 4, 8, cudaMemcpyDeviceToHost));
    if (kept > n_items) fail("internal: pruned item count exceeds 6 per edge");
    n_items = kept;
  }
  ws_bytes = mhb_sort_workspace_bytes(std::max<uint64_t>(n_items, 1), W2);
  d_ws = pool.get<char>(ws_bytes);
  PeerBuf ps;
  const Plan P2 = partition_and_exchange(X, 1, d_sa, n_items, W2, top2, d_hist, d_ws, ws_bytes, pool, &ps);
  pool.drop(d_sa);
  pool.drop(d_ws);
  const uint64_t n_own2 = P2.recv_tot[r];
  std::vector<uint8_t> sdbg_bytes;
  std::vector<uint64_t> table(65536 * 4, 0);
  uint64_t totals[16] = {0};
  {
    uint32_t *d_tmp = pool.get<uint32_t>(n_own2 * W2 + 16);
    const size_t sb = mhb_sort_workspace_bytes(std::max<uint64_t>(n_own2, 1), W2), eb = mhb_s2s_emit_scratch_bytes(n_own2, k);
    void *d_s = pool.get<char>(sb);
    int in_b = 0;
    CKM(mhb_sort_records_relaxed(nullptr, (uint32_t *)ps.mine, d_tmp, n_own2, W2, sbytes, n_ssort, nullptr, d_s, sb, &in_b));
    pool.drop(d_s);

    void *d_e = pool.get<char>(eb);
    const uint64_t cap_b = n_own2 * (4ull + 4ull * wpt) + 16;
    uint8_t *d_out = pool.get<uint8_t>(cap_b);
    uint64_t *d_table = pool.get<uint64_t>(65536 * 4), *d_tot = pool.get<uint64_t>(16);
    CKC(cudaMemset(d_table, 0, 65536 * 32));
    CKC(cudaMemset(d_tot, 0, 128));
    CKM(mhb_s2s_emit(nullptr, in_b ? d_tmp : (uint32_t *)ps.mine, n_own2, k, d_out, cap_b, d_table, d_tot, d_e, eb));
    CKC(cudaMemcpy(totals, d_tot, sizeof(totals), cudaMemcpyDeviceToHost));
    if (totals[0] > cap_b) fail("internal: SdBG byte stream exceeds capacity");
    sdbg_bytes.resize(totals[0]);
    if (totals[0]) CKC(cudaMemcpy(sdbg_bytes.data(), d_out, totals[0], cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(table.data(), d_table, 65536 * 32, cudaMemcpyDeviceToHost));
  }
  close_peers(X, &ps);
  memcpy(C->sdbg_totals[r], totals, sizeof(totals));

  // ---- files: my bucket range of the edges and of the SdBG; the tables go to rank 0 ----
  std::vector<uint32_t> edges(n_solid * WE);
  if (n_solid) CKC(cudaMemcpy(edges.data(), d_edges, n_solid * WE * 4, cudaMemcpyDeviceToHost));
  write_file(J.prefix + ".edges." + std::to_string(r), edges.data(), edges.size() * 4);
  write_file(J.prefix + ".sdbg." + std::to_string(r), sdbg_bytes.data(), sdbg_bytes.size());
  {
    std::vector<int64_t> cnt(65536, 0);
    for (uint64_t i = 0; i < n_solid; ++i) cnt[edges[i * WE] >> 16]++;
    X.publish("ecnt", cnt.data(), 65536 * 8);
    X.publish("stab", table.data(), 65536 * 32);
    // `.cand`: my candidate reads in the REVERSED orientation KmerCounter holds them in (kmer_counter.cpp:387-401)
    std::vector<uint32_t> rec((size_t)n_cand * stride, 0);
    for (uint64_t c = 0; c < n_cand; ++c) {
      const uint32_t *src = my_bin + cand_ids[c] * stride;
      uint32_t *dst = rec.data() + c * stride;
      dst[0] = L;
      for (uint32_t i = 0; i < L; ++i) dst[1 + (i >> 4)] |= base_at(src + 1, L - 1 - i) << (30 - 2 * (i & 15));
    }
    X.publish("candrev", rec.data(), rec.size() * 4);
  }
  X.barrier();
  if (r == 0) {
    // merged P.edges.info (edge_io_meta.h:25-44): bucket -> (file = owner rank, offset inside that file, count)
    std::vector<std::vector<int64_t>> ec(W);
    uint64_t n_edges = 0;
    for (int o = 0; o < W; ++o) {
      const std::vector<char> v = X.fetch("ecnt", o);
      ec[o].assign((const int64_t *)v.data(), (const int64_t *)v.data() + 65536);
      n_edges += C->n_solid[o];
    }
    FILE *g = fopen((J.prefix + ".edges.info").c_str(), "w");
    if (!g) fail("cannot open %s.edges.info", J.prefix.c_str());
    fprintf(g, "kmer_size %u\nwords_per_edge %u\nnum_files %d\nnum_buckets %d\nnum_edges %llu\nis_sorted 1\n", k, WE, W,
            MHB_NUM_BUCKETS, (unsigned long long)n_edges);
    std::vector<int64_t> off(W, 0);
    for (int b = 0; b < MHB_NUM_BUCKETS; ++b) {
      int who = -1;
      for (int o = 0; o < W; ++o)
        if (ec[o][b]) {
          if (who >= 0) fail("bucket %d landed on two ranks", b);
          who = o;
        }
      if (who < 0) fprintf(g, "%d -1 0 0\n", b);
      else {
        fprintf(g, "%d %d %lld %lld\n", b, who, (long long)off[who], (long long)ec[who][b]);
        off[who] += ec[who][b];
      }
    }
    fclose(g);
    // merged P.sdbg_info (sdbg_meta.cpp:44-61): records ordered by (file, starting offset), unused ones last
    g = fopen((J.prefix + ".sdbg_info").c_str(), "w");
    if (!g) fail("cannot open %s.sdbg_info", J.prefix.c_str());
    fprintf(g, "k %u\nwords_per_tip_label %u\nnum_buckets %d\nnum_files %d\n", k, wpt, MHB_NUM_BUCKETS, W);
    int used = 0;
    uint64_t w_count[9] = {0}, items = 0, tips = 0, ones = 0;
    for (int o = 0; o < W; ++o) {
      const std::vector<char> v = X.fetch("stab", o);
      const uint64_t *t = (const uint64_t *)v.data();
      for (int b = 0; b < MHB_NUM_BUCKETS; ++b)
        if (t[4 * b + 1]) {
          fprintf(g, "%d %d %llu %llu %llu %llu\n", b, o, (unsigned long long)t[4 * b], (unsigned long long)t[4 * b + 1],
                  (unsigned long long)t[4 * b + 2], (unsigned long long)t[4 * b + 3]);
          ++used;
        }
      items += C->sdbg_totals[o][1];
      tips += C->sdbg_totals[o][2];
      for (int i = 0; i < 9; ++i) w_count[i] += C->sdbg_totals[o][4 + i];
      ones += C->sdbg_totals[o][13];
    }
    for (int i = used; i < MHB_NUM_BUCKETS; ++i) fprintf(g, "18446744073709551615 18446744073709551615 0 0 0 0\n");
    fclose(g);
    // P.cand (rank order = read order: the reads were dealt in contiguous blocks) and P.counting (global histogram)
    FILE *f = fopen((J.prefix + ".cand").c_str(), "wb");
    if (!f) fail("cannot open %s.cand", J.prefix.c_str());
    uint64_t n_cand_tot = 0, has_tips = 0, n_mercy_tot = 0;
    for (int o = 0; o < W; ++o) {
      const std::vector<char> v = X.fetch("candrev", o);
      if (!v.empty()) fwrite(v.data(), 1, v.size(), f);
      n_cand_tot += C->n_cand[o];
      has_tips += C->has_tips[o];
      n_mercy_tot += C->n_mercy[o];
    }
    fclose(f);
    std::vector<uint64_t> mul(65536, 0);
    for (int o = 0; o < W; ++o) {
      const std::vector<char> v = X.fetch("mul", o);
      const uint64_t *h = (const uint64_t *)v.data();
      for (int i = 0; i < 65536; ++i) mul[i] += h[i];
    }
    f = fopen((J.prefix + ".counting").c_str(), "w");
    if (!f) fail("cannot open %s.counting", J.prefix.c_str());
    for (int i = 1; i <= MHB_MAX_MUL; ++i) fprintf(f, "%d %lld\n", i, (long long)mul[i]);
    fclose(f);
    f = fopen((J.prefix + ".sdbg_fused").c_str(), "w");
    if (f) {
      fprintf(f, "%u 1 %d\n", k, W);
      fclose(f);
    }
    XINFO("Total number of candidate reads: %llu (%llu)\n", (unsigned long long)n_cand_tot, (unsigned long long)has_tips);
    XINFO("Total number of solid edges: %llu\n", (unsigned long long)n_edges);
    XINFO("Number of mercy edges: %llu\n", (unsigned long long)n_mercy_tot);
    XINFO("Number of $ A C G T A- C- G- T-:\n");
    XINFO("%llu %llu %llu %llu %llu %llu %llu %llu %llu\n", (unsigned long long)w_count[0], (unsigned long long)w_count[1],
          (unsigned long long)w_count[2], (unsigned long long)w_count[3], (unsigned long long)w_count[4],
          (unsigned long long)w_count[5], (unsigned long long)w_count[6], (unsigned long long)w_count[7],
          (unsigned long long)w_count[8]);
    XINFO("Total number of edges: %llu\n", (unsigned long long)items);
    XINFO("Total number of ONEs: %llu\n", (unsigned long long)ones);
    XINFO("Total number of $v edges: %llu\n", (unsigned long long)tips);
  }
  X.barrier();
  for (const char *t : {"mul", "tips", "cand", "planes", "ecnt", "stab", "candrev"}) X.cleanup(t);
}

}  // namespace

extern "C" int mhb_count_run_multi(const mhb_count_opts *o, int n_gpus) {
  if (!o || !o->read_lib_file || !o->read_lib_file[0]) return mhb_set_error(MHB_ERR_ARG, "No read library configuration file!");
  if (o->host_mem == 0) return mhb_set_error(MHB_ERR_ARG, "Please specify the host memory!");
  if (n_gpus <= 1) return mhb_count_run(o);
  if (n_gpus > kMaxRanks) return mhb_set_error(MHB_ERR_ARG, "at most %d GPUs of one node are supported", kMaxRanks);
  const std::string lib = o->read_lib_file, prefix = o->output_prefix ? o->output_prefix : "out";
  const double t0 = now_s();
  long long total_bases = 0, n_reads = 0;
  {
    std::ifstream is(lib + ".lib_info");
    if (!(is >> total_bases >> n_reads)) return mhb_set_error(MHB_ERR_IO, "cannot read %s.lib_info", lib.c_str());
  }
  std::vector<uint32_t> bin;
  {
    FILE *f = fopen((lib + ".bin").c_str(), "rb");
    if (!f) return mhb_set_error(MHB_ERR_IO, "cannot open %s.bin", lib.c_str());
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    bin.resize(((size_t)sz + 3) / 4 + 16, 0);
    const size_t got = sz ? fread(bin.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return mhb_set_error(MHB_ERR_IO, "short read on %s.bin", lib.c_str());
    bin.resize(((size_t)sz + 3) / 4);
  }
  // the partitioned build deals contiguous blocks of a FIXED-length library to the GPUs; anything else: one GPU
  const uint32_t L = (n_reads > 0 && !bin.empty()) ? bin[0] : 0;
  const uint64_t stride = 1 + div_ceil(L, 16);
  bool fixed = L > 0 && bin.size() == (uint64_t)n_reads * stride && o->k >= 12 && (uint64_t)n_reads >= (uint64_t)n_gpus;
  for (long long i = 0; fixed && i < n_reads; ++i) fixed = bin[(uint64_t)i * stride] == L;
  if (!fixed) {
    XINFO("variable-length or tiny library: running on one GPU\n");
    return mhb_count_run(o);
  }
  XINFO("%lld reads, %lld bases; k = %u, m = %d; %d GPUs\n", n_reads, total_bases, o->k, o->m, n_gpus);

  Control *C = (Control *)mmap(nullptr, sizeof(Control), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (C == MAP_FAILED) return mhb_set_error(MHB_ERR_NOMEM, "mmap of the control block failed");
  memset(C, 0, sizeof(Control));
  C->world = n_gpus;
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&C->bar, &ba, (unsigned)n_gpus);
  Job J{o->k, o->m, bin.data(), (uint64_t)n_reads, L, prefix};
  const std::string xbase = "/dev/shm/mhb_" + std::to_string((long long)getpid());
  std::vector<pid_t> pids;
  fflush(nullptr);
  for (int r = 0; r < n_gpus; ++r) {
    const pid_t p = fork();
    if (p < 0) {
      for (pid_t q : pids) kill(q, SIGKILL);
      munmap(C, sizeof(Control));
      return mhb_set_error(MHB_ERR_NOMEM, "fork failed");
    }
    if (p == 0) {
      Exchange X{xbase, r, n_gpus, C};
      int rc = 0;
      try {
        worker(J, X);
      } catch (const Fail &e) {
        snprintf(C->err[r], sizeof(C->err[r]), "rank %d: %s", r, e.msg.c_str());
        rc = 1;
      }
      fflush(nullptr);
      _exit(rc);
    }
    pids.push_back(p);
  }
  // a worker that dies would leave the others at a barrier: the first abnormal exit takes the rest down
  int failed = 0;
  for (size_t done = 0; done < pids.size(); ++done) {
    int st = 0;
    const pid_t p = wait(&st);
    if (p < 0) break;
    if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0) && !failed) {
      failed = 1;
      for (pid_t q : pids)
        if (q != p) kill(q, SIGKILL);
    }
  }
  int rc = MHB_OK;
  if (failed) {
    std::string msg;
    for (int r = 0; r < n_gpus; ++r)
      if (C->err[r][0]) msg += std::string(msg.empty() ? "" : "; ") + C->err[r];
    rc = mhb_set_error(MHB_ERR_CUDA, "multi-GPU count failed: %s", msg.empty() ? "a worker process died" : msg.c_str());
    for (int r = 0; r < n_gpus; ++r)
      for (const char *t : {"mul", "tips", "cand", "planes", "ecnt", "stab", "candrev"})
        unlink((xbase + "." + t + "." + std::to_string(r)).c_str());
  }
  pthread_barrier_destroy(&C->bar);
  munmap(C, sizeof(Control));
  if (!rc) XINFO("count (+ k_min SdBG) on %d GPUs done. Time elapsed: %.4f\n", n_gpus, now_s() - t0);
  return rc;
}
