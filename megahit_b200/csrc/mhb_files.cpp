// mhb_files.cpp -- file-level C ABI (include/mhb.h, layer 3): the `count` and `seq2sdbg` sub-commands on
// the reference's on-disk formats.  Host-side IO only; all sorting/counting/mercy-edge search/emission runs on the
// GPU through mhb_count_host / mhb_mercy_host / mhb_s2s_host.
//
// Formats follow voutcn/megahit v1.2.9 (paths relative to src/):
//   read library   sequence/io/sequence_lib.cpp:93-118, sequence/sequence_package.h:224-240
//   edges          sequence/io/edge/edge_io_meta.h:25-70, edge_writer.h:68-111, edge_reader.h:40-138
//   candidates     sorting/kmer_counter.cpp:383-401 ; counting: sorting/edge_counter.h:44-52
//   contigs        sequence/io/contig/contig_reader.h:52-119
//   SdBG           sdbg/sdbg_writer.cpp:25-79, sdbg/sdbg_meta.cpp:12-61
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <algorithm>
#include <fstream>
#include <string>
#include <vector>

#include "mhb.h"
#include "mhb_bits.cuh"
#include "mhb_internal.h"

using namespace mhb;

namespace {

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

bool read_file(const std::string &path, std::vector<uint32_t> *out, bool must_exist = true) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    if (must_exist) mhb_set_error(MHB_ERR_IO, "cannot open %s", path.c_str());
    return false;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(((size_t)sz + 3) / 4 + 4, 0);  // padded so the image can be handed to the device as is
  const size_t got = sz ? fread(out->data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  if (got != (size_t)sz) {
    mhb_set_error(MHB_ERR_IO, "short read on %s", path.c_str());
    return false;
  }
  out->resize(((size_t)sz + 3) / 4);
  return true;
}

// host container mirroring what SeqPackage holds for seq2sdbg: word-aligned package-orientation sequences
struct HostSeqs {
  std::vector<uint32_t> words;
  std::vector<uint64_t> word_off{0};
  std::vector<uint32_t> len;
  std::vector<uint16_t> mult;
  size_t size() const { return len.size(); }
  void append_packed(const uint32_t *w, uint32_t L, uint16_t m) {  // already left-aligned, tail bits may be dirty
    const uint32_t nw = div_ceil(L, 16);
    const size_t at = words.size();
    words.insert(words.end(), w, w + nw);
    if (L % 16) words[at + nw - 1] &= top_mask(2 * (L % 16));
    word_off.push_back(words.size());
    len.push_back(L);
    mult.push_back(m);
  }
  void append_ascii(const char *s, uint32_t L, bool reverse, uint16_t m) {  // sequence_package.h:245-273
    static const struct Map {
      uint8_t v[256];
      Map() {
        memset(v, 0, sizeof(v));
        const char *a = "ACGTNacgtn", *b = "0123201232";
        for (int i = 0; i < 10; ++i) v[(int)a[i]] = b[i] - '0';
      }
    } map;
    const uint32_t nw = div_ceil(L, 16);
    const size_t at = words.size();
    words.resize(at + nw, 0);
    for (uint32_t i = 0; i < L; ++i) {
      const uint8_t c = map.v[(uint8_t)s[reverse ? L - 1 - i : i]];
      words[at + (i >> 4)] |= (uint32_t)c << (30 - 2 * (i & 15));
    }
    word_off.push_back(words.size());
    len.push_back(L);
    mult.push_back(m);
  }
};

// ------------------------------------------------------------------------------------------------
// edges
// ------------------------------------------------------------------------------------------------
struct EdgeMeta {
  uint32_t kmer_size = 0, words_per_edge = 0, num_files = 0, num_buckets = 0;
  int64_t num_edges = 0;
  int is_sorted = 1;
  struct B {
    int file_id;
    int64_t off, cnt;
  };
  std::vector<B> buckets;
};

bool scan_field(std::istream &is, const char *name, long long *v) {  // utils.h ScanField
  std::string s;
  if (!(is >> s) || s != name || !(is >> *v)) {
    mhb_set_error(MHB_ERR_IO, "Invalid format. Expect field %s", name);
    return false;
  }
  return true;
}

int read_edge_meta(const std::string &prefix, EdgeMeta *m) {
  std::ifstream is(prefix + ".edges.info");
  if (!is) return mhb_set_error(MHB_ERR_IO, "cannot open %s.edges.info", prefix.c_str());
  long long v[6];
  const char *names[6] = {"kmer_size", "words_per_edge", "num_files", "num_buckets", "num_edges", "is_sorted"};
  for (int i = 0; i < 6; ++i)
    if (!scan_field(is, names[i], &v[i])) return MHB_ERR_IO;
  m->kmer_size = (uint32_t)v[0];
  m->words_per_edge = (uint32_t)v[1];
  m->num_files = (uint32_t)v[2];
  m->num_buckets = (uint32_t)v[3];
  m->num_edges = v[4];
  m->is_sorted = (int)v[5];
  m->buckets.resize(m->num_buckets);
  for (uint32_t i = 0; i < m->num_buckets; ++i) {
    long long id;
    is >> id >> m->buckets[i].file_id >> m->buckets[i].off >> m->buckets[i].cnt;
    if (!is || id != (long long)i) return mhb_set_error(MHB_ERR_IO, "Invalid format: bucket id not matched!");
    if (m->buckets[i].file_id >= (int)m->num_files)
      return mhb_set_error(MHB_ERR_IO, "Record ID %d is greater than number of files %u", m->buckets[i].file_id, m->num_files);
  }
  return MHB_OK;
}

// All edges in reader order (edge_reader.h:40-138): bucket order when sorted, file order otherwise.
int read_all_edges(const std::string &prefix, EdgeMeta *meta, std::vector<uint32_t> *edges) {
  if (int rc = read_edge_meta(prefix, meta)) return rc;
  const uint32_t W = meta->words_per_edge;
  std::vector<std::vector<uint32_t>> files(meta->num_files);
  for (uint32_t i = 0; i < meta->num_files; ++i)
    if (!read_file(prefix + ".edges." + std::to_string(i), &files[i])) return MHB_ERR_IO;
  edges->clear();
  if (!meta->is_sorted) {
    if (files.empty() || files[0].size() < (size_t)meta->num_edges * W) return mhb_set_error(MHB_ERR_IO, "%s.edges.0 is truncated", prefix.c_str());
    edges->assign(files[0].begin(), files[0].begin() + (size_t)meta->num_edges * W);
    return MHB_OK;
  }
  edges->reserve((size_t)meta->num_edges * W);
  for (const auto &b : meta->buckets) {
    if (b.file_id < 0 || b.cnt == 0) continue;
    const auto &f = files[b.file_id];
    if ((size_t)(b.off + b.cnt) * W > f.size()) return mhb_set_error(MHB_ERR_IO, "%s.edges.%d is truncated", prefix.c_str(), b.file_id);
    edges->insert(edges->end(), f.begin() + (size_t)b.off * W, f.begin() + (size_t)(b.off + b.cnt) * W);
  }
  return MHB_OK;
}

int write_edges(const std::string &prefix, uint32_t k, const uint32_t *edges, uint64_t n) {
  const uint32_t W = words_per_edge(k);
  FILE *f = fopen((prefix + ".edges.0").c_str(), "wb");
  if (!f) return mhb_set_error(MHB_ERR_IO, "cannot open %s.edges.0 for writing", prefix.c_str());
  if (n && fwrite(edges, 4 * (size_t)W, n, f) != n) {
    fclose(f);
    return mhb_set_error(MHB_ERR_IO, "write to %s.edges.0 failed", prefix.c_str());
  }
  fclose(f);
  std::vector<int64_t> cnt(MHB_NUM_BUCKETS, 0), off(MHB_NUM_BUCKETS, 0);
  for (uint64_t i = 0; i < n; ++i) cnt[edges[i * W] >> 16]++;
  int64_t acc = 0;
  for (int b = 0; b < MHB_NUM_BUCKETS; ++b) {
    off[b] = acc;
    acc += cnt[b];
  }
  FILE *g = fopen((prefix + ".edges.info").c_str(), "w");
  if (!g) return mhb_set_error(MHB_ERR_IO, "cannot open %s.edges.info for writing", prefix.c_str());
  fprintf(g, "kmer_size %u\nwords_per_edge %u\nnum_files 1\nnum_buckets %d\nnum_edges %lld\nis_sorted 1\n", k, W,
          MHB_NUM_BUCKETS, (long long)n);
  for (int b = 0; b < MHB_NUM_BUCKETS; ++b) {
    if (cnt[b]) fprintf(g, "%d 0 %lld %lld\n", b, (long long)off[b], (long long)cnt[b]);
    else fprintf(g, "%d -1 0 0\n", b);
  }
  fclose(g);
  return MHB_OK;
}

// ------------------------------------------------------------------------------------------------
// contigs (contig_reader.h:52-119): FASTA with "flag=F multi=M len=N" comments
// ------------------------------------------------------------------------------------------------
int read_contigs(const std::string &path, uint32_t min_len, uint32_t k_from, uint32_t k_to, bool reverse,
                 HostSeqs *out, int64_t *n_read, unsigned discard_flag = 0) {
  *n_read = 0;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return MHB_OK;  // the reference opens a missing file as an empty stream
  {
    const int c0 = fgetc(f), c1 = fgetc(f);
    if (c0 == 0x1f && c1 == 0x8b) {
      fclose(f);
      return mhb_set_error(MHB_ERR_IO, "%s is gzip-compressed: the GPU seq2sdbg reads plain FASTA contigs only", path.c_str());
    }
    rewind(f);
  }
  const bool extend_loop = k_from < k_to;
  std::string header, seq, line;
  char *buf = nullptr;
  size_t cap = 0;
  ssize_t got;
  bool have = false;
  auto flush = [&]() {
    if (!have) return;
    have = false;
    if (seq.size() < min_len) return;
    const size_t sp = header.find_first_of(" \t");
    const std::string comment = sp == std::string::npos ? "" : header.substr(header.find_first_not_of(" \t", sp));
    const unsigned flag = comment.size() > 5 ? (unsigned)(comment[5] - '0') : 0u;
    if (discard_flag & flag) return;  // contig_reader.h:66-69
    const double m = comment.size() > 13 ? atof(comment.c_str() + 13) : 0.0;
    const uint16_t mult = (uint16_t)(int32_t)(m + .5);
    if (extend_loop && (flag & 2u)) {  // contig_flag::kLoop, contig_reader.h:73-86
      if (seq.size() < k_to + 1u) return;
      std::string ss(seq);
      for (uint32_t i = k_from; i < k_to; ++i) ss.push_back(ss[i]);
      out->append_ascii(ss.data(), (uint32_t)ss.size(), reverse, mult);
    } else {
      out->append_ascii(seq.data(), (uint32_t)seq.size(), reverse, mult);
    }
    ++*n_read;
  };
  while ((got = getline(&buf, &cap, f)) >= 0) {
    while (got > 0 && (buf[got - 1] == '\n' || buf[got - 1] == '\r')) --got;
    if (got > 0 && buf[0] == '>') {
      flush();
      header.assign(buf + 1, got - 1);
      seq.clear();
      have = true;
    } else if (have) {
      seq.append(buf, got);
    }
  }
  flush();
  free(buf);
  fclose(f);
  return MHB_OK;
}

// sdbg_writer.cpp:25-79 + sdbg_meta.cpp:44-61: one file, buckets in id order
int write_sdbg_single(const std::string &prefix, uint32_t k, uint32_t words_per_tip_label, uint64_t n_items, uint64_t n_bytes,
                      const uint8_t *bytes, const uint64_t *bucket_table) {
  int rc = MHB_OK;
  FILE *f = fopen((prefix + ".sdbg.0").c_str(), "wb");
  if (!f) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.sdbg.0", prefix.c_str());
  else {
    if (n_bytes && fwrite(bytes, 1, n_bytes, f) != n_bytes) rc = mhb_set_error(MHB_ERR_IO, "write failed");
    fclose(f);
  }
  if (!rc) {
    FILE *g = fopen((prefix + ".sdbg_info").c_str(), "w");
    if (!g) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.sdbg_info", prefix.c_str());
    else {
      fprintf(g, "k %u\nwords_per_tip_label %u\nnum_buckets %d\nnum_files %d\n", k, words_per_tip_label, MHB_NUM_BUCKETS,
              n_items ? 1 : 0);
      int empty = 0;
      for (int b = 0; b < MHB_NUM_BUCKETS; ++b) {
        const uint64_t *t = bucket_table + 4 * (size_t)b;
        if (t[1]) fprintf(g, "%d 0 %llu %llu %llu %llu\n", b, (unsigned long long)t[0], (unsigned long long)t[1],
                          (unsigned long long)t[2], (unsigned long long)t[3]);
        else ++empty;
      }
      for (int i = 0; i < empty; ++i) fprintf(g, "18446744073709551615 18446744073709551615 0 0 0 0\n");
      fclose(g);
    }
  }
  return rc;
}

}  // namespace

// ================================================================================================
// count
// ================================================================================================
extern "C" int mhb_count_run(const mhb_count_opts *o) {
  if (!o || !o->read_lib_file || !o->read_lib_file[0]) return mhb_set_error(MHB_ERR_ARG, "No read library configuration file!");
  if (o->host_mem == 0) return mhb_set_error(MHB_ERR_ARG, "Please specify the host memory!");
  const std::string lib = o->read_lib_file, prefix = o->output_prefix ? o->output_prefix : "out";
  const double t0 = now_s();
  long long total_bases = 0, n_reads = 0;
  {
    std::ifstream is(lib + ".lib_info");
    if (!(is >> total_bases >> n_reads)) return mhb_set_error(MHB_ERR_IO, "cannot read %s.lib_info", lib.c_str());
  }
  std::vector<uint32_t> bin;
  if (!read_file(lib + ".bin", &bin)) return MHB_ERR_IO;
  XINFO("%lld reads, %lld bases; k = %u, m = %d\n", n_reads, total_bases, o->k, o->m);

  mhb_count_args a;
  memset(&a, 0, sizeof(a));
  a.k = o->k;
  a.m = o->m;
  a.bin = bin.data();
  a.bin_words = bin.size();
  a.n_reads = (uint64_t)n_reads;
  a.want_mercy = 1;
  std::vector<char> resbuf(sizeof(mhb_count_result));
  mhb_count_result *res = reinterpret_cast<mhb_count_result *>(resbuf.data());
  if (int rc = mhb_count_host(&a, res)) return rc;
  XINFO("GPU count: %llu edge records, h2d %.2f ms, extract %.2f ms, sort %.2f ms (%u passes), count %.2f ms, mercy %.2f ms, d2h %.2f ms\n",
        (unsigned long long)res->n_edge_records, res->t_h2d_ms, res->t_extract_ms, res->t_sort_ms, res->n_sort_passes,
        res->t_count_ms, res->t_mercy_ms, res->t_d2h_ms);

  int rc = write_edges(prefix, o->k, res->edges, res->n_solid);
  if (!rc) {  // kmer_counter.cpp:383-401: candidate reads, in the reversed orientation KmerCounter holds them
    FILE *f = fopen((prefix + ".cand").c_str(), "wb");
    if (!f) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.cand", prefix.c_str());
    else {
      uint64_t r = 0;
      size_t pos = 0;
      std::vector<uint32_t> rec;
      for (uint64_t c = 0; c < res->n_cand; ++c) {
        while (r < res->cand_ids[c]) {
          pos += 1 + div_ceil(bin[pos], 16);
          ++r;
        }
        const uint32_t L = bin[pos], nw = div_ceil(L, 16);
        rec.assign(nw + 1, 0);
        rec[0] = L;
        for (uint32_t i = 0; i < L; ++i)  // sequence_package.h:284-295 (reverse, no complement)
          rec[1 + (i >> 4)] |= base_at(&bin[pos + 1], L - 1 - i) << (30 - 2 * (i & 15));
        fwrite(rec.data(), 4, nw + 1, f);
      }
      fclose(f);
    }
  }
  if (!rc) {
    FILE *f = fopen((prefix + ".counting").c_str(), "w");
    if (!f) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.counting", prefix.c_str());
    else {
      for (int i = 1; i <= MHB_MAX_MUL; ++i) fprintf(f, "%d %lld\n", i, (long long)res->counting[i]);
      fclose(f);
    }
  }
  XINFO("Total number of candidate reads: %llu (%llu)\n", (unsigned long long)res->n_cand, (unsigned long long)res->n_has_tips);
  XINFO("Total number of solid edges: %llu\n", (unsigned long long)res->n_solid);
  XINFO("count done. Time elapsed: %.4f\n", now_s() - t0);
  mhb_free(res->edges);
  mhb_free(res->cand_ids);
  return rc;
}

// ================================================================================================
// seq2sdbg
// ================================================================================================
extern "C" int mhb_seq2sdbg_run(const mhb_seq2sdbg_opts *o) {
  auto S = [](const char *s) { return std::string(s ? s : ""); };
  if (!o) return mhb_set_error(MHB_ERR_ARG, "null options");
  const std::string input = S(o->input_prefix), contig = S(o->contig), bubble = S(o->bubble), addi = S(o->addi_contig),
                    local = S(o->local_contig), prefix = S(o->output_prefix);
  if (input.empty() && contig.empty() && addi.empty()) return mhb_set_error(MHB_ERR_ARG, "No input files!");
  if (o->k < 9) return mhb_set_error(MHB_ERR_ARG, "kmer size must be >= 9!");
  if (o->host_mem == 0) return mhb_set_error(MHB_ERR_ARG, "Please specify the host memory!");
  const uint32_t k = o->k;
  const double t0 = now_s();

  // A multi-GPU `count` (mhb_count_run_multi) has already built this very graph - same prefix, same k, mercy edges
  // included - while the solid edges were on the devices, and left a marker: nothing to do.
  if (!input.empty() && input == prefix && contig.empty() && addi.empty() && local.empty() && o->need_mercy) {
    std::ifstream mk(prefix + ".sdbg_fused");
    unsigned mk_k = 0, mk_mercy = 0, mk_files = 0;
    if (mk && (mk >> mk_k >> mk_mercy >> mk_files) && mk_k == k && mk_mercy == 1) {
      bool all = std::ifstream(prefix + ".sdbg_info").good();
      for (unsigned i = 0; i < mk_files; ++i) all = all && std::ifstream(prefix + ".sdbg." + std::to_string(i)).good();
      if (all) {
        XINFO("SdBG for k = %u was built by the %u-GPU count stage; nothing to do\n", k, mk_files);
        XINFO("seq2sdbg done. Time elapsed: %.4f\n", now_s() - t0);
        return MHB_OK;
      }
    }
  }

  HostSeqs seqs;
  if (!input.empty()) {  // seq_to_sdbg.cpp:424-434
    EdgeMeta meta;
    std::vector<uint32_t> edges;
    if (int rc = read_all_edges(input, &meta, &edges)) return rc;
    if (meta.kmer_size != k) return mhb_set_error(MHB_ERR_ARG, "edges were built for k=%u, not %u", meta.kmer_size, k);
    const uint32_t W = meta.words_per_edge;
    const size_t n = edges.size() / W;
    seqs.words.reserve(n * div_ceil(k + 1, 16) * 5 / 4);
    for (size_t i = 0; i < n; ++i) seqs.append_packed(&edges[i * W], k + 1, (uint16_t)(edges[i * W + W - 1] & 0xFFFF));
    XINFO("Read %zu edges.\n", n);
    if (o->need_mercy) {  // :436-450
      const double t1 = now_s();
      std::vector<uint32_t> cand;
      if (!meta.is_sorted) return mhb_set_error(MHB_ERR_ARG, "--need_mercy needs sorted edges");
      read_file(input + ".cand", &cand, false);
      // GenMercyEdges (seq_to_sdbg.cpp:171-357) on the device: k_mercy_probe / k_mercy_emit through the host-level ABI
      uint32_t *mercy = nullptr;
      uint64_t nm = 0, nr = 0;
      if (int rc = mhb_mercy_host(k, edges.data(), n, cand.data(), cand.size(), &mercy, &nm, &nr)) return rc;
      for (uint64_t i = 0; i < nm; ++i) seqs.append_packed(mercy + i * W, k + 1, 1);
      mhb_free(mercy);
      XINFO("Number of reads: %lld, Number of mercy edges: %lld\n", (long long)nr, (long long)nm);
      XINFO("Adding mercy Done. Time elapsed: %.4f\n", now_s() - t1);
    }
  }
  int64_t nr = 0;
  if (!contig.empty()) {  // :452-476
    if (int rc = read_contigs(contig, k + 1, o->k_from, k, true, &seqs, &nr)) return rc;
    XINFO("Read %lld contigs from %s.\n", (long long)nr, contig.c_str());
    if (int rc = read_contigs(bubble, k + 1, 0, 0, true, &seqs, &nr)) return rc;
    XINFO("Read %lld contigs from %s.\n", (long long)nr, bubble.c_str());
  }
  if (!addi.empty()) {
    if (int rc = read_contigs(addi, k + 1, 0, 0, true, &seqs, &nr)) return rc;
    XINFO("Read %lld contigs from %s.\n", (long long)nr, addi.c_str());
  }
  if (!local.empty()) {
    if (int rc = read_contigs(local, k + 1, 0, 0, true, &seqs, &nr)) return rc;
    XINFO("Read %lld contigs from %s.\n", (long long)nr, local.c_str());
  }

  mhb_s2s_args a;
  memset(&a, 0, sizeof(a));
  a.k = k;
  if (seqs.words.empty()) seqs.words.push_back(0);
  a.words = seqs.words.data();
  a.word_off = seqs.word_off.data();
  a.len = seqs.len.data();
  a.mult = seqs.mult.data();
  a.n_seqs = seqs.size();
  std::vector<char> resbuf(sizeof(mhb_s2s_result));
  mhb_s2s_result *res = reinterpret_cast<mhb_s2s_result *>(resbuf.data());
  if (int rc = mhb_s2s_host(&a, res)) return rc;
  XINFO("GPU seq2sdbg: %llu sort items, extract %.2f ms, sort %.2f ms, emit %.2f ms\n", (unsigned long long)res->n_records,
        res->t_extract_ms, res->t_sort_ms, res->t_emit_ms);

  const int rc = write_sdbg_single(prefix, k, res->words_per_tip_label, res->n_items, res->n_bytes, res->bytes, res->bucket_table);
  XINFO("Number of $ A C G T A- C- G- T-:\n");
  XINFO("");
  for (int i = 0; i < 9; ++i) fprintf(stderr, "%llu ", (unsigned long long)res->w_count[i]);
  fprintf(stderr, "\n");
  XINFO("Total number of edges: %llu\n", (unsigned long long)res->n_items);
  XINFO("Total number of ONEs: %llu\n", (unsigned long long)res->ones_in_last);
  XINFO("Total number of $v edges: %llu\n", (unsigned long long)res->n_tips);
  XINFO("seq2sdbg done. Time elapsed: %.4f\n", now_s() - t0);
  mhb_free(res->bytes);
  return rc;
}

// ================================================================================================
// read2sdbg (main_read2sdbg, main_sdbg_build.cpp:88-156)
// ================================================================================================
extern "C" int mhb_read2sdbg_run(const mhb_read2sdbg_opts *o) {
  if (!o || !o->read_lib_file || !o->read_lib_file[0]) return mhb_set_error(MHB_ERR_ARG, "No input file!");
  if (o->host_mem == 0) return mhb_set_error(MHB_ERR_ARG, "Please specify the host memory!");
  const std::string lib = o->read_lib_file, prefix = o->output_prefix ? o->output_prefix : "out";
  const double t0 = now_s();
  long long total_bases = 0, n_reads = 0;
  {
    std::ifstream is(lib + ".lib_info");
    if (!(is >> total_bases >> n_reads)) return mhb_set_error(MHB_ERR_IO, "cannot read %s.lib_info", lib.c_str());
  }
  std::vector<uint32_t> bin;
  if (!read_file(lib + ".bin", &bin)) return MHB_ERR_IO;
  XINFO("%lld reads, %lld total bases; k = %u, m = %d, need_mercy = %d\n", n_reads, total_bases, o->k, o->m, o->need_mercy);
  // the candidate files stage 1 hands to stage 2 inside the reference process (read_to_sdbg_s1.cpp:111-126: 1, 2, 4 .. 64
  // files by read count); here the candidates never leave the device (three bit planes), the files are created empty so
  // that whatever cleans up after the reference finds them
  int n_mercy_files = 1;
  while (n_mercy_files * 10485760LL < n_reads && n_mercy_files < 64) n_mercy_files <<= 1;
  for (int i = 0; i < n_mercy_files; ++i) {
    FILE *f = fopen((prefix + ".mercy_cand." + std::to_string(i)).c_str(), "wb");
    if (!f) return mhb_set_error(MHB_ERR_IO, "cannot open %s.mercy_cand.%d", prefix.c_str(), i);
    fclose(f);
  }
  mhb_build_args a;
  memset(&a, 0, sizeof(a));
  a.k = o->k;
  a.m = o->m;
  a.bin = bin.data();
  a.bin_words = bin.size();
  a.n_reads = (uint64_t)n_reads;
  a.need_mercy = o->need_mercy;
  mhb_build_result res;
  if (int rc = mhb_read2sdbg_host(&a, &res)) return rc;
  XINFO("GPU read2sdbg: %llu (k+1)-mer positions, bucket partition %.2f ms, kmsort %.2f ms, %llu mercy edges, %llu sort items, total %.2f ms\n",
        (unsigned long long)res.n_edge_records, res.t_count_ms, res.t_mercy_ms, (unsigned long long)res.n_mercy,
        (unsigned long long)res.n_sort_items, res.t_total_ms);
  int rc = MHB_OK;
  if (o->m > 1) {  // Read2SdbgS1::Lv0Postprocess, read_to_sdbg_s1.cpp:557-566
    FILE *f = fopen((prefix + ".counting").c_str(), "w");
    if (!f) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.counting", prefix.c_str());
    else {
      for (int i = 1; i <= MHB_MAX_MUL; ++i) fprintf(f, "%d %lld\n", i, (long long)res.counting[i]);
      fclose(f);
    }
    if (o->need_mercy) XINFO("Number mercy: %llu\n", (unsigned long long)res.n_mercy);
  }
  if (!rc) rc = write_sdbg_single(prefix, o->k, res.words_per_tip_label, res.n_items, res.n_bytes, res.bytes, res.bucket_table);
  XINFO("Number of $ A C G T A- C- G- T-:\n");
  XINFO("");
  for (int i = 0; i < 9; ++i) fprintf(stderr, "%llu ", (unsigned long long)res.w_count[i]);
  fprintf(stderr, "\n");
  XINFO("Total number of edges: %llu\n", (unsigned long long)res.n_items);
  XINFO("Total number of ONEs: %llu\n", (unsigned long long)res.ones_in_last);
  XINFO("Total number of $v edges: %llu\n", (unsigned long long)res.n_tips);
  XINFO("read2sdbg done. Time elapsed: %.4f\n", now_s() - t0);
  mhb_free(res.bytes);
  mhb_free(res.bucket_table);
  mhb_free(res.counting);
  return rc;
}

// ================================================================================================
// iterate (main_iterate, main_iterate.cpp:196-221)
// ================================================================================================
extern "C" int mhb_iterate_run(const mhb_iterate_opts *o) {
  auto S = [](const char *s) { return std::string(s ? s : ""); };
  if (!o) return mhb_set_error(MHB_ERR_ARG, "null options");
  const std::string contig = S(o->contig_file), bubble = S(o->bubble_file), reads = S(o->read_file), prefix = S(o->output_prefix);
  if (contig.empty()) return mhb_set_error(MHB_ERR_ARG, "No contig file!");
  if (bubble.empty()) return mhb_set_error(MHB_ERR_ARG, "No bubble file!");
  if (reads.empty()) return mhb_set_error(MHB_ERR_ARG, "No reads file!");
  if (o->k == 0) return mhb_set_error(MHB_ERR_ARG, "Invalid kmer size!");
  if (o->step == 0 || o->step > 28 || (o->step & 1)) return mhb_set_error(MHB_ERR_ARG, "Invalid step size!");
  if (prefix.empty()) return mhb_set_error(MHB_ERR_ARG, "No output prefix!");
  const double t0 = now_s();
  // the flank index reads contigs and bubbles in file orientation, without the standalone and loop ones
  // (async_sequence_reader.h:82-101: SetDiscardFlag(kLoop | kStandalone), reverse = false)
  HostSeqs seqs;
  int64_t nr = 0;
  if (int rc = read_contigs(contig, 0, 0, 0, false, &seqs, &nr, 3u)) return rc;
  XINFO("Read %lld contigs\n", (long long)nr);
  if (int rc = read_contigs(bubble, 0, 0, 0, false, &seqs, &nr, 3u)) return rc;
  XINFO("Read %lld contigs\n", (long long)nr);
  std::vector<uint32_t> bin;
  if (!read_file(reads, &bin)) return MHB_ERR_IO;
  uint64_t n_reads = 0;
  for (size_t pos = 0; pos < bin.size(); ++n_reads) pos += 1 + div_ceil(bin[pos], 16);  // binary_reader.h:23-53
  mhb_iterate_args a;
  memset(&a, 0, sizeof(a));
  a.k = o->k;
  a.step = o->step;
  if (seqs.words.empty()) seqs.words.push_back(0);
  a.contig_words = seqs.words.data();
  a.contig_word_off = seqs.word_off.data();
  a.contig_len = seqs.len.data();
  a.n_contigs = seqs.size();
  a.bin = bin.data();
  a.bin_words = bin.size();
  a.n_reads = n_reads;
  mhb_iterate_result res;
  if (int rc = mhb_iterate_host(&a, &res)) return rc;
  XINFO("Number of flank kmers: %llu\n", (unsigned long long)res.n_flanks);
  XINFO("Total: %llu, aligned: %llu. Iterative edges: %llu\n", (unsigned long long)n_reads,
        (unsigned long long)res.n_aligned_reads, (unsigned long long)res.n_edges);
  int rc = MHB_OK;
  FILE *f = fopen((prefix + ".edges.0").c_str(), "wb");
  if (!f) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.edges.0 for writing", prefix.c_str());
  else {
    if (res.n_edges && fwrite(res.edges, 4 * (size_t)res.words_per_edge, res.n_edges, f) != res.n_edges)
      rc = mhb_set_error(MHB_ERR_IO, "write to %s.edges.0 failed", prefix.c_str());
    fclose(f);
  }
  if (!rc) {  // edge_writer.h:94-99 / edge_io_meta.h:25-44, unordered
    FILE *g = fopen((prefix + ".edges.info").c_str(), "w");
    if (!g) rc = mhb_set_error(MHB_ERR_IO, "cannot open %s.edges.info for writing", prefix.c_str());
    else {
      fprintf(g, "kmer_size %u\nwords_per_edge %u\nnum_files 1\nnum_buckets 0\nnum_edges %llu\nis_sorted 0\n", o->k + o->step,
              res.words_per_edge, (unsigned long long)res.n_edges);
      fclose(g);
    }
  }
  XINFO("iterate done. Time elapsed: %.4f (GPU %.2f ms)\n", now_s() - t0, res.t_total_ms);
  mhb_free(res.edges);
  return rc;
}
