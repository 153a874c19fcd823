/*
 * mhb_oracle -- TEST INFRASTRUCTURE ONLY.
 *
 * A deliberately slow, plain-C, one-base-at-a-time restatement of the two sorting engines on the
 * SdBG-construction path of voutcn/megahit v1.2.9 (`megahit_core count` and `megahit_core seq2sdbg`).
 * It exists to CHECK the CUDA path; nothing in megahit_b200/ may link, import or call it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity pin: this restatement is itself checked (tests/test_oracle_vs_reference.py) against outputs of
 * the unmodified reference binary (oracle/_ref/megahit_core_ref, built by oracle/Makefile from the
 * sources under /root/reference) committed as fixtures under tests/golden/ by oracle/gen_golden.py.
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it follows.
 */
#ifndef MHB_ORACLE_H
#define MHB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHBO_NUM_BUCKETS 65536
#define MHBO_MAX_MUL 65535

/* A set of 2-bit packed sequences in "package orientation" (what SeqPackage holds after loading:
 * reads already reversed for count, contigs reversed, edges as is).  Base i of a sequence sits at
 * bits 31-2(i%16)..30-2(i%16) of word i/16 (sequence_package.h:73).  Unlike SeqPackage every sequence
 * starts on a word boundary; nothing on the path depends on the physical packing. */
typedef struct {
  const uint32_t *words;
  const uint64_t *word_off; /* n + 1 entries */
  const uint32_t *len;      /* n entries (bases) */
  uint64_t n;
} mhbo_seqs;

typedef struct {
  uint64_t n_records;      /* (k+1)-mer occurrences = sum max(0, len-k) */
  uint64_t n_distinct;     /* distinct canonical (k+1)-mers */
  uint64_t n_solid;        /* multiplicity >= m */
  uint32_t words_per_edge; /* ceil((2(k+1)+16)/32) */
  uint32_t *edges;         /* n_solid * words_per_edge, ascending; malloc'ed */
  uint32_t *first_0_out;   /* per read, 0xFFFFFFFF = unset; malloc'ed */
  uint32_t *last_0_in;     /* per read, 0xFFFFFFFF = unset; malloc'ed */
  int64_t counting[MHBO_MAX_MUL + 1];
} mhbo_count_out;

typedef struct {
  uint64_t n_records;       /* sort items generated */
  uint64_t n_items;         /* SdBG items emitted */
  uint32_t words_per_tip_label;
  uint64_t bucket_items[MHBO_NUM_BUCKETS];
  uint64_t bucket_tips[MHBO_NUM_BUCKETS];
  uint64_t bucket_large_mul[MHBO_NUM_BUCKETS];
  uint64_t bucket_byte_off[MHBO_NUM_BUCKETS + 1];
  uint64_t w_count[9];
  uint64_t ones_in_last;
  uint8_t *bytes; /* item stream in bucket-id order, malloc'ed; bucket_byte_off[65536] bytes */
} mhbo_sdbg_out;

/* kmer_counter.cpp:60-414 */
int mhbo_count(const mhbo_seqs *reads, uint32_t k, int32_t m, mhbo_count_out *out);
void mhbo_count_free(mhbo_count_out *out);

/* seq_to_sdbg.cpp:530-806 (+ sdbg_writer.cpp:25-58 for the byte format) */
int mhbo_seq2sdbg(const mhbo_seqs *seqs, const uint16_t *mult, uint32_t k, mhbo_sdbg_out *out);
void mhbo_sdbg_free(mhbo_sdbg_out *out);

/* seq_to_sdbg.cpp:100-357: mercy edges.  `edges` = n_edges sorted (k+1)-mers, wpe words each (the
 * multiplicity bits are ignored); `cand` = candidate reads exactly as stored in P.cand (package
 * orientation).  Returns malloc'ed (k+1)-mers, one ceil((k+1)/16)-word record per mercy edge, in read
 * order (the reference's order is thread-schedule dependent and irrelevant downstream). */
int mhbo_gen_mercy(const uint32_t *edges, uint64_t n_edges, uint32_t wpe, const mhbo_seqs *cand,
                   uint32_t k, uint32_t **mercy_out, uint64_t *n_mercy_out);

/* binary_reader.h:23-53 + sequence_package.h:275-306: parse a .lib.bin / .cand image into word-aligned
 * package-orientation sequences (reversed, NOT complemented, when reverse != 0).
 * Two-call protocol: first with words == NULL to size the outputs. */
int mhbo_unpack_bin(const uint8_t *bin, uint64_t bin_bytes, int reverse, uint64_t *n_seqs,
                    uint64_t *n_words, uint32_t *words, uint64_t *word_off, uint32_t *len);

/* main_read2sdbg (main_sdbg_build.cpp:88-156): read_to_sdbg_s1.cpp (only when m > 1) + the mercy step of
 * Read2SdbgS2::Initialize (read_to_sdbg_s2.cpp:117-263, when need_mercy) + read_to_sdbg_s2.cpp, with kmsort's tie
 * order reproduced (mhb_oracle_r2s.c).  counting[65536] = what stage 1 dumps to P.counting (all zero for m == 1).
 * *is_solid_out (optional, malloc'ed, one bit per base of the package, bit i of byte i/8) = the solid-edge marker after
 * the mercy step; *n_bases_out its size in bits. */
int mhbo_read2sdbg(const mhbo_seqs *reads, uint32_t k, int32_t m, int need_mercy, mhbo_sdbg_out *out, int64_t *counting,
                   uint64_t *n_mercy_out, uint8_t **is_solid_out, uint64_t *n_bases_out);

/* `megahit_core iterate` (main_iterate.cpp, iterate/contig_flank_index.h, iterate/kmer_collector.h): the iterative
 * edges for k + step from contigs (flag-filtered, file orientation) and reads (file orientation); ascending unique
 * `.edges` records with multiplicity 0 (mhb_oracle_iter.c).  *n_aligned_out = reads that produced at least one edge. */
int mhbo_iterate(const mhbo_seqs *contigs, const mhbo_seqs *reads, uint32_t k, uint32_t step, uint32_t **edges_out,
                 uint64_t *n_edges_out, uint64_t *n_aligned_out);

/* pieces of the read2sdbg restatement, exported so that the CPU tests can check the device code's __host__ __device__
 * building blocks one by one */
unsigned mhbo_s1_read_records(const uint32_t *w, unsigned L, unsigned k, uint64_t base_off, uint32_t *out);
void mhbo_kmsort(uint32_t *recs, int64_t n, unsigned nw, unsigned rw);
void mhbo_s2_record(const uint32_t *w, unsigned k, unsigned i, unsigned strand, unsigned type, uint32_t *rec, int *palindrome);

void mhbo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
