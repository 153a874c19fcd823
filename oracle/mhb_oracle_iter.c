/*
 * mhb_oracle_iter.c -- TEST INFRASTRUCTURE ONLY (see mhb_oracle.h).
 *
 * Plain-C restatement of `megahit_core iterate` (main_iterate.cpp:117-221, iterate/contig_flank_index.h:16-221,
 * iterate/kmer_collector.h:37-79): the (k+step+1)-mers of the reads that are spanned by contig flanks, i.e. the
 * "iterative edges" the next, larger k starts from.  One base at a time; the flank index is a sorted array, the
 * collector a sort + unique.  The reference writes the collected set in hash-table order (P.edges.0, `is_sorted 0`);
 * the canonical form compared here is the ascending set of records.  Every multiplicity is 0: FeedBatchContigs never
 * stores the contig multiplicity (FlankInfo{ext_seq, ext_len} leaves `mul` zero-initialised, :66), so
 * min(kMaxMul, int(mul + 0.5)) = 0 for every edge (:208-211) - checked on the reference's own outputs.
 * Citations relative to /root/reference/src.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>

#include "mhb_oracle.h"

static inline unsigned ibase(const uint32_t *w, uint64_t i) { return (w[i >> 4] >> (30 - 2 * (i & 15))) & 3u; }

typedef struct {
  uint8_t *key; /* k+1 bases */
  uint64_t ext_seq;
  unsigned ext_len;
} flank;

static unsigned g_klen;
static int flank_cmp(const void *a, const void *b) {
  const flank *x = (const flank *)a, *y = (const flank *)b;
  int c = memcmp(x->key, y->key, g_klen);
  if (c) return c;
  /* larger (ext_len, ext_seq) first: that entry survives (contig_flank_index.h:67-75) */
  if (x->ext_len != y->ext_len) return x->ext_len > y->ext_len ? -1 : 1;
  if (x->ext_seq != y->ext_seq) return x->ext_seq > y->ext_seq ? -1 : 1;
  return 0;
}

static const flank *find_flank(const flank *t, int64_t n, const uint8_t *key, unsigned klen) {
  int64_t lo = 0, hi = n - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) / 2;
    int c = memcmp(t[mid].key, key, klen);
    if (c == 0) return &t[mid];
    if (c < 0)
      lo = mid + 1;
    else
      hi = mid - 1;
  }
  return NULL;
}

static unsigned g_words;
static int words_cmp(const void *a, const void *b) {
  const uint32_t *x = (const uint32_t *)a, *y = (const uint32_t *)b;
  for (unsigned i = 0; i < g_words; ++i)
    if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  return 0;
}

/* contigs / reads: word-aligned packed sequences in FILE orientation (iterate loads both with reverse = false,
 * async_sequence_reader.h:51,97); contigs already filtered by flag (kLoop | kStandalone discarded, :87).
 * Output: malloc'ed ascending unique `.edges` records (ceil((2(k+step+1)+16)/32) words, multiplicity 0). */
int mhbo_iterate(const mhbo_seqs *contigs, const mhbo_seqs *reads, uint32_t k, uint32_t step, uint32_t **edges_out,
                 uint64_t *n_edges_out, uint64_t *n_aligned_out) {
  const unsigned K1 = k + 1, KN = k + step + 1, W = (2 * KN + 16 + 31) / 32;
  /* ---- FeedBatchContigs, contig_flank_index.h:29-86 ---- */
  flank *tab = (flank *)malloc((size_t)(2 * contigs->n + 1) * sizeof(flank));
  uint8_t *keys = (uint8_t *)malloc((size_t)(2 * contigs->n + 1) * K1);
  int64_t nt = 0;
  for (uint64_t c = 0; c < contigs->n; ++c) {
    const uint32_t *w = contigs->words + contigs->word_off[c];
    const unsigned L = contigs->len[c];
    if (L < K1) continue;
    for (int strand = 0; strand < 2; ++strand) {
      uint8_t *key = keys + (size_t)nt * K1;
#define CH(j) (strand == 0 ? ibase(w, (j)) : 3u ^ ibase(w, L - 1 - (j)))
      for (unsigned j = 0; j < K1; ++j) key[j] = (uint8_t)CH(j);
      int pal = 1; /* Kmer::IsPalindrome: equal to its reverse complement */
      for (unsigned j = 0; j < K1 && pal; ++j) pal = key[j] == (3u ^ key[K1 - 1 - j]);
      if (!pal) {
        unsigned ext_len = step - 1 < L - K1 ? step - 1 : L - K1;
        uint64_t ext_seq = 0;
        for (unsigned j = 0; j < ext_len; ++j) ext_seq |= (uint64_t)CH(K1 + j) << (2 * j);
        tab[nt].key = key;
        tab[nt].ext_seq = ext_seq;
        tab[nt].ext_len = ext_len;
        ++nt;
      }
#undef CH
      if (L == K1) break; /* :82-84 */
    }
  }
  g_klen = K1;
  qsort(tab, (size_t)nt, sizeof(flank), flank_cmp);
  int64_t nu = 0;
  for (int64_t i = 0; i < nt; ++i)
    if (nu == 0 || memcmp(tab[nu - 1].key, tab[i].key, K1) != 0) tab[nu++] = tab[i];

  /* ---- FindNextKmersFromReads, :88-215 ---- */
  uint64_t cap = 1024, ne = 0, aligned = 0;
  uint32_t *out = (uint32_t *)malloc(cap * W * 4);
  uint32_t maxlen = 1;
  for (uint64_t r = 0; r < reads->n; ++r)
    if (reads->len[r] > maxlen) maxlen = reads->len[r];
  uint8_t *exist = (uint8_t *)malloc(maxlen + 1), *fk = (uint8_t *)malloc(K1), *rk = (uint8_t *)malloc(K1);
  uint8_t *nf = (uint8_t *)malloc(KN), *nr = (uint8_t *)malloc(KN);
  for (uint64_t r = 0; r < reads->n; ++r) {
    const uint32_t *w = reads->words + reads->word_off[r];
    const unsigned L = reads->len[r];
    if (L < KN) continue;
    memset(exist, 0, L);
    unsigned cur = 0;
    while (cur + K1 <= L) {
      unsigned next = cur + 1;
      if (!exist[cur]) {
        for (unsigned j = 0; j < K1; ++j) {
          fk[j] = (uint8_t)ibase(w, cur + j);
          rk[j] = (uint8_t)(3u ^ ibase(w, cur + K1 - 1 - j));
        }
        const flank *f = find_flank(tab, nu, fk, K1);
        if (f) {
          exist[cur] = 1;
          for (unsigned j = 0; j < f->ext_len && cur + K1 + j < L; ++j, ++next) {
            if (ibase(w, cur + K1 + j) == ((f->ext_seq >> (2 * j)) & 3u))
              exist[cur + j + 1] = 1;
            else
              break;
          }
        }
        f = find_flank(tab, nu, rk, K1);
        if (f) {
          exist[cur] = 1;
          for (unsigned j = 0; j < f->ext_len && cur >= j + 1; ++j) {
            if ((3u ^ ibase(w, cur - 1 - j)) == ((f->ext_seq >> (2 * j)) & 3u))
              exist[cur - 1 - j] = 1;
            else
              break;
          }
        }
      }
      if (next + K1 <= L)
        cur = next;
      else
        break;
    }
    int success = 0;
    unsigned acc = 0;
    for (unsigned j = 0; j + k < L; ++j) { /* :177-212 */
      acc = exist[j] ? acc + 1 : 0;
      if (acc >= step + 1) {
        const unsigned s0 = j + K1 - KN;
        for (unsigned i = 0; i < KN; ++i) {
          nf[i] = (uint8_t)ibase(w, s0 + i);
          nr[i] = (uint8_t)(3u ^ ibase(w, s0 + KN - 1 - i));
        }
        const uint8_t *can = memcmp(nf, nr, KN) < 0 ? nf : nr; /* new_kmer < new_rkmer ? new_kmer : new_rkmer */
        if (ne == cap) {
          cap *= 2;
          out = (uint32_t *)realloc(out, cap * W * 4);
        }
        uint32_t *rec = out + ne * W;
        memset(rec, 0, 4 * W);
        /* KmerCollector::WriteToFile, kmer_collector.h:50-69: base KN-1-j of the k-mer goes to position j */
        for (unsigned i = 0; i < KN; ++i) rec[i >> 4] |= (uint32_t)can[KN - 1 - i] << (30 - 2 * (i & 15));
        ++ne;
        success = 1;
      }
    }
    aligned += (uint64_t)success;
  }
  g_words = W;
  qsort(out, (size_t)ne, 4 * W, words_cmp);
  uint64_t nuq = 0;
  for (uint64_t i = 0; i < ne; ++i)
    if (nuq == 0 || memcmp(out + (nuq - 1) * W, out + i * W, 4 * W) != 0) memmove(out + nuq++ * W, out + i * W, 4 * W);
  free(tab);
  free(keys);
  free(exist);
  free(fk);
  free(rk);
  free(nf);
  free(nr);
  *edges_out = out;
  *n_edges_out = nuq;
  if (n_aligned_out) *n_aligned_out = aligned;
  return 0;
}
