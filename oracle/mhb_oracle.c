/*
 * mhb_oracle.c -- TEST INFRASTRUCTURE ONLY (see mhb_oracle.h).
 *
 * Plain-C restatement of megahit v1.2.9's `count` and `seq2sdbg` engines.  Written for obviousness, not
 * speed: sequences are handled one base at a time and records are sorted with qsort.  Reference
 * citations are relative to /root/reference/src.
 */
#define _GNU_SOURCE
#include "mhb_oracle.h"

#include <stdlib.h>
#include <string.h>

#define SENTINEL 4u
#define SENT_OFF 0xFFFFFFFFu

static inline unsigned div_ceil(unsigned a, unsigned b) { return (a + b - 1) / b; }

/* base i of a word-aligned packed sequence (kmcompactvector.h:53-57, big-endian-in-word) */
static inline unsigned seq_base(const uint32_t *w, uint64_t i) {
  return (w[i >> 4] >> (30 - 2 * (i & 15))) & 3u;
}
static inline void put_base(uint32_t *w, uint64_t i, unsigned c) {
  w[i >> 4] |= (uint32_t)(c & 3u) << (30 - 2 * (i & 15));
}

void mhbo_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * .bin / .cand parsing: binary_reader.h:23-53 (u32 length + ceil(len/16) words per read) and
 * sequence_package.h:275-306 (reverse = base order flipped, no complement; zero-length reads become a
 * fake 1-base 'A': :276-281).
 * ---------------------------------------------------------------------------------------------- */
int mhbo_unpack_bin(const uint8_t *bin, uint64_t bin_bytes, int reverse, uint64_t *n_seqs,
                    uint64_t *n_words, uint32_t *words, uint64_t *word_off, uint32_t *len) {
  uint64_t pos = 0, ns = 0, nw = 0;
  while (pos + 4 <= bin_bytes) {
    uint32_t l;
    memcpy(&l, bin + pos, 4);
    pos += 4;
    uint64_t src_words = div_ceil(l, 16);
    if (pos + 4 * src_words > bin_bytes) return -1;
    uint32_t eff = l == 0 ? 1 : l;
    uint64_t dst_words = div_ceil(eff, 16);
    if (words) {
      const uint32_t *src = (const uint32_t *)(bin + pos);
      uint32_t *dst = words + nw;
      memset(dst, 0, 4 * dst_words);
      for (uint32_t i = 0; i < l; ++i) {
        uint32_t w;
        uint64_t si = reverse ? (uint64_t)(l - 1 - i) : i;
        memcpy(&w, &src[si >> 4], 4);
        put_base(dst, i, (w >> (30 - 2 * (si & 15))) & 3u);
      }
      word_off[ns] = nw;
      len[ns] = eff;
    }
    pos += 4 * src_words;
    nw += dst_words;
    ++ns;
  }
  if (words) word_off[ns] = nw;
  *n_seqs = ns;
  *n_words = nw;
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * record comparison: ascending unsigned lexicographic on the first nw words (kmsort_selector.cpp:18-27)
 * ---------------------------------------------------------------------------------------------- */
static int cmp_words(const void *a, const void *b, void *arg) {
  unsigned nw = *(const unsigned *)arg;
  const uint32_t *x = (const uint32_t *)a, *y = (const uint32_t *)b;
  for (unsigned i = 0; i < nw; ++i) {
    if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  }
  return 0;
}

/* ================================================================================================
 * count  (kmer_counter.cpp)
 * ============================================================================================== */
int mhbo_count(const mhbo_seqs *reads, uint32_t k, int32_t m, mhbo_count_out *out) {
  const unsigned K1 = k + 1;
  const unsigned W = div_ceil(K1 * 2, 32);        /* words_per_substr_, kmer_counter.cpp:77-78 */
  const unsigned WPE = div_ceil(K1 * 2 + 16, 32); /* words_per_edge_,   kmer_counter.cpp:79-80 */
  const unsigned STRIDE = W + 3;                  /* key | read id | offset | strand<<6|prev<<3|next */
  memset(out, 0, sizeof(*out));
  out->words_per_edge = WPE;

  uint64_t n_rec = 0;
  for (uint64_t r = 0; r < reads->n; ++r) {
    if (reads->len[r] >= K1) n_rec += reads->len[r] - k; /* kmer_counter.cpp:124 skips short reads */
  }
  out->n_records = n_rec;
  out->first_0_out = (uint32_t *)malloc(sizeof(uint32_t) * (reads->n ? reads->n : 1));
  out->last_0_in = (uint32_t *)malloc(sizeof(uint32_t) * (reads->n ? reads->n : 1));
  for (uint64_t r = 0; r < reads->n; ++r) { /* kmer_counter.cpp:86-89 */
    out->first_0_out[r] = SENT_OFF;
    out->last_0_in[r] = SENT_OFF;
  }
  uint32_t *rec = (uint32_t *)calloc((size_t)(n_rec ? n_rec : 1) * STRIDE, sizeof(uint32_t));
  if (!rec) return -2;

  /* Lv1FillOffsets + Lv2ExtractSubString, kmer_counter.cpp:158-252 */
  uint64_t ri = 0;
  for (uint64_t r = 0; r < reads->n; ++r) {
    const uint32_t *w = reads->words + reads->word_off[r];
    const unsigned L = reads->len[r];
    if (L < K1) continue;
    for (unsigned pos = 0; pos + K1 <= L; ++pos) {
      /* canonical choice: rev_edge.cmp(edge) < 0 ? rc : fwd  (kmer_counter.cpp:182) */
      int rc_less = 0;
      for (unsigned i = 0; i < K1; ++i) {
        unsigned f = seq_base(w, pos + i);
        unsigned c = 3u - seq_base(w, pos + K1 - 1 - i);
        if (c != f) {
          rc_less = c < f;
          break;
        }
      }
      uint32_t *rp = rec + ri * STRIDE;
      for (unsigned i = 0; i < K1; ++i) {
        unsigned b = rc_less ? 3u - seq_base(w, pos + K1 - 1 - i) : seq_base(w, pos + i);
        put_base(rp, i, b);
      }
      unsigned prev = pos > 0 ? seq_base(w, pos - 1) : SENTINEL;        /* :223-227 */
      unsigned next = pos + K1 < L ? seq_base(w, pos + K1) : SENTINEL;  /* :229-233 */
      unsigned p = prev, n = next;
      if (rc_less) { /* :243-248: swap + complement on the rc strand */
        p = next == SENTINEL ? SENTINEL : 3u - next;
        n = prev == SENTINEL ? SENTINEL : 3u - prev;
      }
      rp[W] = (uint32_t)r;
      rp[W + 1] = pos;
      rp[W + 2] = ((unsigned)rc_less << 6) | (p << 3) | n;
      ++ri;
    }
  }

  unsigned nw = W;
  qsort_r(rec, (size_t)n_rec, sizeof(uint32_t) * STRIDE, cmp_words, &nw);

  /* Lv2Postprocess, kmer_counter.cpp:254-381 -- pass 1 counts solid edges, pass 2 writes them */
  for (int pass = 0; pass < 2; ++pass) {
    uint64_t n_solid = 0, n_distinct = 0;
    if (pass == 1) out->edges = (uint32_t *)calloc((size_t)(out->n_solid ? out->n_solid : 1) * WPE, 4);
    for (uint64_t i = 0, to; i < n_rec; i = to) {
      to = i + 1;
      while (to < n_rec && cmp_words(rec + i * STRIDE, rec + to * STRIDE, &nw) == 0) ++to;
      int64_t count = (int64_t)(to - i);
      ++n_distinct;
      if (pass == 0) {
        int64_t cp[5] = {0, 0, 0, 0, 0}, cn[5] = {0, 0, 0, 0, 0};
        for (uint64_t j = i; j < to; ++j) {
          unsigned pn = rec[j * STRIDE + W + 2] & 63u;
          cp[pn >> 3]++;
          cn[pn & 7]++;
        }
        int has_in = 0, has_out = 0;
        for (int j = 0; j < 4; ++j) { /* :297-305 */
          if (cp[j] >= m) has_in = 1;
          if (cn[j] >= m) has_out = 1;
        }
        if (count >= m && (!has_in || !has_out)) {
          for (uint64_t j = i; j < to; ++j) {
            uint32_t read = rec[j * STRIDE + W], off = rec[j * STRIDE + W + 1];
            unsigned strand = (rec[j * STRIDE + W + 2] >> 6) & 1u;
            /* :307-335 (no in): strand 0 -> last = max, strand 1 -> first = min(offset+1)
             * :338-367 (no out): strand 0 -> first = min(offset+1), strand 1 -> last = max */
            for (int which = 0; which < 2; ++which) {
              if (which == 0 && has_in) continue;
              if (which == 1 && has_out) continue;
              int upd_last = (which == 0) ? (strand == 0) : (strand == 1);
              if (upd_last) {
                uint32_t old = out->last_0_in[read];
                if (old == SENT_OFF || old < off) out->last_0_in[read] = off;
              } else {
                if (out->first_0_out[read] > off + 1) out->first_0_out[read] = off + 1;
              }
            }
          }
        }
        out->counting[count > MHBO_MAX_MUL ? MHBO_MAX_MUL : count]++; /* edge_counter.h:30-33 */
      }
      if (count >= m) {
        if (pass == 1) { /* PackEdge, kmer_counter.cpp:32-52 (key tail bits are already zero) */
          uint32_t *e = out->edges + n_solid * WPE;
          for (unsigned x = 0; x < W && x < WPE; ++x) e[x] = rec[i * STRIDE + x];
          e[WPE - 1] |= (uint32_t)(count > MHBO_MAX_MUL ? MHBO_MAX_MUL : count);
        }
        ++n_solid;
      }
    }
    out->n_solid = n_solid;
    out->n_distinct = n_distinct;
  }
  free(rec);
  return 0;
}

void mhbo_count_free(mhbo_count_out *out) {
  free(out->edges);
  free(out->first_0_out);
  free(out->last_0_in);
  out->edges = out->first_0_out = out->last_0_in = NULL;
}

/* ================================================================================================
 * seq2sdbg  (seq_to_sdbg.cpp)
 * ============================================================================================== */
static inline int s2s_a(const uint32_t *item, unsigned W, unsigned k) { /* Extract_a, :71-87 */
  if ((item[W - 1] >> 19) & 1u) return (int)seq_base(item, k - 1);
  return (int)SENTINEL;
}
static inline int s2s_b(const uint32_t *item, unsigned W) { return (item[W - 1] >> 16) & 7u; } /* :89-92 */

static int diff_km1(const uint32_t *x, const uint32_t *y, unsigned k) { /* IsDiffKMinusOneMer :46-69 */
  for (unsigned i = 0; i + 1 < k; ++i) {
    if (seq_base(x, i) != seq_base(y, i)) return 1;
  }
  return 0;
}

int mhbo_seq2sdbg(const mhbo_seqs *seqs, const uint16_t *mult, uint32_t k, mhbo_sdbg_out *out) {
  const unsigned W = div_ceil(k * 2 + 3 + 1 + 16, 32); /* words_per_substr_, seq_to_sdbg.cpp:510-512 */
  const unsigned WPT = div_ceil(k, 16);                /* sdbg_writer.h:41 */
  memset(out, 0, sizeof(*out));
  out->words_per_tip_label = WPT;

  uint64_t n_rec = 0;
  for (uint64_t s = 0; s < seqs->n; ++s) {
    if (seqs->len[s] >= k + 1) n_rec += 2ull * (seqs->len[s] - k + 2); /* :530-577 */
  }
  out->n_records = n_rec;
  uint32_t *rec = (uint32_t *)calloc((size_t)(n_rec ? n_rec : 1) * W, sizeof(uint32_t));
  if (!rec) return -2;

  /* Lv2ExtractSubString, seq_to_sdbg.cpp:630-700 */
  uint64_t ri = 0;
  for (uint64_t s = 0; s < seqs->n; ++s) {
    const uint32_t *w = seqs->words + seqs->word_off[s];
    const int L = (int)seqs->len[s];
    if (L < (int)k + 1) continue;
    for (int strand = 0; strand < 2; ++strand) {
      for (int offset = 0; offset <= L - (int)k + 1; ++offset) {
        unsigned nc = k - (unsigned)(offset + (int)k > L);
        int counting = 0;
        if (offset > 0 && offset + (int)k <= L) counting = mult[s]; /* :641-643 */
        uint32_t *rp = rec + ri * W;
        unsigned prev;
        if (strand == 0) {
          prev = offset == 0 ? SENTINEL : seq_base(w, (uint64_t)offset - 1);
          for (unsigned i = 0; i < nc; ++i) put_base(rp, i, seq_base(w, (uint64_t)offset + i));
        } else {
          prev = offset == 0 ? SENTINEL : 3u - seq_base(w, (uint64_t)(L - 1 - offset + 1)); /* :678 */
          int off2 = L - 1 - offset - ((int)k - 1);                                          /* :681 */
          if (off2 < 0) off2 = 0;                                                            /* :683-686 */
          for (unsigned i = 0; i < nc; ++i)
            put_base(rp, i, 3u - seq_base(w, (uint64_t)off2 + nc - 1 - i));
        }
        int stored = MHBO_MAX_MUL - counting;
        if (stored < 0) stored = 0;
        rp[W - 1] |= (uint32_t)(nc == k) << 19; /* :664-670 */
        rp[W - 1] |= prev << 16;
        rp[W - 1] |= (uint32_t)stored;
        ++ri;
      }
    }
  }

  unsigned nw = W;
  qsort_r(rec, (size_t)n_rec, sizeof(uint32_t) * W, cmp_words, &nw);

  /* Lv2Postprocess :702-789 + SdbgWriter::Write sdbg_writer.cpp:25-58.  Pass 0 sizes, pass 1 writes. */
  for (int pass = 0; pass < 2; ++pass) {
    uint64_t byte_pos = 0;
    int cur_bucket = -1;
    if (pass == 1) out->bytes = (uint8_t *)malloc(out->bucket_byte_off[MHBO_NUM_BUCKETS] + 1);
    for (uint64_t start = 0, end; start < n_rec; start = end) {
      end = start + 1;
      const uint32_t *item = rec + start * W;
      while (end < n_rec && !diff_km1(item, rec + end * W, k)) ++end;

      int has_solid_a = 0, has_solid_b = 0, outputed_b = 0;
      int64_t last_a[4] = {-1, -1, -1, -1};
      for (uint64_t i = start; i < end; ++i) { /* :724-738 */
        int a = s2s_a(rec + i * W, W, k), b = s2s_b(rec + i * W, W);
        if (a != (int)SENTINEL && b != (int)SENTINEL) {
          has_solid_a |= 1 << a;
          has_solid_b |= 1 << b;
        }
        if (a != (int)SENTINEL && (b != (int)SENTINEL || !(has_solid_a & (1 << a)))) last_a[a] = (int64_t)i;
      }
      for (uint64_t i = start, j; i < end; i = j) { /* :740-786 */
        const uint32_t *cur = rec + i * W;
        int a = s2s_a(cur, W, k), b = s2s_b(cur, W);
        j = i + 1;
        while (j < end && s2s_a(rec + j * W, W, k) == a && s2s_b(rec + j * W, W) == b) ++j;
        int is_dollar = 0;
        if (a == (int)SENTINEL) {
          if (has_solid_b & (1 << b)) continue;
          is_dollar = 1;
        }
        if (b == (int)SENTINEL) {
          if (has_solid_a & (1 << a)) continue;
        }
        int w = (b == (int)SENTINEL) ? 0 : ((outputed_b & (1 << b)) ? b + 5 : b + 1);
        int last = (a == (int)SENTINEL) ? 0 : (last_a[a] == (int64_t)j - 1);
        outputed_b |= 1 << b;
        unsigned mul = MHBO_MAX_MUL - (cur[W - 1] & 0xFFFFu);
        unsigned bucket = cur[0] >> 16;

        if (pass == 0) {
          out->bucket_items[bucket]++;
          out->w_count[w]++;
          out->ones_in_last += (unsigned)last;
          if (mul > 254) out->bucket_large_mul[bucket]++;
          if (is_dollar) out->bucket_tips[bucket]++;
        } else {
          if ((int)bucket != cur_bucket) {
            cur_bucket = (int)bucket;
            byte_pos = out->bucket_byte_off[bucket];
          }
          uint8_t *p = out->bytes + byte_pos;
          p[0] = (uint8_t)(w | (last << 4) | (is_dollar << 5)); /* sdbg_item.h:14-24 */
          p[1] = (uint8_t)(mul > 255 ? 255 : mul);              /* sdbg_writer.cpp:38 */
          byte_pos += 2;
          if (mul > 254) { /* :45-50 */
            uint16_t m16 = (uint16_t)mul;
            memcpy(out->bytes + byte_pos, &m16, 2);
            byte_pos += 2;
          }
          if (is_dollar) { /* :52-57: raw leading words of the sort record */
            memcpy(out->bytes + byte_pos, cur, 4 * WPT);
            byte_pos += 4 * WPT;
          }
        }
      }
    }
    if (pass == 0) {
      uint64_t acc = 0;
      for (unsigned b = 0; b < MHBO_NUM_BUCKETS; ++b) {
        out->bucket_byte_off[b] = acc;
        acc += 2 * out->bucket_items[b] + 2 * out->bucket_large_mul[b] + 4ull * WPT * out->bucket_tips[b];
        out->n_items += out->bucket_items[b];
      }
      out->bucket_byte_off[MHBO_NUM_BUCKETS] = acc;
    }
  }
  free(rec);
  return 0;
}

void mhbo_sdbg_free(mhbo_sdbg_out *out) {
  free(out->bytes);
  out->bytes = NULL;
}

/* ================================================================================================
 * mercy edges  (seq_to_sdbg.cpp:100-357)
 * ============================================================================================== */
#define LUT_PREFIX 12u
#define LUT_SIZE (1u << (2 * LUT_PREFIX))

typedef struct {
  const uint32_t *edges;
  uint64_t n;
  unsigned wpe;
  int64_t *lut; /* 2 * LUT_SIZE, [first,last] per 12-mer prefix, -1 = empty (InitLookupTable :100-127) */
} edge_index;

static inline unsigned edge_base(const edge_index *ix, int64_t e, unsigned i) {
  return seq_base(ix->edges + (uint64_t)e * ix->wpe, i);
}

/* BinarySearchKmer :132-161 -- km[0..ksz) compared with the first ksz bases of each edge.  The exact
 * probe sequence matters: with ksz == k several edges can match and the caller looks at base k of the
 * one that was hit. */
static int64_t search_kmer(const edge_index *ix, const uint8_t *km, unsigned ksz) {
  uint32_t prefix = 0;
  for (unsigned i = 0; i < LUT_PREFIX; ++i) prefix = (prefix << 2) | (i < ksz ? km[i] : 0u);
  int64_t l = ix->lut[2 * (uint64_t)prefix];
  if (l == -1) return -1;
  int64_t r = ix->lut[2 * (uint64_t)prefix + 1];
  while (l <= r) {
    int64_t mid = (l + r) / 2;
    int cmp = 0;
    for (unsigned i = 0; i < ksz; ++i) {
      unsigned eb = edge_base(ix, mid, i);
      if (km[i] != eb) {
        cmp = km[i] < eb ? -1 : 1;
        break;
      }
    }
    if (cmp > 0) l = mid + 1;
    else if (cmp < 0) r = mid - 1;
    else return mid;
  }
  return -1;
}

static int cmp_bases(const uint8_t *x, const uint8_t *y, unsigned n) {
  for (unsigned i = 0; i < n; ++i) {
    if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  }
  return 0;
}

int mhbo_gen_mercy(const uint32_t *edges, uint64_t n_edges, uint32_t wpe, const mhbo_seqs *cand,
                   uint32_t k, uint32_t **mercy_out, uint64_t *n_mercy_out) {
  const unsigned WM = div_ceil(k + 1, 16);
  edge_index ix = {edges, n_edges, wpe, NULL};
  ix.lut = (int64_t *)malloc(sizeof(int64_t) * 2 * LUT_SIZE);
  memset(ix.lut, 0xFF, sizeof(int64_t) * 2 * LUT_SIZE);
  if (n_edges > 0) { /* :100-127 */
    uint32_t cur = edges[0] >> 8;
    ix.lut[2 * (uint64_t)cur] = 0;
    for (uint64_t i = 1; i < n_edges; ++i) {
      uint32_t p = edges[i * wpe] >> 8;
      if (p > cur) {
        ix.lut[2 * (uint64_t)cur + 1] = (int64_t)i - 1;
        cur = p;
        ix.lut[2 * (uint64_t)cur] = (int64_t)i;
      }
    }
    ix.lut[2 * (uint64_t)cur + 1] = (int64_t)n_edges - 1;
  }

  uint64_t cap = 1024, nm = 0;
  uint32_t *mercy = (uint32_t *)malloc(sizeof(uint32_t) * cap * WM);
  uint8_t *km = (uint8_t *)malloc(k + 2), *rv = (uint8_t *)malloc(k + 2);
  uint8_t *has_in = NULL, *has_out = NULL;
  unsigned flag_cap = 0;

  for (uint64_t r = 0; r < cand->n; ++r) {
    const uint32_t *w = cand->words + cand->word_off[r];
    const unsigned L = cand->len[r];
    if (L < k + 2) continue; /* :206-208 */
    if (L + 2 > flag_cap) {
      flag_cap = L + 2;
      has_in = (uint8_t *)realloc(has_in, flag_cap);
      has_out = (uint8_t *)realloc(has_out, flag_cap);
    }
    memset(has_in, 0, L + 2);
    memset(has_out, 0, L + 2);
    /* km = k-mer at position i (slot k kept 0), rv = its reverse complement (slot k kept 0) */
    for (unsigned i = 0; i < k; ++i) {
      km[i] = (uint8_t)seq_base(w, i);
      rv[i] = (uint8_t)(3u - seq_base(w, k - 1 - i));
    }
    km[k] = rv[k] = 0;

    for (unsigned i = 0; i + k <= L; ++i) { /* :224-307 */
      if (!has_in[i]) {
        if (search_kmer(&ix, rv, k) != -1) {
          has_in[i] = 1;
        } else {
          rv[k] = 3;                 /* rev_kmer.SetBase(k, 3) */
          memmove(km + 1, km, k);    /* kmer.ShiftPreappend(0, k+1) */
          for (unsigned c = 0; c < 4; ++c) {
            km[0] = (uint8_t)c;
            if (cmp_bases(km, rv, k + 1) > 0) break;
            if (search_kmer(&ix, km, k + 1) != -1) {
              has_in[i] = 1;
              break;
            }
          }
          rv[k] = 0;
          memmove(km, km + 1, k);    /* kmer.ShiftAppend(0, k+1) */
          km[k] = 0;
        }
      }
      int64_t edge_id = search_kmer(&ix, km, k);
      if (edge_id != -1) {
        has_out[i] = 1;
        if (i + k < L && edge_base(&ix, edge_id, k) == seq_base(w, i + k)) has_in[i + 1] = 1;
      } else {
        km[k] = 3;
        unsigned next_char = i + k < L ? 3u - seq_base(w, i + k) : 0u;
        memmove(rv + 1, rv, k); /* rev_kmer.ShiftPreappend(next_char, k+1) */
        rv[0] = (uint8_t)next_char;
        if (cmp_bases(rv, km, k + 1) <= 0 && search_kmer(&ix, rv, k + 1) != -1) {
          has_out[i] = 1;
          has_in[i + 1] = 1;
        } else {
          for (unsigned c = 0; c < 4; ++c) {
            if (c == next_char) continue;
            rv[0] = (uint8_t)c;
            if (cmp_bases(rv, km, k + 1) > 0) break;
            if (search_kmer(&ix, rv, k + 1) != -1) {
              has_out[i] = 1;
              break;
            }
          }
        }
        km[k] = 0;
        memmove(rv, rv + 1, k); /* rev_kmer.ShiftAppend(0, k+1) */
        rv[k] = 0;
      }
      if (i + k < L) { /* :301-306 */
        unsigned nc = seq_base(w, i + k);
        memmove(km, km + 1, k - 1);
        km[k - 1] = (uint8_t)nc;
        memmove(rv + 1, rv, k - 1);
        rv[0] = (uint8_t)(3u - nc);
      }
    }

    int last_no_out = -1; /* :310-352 */
    for (unsigned i = 0; i + k <= L; ++i) {
      switch (has_in[i] | (has_out[i] << 1)) {
        case 1:
          last_no_out = (int)i;
          break;
        case 2:
          if (last_no_out >= 0) {
            for (unsigned j = (unsigned)last_no_out; j < i; ++j) {
              if (nm == cap) {
                cap *= 2;
                mercy = (uint32_t *)realloc(mercy, sizeof(uint32_t) * cap * WM);
              }
              uint32_t *e = mercy + nm * WM;
              memset(e, 0, 4 * WM);
              for (unsigned x = 0; x < k + 1; ++x) put_base(e, x, seq_base(w, j + x));
              ++nm;
            }
          }
          last_no_out = -1;
          break;
        case 3:
          last_no_out = -1;
          break;
        default:
          break;
      }
    }
  }
  free(km);
  free(rv);
  free(has_in);
  free(has_out);
  free(ix.lut);
  *mercy_out = mercy;
  *n_mercy_out = nm;
  return 0;
}
