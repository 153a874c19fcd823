/*
 * mhb_oracle_r2s.c -- TEST INFRASTRUCTURE ONLY (see mhb_oracle.h).
 *
 * Plain-C restatement of megahit v1.2.9's 1-pass SdBG build, `megahit_core read2sdbg`
 * (sorting/read_to_sdbg_s1.cpp, sorting/read_to_sdbg_s2.cpp, driven by main_sdbg_build.cpp:88-156), INCLUDING
 * the permutation kmlib::kmsort (kmlib/kmsort.h:22-117) leaves among records with equal keys: stage 1 reads the
 * prev/next characters of the FIRST record of every (k-1)-mer group for all members of the group
 * (read_to_sdbg_s1.cpp:393-401), so its output depends on which of several tied records the unstable sort puts
 * first.  The reference is deterministic all the same (bucket contents arrive in global read order,
 * base_engine.cpp:323-348, and kmsort is a deterministic function of its input); this file reproduces that order.
 * One base at a time, qsort where order among ties cannot matter.  Citations relative to /root/reference/src.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>

#include "mhb_oracle.h"

#define SENTINEL 4u

static inline unsigned div_ceil_u(unsigned a, unsigned b) { return (a + b - 1) / b; }
static inline unsigned sbase(const uint32_t *w, uint64_t i) { return (w[i >> 4] >> (30 - 2 * (i & 15))) & 3u; }
static inline void pbase(uint32_t *w, uint64_t i, unsigned c) { w[i >> 4] |= (uint32_t)(c & 3u) << (30 - 2 * (i & 15)); }
static inline unsigned comp_or_sentinel(unsigned c) { return c == SENTINEL ? SENTINEL : 3u - c; }

/* ------------------------------------------------------------------------------------------------
 * kmlib::kmsort on records of `rw` words whose first `nw` words are the key (kmsort_selector.cpp:16-40:
 * n_bytes = 4*nw - 2 radix bytes, byte kb = (data[nw-1-kb/4] >> 8*(kb%4)) & 0xFF, operator< on the nw key words)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  unsigned nw, rw;
} recfmt;

static inline unsigned kth_byte(const uint32_t *r, const recfmt *f, int kb) {
  return (r[f->nw - 1 - (unsigned)kb / 4] >> (((unsigned)kb % 4) * 8)) & 0xFFu;
}
static int key_less(const uint32_t *x, const uint32_t *y, const recfmt *f) {
  for (unsigned i = 0; i < f->nw; ++i) {
    if (x[i] != y[i]) return x[i] < y[i];
  }
  return 0;
}

/* insert_sort_core, kmsort.h:22-35 (stable) */
static void kms_insertion(uint32_t *a, int64_t n, const recfmt *f) {
  uint32_t tmp[64];
  const unsigned rw = f->rw;
  for (int64_t i = 1; i < n; ++i) {
    if (key_less(a + i * rw, a + (i - 1) * rw, f)) {
      memcpy(tmp, a + i * rw, 4 * rw);
      memcpy(a + i * rw, a + (i - 1) * rw, 4 * rw);
      int64_t j = i - 1;
      for (; j > 0 && key_less(tmp, a + (j - 1) * rw, f); --j) memcpy(a + j * rw, a + (j - 1) * rw, 4 * rw);
      memcpy(a + j * rw, tmp, 4 * rw);
    }
  }
}

/* radix_sort_core, kmsort.h:43-101: in-place American-flag permutation on byte `kb`, then the children */
static void kms_radix(uint32_t *a, int64_t n, const recfmt *f, int kb) {
  int64_t count[256], begin[257], last[256];
  uint32_t swapper[64], other[64];
  const unsigned rw = f->rw;
  memset(count, 0, sizeof(count));
  for (int64_t i = 0; i < n; ++i) ++count[kth_byte(a + i * rw, f, kb)];
  begin[0] = 0;
  for (int i = 0; i < 256; ++i) begin[i + 1] = begin[i] + count[i];
  for (int i = 0; i < 256; ++i) last[i] = begin[i];
  for (int i = 0; i < 256; ++i) {
    const int64_t end = begin[i] + count[i];
    if (end == n) { /* :66-69: the last populated bin is in place once all others are */
      last[i] = n;
      break;
    }
    while (last[i] != end) {
      memcpy(swapper, a + last[i] * rw, 4 * rw);
      unsigned tag = kth_byte(swapper, f, kb);
      if (tag != (unsigned)i) {
        do { /* :75-79 */
          memcpy(other, a + last[tag] * rw, 4 * rw);
          memcpy(a + last[tag] * rw, swapper, 4 * rw);
          memcpy(swapper, other, 4 * rw);
          ++last[tag];
        } while ((tag = kth_byte(swapper, f, kb)) != (unsigned)i);
        memcpy(a + last[i] * rw, swapper, 4 * rw);
      }
      ++last[i];
    }
  }
  if (kb > 0) { /* :84-100 */
    for (int i = 0; i < 256; ++i) {
      if (count[i] > 64)
        kms_radix(a + begin[i] * rw, count[i], f, kb - 1);
      else if (count[i] > 1)
        kms_insertion(a + begin[i] * rw, count[i], f);
    }
  }
}

/* radix_sort_entry, kmsort.h:103-117 */
static void kms_sort(uint32_t *a, int64_t n, const recfmt *f) {
  if (n <= 1) return;
  if (n <= 64)
    kms_insertion(a, n, f);
  else
    kms_radix(a, n, f, 4 * (int)f->nw - 2 - 1);
}

/* ------------------------------------------------------------------------------------------------
 * stage 1 (read_to_sdbg_s1.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t *v;
  uint64_t n, cap;
} u64vec;
static int u64vec_push(u64vec *x, uint64_t val) {
  if (x->n == x->cap) {
    uint64_t nc = x->cap ? 2 * x->cap : 1024;
    uint64_t *nv = (uint64_t *)realloc(x->v, nc * 8);
    if (!nv) return -1;
    x->v = nv;
    x->cap = nc;
  }
  x->v[x->n++] = val;
  return 0;
}

static inline int bit_get(const uint8_t *b, uint64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
static inline void bit_set(uint8_t *b, uint64_t i) { b[i >> 3] |= (uint8_t)(1u << (i & 7)); }

/* (k-1)-mer at p vs its reverse complement: <0, 0, >0 as GenericKmer::cmp (read_to_sdbg_s1.cpp:254) */
static int cmp_fwd_rc(const uint32_t *w, unsigned p, unsigned kk) {
  for (unsigned i = 0; i < kk; ++i) {
    unsigned f = sbase(w, p + i), r = 3u - sbase(w, p + kk - 1 - i);
    if (f != r) return f < r ? -1 : 1;
  }
  return 0;
}

/* Lv2ExtractSubString, read_to_sdbg_s1.cpp:295-366: the record of ((k-1)-mer position p, strand) */
static void s1_record(const uint32_t *w, unsigned L, unsigned k, unsigned p, unsigned strand, uint64_t full_offset,
                      unsigned nw, uint32_t *rec) {
  const unsigned kk = k - 1;
  unsigned head, prev, tail, next;
  if (p > 1) {
    head = sbase(w, p - 1);
    prev = sbase(w, p - 2);
  } else {
    prev = SENTINEL;
    head = p > 0 ? sbase(w, p - 1) : SENTINEL;
  }
  if (p + k < L) {
    tail = sbase(w, p + k - 1);
    next = sbase(w, p + k);
  } else {
    next = SENTINEL;
    tail = p + k - 1 < L ? sbase(w, p + k - 1) : SENTINEL;
  }
  memset(rec, 0, 4 * (nw + 2));
  uint64_t info;
  if (strand == 0) {
    for (unsigned i = 0; i < kk; ++i) pbase(rec, i, sbase(w, p + i));
    rec[nw - 1] |= (head << 3) | tail;
    info = (full_offset << 6) | (prev << 3) | next;
  } else {
    for (unsigned i = 0; i < kk; ++i) pbase(rec, i, 3u - sbase(w, p + kk - 1 - i));
    rec[nw - 1] |= (comp_or_sentinel(tail) << 3) | comp_or_sentinel(head);
    info = (full_offset << 6) | (comp_or_sentinel(next) << 3) | comp_or_sentinel(prev);
  }
  rec[nw] = (uint32_t)(info >> 32); /* DecomposeUint64: high word first (utils.h:58-62) */
  rec[nw + 1] = (uint32_t)info;
}

/* The stage-1 records of one read in the order the reference appends them to their buckets (Lv1FillOffsets
 * :207-293: first (k-1)-mer on both strands, inner ones on the canonical strand, last one on both strands).
 * out: (L - k + 4) records of ceil((2(k-1)+6)/32) + 2 words.  Returns their number (0 for L < k + 1). */
unsigned mhbo_s1_read_records(const uint32_t *w, unsigned L, unsigned k, uint64_t base_off, uint32_t *out) {
  const unsigned kk = k - 1, nw = div_ceil_u(2 * kk + 6, 32), rw = nw + 2;
  unsigned n = 0;
  if (L < k + 1) return 0;
  for (unsigned p = 0; p + kk <= L; ++p) {
    unsigned strands[2], ns = 0;
    if (p == 0 || p + kk == L) { /* :239-245, :286-292 */
      strands[ns++] = 0;
      strands[ns++] = 1;
    } else {
      const int c = cmp_fwd_rc(w, p, kk);
      if (c > 0) {
        strands[ns++] = 1;
      } else if (c < 0) {
        strands[ns++] = 0;
      } else { /* palindrome :263-279 */
        const unsigned prev = sbase(w, p - 1), next = sbase(w, p + kk);
        strands[ns++] = prev <= 3u - next ? 0 : 1;
      }
    }
    for (unsigned s = 0; s < ns; ++s) s1_record(w, L, k, p, strands[s], ((base_off + p) << 1) | strands[s], nw, out + (size_t)(n++) * rw);
  }
  return n;
}

/* kmlib::kmsort on n records of rw words keyed on their first nw words (exported for the CPU tests of the device code) */
void mhbo_kmsort(uint32_t *recs, int64_t n, unsigned nw, unsigned rw) {
  const recfmt f = {nw, rw};
  kms_sort(recs, n, &f);
}

/* stage-2 item in the reference's own layout (ceil((2k+4)/32) words) + whether the edge at i is a palindrome */
void mhbo_s2_record(const uint32_t *w, unsigned k, unsigned i, unsigned strand, unsigned type, uint32_t *rec, int *palindrome);

/* IsDiffKMinusOneMer (read_to_sdbg_s1.cpp:40-63) */
static int diff_km1(const uint32_t *x, const uint32_t *y, unsigned k) {
  for (unsigned i = 0; i + 1 < k; ++i) {
    if (sbase(x, i) != sbase(y, i)) return 1;
  }
  return 0;
}

/* read id of a full base offset: last r with base_off[r] <= off */
static uint64_t read_of_offset(const uint64_t *base_off, uint64_t n, uint64_t off) {
  uint64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (base_off[mid] <= off)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

/* Lv2Postprocess, read_to_sdbg_s1.cpp:368-555, on one sorted bucket */
static int s1_postprocess(const uint32_t *recs, int64_t n, unsigned k, int m, unsigned nw, const uint64_t *base_off,
                          uint64_t n_reads, uint8_t *is_solid, u64vec *cand, int64_t *counting) {
  const unsigned rw = nw + 2;
  for (int64_t i = 0, end; i < n; i = end) {
    const uint32_t *first = recs + i * rw;
    int64_t cph[5][5], ctn[5][5], cht[64];
    memset(cph, 0, sizeof(cph));
    memset(ctn, 0, sizeof(ctn));
    memset(cht, 0, sizeof(cht));
    /* :393-401: prev/next are taken from the FIRST item for every member of the group */
    const unsigned pn_first = first[nw + 1] & 63u;
    end = i;
    while (end < n && (end == i || !diff_km1(first, recs + end * rw, k))) {
      const unsigned ht = recs[end * rw + nw - 1] & 63u;
      cph[pn_first >> 3][ht >> 3]++;
      ctn[ht & 7][pn_first & 7]++;
      cht[ht]++;
      ++end;
    }
    int has_in = 0, has_out = 0, l_has_out = 0, r_has_in = 0;
    for (int j = 0; j < 4; ++j) {
      for (int x = 0; x < 4; ++x) {
        if (cph[x][j] >= m) has_in |= 1 << j;
        if (ctn[j][x] >= m) has_out |= 1 << j;
        if (cht[(j << 3) | x] >= m) {
          l_has_out |= 1 << j;
          r_has_in |= 1 << x;
        }
      }
    }
    int64_t q = i;
    while (q < end) {
      const unsigned ht = recs[q * rw + nw - 1] & 63u;
      const unsigned head = ht >> 3, tail = ht & 7;
      const int both = head != SENTINEL && tail != SENTINEL;
      if (both) counting[cht[ht] > MHBO_MAX_MUL ? MHBO_MAX_MUL : cht[ht]]++; /* edge_counter.h:29-32 */
      const int solid = both && cht[ht] >= m;
      for (int64_t j = 0; j < cht[ht]; ++j, ++q) {
        const uint64_t info = (((uint64_t)recs[q * rw + nw] << 32) | recs[q * rw + nw + 1]) >> 6;
        const unsigned strand = (unsigned)(info & 1);
        const uint64_t pos = info >> 1; /* full offset of the (k-1)-mer */
        const uint64_t r = read_of_offset(base_off, n_reads, pos);
        const int64_t offset = (int64_t)(pos - base_off[r]) - 1;
        const int64_t l_off = strand == 0 ? offset : offset + 1, r_off = strand == 0 ? offset + 1 : offset;
#define CAND(off, code)                                                               \
  if (u64vec_push(cand, ((base_off[r] + (uint64_t)(off)) << 2) | (uint64_t)(code))) return -2;
        if (solid) {
          bit_set(is_solid, pos - 1); /* :441 */
          if (!(has_in & (1 << head))) CAND(l_off, 1 + strand);
          if (!(has_out & (1 << tail))) CAND(r_off, 2 - strand);
        } else { /* :461-551 */
          if (l_has_out & (1 << head)) {
            if (has_in & (1 << head)) {
              CAND(l_off, 0);
            } else {
              CAND(l_off, 1 + strand);
            }
          } else if (has_in & (1 << head)) {
            CAND(l_off, 2 - strand);
          }
          if (r_has_in & (1 << tail)) {
            if (has_out & (1 << tail)) {
              CAND(r_off, 0);
            } else {
              CAND(r_off, 2 - strand);
            }
          } else if (has_out & (1 << tail)) {
            CAND(r_off, 1 + strand);
          }
        }
#undef CAND
      }
    }
  }
  return 0;
}

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

static int stage1(const mhbo_seqs *reads, unsigned k, int m, const uint64_t *base_off, uint8_t *is_solid, u64vec *cand,
                  int64_t *counting) {
  const unsigned kk = k - 1, nw = div_ceil_u(2 * kk + 6, 32), rw = nw + 2;
  uint64_t *bsize = (uint64_t *)calloc(MHBO_NUM_BUCKETS + 1, 8);
  if (!bsize) return -2;
  uint32_t *read_recs = NULL;
  unsigned read_cap = 0;
  uint32_t *recs = NULL;
  /* two sweeps in global read order: bucket sizes (Lv0CalcBucketSize :137-205), then the records in place
   * (Lv1FillOffsets :207-293 + Lv2ExtractSubString): a bucket holds its records in read order */
  uint64_t *cursor = NULL;
  int rc = 0;
  for (int sweep = 0; sweep < 2 && rc == 0; ++sweep) {
    if (sweep == 1) {
      uint64_t acc = 0;
      cursor = (uint64_t *)malloc(MHBO_NUM_BUCKETS * 8);
      for (unsigned b = 0; b < MHBO_NUM_BUCKETS; ++b) {
        cursor[b] = acc;
        uint64_t c = bsize[b];
        bsize[b] = acc;
        acc += c;
      }
      bsize[MHBO_NUM_BUCKETS] = acc;
      recs = (uint32_t *)malloc((size_t)(acc ? acc : 1) * rw * 4);
      if (!recs || !cursor) rc = -2;
    }
    for (uint64_t r = 0; r < reads->n && rc == 0; ++r) {
      const unsigned L = reads->len[r];
      if (L < k + 1) continue;
      const uint32_t *w = reads->words + reads->word_off[r];
      if (L - k + 4 > read_cap) {
        read_cap = 2 * (L - k + 4);
        read_recs = (uint32_t *)realloc(read_recs, (size_t)read_cap * rw * 4);
        if (!read_recs) {
          rc = -2;
          break;
        }
      }
      const unsigned ne = mhbo_s1_read_records(w, L, k, base_off[r], read_recs);
      for (unsigned e = 0; e < ne; ++e) {
        const uint32_t *t = read_recs + (size_t)e * rw;
        const unsigned bucket = t[0] >> 16;
        if (sweep == 0) {
          bsize[bucket]++;
        } else {
          memcpy(recs + cursor[bucket]++ * rw, t, 4 * rw);
        }
      }
    }
  }
  free(read_recs);
  if (rc == 0) {
    const recfmt f = {nw, rw};
    for (unsigned b = 0; b < MHBO_NUM_BUCKETS && rc == 0; ++b) {
      const int64_t n = (int64_t)(bsize[b + 1] - bsize[b]);
      if (n == 0) continue;
      kms_sort(recs + bsize[b] * rw, n, &f); /* base_engine.cpp:344 */
      rc = s1_postprocess(recs + bsize[b] * rw, n, k, m, nw, base_off, reads->n, is_solid, cand, counting);
    }
  }
  free(bsize);
  free(cursor);
  free(recs);
  return rc;
}

/* Read2SdbgS2::Initialize, read_to_sdbg_s2.cpp:117-263: mercy edges become solid bits */
static uint64_t add_mercy(const mhbo_seqs *reads, unsigned k, const uint64_t *base_off, uint8_t *is_solid, u64vec *cand) {
  uint64_t num_mercy = 0, maxlen = 1;
  for (uint64_t r = 0; r < reads->n; ++r)
    if (reads->len[r] > maxlen) maxlen = reads->len[r];
  uint8_t *no_in = (uint8_t *)malloc(3 * (maxlen + 2));
  uint8_t *no_out = no_in + maxlen + 2, *has_solid = no_out + maxlen + 2;
  qsort(cand->v, cand->n, 8, cmp_u64);
  uint64_t ci = 0;
  while (ci < cand->n) {
    const uint64_t r = read_of_offset(base_off, reads->n, cand->v[ci] >> 2);
    const uint64_t L = reads->len[r];
    int64_t first_0_out = (int64_t)maxlen + 1, last_0_in = -1;
    memset(no_in, 0, 3 * (maxlen + 2));
    while (ci < cand->n && (cand->v[ci] >> 2) < base_off[r] + L) {
      const int64_t off = (int64_t)((cand->v[ci] >> 2) - base_off[r]);
      const unsigned code = (unsigned)(cand->v[ci] & 3);
      if (code == 2) {
        no_out[off] = 1;
        if (off < first_0_out) first_0_out = off;
      } else if (code == 1) {
        no_in[off] = 1;
        if (off > last_0_in) last_0_in = off;
      }
      has_solid[off] = 1;
      ++ci;
    }
    if (last_0_in < first_0_out) continue;
    for (uint64_t i = 0; i + k < L; ++i)
      if (bit_get(is_solid, base_off[r] + i)) has_solid[i] = has_solid[i + 1] = 1;
    int64_t last_no_out = -1;
    for (uint64_t i = 0; i + k <= L; ++i) {
      if (no_in[i] && last_no_out != -1) {
        for (uint64_t j = (uint64_t)last_no_out; j < i; ++j) bit_set(is_solid, base_off[r] + j);
        num_mercy += i - (uint64_t)last_no_out;
      }
      if (has_solid[i]) last_no_out = -1;
      if (no_out[i]) last_no_out = (int64_t)i;
    }
  }
  free(no_in);
  return num_mercy;
}

/* ------------------------------------------------------------------------------------------------
 * stage 2 (read_to_sdbg_s2.cpp)
 * ---------------------------------------------------------------------------------------------- */
/* Lv2ExtractSubString :438-519: the sort item of (edge position i, strand, edge_type 0 = left $, 1 = solid, 2 = right $) */
static void s2_record(const uint32_t *w, unsigned k, unsigned i, unsigned strand, unsigned type, unsigned W, uint32_t *rec) {
  unsigned nc = k, prev = SENTINEL;
  memset(rec, 0, 4 * W);
  if (strand == 0) {
    unsigned off = i;
    if (type == 1) {
      prev = sbase(w, i);
      off = i + 1;
    } else if (type == 2) {
      prev = sbase(w, i + 1);
      off = i + 2;
      nc = k - 1;
    }
    for (unsigned j = 0; j < nc; ++j) pbase(rec, j, sbase(w, off + j));
  } else {
    unsigned off = i;
    if (type == 0) {
      nc = k - 1;
      prev = 3u - sbase(w, i + k - 1);
    } else if (type == 1) {
      prev = 3u - sbase(w, i + k);
    } else {
      off = i + 1;
    }
    for (unsigned j = 0; j < nc; ++j) pbase(rec, j, 3u - sbase(w, off + nc - 1 - j));
  }
  rec[W - 1] |= (uint32_t)(nc == k) << 3;
  rec[W - 1] |= prev;
}

void mhbo_s2_record(const uint32_t *w, unsigned k, unsigned i, unsigned strand, unsigned type, uint32_t *rec, int *palindrome) {
  s2_record(w, k, i, strand, type, div_ceil_u(2 * k + 4, 32), rec);
  int pal = 1;
  for (unsigned j = 0; j <= k && pal; ++j) pal = sbase(w, i + j) == 3u - sbase(w, i + k - j);
  *palindrome = pal;
}

static int cmp_words_r(const void *a, const void *b, void *arg) {
  unsigned nw = *(const unsigned *)arg;
  const uint32_t *x = (const uint32_t *)a, *y = (const uint32_t *)b;
  for (unsigned i = 0; i < nw; ++i)
    if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  return 0;
}

static inline int s2_a(const uint32_t *item, unsigned W, unsigned k) { /* Extract_a :69-83 */
  return ((item[W - 1] >> 3) & 1u) ? (int)sbase(item, k - 1) : (int)SENTINEL;
}
static inline int s2_b(const uint32_t *item, unsigned W) { return (int)(item[W - 1] & 7u); } /* :85-88 */

static int stage2(const mhbo_seqs *reads, unsigned k, int m, const uint64_t *base_off, const uint8_t *is_solid,
                  mhbo_sdbg_out *out) {
  const unsigned W = div_ceil_u(2 * k + 4, 32), WPT = div_ceil_u(k, 16);
  out->words_per_tip_label = WPT;
  const int sure = m == 1;
  uint32_t *rec = NULL;
  uint64_t n_rec = 0;
  uint32_t tmp[64];
  /* Lv0CalcBucketSize :265-345 / Lv1FillOffsets :347-436: two sweeps, count then fill */
  for (int sweep = 0; sweep < 2; ++sweep) {
    uint64_t ri = 0;
    for (uint64_t r = 0; r < reads->n; ++r) {
      const unsigned L = reads->len[r];
      if (L < k + 1) continue;
      const uint32_t *w = reads->words + reads->word_off[r];
      for (unsigned i = 0; i + k < L; ++i) {
        if (!(sure || bit_get(is_solid, base_off[r] + i))) continue;
        int pal = 1;
        for (unsigned j = 0; j <= k && pal; ++j) pal = sbase(w, i + j) == 3u - sbase(w, i + k - j);
        unsigned types[3], nt = 0;
        if (i == 0 || !(sure || bit_get(is_solid, base_off[r] + i - 1))) types[nt++] = 0;
        types[nt++] = 1;
        if (i + k == L - 1 || !(sure || bit_get(is_solid, base_off[r] + i + 1))) types[nt++] = 2;
        for (unsigned t = 0; t < nt; ++t) {
          for (unsigned strand = 0; strand < (pal ? 1u : 2u); ++strand) {
            if (sweep == 1) {
              s2_record(w, k, i, strand, types[t], W, tmp);
              memcpy(rec + ri * W, tmp, 4 * W);
            }
            ++ri;
          }
        }
      }
    }
    if (sweep == 0) {
      n_rec = ri;
      rec = (uint32_t *)malloc((size_t)(n_rec ? n_rec : 1) * W * 4);
      if (!rec) return -2;
    }
  }
  out->n_records = n_rec;
  unsigned nwq = W;
  qsort_r(rec, (size_t)n_rec, 4 * W, cmp_words_r, &nwq); /* whole-record keys: ties are identical records */

  /* Lv2Postprocess :521-614 + SdbgWriter::Write (sdbg_writer.cpp:25-58).  Pass 0 sizes, pass 1 writes. */
  for (int pass = 0; pass < 2; ++pass) {
    uint64_t byte_pos = 0;
    int cur_bucket = -1;
    if (pass == 1) out->bytes = (uint8_t *)malloc(out->bucket_byte_off[MHBO_NUM_BUCKETS] + 1);
    for (uint64_t start = 0, end; start < n_rec; start = end) {
      end = start + 1;
      while (end < n_rec && !diff_km1(rec + start * W, rec + end * W, k)) ++end;
      int has_solid_a = 0, has_solid_b = 0, outputed_b = 0;
      int64_t last_a[4] = {-1, -1, -1, -1};
      for (uint64_t i = start; i < end; ++i) { /* :539-553 */
        int a = s2_a(rec + i * W, W, k), b = s2_b(rec + i * W, W);
        if (a != (int)SENTINEL && b != (int)SENTINEL) {
          has_solid_a |= 1 << a;
          has_solid_b |= 1 << b;
        }
        if (a != (int)SENTINEL && (b != (int)SENTINEL || !(has_solid_a & (1 << a)))) last_a[a] = (int64_t)i;
      }
      for (uint64_t i = start, j; i < end; i = j) { /* :555-611 */
        const uint32_t *cur = rec + i * W;
        int a = s2_a(cur, W, k), b = s2_b(cur, W);
        j = i + 1;
        while (j < end && s2_a(rec + j * W, W, k) == a && s2_b(rec + j * W, W) == b) ++j;
        int is_dollar = 0;
        uint64_t count = j - i > MHBO_MAX_MUL ? MHBO_MAX_MUL : j - i;
        if (a == (int)SENTINEL) {
          if (has_solid_b & (1 << b)) continue;
          is_dollar = 1;
        }
        if (b == (int)SENTINEL && (has_solid_a & (1 << a))) continue;
        int w = (b == (int)SENTINEL) ? 0 : ((outputed_b & (1 << b)) ? b + 5 : b + 1);
        int last = (a == (int)SENTINEL) ? 0 : (last_a[a] == (int64_t)j - 1);
        outputed_b |= 1 << b;
        unsigned mul = (unsigned)count, bucket = cur[0] >> 16;
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
          p[0] = (uint8_t)(w | (last << 4) | (is_dollar << 5));
          p[1] = (uint8_t)(mul > 255 ? 255 : mul);
          byte_pos += 2;
          if (mul > 254) {
            uint16_t m16 = (uint16_t)mul;
            memcpy(out->bytes + byte_pos, &m16, 2);
            byte_pos += 2;
          }
          if (is_dollar) { /* :602-606: tip label = the item's raw leading words */
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

/* main_read2sdbg, main_sdbg_build.cpp:88-156 */
int mhbo_read2sdbg(const mhbo_seqs *reads, uint32_t k, int32_t m, int need_mercy, mhbo_sdbg_out *out, int64_t *counting,
                   uint64_t *n_mercy_out, uint8_t **is_solid_out, uint64_t *n_bases_out) {
  if (k < 9 || k > 255 || m < 1) return -1;
  memset(out, 0, sizeof(*out));
  memset(counting, 0, (MHBO_MAX_MUL + 1) * sizeof(int64_t));
  uint64_t *base_off = (uint64_t *)malloc((reads->n + 1) * 8);
  if (!base_off) return -2;
  base_off[0] = 0;
  for (uint64_t r = 0; r < reads->n; ++r) base_off[r + 1] = base_off[r] + reads->len[r];
  const uint64_t n_bases = base_off[reads->n];
  uint8_t *is_solid = (uint8_t *)calloc(n_bases / 8 + 2, 1);
  u64vec cand = {NULL, 0, 0};
  int rc = is_solid ? 0 : -2;
  uint64_t n_mercy = 0;
  if (rc == 0 && m > 1) { /* :141-147: stage 1 only when the threshold can reject anything */
    rc = stage1(reads, k, m, base_off, is_solid, &cand, counting);
    if (rc == 0 && need_mercy) n_mercy = add_mercy(reads, k, base_off, is_solid, &cand);
  }
  if (rc == 0) rc = stage2(reads, k, m, base_off, is_solid, out);
  if (n_mercy_out) *n_mercy_out = n_mercy;
  if (n_bases_out) *n_bases_out = n_bases;
  if (is_solid_out && rc == 0) {
    *is_solid_out = is_solid;
    is_solid = NULL;
  }
  free(cand.v);
  free(is_solid);
  free(base_off);
  return rc;
}
