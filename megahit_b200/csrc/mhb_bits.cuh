// mhb_bits.cuh -- 2-bit sequence arithmetic shared by host and device code.
//
// Packing convention (voutcn/megahit src/kmlib/kmcompactvector.h:53-57, kBigEndian): base i of a
// sequence sits in word i/16 at bits (31-2(i%16), 30-2(i%16)); A=0 C=1 G=2 T=3.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MHB_HD __host__ __device__ __forceinline__
#else
#define MHB_HD inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

namespace mhb {

MHB_HD u32 div_ceil(u32 a, u32 b) { return (a + b - 1) / b; }

// reverse the order of the sixteen 2-bit groups of a word (kmlib/kmbit.h:82-89 Reverse<2>)
MHB_HD u32 rev2(u32 x) {
#if defined(__CUDA_ARCH__)
  x = __brev(x);
#else
  x = ((x >> 16) | (x << 16));
  x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
  x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
  x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
#endif
  return ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
}

// (hi:lo) << s, upper 32 bits; s in [0,31]
MHB_HD u32 fshl(u32 hi, u32 lo, u32 s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

MHB_HD u32 base_at(const u32 *w, u32 i) { return (w[i >> 4] >> (30 - 2 * (i & 15))) & 3u; }

// keep the top `bits` bits of a word (bits in [0,32])
MHB_HD u32 top_mask(u32 bits) { return bits >= 32 ? 0xFFFFFFFFu : (bits == 0 ? 0u : ~(0xFFFFFFFFu >> bits)); }

// Substring [start, start+nb) of a word-aligned sequence with `nwords` words, left-aligned into W
// words, tail bits zero (the result CopySubstring produces, sequence/copy_substr.h:53-101).
template <int W>
MHB_HD void load_sub(const u32 *s, u32 nwords, u32 start, u32 nb, u32 (&out)[W]) {
  const u32 w0 = start >> 4, sh = (start & 15) * 2;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    u32 a = (w0 + j < nwords) ? s[w0 + j] : 0u;
    u32 b = (sh && w0 + j + 1 < nwords) ? s[w0 + j + 1] : 0u;
    u32 v = fshl(a, b, sh);
    int keep = (int)(2 * nb) - 32 * j;
    out[j] = keep <= 0 ? 0u : (v & top_mask((u32)keep));
  }
}

// r[idx] with a runtime index, compiled to selects (keeps r in registers); 0 when idx >= W
template <int W>
MHB_HD u32 pick(const u32 (&r)[W], u32 idx) {
  u32 v = 0;
#pragma unroll
  for (int j = 0; j < W; ++j) v = (idx == (u32)j) ? r[j] : v;
  return v;
}

// Reverse (no complement) of an nb-base string held left-aligned in W words; result left-aligned.
template <int W>
MHB_HD void reverse_sub(const u32 (&in)[W], u32 nb, u32 (&out)[W]) {
  u32 r[W];
#pragma unroll
  for (int i = 0; i < W; ++i) r[i] = rev2(in[W - 1 - i]);
  // the full reversal leaves the string right-aligned: shift left by t bits
  const u32 t = 32u * W - 2u * nb;
  const u32 tw = t >> 5, ts = t & 31;
#pragma unroll
  for (int i = 0; i < W; ++i) out[i] = fshl(pick<W>(r, i + tw), pick<W>(r, i + tw + 1), ts);
}

// Complement of an nb-base string (tail bits stay zero).
template <int W>
MHB_HD void complement_sub(const u32 (&in)[W], u32 nb, u32 (&out)[W]) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    int keep = (int)(2 * nb) - 32 * j;
    out[j] = keep <= 0 ? 0u : (~in[j] & top_mask((u32)keep));
  }
}

template <int W>
MHB_HD bool less_words(const u32 (&a)[W], const u32 (&b)[W]) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (a[j] != b[j]) return a[j] < b[j];
  }
  return false;
}

// ---- geometry -------------------------------------------------------------------------------
MHB_HD u32 count_key_words(u32 k) { return div_ceil(2 * (k + 1), 32); }
MHB_HD u32 count_record_words(u32 k) { return div_ceil(2 * (k + 1) + 6, 32); }
MHB_HD u32 words_per_edge(u32 k) { return div_ceil(2 * (k + 1) + 16, 32); }
MHB_HD u32 s2s_record_words(u32 k) { return div_ceil(2 * k + 20, 32); }
MHB_HD u32 words_per_tip_label(u32 k) { return div_ceil(k, 16); }

}  // namespace mhb
