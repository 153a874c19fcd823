// mhb_count.cuh -- `count` stage kernels: edge extraction, solid-edge counting, mercy bookkeeping.
// Reference: voutcn/megahit src/sorting/kmer_counter.cpp (line numbers cited per kernel).
#pragma once
#include "mhb.h"
#include "mhb_kernels.cuh"

namespace mhb {

// ------------------------------------------------------------------------------------------------
// Streaming reads through shared memory with 1-D TMA (cp.async.bulk + mbarrier), two stages.
// A CTA walks batches of kReadsPerBatch consecutive reads (static round-robin over the grid); warp w
// handles reads w, w+NW, ... of the batch and calls f(read_id, s, nwords, L) with s pointing at the
// read's packed words (shared memory, or global memory for batches that do not fit a stage).
// ------------------------------------------------------------------------------------------------
static constexpr int kReadsPerBatch = 64;
static constexpr int kStageWords = 4096;  // 16 KiB per stage
static constexpr int kExtractThreads = 256;

struct ReadsView {
  const u32 *bin;
  u64 bin_words;
  u64 n_reads;
  u32 fixed_len;
  u32 fixed_stride;  // 1 + ceil(fixed_len/16)
  const u64 *rec_off;
  const u64 *edge_off;
  __device__ __forceinline__ u64 rec_start(u64 r) const { return fixed_len ? r * fixed_stride : rec_off[r]; }
};

template <class F>
__device__ __forceinline__ void for_each_read(const ReadsView &rv, F &&f) {
  __shared__ __align__(16) u32 s_stage[2][kStageWords];
  __shared__ __align__(8) u64 s_bar[2];
  const u32 tid = threadIdx.x, warp = tid >> 5;
  constexpr int NW = kExtractThreads / 32;
  const u64 n_batches = (rv.n_reads + kReadsPerBatch - 1) / kReadsPerBatch;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // geometry of a batch: [a0, a1) = 16-byte aligned word range holding its records
  auto geom = [&](u64 b, u64 &r0, u64 &r1, u64 &a0, u64 &a1) {
    r0 = b * kReadsPerBatch;
    r1 = r0 + kReadsPerBatch < rv.n_reads ? r0 + kReadsPerBatch : rv.n_reads;
    const u64 w0 = rv.rec_start(r0);
    const u64 w1 = r1 < rv.n_reads ? rv.rec_start(r1) : rv.bin_words;
    a0 = w0 & ~3ull;
    a1 = (w1 + 3) & ~3ull;
  };
  auto issue = [&](u64 b, int st) {
    u64 r0, r1, a0, a1;
    geom(b, r0, r1, a0, a1);
    if (a1 - a0 <= (u64)kStageWords && tid == 0) {
      const u32 bytes = (u32)(a1 - a0) * 4;
      mbar_expect_tx(&s_bar[st], bytes);
      bulk_g2s(&s_stage[st][0], rv.bin + a0, bytes, &s_bar[st]);
    }
  };

  u32 parity[2] = {0, 0};
  u64 b = blockIdx.x;
  int st = 0;
  if (b < n_batches) issue(b, 0);
  for (; b < n_batches; b += gridDim.x, st ^= 1) {
    const u64 nb = b + gridDim.x;
    if (nb < n_batches) issue(nb, st ^ 1);  // stage st^1 was released by the __syncthreads below
    u64 r0, r1, a0, a1;
    geom(b, r0, r1, a0, a1);
    const bool staged = (a1 - a0) <= (u64)kStageWords;
    if (staged) {
      mbar_wait(&s_bar[st], parity[st]);
      parity[st] ^= 1;
    }
    for (u64 r = r0 + warp; r < r1; r += NW) {
      const u64 w = rv.rec_start(r);
      const u32 *rec = staged ? &s_stage[st][w - a0] : rv.bin + w;
      const u32 L = rec[0];
      f(r, rec + 1, div_ceil(L, 16), L);
    }
    __syncthreads();  // everyone is done with stage st before it is refilled
  }
}

// ------------------------------------------------------------------------------------------------
// K-extract (A1-A3; kmer_counter.cpp:114-252): one record per (k+1)-mer occurrence.
// ------------------------------------------------------------------------------------------------
template <int W, int WR>
__global__ void __launch_bounds__(kExtractThreads)
    k_count_extract(ReadsView rv, u32 k, u32 *__restrict__ records, u64 *hist, int hist_byte) {
  __shared__ u32 s_hist[256];
  for (int i = threadIdx.x; i < 256; i += kExtractThreads) s_hist[i] = 0;
  const u32 lane = lane_id();
  const u32 K1 = k + 1;
  for_each_read(rv, [&](u64 r, const u32 *s, u32 nwords, u32 L) {
    if (L < K1) return;  // kmer_counter.cpp:124
    const u32 n_e = L - k;
    const u64 base = rv.fixed_len ? r * (u64)(rv.fixed_len - k) : rv.edge_off[r];
    for (u32 q0 = 0; q0 < n_e; q0 += 32) {
      const u32 q = q0 + lane;
      if (q < n_e) {
        u32 rec[WR], strand;
        make_count_record<W, WR>(s, nwords, L, k, q, rec, strand);
        st_rec<WR>(records, base + q, rec);
        if (hist) atomicAdd(&s_hist[rec_byte<WR>(rec, hist_byte)], 1u);
      }
    }
  });
  __syncthreads();
  if (hist)
    for (int i = threadIdx.x; i < 256; i += kExtractThreads)
      if (s_hist[i]) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)s_hist[i]);
}

// ------------------------------------------------------------------------------------------------
// K-count (A5/A6; kmer_counter.cpp:254-381).  v1: one thread per run head walks its run.
// info[i] = 0 for non-heads, else 1 | solid<<1 | no_in<<2 | no_out<<3 | min(count,65535)<<8
// ------------------------------------------------------------------------------------------------
template <int WR>
__device__ __forceinline__ bool same_edge(const u32 (&a)[WR], const u32 (&b)[WR]) {
  bool eq = true;
#pragma unroll
  for (int j = 0; j < WR; ++j) {
    const u32 m = (j == WR - 1) ? ~63u : ~0u;
    eq = eq && ((a[j] & m) == (b[j] & m));
  }
  return eq;
}

static constexpr int kMulHistSmem = 1024;

template <int WR>
__global__ void __launch_bounds__(256)
    k_count_mark(const u32 *__restrict__ recs, u64 n, int m, u32 *__restrict__ info, u64 *mul_hist) {
  __shared__ u32 s_hist[kMulHistSmem];
  for (int i = threadIdx.x; i < kMulHistSmem; i += 256) s_hist[i] = 0;
  __syncthreads();
  const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    u32 r[WR], p[WR];
    ld_rec<WR>(recs, i, r);
    bool head = true;
    if (i > 0) {
      ld_rec<WR>(recs, i - 1, p);
      head = !same_edge<WR>(r, p);
    }
    u32 word = 0;
    if (head) {
      u32 count = 0, cp[5] = {0, 0, 0, 0, 0}, cn[5] = {0, 0, 0, 0, 0};
      u32 x[WR];
#pragma unroll
      for (int j = 0; j < WR; ++j) x[j] = r[j];
      u64 j = i;
      while (true) {
        const u32 pn = x[WR - 1] & 63u;  // kmer_counter.cpp:288-295
        count += count < 0x40000000u ? 1u : 0u;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          cp[c] += ((pn >> 3) == (u32)c && cp[c] < 0x40000000u) ? 1u : 0u;
          cn[c] += ((pn & 7u) == (u32)c && cn[c] < 0x40000000u) ? 1u : 0u;
        }
        if (++j >= n) break;
        ld_rec<WR>(recs, j, x);
        if (!same_edge<WR>(x, r)) break;
      }
      bool has_in = false, has_out = false;  // :297-305
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        has_in = has_in || (long long)cp[c] >= (long long)m;
        has_out = has_out || (long long)cn[c] >= (long long)m;
      }
      const bool solid = (long long)count >= (long long)m;
      const u32 c16 = count > 65535u ? 65535u : count;
      word = 1u | (solid ? 2u : 0u) | (has_in ? 0u : 4u) | (has_out ? 0u : 8u) | (c16 << 8);
      if (c16 < (u32)kMulHistSmem) atomicAdd(&s_hist[c16], 1u);  // edge_counter.h:30-33
      else atomicAdd((unsigned long long *)&mul_hist[c16], 1ull);
    }
    info[i] = word;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kMulHistSmem; c += 256)
    if (s_hist[c]) atomicAdd((unsigned long long *)&mul_hist[c], (unsigned long long)s_hist[c]);
}

// ordered compaction helpers: per-block totals, single-block scan of the totals, apply
static constexpr int kCompactThreads = 256, kCompactItems = 8, kCompactTile = kCompactThreads * kCompactItems;

__global__ void __launch_bounds__(kCompactThreads) k_solid_block_totals(const u32 *info, u64 n, u64 *btot) {
  __shared__ u32 s_scan[kCompactThreads / 32 + 1];
  const u64 base = (u64)blockIdx.x * kCompactTile + (u64)threadIdx.x * kCompactItems;
  u32 c = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j)
    if (base + j < n) c += (info[base + j] >> 1) & 1u;
  u32 total;
  block_excl_scan<kCompactThreads>(c, s_scan, total);
  if (threadIdx.x == 0) btot[blockIdx.x] = total;
}

// in-place exclusive scan of nb u64 values by ONE block; writes the grand total
__global__ void __launch_bounds__(1024) k_scan_u64(u64 *v, u64 nb, u64 *total_out) {
  __shared__ u64 s_w[33];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const u32 lane = lane_id(), warp = threadIdx.x >> 5;
  for (u64 base = 0; base < nb; base += 1024) {
    const u64 i = base + threadIdx.x;
    const u64 x = i < nb ? v[i] : 0;
    u64 inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u64 t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (u32)d) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u64 w = s_w[lane], winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= (u32)d) winc += t;
      }
      s_w[lane] = winc - w;
      if (lane == 31) s_w[32] = winc;
    }
    __syncthreads();
    if (i < nb) v[i] = s_carry + s_w[warp] + inc - x;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_w[32];
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

// PackEdge (kmer_counter.cpp:32-52) for every solid run head, in sorted order
template <int WR>
__global__ void __launch_bounds__(kCompactThreads)
    k_count_emit(const u32 *__restrict__ recs, const u32 *__restrict__ info, u64 n, u32 k, const u64 *btot,
                 u32 *__restrict__ edges, uint8_t *__restrict__ aux, u64 capacity) {
  __shared__ u32 s_scan[kCompactThreads / 32 + 1];
  const u64 base = (u64)blockIdx.x * kCompactTile + (u64)threadIdx.x * kCompactItems;
  u32 w[kCompactItems];
  u32 c = 0;
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    w[j] = base + j < n ? info[base + j] : 0u;
    c += (w[j] >> 1) & 1u;
  }
  u32 total;
  const u32 excl = block_excl_scan<kCompactThreads>(c, s_scan, total);
  if (c == 0) return;
  u64 pos = btot[blockIdx.x] + excl;
  const u32 W = count_key_words(k), WE = words_per_edge(k);
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    if ((w[j] >> 1) & 1u) {
      if (pos < capacity) {
        u32 r[WR];
        ld_rec<WR>(recs, base + j, r);
        r[WR - 1] &= ~63u;  // drop prev/next; what is left of the key words has zero tail bits
        u32 *e = edges + pos * WE;
        for (u32 x = 0; x < WE; ++x) e[x] = (x < W && x < (u32)WR) ? pick<WR>(r, x) : 0u;
        e[WE - 1] |= w[j] >> 8;  // min(count, kMaxMul)
        aux[pos] = (uint8_t)((w[j] >> 2) & 3u);
      }
      ++pos;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tip set: open-addressing hash set of the solid edges that lack an incoming or outgoing solid
// neighbour (aux != 0).  Layout: u64 capacity (power of two), u64 pad, then capacity entries of
// (1 + W) words: flags (0 = empty) followed by the (k+1)-mer.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline u64 tipset_capacity(u64 n_tip) {
  u64 c = 1024;
  while (c < 2 * n_tip) c <<= 1;
  return c;
}

template <int W>
__device__ __forceinline__ u64 hash_key(const u32 (&key)[W]) {
  u64 h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    h ^= key[j];
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  return h;
}

__global__ void k_count_tips(const uint8_t *aux, u64 n, unsigned long long *out) {
  u64 c = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    c += aux[i] != 0;
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane_id() == 0 && c) atomicAdd(out, (unsigned long long)c);
}

template <int W>
__global__ void k_tipset_insert(const u32 *__restrict__ edges, const uint8_t *__restrict__ aux, u64 n, u32 k,
                                u32 *table, u64 cap) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || aux[i] == 0) return;
  const u32 WE = words_per_edge(k);
  u32 key[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const u32 keep = 2 * (k + 1) - 32 * j;
    key[j] = edges[i * WE + j] & top_mask(keep > 32 ? 32 : keep);
  }
  u64 slot = hash_key<W>(key) & (cap - 1);
  while (true) {
    u32 *e = table + slot * (W + 1);
    if (atomicCAS(e, 0u, (u32)aux[i]) == 0u) {
#pragma unroll
      for (int j = 0; j < W; ++j) e[1 + j] = key[j];
      return;
    }
    slot = (slot + 1) & (cap - 1);
  }
}

// K-mercy (kmer_counter.cpp:307-367): per read min/max over the occurrences of tip edges.
template <int W, int WR>
__global__ void __launch_bounds__(kExtractThreads)
    k_mark_mercy(ReadsView rv, u32 k, const u32 *__restrict__ table, u64 cap, u32 *first_0_out, u32 *last_0_in) {
  const u32 lane = lane_id();
  const u32 K1 = k + 1;
  for_each_read(rv, [&](u64 r, const u32 *s, u32 nwords, u32 L) {
    u32 first = 0xFFFFFFFFu;
    long long last = -1;
    if (L >= K1) {
      const u32 n_e = L - k;
      for (u32 q0 = 0; q0 < n_e; q0 += 32) {
        const u32 q = q0 + lane;
        if (q < n_e) {
          u32 rec[WR], strand;
          make_count_record<W, WR>(s, nwords, L, k, q, rec, strand);
          u32 key[W];
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const u32 keep = 2 * K1 - 32 * j;
            key[j] = rec[j] & top_mask(keep > 32 ? 32 : keep);
          }
          u64 slot = hash_key<W>(key) & (cap - 1);
          u32 flags = 0;
          while (true) {
            const u32 *e = table + slot * (W + 1);
            const u32 f = e[0];
            if (f == 0) break;
            bool eq = true;
#pragma unroll
            for (int j = 0; j < W; ++j) eq = eq && e[1 + j] == key[j];
            if (eq) {
              flags = f;
              break;
            }
            slot = (slot + 1) & (cap - 1);
          }
          if (flags) {
            const u32 off = L - K1 - q;  // offset in the reversed (package) read
            const bool upd_last_in = (flags & 1u) && strand == 0;   // no in,  strand 0 -> last
            const bool upd_first_in = (flags & 1u) && strand == 1;  // no in,  strand 1 -> first
            const bool upd_first_out = (flags & 2u) && strand == 0; // no out, strand 0 -> first
            const bool upd_last_out = (flags & 2u) && strand == 1;  // no out, strand 1 -> last
            if (upd_last_in || upd_last_out) last = last > (long long)off ? last : (long long)off;
            if (upd_first_in || upd_first_out) first = first < off + 1 ? first : off + 1;
          }
        }
      }
    }
    for (int d = 16; d; d >>= 1) {
      const u32 f2 = __shfl_xor_sync(0xffffffffu, first, d);
      const long long l2 = __shfl_xor_sync(0xffffffffu, last, d);
      first = first < f2 ? first : f2;
      last = last > l2 ? last : l2;
    }
    if (lane == 0) {
      first_0_out[r] = first;
      last_0_in[r] = last < 0 ? 0xFFFFFFFFu : (u32)last;
    }
  });
}

}  // namespace mhb
