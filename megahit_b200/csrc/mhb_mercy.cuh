// mhb_mercy.cuh -- mercy-edge generation on the device (A11; seq_to_sdbg.cpp:100-357 GenMercyEdges).
//
// For every candidate read (kmer_counter.cpp:390-401) and every k-mer position i the reference decides
// has_in[i] / has_out[i] by binary searches in the SORTED solid-edge array and then adds every (k+1)-mer
// between a "has in, no out" position and the next "has out, no in" position as a multiplicity-1 edge.
// The reference walks a read left to right, but position i only feeds position i+1 through the flag
// "this out-search also proves has_in[i+1]" (N_i below), so all positions are independent:
//     has_in[i] = N_{i-1} | A_i        has_out[i] = O_i
// where A_i is the outcome of the in-searches (seq_to_sdbg.cpp:225-251) and (O_i, N_i) of the
// out-searches (:254-298).  k_mercy_probe evaluates (A, O, N) for every (read, position) in parallel;
// k_mercy_emit replays the cheap sequential scan (:310-352) per read.
//
// Bit-exactness note: a k-mer can prefix several edges and the reference looks at base k of the one its
// binary search happens to hit (:258-262), so the probe sequence is reproduced exactly: bounds from the
// 12-mer look-up table (InitLookupTable :100-127, here two bisections on the 24-bit prefix), then
// mid = (l + r) / 2 with three-way compares (BinarySearchKmer :132-161).
#pragma once
#include "mhb.h"
#include "mhb_kernels.cuh"

namespace mhb {

// candidate reads: first_0_out / last_0_in both set and last > first (kmer_counter.cpp:395-401)
static __global__ void k_cand_flags(const u32 *first, const u32 *last, u64 n_reads, u32 *flag) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const u32 f = first[r], l = last[r];
  flag[r] = (f != MHB_SENTINEL_OFFSET && l != MHB_SENTINEL_OFFSET && l > f) ? 1u : 0u;
}
static __global__ void k_cand_compact(const u32 *flag, const u64 *off, u64 n_reads, u64 *ids) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads && flag[r]) ids[off[r]] = r;
}

// The sorted solid edges as up to 16 segments, segment o holding every edge whose leading byte maps to o (one
// segment = everything on a single GPU; several = the bucket ranges of the ranks of a multi-GPU build, each in its
// owner's memory and read through CUDA IPC peer pointers - no gather of the edges is needed).
struct EdgeSegs {
  const u32 *ptr[16];
  long long n[16];
  const uint2 *lut[16];  // per segment: [first,last] edge index of every 12-base prefix (0xFFFFFFFF = none)
  uint8_t owner[256];
};

// InitLookupTable (seq_to_sdbg.cpp:100-127) for one segment: lut[p] = {first, last} index of the edges whose first
// 12 bases are p.  lut must be pre-filled with 0xFF.
static constexpr u32 kLutEntries = 1u << 24;
static __global__ void k_edge_lut(const u32 *__restrict__ edges, u64 n, u32 we, uint2 *lut) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 p = edges[i * we] >> 8;
    if (i == 0 || (edges[(i - 1) * we] >> 8) != p) lut[p].x = (u32)i;
    if (i + 1 == n || (edges[(i + 1) * we] >> 8) != p) lut[p].y = (u32)i;
  }
}

template <int WM>  // WM = ceil((k+1)/16) words
struct MercyOps {
  // compare the first nb bases of x and y (both left-aligned, arbitrary tail bits)
  __device__ static int cmp(const u32 (&x)[WM], const u32 *y, u32 nb) {
#pragma unroll
    for (int j = 0; j < WM; ++j) {
      const int keep = (int)(2 * nb) - 32 * j;
      if (keep <= 0) break;
      const u32 m = top_mask(keep > 32 ? 32u : (u32)keep);
      const u32 a = x[j] & m, b = y[j] & m;
      if (a != b) return a < b ? -1 : 1;
    }
    return 0;
  }
  // BinarySearchKmer (seq_to_sdbg.cpp:132-161).  Edges sharing a 12-base prefix share their leading byte, hence
  // their segment, so searching inside the owning segment probes exactly the edges the reference would probe.
  // Returns the matching edge record or nullptr.
  __device__ static const u32 *search(const EdgeSegs &sg, u32 we, const u32 (&km)[WM], u32 ksz) {
    const u32 seg = sg.owner[km[0] >> 24];
    const u32 *edges = sg.ptr[seg];
    const long long n = sg.n[seg];
    if (n == 0) return nullptr;
    const uint2 lr = sg.lut[seg][km[0] >> 8];  // 12 bases (kLookUpPrefixLength)
    if (lr.x == 0xFFFFFFFFu) return nullptr;   // lookup_table[...] == -1
    long long l = lr.x, r = lr.y;
    while (l <= r) {
      const long long mid = (l + r) / 2;
      const int c = cmp(km, edges + mid * we, ksz);
      if (c > 0) l = mid + 1;
      else if (c < 0) r = mid - 1;
      else return edges + mid * we;
    }
    return nullptr;
  }
  // x (nb bases) -> c followed by x's first nb bases (nb+1 bases)   [Kmer::ShiftPreappend, kmer.h:151-166]
  __device__ static void preappend(const u32 (&x)[WM], u32 c, u32 nb, u32 (&out)[WM]) {
#pragma unroll
    for (int j = WM - 1; j >= 0; --j) out[j] = (x[j] >> 2) | (j > 0 ? (x[j - 1] << 30) : (c << 30));
#pragma unroll
    for (int j = 0; j < WM; ++j) {
      const int keep = (int)(2 * (nb + 1)) - 32 * j;
      out[j] &= keep <= 0 ? 0u : top_mask(keep > 32 ? 32u : (u32)keep);
    }
  }
  __device__ static void set_base(u32 (&x)[WM], u32 idx, u32 c) {  // Kmer::SetBase kmer.h:187-193
    const u32 sh = 30 - 2 * (idx & 15);
#pragma unroll
    for (int j = 0; j < WM; ++j)
      if ((u32)j == (idx >> 4)) x[j] = (x[j] & ~(3u << sh)) | (c << sh);
  }
};

// base p of the PACKAGE-orientation (reversed) read whose file-orientation words are s
__device__ __forceinline__ u32 pkg_base(const u32 *s, u32 L, u32 p) { return base_at(s, L - 1 - p); }

// bits layout per candidate read: 3 planes (A, O, N) of `words_per_read` u32 each
template <int WM>
__global__ void __launch_bounds__(256)
    k_mercy_probe(ReadsView rv, const u64 *__restrict__ cand_ids, u64 n_cand, u32 k, const EdgeSegs sg, u32 we,
                  u32 *__restrict__ bits, u32 words_per_read) {
  using Ops = MercyOps<WM>;
  const u32 lane = lane_id();
  for (u64 c = (u64)blockIdx.x * 8 + (threadIdx.x >> 5); c < n_cand; c += (u64)gridDim.x * 8) {
    const u64 r = cand_ids[c];
    const u32 *rec0 = rv.bin + rv.rec_start(r);
    const u32 L = rec0[0];
    const u32 *s = rec0 + 1;
    const u32 nwords = div_ceil(L, 16);
    u32 *bA = bits + c * 3ull * words_per_read, *bO = bA + words_per_read, *bN = bO + words_per_read;
    const u32 npos = L >= k + 2 ? L - k + 1 : 0;  // positions i with i + k <= L; reads shorter than k+2 are skipped (:206)
    for (u32 i0 = 0; i0 < words_per_read * 32; i0 += 32) {
      const u32 i = i0 + lane;
      bool A = false, O = false, N = false;
      if (i < npos) {
        // km = pkg[i, i+k) ; rvk = its reverse complement.  pkg[i, i+k) = reverse(orig[L-i-k, L-i))
        u32 S[WM], km[WM], rvk[WM];
        load_sub<WM>(s, nwords, L - i - k, k, S);
        reverse_sub<WM>(S, k, km);
        complement_sub<WM>(S, k, rvk);  // rc(reverse(S)) = complement(S)
        // ---- has_in searches (:225-251) ----
        if (Ops::search(sg, we, rvk, k) != nullptr) {
          A = true;
        } else {
          u32 rv1[WM], km1[WM];
#pragma unroll
          for (int j = 0; j < WM; ++j) rv1[j] = rvk[j];
          Ops::set_base(rv1, k, 3);
          Ops::preappend(km, 0, k, km1);
          for (u32 ch = 0; ch < 4; ++ch) {
            Ops::set_base(km1, 0, ch);
            if (Ops::cmp(km1, rv1, k + 1) > 0) break;
            if (Ops::search(sg, we, km1, k + 1) != nullptr) {
              A = true;
              break;
            }
          }
        }
        // ---- has_out searches (:254-298) ----
        const u32 *e = Ops::search(sg, we, km, k);
        if (e != nullptr) {
          O = true;
          if (i + k < L && base_at(e, k) == pkg_base(s, L, i + k)) N = true;
        } else {
          u32 km1[WM], rv1[WM];
#pragma unroll
          for (int j = 0; j < WM; ++j) km1[j] = km[j];
          Ops::set_base(km1, k, 3);
          const u32 next_char = i + k < L ? 3u - pkg_base(s, L, i + k) : 0u;
          Ops::preappend(rvk, next_char, k, rv1);
          if (Ops::cmp(rv1, km1, k + 1) <= 0 && Ops::search(sg, we, rv1, k + 1) != nullptr) {
            O = true;
            N = true;
          } else {
            for (u32 ch = 0; ch < 4; ++ch) {
              if (ch == next_char) continue;
              Ops::set_base(rv1, 0, ch);
              if (Ops::cmp(rv1, km1, k + 1) > 0) break;
              if (Ops::search(sg, we, rv1, k + 1) != nullptr) {
                O = true;
                break;
              }
            }
          }
        }
      }
      const u32 mA = __ballot_sync(0xffffffffu, A), mO = __ballot_sync(0xffffffffu, O), mN = __ballot_sync(0xffffffffu, N);
      if (lane == 0) {
        bA[i0 >> 5] = mA;
        bO[i0 >> 5] = mO;
        bN[i0 >> 5] = mN;
      }
    }
  }
}

// seq_to_sdbg.cpp:310-352 per candidate read.  WRITE=false: count mercy edges; WRITE=true: emit them as
// `.edges`-format records (we words, multiplicity 1) at out + off[c] * we.
template <bool WRITE>
__global__ void k_mercy_emit(ReadsView rv, const u64 *__restrict__ cand_ids, u64 n_cand, u32 k,
                             const u32 *__restrict__ bits, u32 words_per_read, u32 *count, const u64 *off, u32 *out,
                             u32 we) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const u64 r = cand_ids[c];
  const u32 *rec0 = rv.bin + rv.rec_start(r);
  const u32 L = rec0[0];
  const u32 *s = rec0 + 1;
  const u32 *bA = bits + c * 3ull * words_per_read, *bO = bA + words_per_read, *bN = bO + words_per_read;
  const u32 npos = L >= k + 2 ? L - k + 1 : 0;
  int last_no_out = -1;
  u32 n_out = 0;
  u64 at = WRITE ? off[c] : 0;
  for (u32 i = 0; i < npos; ++i) {
    const bool has_in = ((bA[i >> 5] >> (i & 31)) & 1u) || (i > 0 && ((bN[(i - 1) >> 5] >> ((i - 1) & 31)) & 1u));
    const bool has_out = (bO[i >> 5] >> (i & 31)) & 1u;
    const int code = (has_in ? 1 : 0) | (has_out ? 2 : 0);
    if (code == 1) {
      last_no_out = (int)i;
    } else if (code == 2) {
      if (last_no_out >= 0) {
        for (u32 j = (u32)last_no_out; j < i; ++j) {
          if (WRITE) {
            u32 *e = out + (at + n_out) * we;
            for (u32 x = 0; x < we; ++x) e[x] = 0;
            for (u32 x = 0; x < k + 1; ++x) e[x >> 4] |= pkg_base(s, L, j + x) << (30 - 2 * (x & 15));
            e[we - 1] |= 1u;  // multiplicity 1 (seq_to_sdbg.cpp:354)
          }
          ++n_out;
        }
      }
      last_no_out = -1;
    } else if (code == 3) {
      last_no_out = -1;
    }
  }
  if (!WRITE) count[c] = n_out;
}

// Layout of the mercy scratch shared by the count and the write half (mhb_mercy_edges_count / _write)
struct MercyScratch {
  u64 *total;
  u32 *bits, *count;
  u64 *off, *bsum;
  u32 wpr;
};

}  // namespace mhb

// host helpers defined in mhb_device.cu, shared with mhb_multi.cu
mhb::ReadsView make_reads_view(const mhb_dev_reads *r);
int check_reads(const mhb_dev_reads *r, uint32_t k);
int scan32(cudaStream_t st, const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *total_dev, uint64_t *bsum);
size_t mercy_core_scratch(uint64_t n_cand, uint32_t max_read_len);
mhb::MercyScratch mercy_scratch_layout(void *scratch, uint64_t n_cand, uint32_t max_read_len);
