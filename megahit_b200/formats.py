"""On-disk formats on the SdBG-construction path (numpy readers/writers + canonical streams).

All citations are relative to /root/reference/src (voutcn/megahit v1.2.9).

* read library  ``L.bin`` / ``L.lib_info``  -- sequence_package.h:224-240, sequence_lib.cpp:84-110
* edges         ``P.edges.<i>`` / ``P.edges.info`` -- edge_io_meta.h:25-70, edge_writer.h:68-111
* SdBG          ``P.sdbg.<i>`` / ``P.sdbg_info``  -- sdbg_writer.cpp:25-79, sdbg_meta.cpp:12-61

"Canonical streams" (SURVEY.md 8c): the physical split of buckets over files depends on thread
scheduling in the reference, so equality is defined on the bucket-id ordered concatenation.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass, field

import numpy as np

NUM_BUCKETS = 65536
NULL_ID = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------
# read library
# ----------------------------------------------------------------------------------------------
def pack_reads_fixed(bases: np.ndarray) -> np.ndarray:
    """bases: (n, L) uint8 in {0..3}.  Returns (n, 1 + ceil(L/16)) uint32 rows = the ``.bin`` records
    (u32 length followed by the big-endian-in-word 2-bit packing, sequence_package.h:224-240)."""
    n, L = bases.shape
    W = (L + 15) // 16
    pad = np.zeros((n, W * 16), dtype=np.uint32)
    pad[:, :L] = bases
    pad = pad.reshape(n, W, 16)
    shifts = (30 - 2 * np.arange(16, dtype=np.uint32)).astype(np.uint32)
    words = np.bitwise_or.reduce(pad << shifts, axis=2).astype(np.uint32)
    out = np.empty((n, W + 1), dtype=np.uint32)
    out[:, 0] = L
    out[:, 1:] = words
    return out


def pack_read(bases: np.ndarray) -> np.ndarray:
    """One variable-length read -> u32 length + packed words."""
    return pack_reads_fixed(np.asarray(bases, dtype=np.uint8)[None, :])[0] if len(bases) else np.array([0], np.uint32)


def write_lib(prefix: str, bin_words: np.ndarray, n_reads: int, n_bases: int, max_len: int,
              desc: str = "synthetic") -> None:
    """Write ``prefix.bin`` + ``prefix.lib_info`` (one single-end library)."""
    np.ascontiguousarray(bin_words, dtype=np.uint32).tofile(prefix + ".bin")
    with open(prefix + ".lib_info", "w") as f:
        f.write(f"{n_bases} {n_reads}\n{desc}\n0 {n_reads} {max_len} 0\n")


def read_lib_info(prefix: str):
    with open(prefix + ".lib_info") as f:
        total_bases, n_reads = (int(x) for x in f.readline().split())
    return total_bases, n_reads


# ----------------------------------------------------------------------------------------------
# edges
# ----------------------------------------------------------------------------------------------
@dataclass
class EdgesInfo:
    kmer_size: int
    words_per_edge: int
    num_files: int
    num_buckets: int
    num_edges: int
    is_sorted: bool
    buckets: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int64))  # file, offset, count


def parse_edges_info(prefix: str) -> EdgesInfo:
    with open(prefix + ".edges.info") as f:
        toks = f.read().split()
    hdr = {toks[i]: int(toks[i + 1]) for i in range(0, 12, 2)}
    nb = hdr["num_buckets"]
    body = np.array(toks[12:12 + 4 * nb], dtype=np.int64).reshape(nb, 4)
    assert (body[:, 0] == np.arange(nb)).all()
    return EdgesInfo(hdr["kmer_size"], hdr["words_per_edge"], hdr["num_files"], nb, hdr["num_edges"],
                     bool(hdr["is_sorted"]), body[:, 1:].copy())


def canonical_edges(prefix: str) -> np.ndarray:
    """Bucket-id ordered (n_edges, words_per_edge) uint32 array (edge_reader.h:105-138 order)."""
    info = parse_edges_info(prefix)
    W = info.words_per_edge
    files = [np.fromfile(f"{prefix}.edges.{i}", dtype=np.uint32) for i in range(info.num_files)]
    if not info.is_sorted:
        return files[0][: info.num_edges * W].reshape(-1, W)
    parts = []
    for fid, off, cnt in info.buckets:
        if fid >= 0 and cnt > 0:
            parts.append(files[fid][off * W:(off + cnt) * W])
    if not parts:
        return np.zeros((0, W), np.uint32)
    return np.concatenate(parts).reshape(-1, W)


def write_edges(prefix: str, k: int, edges: np.ndarray, num_files: int = 1) -> None:
    """Write sorted edges (n, W) as ``prefix.edges.0`` (+ empty files) and ``prefix.edges.info``."""
    W = (2 * (k + 1) + 16 + 31) // 32
    edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, W)
    edges.tofile(f"{prefix}.edges.0")
    for i in range(1, num_files):
        open(f"{prefix}.edges.{i}", "wb").close()
    b = (edges[:, 0] >> 16).astype(np.int64)
    cnt = np.bincount(b, minlength=NUM_BUCKETS)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    with open(prefix + ".edges.info", "w") as f:
        f.write(f"kmer_size {k}\nwords_per_edge {W}\nnum_files {num_files}\nnum_buckets {NUM_BUCKETS}\n"
                f"num_edges {len(edges)}\nis_sorted 1\n")
        for i in range(NUM_BUCKETS):
            if cnt[i]:
                f.write(f"{i} 0 {off[i]} {cnt[i]}\n")
            else:
                f.write(f"{i} -1 0 0\n")


# ----------------------------------------------------------------------------------------------
# SdBG
# ----------------------------------------------------------------------------------------------
@dataclass
class SdbgInfo:
    k: int
    words_per_tip_label: int
    num_buckets: int
    num_files: int
    records: np.ndarray  # (num_buckets, 6) uint64: bucket, file, start_byte, items, tips, large_mul


def parse_sdbg_info(prefix: str) -> SdbgInfo:
    with open(prefix + ".sdbg_info") as f:
        toks = f.read().split()
    hdr = {toks[i]: int(toks[i + 1]) for i in range(0, 8, 2)}
    nb = hdr["num_buckets"]
    body = np.array([int(t) for t in toks[8:8 + 6 * nb]], dtype=np.uint64).reshape(nb, 6)
    return SdbgInfo(hdr["k"], hdr["words_per_tip_label"], nb, hdr["num_files"], body)


def _sdbg_bucket_nbytes(buf: np.ndarray, start: int, n_items: int, wpt: int) -> int:
    """Walk n_items variable-length items (2 B, +2 B if byte1 == 255, +4*wpt B if tip bit)."""
    pos = start
    for _ in range(n_items):
        b0, b1 = int(buf[pos]), int(buf[pos + 1])
        pos += 2
        if b1 == 255:
            pos += 2
        if b0 & 0x20:
            pos += 4 * wpt
    return pos - start


def canonical_sdbg(prefix: str):
    """Returns (info, stream_bytes, per-bucket table).  stream = for bucket id ascending with items > 0:
    u32 bucket_id, u64 num_items, raw item bytes."""
    info = parse_sdbg_info(prefix)
    files = [np.fromfile(f"{prefix}.sdbg.{i}", dtype=np.uint8) for i in range(info.num_files)]
    recs = info.records[info.records[:, 0] != np.uint64(NULL_ID)]
    recs = recs[np.argsort(recs[:, 0], kind="stable")]
    chunks = []
    table = np.zeros((NUM_BUCKETS, 3), np.uint64)
    for bucket, fid, start, items, tips, large in recs:
        bucket, fid, start, items, tips, large = (int(x) for x in (bucket, fid, start, items, tips, large))
        nbytes = 2 * items + 2 * large + 4 * info.words_per_tip_label * tips
        raw = files[fid][start:start + nbytes]
        assert len(raw) == nbytes, (bucket, fid, start, nbytes, len(raw))
        table[bucket] = (items, tips, large)
        if items:
            chunks.append(np.array([bucket], "<u4").tobytes() + np.array([items], "<u8").tobytes() + raw.tobytes())
    return info, b"".join(chunks), table


def canonical_sdbg_from_arrays(bucket_items, bucket_byte_off, data: bytes) -> bytes:
    chunks = []
    for b in np.nonzero(np.asarray(bucket_items))[0]:
        chunks.append(np.array([b], "<u4").tobytes() + np.array([bucket_items[b]], "<u8").tobytes()
                      + bytes(data[int(bucket_byte_off[b]):int(bucket_byte_off[b + 1])]))
    return b"".join(chunks)


def sha256(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def file_sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def exists_all(prefix: str, suffixes) -> bool:
    return all(os.path.exists(prefix + s) for s in suffixes)
