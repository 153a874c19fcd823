"""ctypes binding of oracle/libmhb_oracle.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing under
megahit_b200/ may import this module (tests/test_layout.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REF_BIN = os.path.join(_HERE, "_ref", "megahit_core_ref")


class _Seqs(C.Structure):
    _fields_ = [("words", C.c_void_p), ("word_off", C.c_void_p), ("len", C.c_void_p), ("n", C.c_uint64)]


class _CountOut(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_distinct", C.c_uint64), ("n_solid", C.c_uint64),
                ("words_per_edge", C.c_uint32), ("edges", C.POINTER(C.c_uint32)),
                ("first_0_out", C.POINTER(C.c_uint32)), ("last_0_in", C.POINTER(C.c_uint32)),
                ("counting", C.c_int64 * 65536)]


class _SdbgOut(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_items", C.c_uint64), ("words_per_tip_label", C.c_uint32),
                ("bucket_items", C.c_uint64 * 65536), ("bucket_tips", C.c_uint64 * 65536),
                ("bucket_large_mul", C.c_uint64 * 65536), ("bucket_byte_off", C.c_uint64 * 65537),
                ("w_count", C.c_uint64 * 9), ("ones_in_last", C.c_uint64), ("bytes", C.POINTER(C.c_uint8))]


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "libmhb_oracle.so")
    src = os.path.join(_HERE, "mhb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "libmhb_oracle.so"))
        _LIB.mhbo_count.restype = C.c_int
        _LIB.mhbo_seq2sdbg.restype = C.c_int
        _LIB.mhbo_gen_mercy.restype = C.c_int
        _LIB.mhbo_unpack_bin.restype = C.c_int
    return _LIB


class Seqs:
    """Word-aligned package-orientation sequences (numpy owned)."""

    def __init__(self, words, word_off, length):
        self.words = np.ascontiguousarray(words, np.uint32)
        if len(self.words) == 0:
            self.words = np.zeros(1, np.uint32)
        self.word_off = np.ascontiguousarray(word_off, np.uint64)
        self.len = np.ascontiguousarray(length, np.uint32)
        self.n = len(self.len)

    def c(self) -> _Seqs:
        lens = self.len if self.n else np.zeros(1, np.uint32)
        self._keep = lens
        return _Seqs(self.words.ctypes.data, self.word_off.ctypes.data, lens.ctypes.data, self.n)

    def base(self, s: int, i: int) -> int:
        w = self.words[int(self.word_off[s]) + (i >> 4)]
        return (int(w) >> (30 - 2 * (i & 15))) & 3

    def bin_bytes(self, ids) -> bytes:
        """`.bin`-format image (u32 len + words) of the given sequences, as SeqPackage::WriteSequences."""
        out = []
        for s in ids:
            a, b = int(self.word_off[s]), int(self.word_off[s + 1])
            out.append(np.array([self.len[s]], "<u4").tobytes() + self.words[a:b].tobytes())
        return b"".join(out)

    @staticmethod
    def concat(parts):
        words = np.concatenate([p.words[: int(p.word_off[-1])] for p in parts]) if parts else np.zeros(0, np.uint32)
        offs = [np.zeros(1, np.uint64)]
        base = 0
        for p in parts:
            offs.append(p.word_off[1:] + np.uint64(base))
            base += int(p.word_off[-1])
        lens = np.concatenate([p.len for p in parts]) if parts else np.zeros(0, np.uint32)
        return Seqs(words, np.concatenate(offs), lens)

    @staticmethod
    def from_fixed(words2d: np.ndarray, length: int):
        """(n, W) packed rows of equal-length sequences (e.g. edges)."""
        n, W = words2d.shape
        return Seqs(words2d.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(W),
                    np.full(n, length, np.uint32))


def unpack_bin(data: bytes, reverse: bool) -> Seqs:
    L = lib()
    buf = np.frombuffer(data, np.uint8)
    if len(buf) == 0:
        return Seqs(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    ns, nw = C.c_uint64(), C.c_uint64()
    rc = L.mhbo_unpack_bin(C.c_void_p(buf.ctypes.data), C.c_uint64(len(buf)), int(reverse), C.byref(ns),
                           C.byref(nw), None, None, None)
    assert rc == 0
    words = np.zeros(max(nw.value, 1), np.uint32)
    off = np.zeros(ns.value + 1, np.uint64)
    ln = np.zeros(max(ns.value, 1), np.uint32)
    rc = L.mhbo_unpack_bin(C.c_void_p(buf.ctypes.data), C.c_uint64(len(buf)), int(reverse), C.byref(ns),
                           C.byref(nw), C.c_void_p(words.ctypes.data), C.c_void_p(off.ctypes.data),
                           C.c_void_p(ln.ctypes.data))
    assert rc == 0
    return Seqs(words[: nw.value], off, ln[: ns.value])


def count(reads: Seqs, k: int, m: int):
    L = lib()
    out = _CountOut()
    s = reads.c()
    rc = L.mhbo_count(C.byref(s), C.c_uint32(k), C.c_int32(m), C.byref(out))
    assert rc == 0, rc
    wpe = out.words_per_edge
    res = {
        "n_records": out.n_records, "n_distinct": out.n_distinct, "n_solid": out.n_solid,
        "words_per_edge": wpe,
        "edges": np.ctypeslib.as_array(out.edges, (max(out.n_solid, 1) * wpe,))[: out.n_solid * wpe].reshape(-1, wpe).copy(),
        "first_0_out": np.ctypeslib.as_array(out.first_0_out, (max(reads.n, 1),))[: reads.n].copy(),
        "last_0_in": np.ctypeslib.as_array(out.last_0_in, (max(reads.n, 1),))[: reads.n].copy(),
        "counting": np.array(out.counting, dtype=np.int64),
    }
    L.mhbo_count_free(C.byref(out))
    return res


def cand_ids(first_0_out: np.ndarray, last_0_in: np.ndarray) -> np.ndarray:
    """kmer_counter.cpp:390-401."""
    S = np.uint32(0xFFFFFFFF)
    return np.nonzero((first_0_out != S) & (last_0_in != S) & (last_0_in > first_0_out))[0]


def counting_text(counting: np.ndarray) -> bytes:
    """edge_counter.h:44-52."""
    return "".join(f"{i} {int(counting[i])}\n" for i in range(1, 65536)).encode()


def seq2sdbg(seqs: Seqs, mult: np.ndarray, k: int):
    L = lib()
    out = _SdbgOut()
    s = seqs.c()
    mult = np.ascontiguousarray(mult, np.uint16)
    if len(mult) == 0:
        mult = np.zeros(1, np.uint16)
    rc = L.mhbo_seq2sdbg(C.byref(s), C.c_void_p(mult.ctypes.data), C.c_uint32(k), C.byref(out))
    assert rc == 0, rc
    nbytes = out.bucket_byte_off[65536]
    res = {
        "n_records": out.n_records, "n_items": out.n_items, "words_per_tip_label": out.words_per_tip_label,
        "bucket_items": np.array(out.bucket_items, np.uint64), "bucket_tips": np.array(out.bucket_tips, np.uint64),
        "bucket_large_mul": np.array(out.bucket_large_mul, np.uint64),
        "bucket_byte_off": np.array(out.bucket_byte_off, np.uint64),
        "w_count": np.array(out.w_count, np.uint64), "ones_in_last": out.ones_in_last,
        "bytes": bytes(np.ctypeslib.as_array(out.bytes, (max(nbytes, 1),))[:nbytes]),
    }
    L.mhbo_sdbg_free(C.byref(out))
    return res


def gen_mercy(edges: np.ndarray, cand: Seqs, k: int) -> np.ndarray:
    L = lib()
    edges = np.ascontiguousarray(edges, np.uint32)
    wpe = edges.shape[1] if edges.ndim == 2 else (2 * (k + 1) + 16 + 31) // 32
    n = edges.shape[0] if edges.ndim == 2 else 0
    ptr = C.POINTER(C.c_uint32)()
    nm = C.c_uint64()
    s = cand.c()
    e = edges if n else np.zeros((1, wpe), np.uint32)
    rc = L.mhbo_gen_mercy(C.c_void_p(e.ctypes.data), C.c_uint64(n), C.c_uint32(wpe), C.byref(s), C.c_uint32(k),
                          C.byref(ptr), C.byref(nm))
    assert rc == 0
    wm = (k + 1 + 15) // 16
    res = np.ctypeslib.as_array(ptr, (max(nm.value, 1) * wm,))[: nm.value * wm].reshape(-1, wm).copy()
    L.mhbo_free(ptr)
    return res


def edges_as_seqs(edges: np.ndarray, k: int):
    """Sorted `.edges` records -> ((k+1)-mer sequences, multiplicities) as SeqToSdbg::Initialize loads them
    (seq_to_sdbg.cpp:424-434, edge_reader.h:40-59)."""
    wpe = edges.shape[1]
    wm = (k + 1 + 15) // 16
    mult = (edges[:, wpe - 1] & 0xFFFF).astype(np.uint16)
    body = edges[:, :wm].copy()
    rem = (k + 1) % 16
    if rem:
        body[:, wm - 1] &= np.uint32((0xFFFFFFFF << (32 - 2 * rem)) & 0xFFFFFFFF)
    return Seqs.from_fixed(body, k + 1), mult


def read2sdbg(reads: Seqs, k: int, m: int, need_mercy: bool, want_solid: bool = False):
    """main_read2sdbg (main_sdbg_build.cpp:88-156) on the oracle: SdBG stream + what stage 1 writes to P.counting."""
    L = lib()
    L.mhbo_read2sdbg.restype = C.c_int
    out = _SdbgOut()
    s = reads.c()
    counting = np.zeros(65536, np.int64)
    nm, nb = C.c_uint64(), C.c_uint64()
    solid = C.POINTER(C.c_uint8)()
    rc = L.mhbo_read2sdbg(C.byref(s), C.c_uint32(k), C.c_int32(m), C.c_int(int(need_mercy)), C.byref(out),
                          C.c_void_p(counting.ctypes.data), C.byref(nm), C.byref(solid) if want_solid else None,
                          C.byref(nb))
    assert rc == 0, rc
    nbytes = out.bucket_byte_off[65536]
    res = {
        "n_records": out.n_records, "n_items": out.n_items, "words_per_tip_label": out.words_per_tip_label,
        "bucket_items": np.array(out.bucket_items, np.uint64), "bucket_tips": np.array(out.bucket_tips, np.uint64),
        "bucket_large_mul": np.array(out.bucket_large_mul, np.uint64),
        "bucket_byte_off": np.array(out.bucket_byte_off, np.uint64),
        "w_count": np.array(out.w_count, np.uint64), "ones_in_last": out.ones_in_last,
        "bytes": bytes(np.ctypeslib.as_array(out.bytes, (max(nbytes, 1),))[:nbytes]),
        "counting": counting, "n_mercy": nm.value, "n_bases": nb.value,
    }
    if want_solid:
        res["is_solid"] = np.ctypeslib.as_array(solid, (nb.value // 8 + 2,)).copy()
        L.mhbo_free(solid)
    L.mhbo_sdbg_free(C.byref(out))
    return res


def iterate(contigs: Seqs, reads: Seqs, k: int, step: int):
    """`megahit_core iterate` on the oracle: ascending unique `.edges` records (multiplicity 0) for k + step."""
    L = lib()
    L.mhbo_iterate.restype = C.c_int
    ptr = C.POINTER(C.c_uint32)()
    ne, na = C.c_uint64(), C.c_uint64()
    cs, rs = contigs.c(), reads.c()
    rc = L.mhbo_iterate(C.byref(cs), C.byref(rs), C.c_uint32(k), C.c_uint32(step), C.byref(ptr), C.byref(ne), C.byref(na))
    assert rc == 0, rc
    W = (2 * (k + step + 1) + 16 + 31) // 32
    res = np.ctypeslib.as_array(ptr, (max(ne.value, 1) * W,))[: ne.value * W].reshape(-1, W).copy()
    L.mhbo_free(ptr)
    return res, na.value
