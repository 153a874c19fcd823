"""Multi-GPU SdBG build: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) for the exchange.

The path shards naturally: both stages are "extract locally -> all records with the same
leading bases must meet -> sort + scan locally -> own a contiguous range of the 65 536 prefix buckets".
The reference already treats prefix buckets as independent (sorting/base_engine.cpp:323-326) and both
on-disk formats map bucket -> (file, offset), so every rank can write its own `.edges.<r>` / `.sdbg.<r>`
with no merge.

Partition key = the record's most significant byte (its first 4 bases), cut into `world` CONTIGUOUS
ranges chosen from the all-reduced 256-bin histogram so that every rank receives about the same number of
records (canonical (k+1)-mers are skewed towards A-prefixes, equal-width ranges would not balance).  A
contiguous range of the top byte is a contiguous range of bucket ids, which is what `.edges.info` /
`.sdbg_info` can express - a whole-edge minimizer hash could not.

Per stage and rank:   extract (+ histogram of the top byte)  ->  one stable radix pass on the top byte
(groups records by destination)  ->  ONE variable-size all-to-all  ->  LSD radix sort of the received
records  ->  count / emit for the owned buckets.  Cross-rank state besides the two record exchanges:
the 256-bin histograms (all-reduce), the tip edges (all-gather, tiny) for the per-read mercy marks, and
the solid edges (all-gather) for the mercy-edge searches.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import lib
from .dev import _ptr, _stream, sort_records


def plan_ranges(hist: np.ndarray, world: int) -> np.ndarray:
    """Cut the 256 top-byte values into `world` contiguous, non-empty ranges with near-equal record counts.
    Returns bounds[world+1] (bounds[0] = 0, bounds[world] = 256); rank r owns [bounds[r], bounds[r+1])."""
    assert 1 <= world <= 256
    cum = np.concatenate([[0], np.cumsum(np.asarray(hist).astype(np.int64))])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        lo, hi = bounds[-1] + 1, 256 - (world - r)  # leave at least one value for every later rank
        target = total * r // world
        cand = np.arange(lo, hi + 1)
        bounds.append(int(cand[np.argmin(np.abs(cum[cand] - target))]))
    bounds.append(256)
    return np.array(bounds, np.int64)


def split_counts(hist: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    cum = np.concatenate([[0], np.cumsum(hist.astype(np.int64))])
    return (cum[bounds[1:]] - cum[bounds[:-1]]).astype(np.int64)


def exchange_records(grouped: torch.Tensor, words: int, send_counts: np.ndarray, group=None) -> torch.Tensor:
    """grouped: int32 tensor of records already grouped by destination rank (ascending); returns the
    records this rank owns (concatenated in source-rank order)."""
    world = dist.get_world_size(group)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=grouped.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.cpu().numpy()
    n_recv = int(recv_counts.sum())
    out = torch.empty(n_recv * words + 4, dtype=torch.int32, device=grouped.device)
    n_send = int(send_counts.sum())
    dist.all_to_all_single(out[: n_recv * words], grouped[: n_send * words],
                           output_split_sizes=[int(c) * words for c in recv_counts],
                           input_split_sizes=[int(c) * words for c in send_counts], group=group)
    assert world == len(send_counts)
    return out, n_recv


class _RawView:
    """int32 torch view over raw (cudaMalloc'ed) device memory, via the CUDA array interface"""

    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (int(nelem),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


class PeerBuffers:
    """One receive buffer per rank for a stage, cudaMalloc'ed by libmhb and opened on every rank through CUDA IPC,
    so that a rank's partition kernel can store records straight into their owner's memory over NVLink."""

    def __init__(self, nbytes: int):
        L = lib.load()
        self.L, self.nbytes = L, int(nbytes)
        p = C.c_void_p()
        lib._check(L.mhb_dev_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        h = (C.c_uint8 * 64)()
        lib._check(L.mhb_ipc_export(C.c_void_p(self.ptr), h))
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, bytes(h))
        self.peers = []
        for r, hb in enumerate(handles):
            if r == dist.get_rank():
                self.peers.append(self.ptr)
            else:
                q = C.c_void_p()
                buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                lib._check(L.mhb_ipc_open(buf, C.byref(q)))
                self.peers.append(q.value)
        dist.barrier()

    def close(self):
        dist.barrier()
        for r, q in enumerate(self.peers):
            if r != dist.get_rank():
                self.L.mhb_ipc_close(C.c_void_p(q))
        dist.barrier()
        self.L.mhb_dev_free(C.c_void_p(self.ptr))


class MultiGpuBuild:
    """count -> mercy -> seq2sdbg across the ranks of the default process group (fixed-length reads)."""

    def __init__(self, n_reads: int, read_len: int, k: int, m: int, device, need_mercy: bool = True):
        self.L = lib.load()
        self.n_reads, self.read_len, self.k, self.m, self.device = n_reads, read_len, k, m, device
        self.need_mercy = need_mercy
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.WR, self.WE, self.W2 = lib.count_record_words(k), lib.words_per_edge(k), lib.s2s_record_words(k)
        self.cbytes, self.sbytes = lib.count_sort_bytes(k), lib.s2s_sort_bytes(k)
        self.n_local = n_reads * (read_len - k) if read_len >= k + 1 else 0
        self.times = {}
        import os
        self.fused = self.world > 1 and not os.environ.get("MHB_MGPU_NCCL_A2A")
        self.peer = {}

    def _mark(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.times.setdefault(name, []).append(ev)

    def _buf(self, name, numel, dtype=torch.int32, slack=1.0):
        """grow-only named device buffer: steady-state steps allocate nothing"""
        if not hasattr(self, "_bufs"):
            self._bufs = {}
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(int(numel * slack) + 64, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def _partition_and_exchange(self, recs, n, words, top_byte, hist_dev, tag):
        """Group the records by owner (one stable radix pass on the top byte) and move them to their owners.
        Fused mode: the pass's scatter stores go straight into the owners' receive buffers over NVLink (CUDA IPC peer
        pointers) - ONE kernel does partition + exchange.  Fallback (MHB_MGPU_NCCL_A2A=1): local pass, then one
        variable-size NCCL all-to-all.  Returns (pointer to the owned records, count, bounds)."""
        L = self.L
        ghist = hist_dev.clone()
        dist.all_reduce(ghist)
        hist_h = hist_dev.cpu().numpy().astype(np.int64)
        bounds = plan_ranges(ghist.cpu().numpy(), self.world)
        send = split_counts(hist_h, bounds)
        if self._timed:
            self._mark(tag + "_plan")
        ws = self._buf(tag + "_ws", L.mhb_sort_workspace_bytes(max(n, 1), words), torch.uint8)
        if not self.fused:
            tmp = self._buf(tag + "_part", recs.numel())
            grouped = sort_records(recs, tmp, n, words, [top_byte], hist_dev, ws)
            if self._timed:
                self._mark(tag + "_partition")
            sc = torch.tensor(send, dtype=torch.int64, device=recs.device)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc)
            recv_counts = rc.cpu().numpy()
            n_recv = int(recv_counts.sum())
            out = self._buf(tag + "_own", n_recv * words + 4, slack=1.15)
            allr = [torch.empty_like(rc[:1]) for _ in range(self.world)]
            dist.all_gather(allr, torch.tensor([n_recv], dtype=torch.int64, device=recs.device))
            self._recv_tot = np.array([int(x.item()) for x in allr], np.int64)
            if self._timed:
                self._mark(tag + "_counts")
            dist.all_to_all_single(out[: n_recv * words], grouped[: int(send.sum()) * words],
                                   output_split_sizes=[int(c) * words for c in recv_counts],
                                   input_split_sizes=[int(c) * words for c in send])
            return out.data_ptr(), n_recv, bounds
        # ---- fused: who sends how much to whom, where my block starts inside every owner's buffer ----
        sc = torch.tensor(send, dtype=torch.int64, device=recs.device)
        allsend = [torch.empty_like(sc) for _ in range(self.world)]
        dist.all_gather(allsend, sc)
        M = torch.stack(allsend).cpu().numpy()  # M[r][o] = records rank r sends to owner o
        recv_tot = M.sum(axis=0)
        self._recv_tot = recv_tot
        need = int(recv_tot.max()) * words * 4 + 64
        pb = self.peer.get(tag)
        if pb is None or pb.nbytes < need:  # same decision on every rank: M is identical everywhere
            if pb is not None:
                pb.close()
            pb = self.peer[tag] = PeerBuffers(int(need * 1.2))
        rb = words * 4
        my_off = M[: self.rank].sum(axis=0)  # my block's first record inside owner o's buffer
        # digit of the partition pass = owning rank: one long contiguous run per destination and tile
        addr = np.zeros(256, np.uint64)
        lut = np.zeros(256, np.uint8)
        for o in range(self.world):
            lut[int(bounds[o]):int(bounds[o + 1])] = o
            addr[o] = np.uint64(pb.peers[o]) + np.uint64(my_off[o]) * np.uint64(rb)
        addr_dev = torch.from_numpy(addr.view(np.int64)).to(recs.device)
        lut_dev = torch.from_numpy(lut).to(recs.device)
        if self._timed:
            self._mark(tag + "_partition")
        dist.barrier()  # every owner is done with what it received in the previous step
        if self._timed:
            self._mark(tag + "_counts")
        lib._check(L.mhb_partition_scatter(_stream(), _ptr(recs), n, words, top_byte, _ptr(lut_dev), _ptr(addr_dev), _ptr(ws),
                                           ws.numel()))
        dist.barrier()  # all ranks' scatter kernels have completed: my buffer is complete
        return pb.ptr, int(recv_tot[self.rank]), bounds

    def close(self):
        for pb in self.peer.values():
            pb.close()
        self.peer = {}

    def _sort_raw(self, ptr_a, n, words, sort_bytes, tag):
        """LSD sort of n records at raw device pointer ptr_a; returns the pointer holding the result."""
        L = self.L
        tmp = self._buf(tag + "_tmp", n * words + 4, slack=1.15)
        ws = self._buf(tag + "_ws2", L.mhb_sort_workspace_bytes(max(n, 1), words), torch.uint8)
        arr = (C.c_uint8 * len(sort_bytes))(*sort_bytes)
        in_b = C.c_int(0)
        lib._check(L.mhb_sort_records(_stream(), C.c_void_p(ptr_a), _ptr(tmp), n, words, arr, len(sort_bytes), None, _ptr(ws),
                                      ws.numel(), C.byref(in_b)))
        return tmp.data_ptr() if in_b.value else ptr_a

    def run(self, bin_dev: torch.Tensor, timed: bool = False) -> dict:
        L, k, m, dev = self.L, self.k, self.m, self.device
        self._timed = timed
        i32 = dict(dtype=torch.int32, device=dev)
        reads = lib.DevReads(bin_dev.data_ptr(), bin_dev.numel(), self.n_reads, self.read_len, None, None)
        if timed:
            self._mark("t0")
        # ---- count stage ----
        n = self.n_local
        a = self._buf("c_a", n * self.WR + 4)
        hist = torch.zeros(256, dtype=torch.int64, device=dev)
        top = self.cbytes[-1]
        lib._check(L.mhb_count_extract(_stream(), C.byref(reads), k, _ptr(a), n, _ptr(hist), top))
        if timed:
            self._mark("extract")
        own, n_own, bounds = self._partition_and_exchange(a, n, self.WR, top, hist, "c")
        if timed:
            self._mark("exchange1")
        srt = self._sort_raw(own, n_own, self.WR, self.cbytes, "c")
        if timed:
            self._mark("sort1")
        cap = n_own // max(1, m) + 1
        # the solid edges live in a buffer every rank can read (CUDA IPC): the mercy searches of other ranks look
        # edges up in their owner's memory over NVLink instead of gathering 12 B x all edges onto every GPU
        lut_bytes = L.mhb_edge_lut_bytes()
        need_e = (int(self._recv_tot.max()) // max(1, m) + 2) * self.WE * 4 + 256 + lut_bytes  # identical on every rank
        pe = self.peer.get("edges")
        if pe is None or pe.nbytes < need_e:
            if pe is not None:
                pe.close()
            pe = self.peer["edges"] = PeerBuffers(int((need_e - lut_bytes) * 1.1) + lut_bytes)
        lut_off = (pe.nbytes - lut_bytes) & ~255  # the 12-mer look-up table sits at the end of the shared buffer
        edges = torch.as_tensor(_RawView(pe.ptr, lut_off // 4), device=dev)
        aux = self._buf("aux", cap, torch.uint8, slack=1.15)
        mul_hist = torch.zeros(65536, dtype=torch.int64, device=dev)
        nsol = torch.zeros(8, dtype=torch.int64, device=dev)
        scratch = self._buf("c_scratch", L.mhb_count_solid_scratch_bytes(n_own), torch.uint8, slack=1.15)
        lib._check(L.mhb_count_solid(_stream(), C.c_void_p(srt), n_own, k, m, _ptr(edges), _ptr(aux), cap, _ptr(mul_hist),
                                     _ptr(nsol), _ptr(scratch), scratch.numel()))
        n_solid = int(nsol[0].item())
        dist.all_reduce(mul_hist)  # edge_counter.h:44-52: `.counting` is a global histogram
        if timed:
            self._mark("count")

        # ---- mercy: tip edges from every rank -> per-read marks -> candidates -> mercy edges ----
        n_mercy = 0
        n_cand = 0
        if self.need_mercy:
            e2 = edges[: n_solid * self.WE].view(-1, self.WE)
            tipmask = aux[:n_solid] != 0
            tips, tipaux = e2[tipmask].contiguous(), aux[:n_solid][tipmask].contiguous()
            lib._check(L.mhb_edge_lut_build(_stream(), C.c_void_p(pe.ptr), n_solid, k, C.c_void_p(pe.ptr + lut_off)))
            cnt = torch.tensor([tips.shape[0], n_solid], dtype=torch.int64, device=dev)
            allcnt = [torch.zeros_like(cnt) for _ in range(self.world)]
            dist.all_gather(allcnt, cnt)  # also orders: every rank's solid edges are complete from here on
            tip_n = [int(c[0].item()) for c in allcnt]
            sol_n = [int(c[1].item()) for c in allcnt]
            mx = max(max(tip_n), 1)
            pad_t = torch.zeros((mx, self.WE), **i32)
            pad_a = torch.zeros(mx, dtype=torch.uint8, device=dev)
            pad_t[: tips.shape[0]] = tips
            pad_a[: tips.shape[0]] = tipaux
            gt = [torch.empty_like(pad_t) for _ in range(self.world)]
            ga = [torch.empty_like(pad_a) for _ in range(self.world)]
            dist.all_gather(gt, pad_t)
            dist.all_gather(ga, pad_a)
            all_t = torch.cat([g[:c] for g, c in zip(gt, tip_n)]).contiguous()
            all_a = torch.cat([g[:c] for g, c in zip(ga, tip_n)]).contiguous()
            n_tip = int(all_t.shape[0])
            need = L.mhb_tipset_bytes(n_tip, k)
            tipset = self._buf("tipset", need, torch.uint8, slack=1.2)
            lib._check(L.mhb_tipset_build(_stream(), _ptr(all_t) if n_tip else None, _ptr(all_a) if n_tip else None,
                                          n_tip, k, _ptr(tipset), need, n_tip))
            first = self._buf("first", self.n_reads + 1)
            last = self._buf("last", self.n_reads + 1)
            lib._check(L.mhb_count_mark_mercy(_stream(), C.byref(reads), k, _ptr(tipset), need, n_tip, _ptr(first),
                                              _ptr(last)))
            cand = self._buf("cand", self.n_reads + 1, torch.int64)
            cs = self._buf("cand_scratch", L.mhb_mercy_candidates_scratch_bytes(self.n_reads), torch.uint8)
            nc = C.c_uint64(0)
            lib._check(L.mhb_mercy_candidates(_stream(), _ptr(first), _ptr(last), self.n_reads, _ptr(cand), C.byref(nc),
                                              _ptr(cs), cs.numel()))
            n_cand = nc.value
            nm = C.c_uint64(0)
            if n_cand:
                seg_p = (C.c_void_p * self.world)(*[C.c_void_p(q) for q in pe.peers])
                seg_n = (C.c_uint64 * self.world)(*sol_n)
                seg_l = (C.c_void_p * self.world)(*[C.c_void_p(q + lut_off) for q in pe.peers])
                lut = np.zeros(256, np.uint8)
                for o in range(self.world):
                    lut[int(bounds[o]):int(bounds[o + 1])] = o
                lut_c = (C.c_uint8 * 256)(*lut.tolist())
                ms = self._buf("mercy_scratch", L.mhb_mercy_edges_scratch_bytes(n_cand, self.read_len) - lut_bytes,
                               torch.uint8, slack=1.2)
                # mercy edges are appended right behind this rank's solid edges (peers only read the solid part)
                lib._check(L.mhb_mercy_edges_segs(_stream(), C.byref(reads), _ptr(cand), n_cand, self.read_len, k, self.world,
                                                  seg_p, seg_n, seg_l, lut_c, C.c_void_p(pe.ptr + n_solid * self.WE * 4),
                                                  cap - n_solid, C.byref(nm), _ptr(ms), ms.numel()))
            n_mercy = nm.value
        seq_edges = edges
        if timed:
            self._mark("mercy")

        # ---- seq2sdbg stage ----
        n_seqs = n_solid + n_mercy
        n_items = n_seqs * 6
        seqs = lib.DevSeqs(seq_edges.data_ptr(), n_seqs * self.WE, n_seqs, k + 1, None, None, None, None, self.WE)
        sa = self._buf("s_a", n_items * self.W2 + 4, slack=1.1)
        hist2 = torch.zeros(256, dtype=torch.int64, device=dev)
        top2 = self.sbytes[-1]
        lib._check(L.mhb_s2s_extract(_stream(), C.byref(seqs), k, _ptr(sa), n_items, _ptr(hist2), top2))
        own2, n_own2, bounds2 = self._partition_and_exchange(sa, n_items, self.W2, top2, hist2, "s")
        if timed:
            self._mark("exchange2")
        srt2 = self._sort_raw(own2, n_own2, self.W2, self.sbytes, "s")
        wpt = (k + 15) // 16
        cap_b = n_own2 * (4 + 4 * wpt) + 16
        out_bytes = self._buf("sdbg", cap_b, torch.uint8, slack=1.1)
        table = torch.zeros(65536 * 4, dtype=torch.int64, device=dev)
        totals = torch.zeros(16, dtype=torch.int64, device=dev)
        es = self._buf("s_scratch", L.mhb_s2s_emit_scratch_bytes(n_own2, k), torch.uint8, slack=1.1)
        lib._check(L.mhb_s2s_emit(_stream(), C.c_void_p(srt2), n_own2, k, _ptr(out_bytes), cap_b, _ptr(table), _ptr(totals),
                                  _ptr(es), es.numel()))
        if timed:
            self._mark("s2s")
        return {"n_solid": n_solid, "n_cand": n_cand, "n_mercy": n_mercy, "edges": edges, "mul_hist": mul_hist,
                "bounds": bounds, "bounds2": bounds2, "n_items_sorted": n_own2, "sdbg_bytes": out_bytes, "table": table,
                "totals": totals, "n_records_owned": n_own}


def gather_sdbg_stream(res: dict) -> bytes | None:
    """Canonical SdBG stream of the whole job on rank 0 (validation only)."""
    tot = res["totals"].cpu().numpy()
    nbytes = int(tot[0])
    data = res["sdbg_bytes"][:nbytes].cpu().numpy().tobytes()
    table = res["table"].cpu().numpy().view(np.uint64).reshape(65536, 4)
    mine = lib.sdbg_stream_from_table(table, data)
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (res["bounds2"].tolist(), mine, int(tot[1])))
    if dist.get_rank() != 0:
        return None
    return b"".join(o[1] for o in objs)  # rank order == bucket order


def bench(args, bin_dev, bin_words, rank, world, device, metric, clocks=None):
    """bench.py's N > 1 arm: weak scaling, `args.reads` reads per GPU, one all-to-all per stage."""
    import json
    import os
    import sys

    k, m, n_reads, L = args.k, args.m, args.reads, 150
    job = MultiGpuBuild(n_reads, L, k, m, device, need_mercy=True)
    for _ in range(max(1, args.warmup)):
        job.run(bin_dev)
    torch.cuda.synchronize()
    dist.barrier()
    job.times.clear()
    if clocks is not None:
        clocks.start()  # nvidia-smi clocks / throttle reasons of this rank's GPU during the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = None
    pass_ms = []
    for _ in range(args.steps):
        res = job.run(bin_dev, timed=True)
        # the most recent traced sort over count-width records of this rank's owned count = this step's sort1 (the
        # fused partition pass is not traced; with the NCCL fallback the partitions are, so search instead of counting)
        for back in range(4):
            pm, nrec, words = lib.sort_pass_ms(back)
            if words == job.WR and nrec == res["n_records_owned"] and len(pm) == len(job.cbytes):
                pass_ms.append(pm)
                break
        else:
            pass_ms.append(lib.sort_pass_ms(1)[0])  # fused mode: [sort1, sort2] per step
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clk = clocks.stop() if clocks is not None else None
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    stage = {}
    names = ["t0", "extract", "c_plan", "c_partition", "c_counts", "exchange1", "sort1", "count", "mercy", "s_plan",
             "s_partition", "s_counts", "exchange2", "s2s"]
    for a, b in zip(names[:-1], names[1:]):
        t = torch.tensor([np.mean([x.elapsed_time(y) for x, y in zip(job.times[a], job.times[b])])], dtype=torch.float64,
                         device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage[b] = float(t.item())
    own = torch.tensor([res["n_records_owned"], res["n_solid"], res["n_mercy"]], dtype=torch.int64, device=device)
    owns = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(owns, own)
    # slowest rank's mean radix pass over the count records
    pm = torch.tensor([float(np.mean([np.mean(x) for x in pass_ms]))], dtype=torch.float64, device=device)
    dist.all_reduce(pm, op=dist.ReduceOp.MAX)

    # ---- e2e: pinned host reads -> device, build, SdBG bytes -> pinned host; device-timed, max over ranks ----
    host_bin = torch.empty(bin_words, dtype=torch.int32).pin_memory()
    host_bin.copy_(bin_dev[:bin_words])
    stage_dev = torch.empty_like(bin_dev)
    out_host = torch.empty(max(1 << 20, int(res["totals"][0].item()) * 2), dtype=torch.uint8).pin_memory()
    e2e_ms = []
    for i in range(1 + max(1, args.e2e_steps)):
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        stage_dev[:bin_words].copy_(host_bin, non_blocking=True)
        r2 = job.run(stage_dev)
        nb = int(r2["totals"][0].item())
        out_host[:nb].copy_(r2["sdbg_bytes"][:nb], non_blocking=True)
        tbl = r2["table"].cpu()
        a1.record()
        torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i > 0:
            e2e_ms.append(float(t.item()))
        del tbl
    nbytes = torch.tensor([nb], dtype=torch.int64, device=device)
    dist.all_reduce(nbytes)
    if rank == 0:
        from .lib import count_record_words
        peak = 6650.0
        src = "fallback (B200_PROFILING.md 6.65 TB/s)"
        pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak, src = float(json.load(open(pk))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        n_edges = world * n_reads * (L - k)
        ms_per_step = float(ms.item())
        S = count_record_words(k) * 4
        n_max = max(int(o[0]) for o in owns)
        ach = 2.0 * n_max * S / (float(pm.item()) * 1e-3) / 1e9
        e2e_v = n_edges / (float(np.mean(e2e_ms)) * 1e-3)
        print(json.dumps({
            "metric": metric, "value": n_edges / (ms_per_step * 1e-3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"synthetic {n_reads}x{L}bp reads PER GPU (30x, 1% subst.), k={k}, m={m}, {world}xB200: "
                                   "top-byte range partition (balanced from the all-reduced histogram), one fused partition+"
                                   "exchange pass per stage storing into the owners' buffers over NVLink (CUDA IPC peer "
                                   "memory), then per-GPU radix sort / count / mercy / seq2sdbg emit",
                       "parallelism": f"bucket-range x{world}", "n_edge_records": n_edges,
                       "records_owned_per_rank": [int(o[0]) for o in owns],
                       "solid_edges_per_rank": [int(o[1]) for o in owns], "mercy_edges_per_rank": [int(o[2]) for o in owns],
                       "l2_note": "inputs (>= 4.9 GB per kernel) exceed the 126 MB L2, no explicit flush needed"},
            "stage_ms_max_over_ranks": stage,
            "roofline": {"bound": "hbm", "kernel": f"k_radix_pass<{S // 4}> (count records, {S} B), slowest rank",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": src, "avg_launch_ms": float(pm.item()),
                         "algorithmic_bytes_per_launch": 2 * n_max * S},
            "cpu_baseline": None, "clocks": clk,
            "e2e": {"value": e2e_v, "unit": "edges/s", "h2d_bytes_per_step": int(world * bin_words * 4),
                    "d2h_bytes_per_step": int(nbytes.item()) + world * 65536 * 32, "ms_per_step": float(np.mean(e2e_ms)),
                    "api": "MultiGpuBuild.run on reads copied from pinned host memory each step; SdBG bytes + bucket "
                           "table copied back to pinned host memory"},
            "gpu_launches": 60,
        }))
    sys.stdout.flush()
    torch.cuda.synchronize()
    job.close()
    dist.barrier()
    dist.destroy_process_group()
