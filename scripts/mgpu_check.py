"""Multi-GPU parity check (run under torchrun on N GPUs): the SdBG stream / edges / counting produced by the
bucket-range partitioned build must be bit-identical to the reference fixtures."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megahit_b200 import formats as F  # noqa: E402
from megahit_b200 import lib, multigpu  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
lib.load().mhb_set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for name, k in (("syn150_k27", 27), ("syn150_klist", 21), ("syn150_klist", 59), ("syn150_klist", 141), ("tandem_k27", 28),
                ("polya_k27", 27)):
    case = os.path.join(ROOT, "tests", "golden", name)
    gold = json.load(open(os.path.join(case, "golden.json")))
    g, m = gold["by_k"][str(k)], gold["m"]
    allw = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32)
    L = int(allw[0])
    stride = 1 + (L + 15) // 16
    rows = allw.reshape(-1, stride)
    per = (len(rows) + world - 1) // world
    mine = rows[rank * per:(rank + 1) * per]
    bin_dev = torch.from_numpy(np.concatenate([mine.reshape(-1), np.zeros(8, np.uint32)]).view(np.int32)).to(dev)
    job = multigpu.MultiGpuBuild(len(mine), L, k, m, dev, need_mercy=True)
    res = job.run(bin_dev)
    info = [None] * world
    dist.all_gather_object(info, (res["n_solid"], res["n_cand"], res["n_mercy"], res["n_items_sorted"]))
    if rank == 0:
        print("   per-rank (n_solid, n_cand, n_mercy, items):", info)
    stream = multigpu.gather_sdbg_stream(res)
    torch.cuda.synchronize()
    edges = res["edges"][: res["n_solid"] * job.WE].cpu().numpy().view(np.uint32).tobytes()
    objs = [None] * world
    dist.all_gather_object(objs, edges)
    if rank == 0:
        e_ok = F.sha256(b"".join(objs)) == g["edges_sha256"] or g["n_solid"] == 0
        s_ok = F.sha256(stream) == g["sdbg_sha256"]
    dist.barrier()
    if rank == 0:
        cnt = res["mul_hist"].cpu().numpy()
        c_ok = F.sha256("".join(f"{i} {int(cnt[i])}\n" for i in range(1, 65536)).encode()) == g["counting_sha256"]
        print(f"{name} k={k}: edges {'OK' if e_ok else 'MISMATCH'} sdbg {'OK' if s_ok else 'MISMATCH'} "
              f"counting {'OK' if c_ok else 'MISMATCH'} bounds={res['bounds'].tolist()} bounds2={res['bounds2'].tolist()}")
        ok = ok and e_ok and s_ok and c_ok
        if not s_ok:
            ref = lib.build_host(allw, len(rows), k, m, need_mercy=True)
            rs = lib.sdbg_stream_from_table(ref["bucket_table"], ref["bytes"])
            print("  single-GPU fused build sha ok:", F.sha256(rs) == g["sdbg_sha256"], "n_mercy single", ref["n_mercy"],
                  "len", len(rs), len(stream))
            # first differing position
            n = min(len(rs), len(stream))
            a, b = np.frombuffer(rs[:n], np.uint8), np.frombuffer(stream[:n], np.uint8)
            d = np.nonzero(a != b)[0]
            print("  first diff at byte", int(d[0]) if len(d) else None, "of", n)
    job.close()
if rank == 0:
    print("MGPU PARITY", "PASS" if ok else "FAIL")
dist.destroy_process_group()
