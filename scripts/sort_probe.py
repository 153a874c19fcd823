"""Micro-benchmark of k_radix_pass: how much of a pass is scatter/bank conflicts vs the fixed pipeline?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megahit_b200 import dev, lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 246_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randint(-2**31, 2**31 - 1, (n * 2 + 4,), generator=g, device="cuda", dtype=torch.int32)
a.view(-1)[1:2 * n:2] &= 0x00FFFFFF            # byte 3 of word 1 (record byte 3) constant zero
a.view(-1)[1:2 * n:2] &= ~0x00F00000            # byte 2 has 16 distinct values
b = torch.empty_like(a)
ws = torch.empty(lib.load().mhb_sort_workspace_bytes(n, 2), dtype=torch.uint8, device="cuda")
for name, bl in (("uniform byte x3", [0, 1, 5]), ("constant byte x3", [3, 3, 3]), ("16-valued byte x3", [2, 2, 2]),
                 ("uniform, already sorted by it x2", [6, 6, 6])):
    for rep in range(2):
        dev.sort_records(a, b, n, 2, bl, None, ws)
        torch.cuda.synchronize()
    ms, _, _ = lib.sort_pass_ms(0)
    print(f"{name:36s} per-pass ms {[round(x, 3) for x in ms]}  -> {2 * n * 8 / (min(ms) * 1e-3) / 1e9:.0f} GB/s")
# memcpy reference
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
b.copy_(a); torch.cuda.synchronize(); t0.record(); b.copy_(a); t1.record(); torch.cuda.synchronize()
print("torch copy of the same buffer: %.3f ms -> %.0f GB/s" % (t0.elapsed_time(t1), 2 * a.numel() * 4 / (t0.elapsed_time(t1) * 1e-3) / 1e9))
