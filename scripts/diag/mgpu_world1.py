"""Diagnostic: the multi-GPU code path with world_size 1 (one GPU), to separate what the path itself costs from what
the second GPU adds.  torchrun --nproc-per-node 1 scripts/diag/mgpu_world1.py"""
import argparse, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from megahit_b200 import lib, multigpu, synth
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
lib.load().mhb_set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
args = argparse.Namespace(k=27, m=2, reads=10_000_000, steps=3, warmup=2, e2e_steps=1)
b2 = synth.synth_reads_torch(args.reads, 150, 5 * args.reads, 0.01, seed=1, device=dev)
bin_dev = torch.cat([b2.reshape(-1), torch.zeros(8, dtype=torch.int32, device=dev)])
multigpu.bench(args, bin_dev, args.reads * b2.shape[1], 0, dist.get_world_size(), dev, bench.METRIC, clocks=bench.ClockSampler(local))
