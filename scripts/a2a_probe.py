"""NCCL all_to_all_single bandwidth probe (run under torchrun)."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_230_000_000  # records of 8 B
x = torch.empty(n * 2, dtype=torch.int32, device="cuda").random_()
y = torch.empty_like(x)
per = (n // world) * 2
splits = [per] * world
splits[-1] = n * 2 - per * (world - 1)
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_to_all_single(y, x, output_split_sizes=splits, input_split_sizes=splits)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if rank == 0 and it:
        sent = (n * 8) * (world - 1) / world
        print(f"world {world} env NCHANNELS={os.environ.get('NCCL_MIN_NCHANNELS')} a2a {ms:.2f} ms; off-GPU bytes/rank {sent/1e9:.2f} GB -> {sent/ms/1e6:.0f} GB/s per direction")
dist.destroy_process_group()
