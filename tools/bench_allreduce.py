"""NCCL all-reduce timing for the training step's gradient exchange (one process per GPU under torchrun):
95 MB fp32 vs 47.6 MB bf16, whole and in 2 / 4 buckets, CUDA events, max over ranks."""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
n = 23_800_000
for dtype in (torch.float32, torch.bfloat16):
    for parts in (1, 2, 4):
        bufs = [torch.ones(n // parts, dtype=dtype, device="cuda") for _ in range(parts)]
        for _ in range(5):
            for b in bufs:
                dist.all_reduce(b)
        torch.cuda.synchronize(); dist.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            for b in bufs:
                dist.all_reduce(b)
        e.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(e) / 20], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            mb = n * bufs[0].element_size() / 1e6
            print(f"all_reduce {mb:.1f} MB {dtype} in {parts} part(s): {t.item():.3f} ms  ({2 * (world - 1) / world * mb / t.item():.1f} GB/s bus)", flush=True)
dist.destroy_process_group()
