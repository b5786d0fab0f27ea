"""Dev tool (torchrun): time of NCCL broadcast / all-reduce for panel-sized FP64 messages, back to back on one stream."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (0.25, 1, 2, 4, 8, 16, 64):
    n = int(mb * (1 << 20) / 8)
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    res = {}
    for name in ("broadcast", "all_reduce"):
        for _ in range(5):
            (dist.broadcast(x, src=0) if name == "broadcast" else dist.all_reduce(x))
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(40):
            if name == "broadcast":
                dist.broadcast(x, src=k % world)
            else:
                dist.all_reduce(x)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 40 * 1e3
    if rank == 0:
        print(f"{mb:6.2f} MB: broadcast {res['broadcast']:.1f} us ({mb / 1024 / (res['broadcast'] * 1e-6):.0f} GB/s)   all_reduce {res['all_reduce']:.1f} us", flush=True)
dist.destroy_process_group()
