"""Development aid (2 ranks): latency of the train step's collectives on this box."""
import os, time, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for n in (3, 290278):
    x = torch.ones(n, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.time()
    for _ in range(100):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"all_reduce {n} floats: {(time.time() - t) * 10:.3f} ms per call", flush=True)
dist.destroy_process_group()
