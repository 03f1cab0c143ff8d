import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev); t.zero_()
h = symm_mem.rendezvous(t, dist.group.WORLD)
if rank == 0:
    print("multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else h.multicast_ptr, "world", h.world_size, flush=True)
    print([a for a in dir(h) if not a.startswith("_")], flush=True)
    try:
        print("backend", symm_mem.get_backend(dev), flush=True)
    except Exception as ex:
        print("backend?", ex)
# built-in NVLS all-reduce op if present
try:
    t.fill_(rank + 1.0)
    dist.barrier(); torch.cuda.synchronize()
    import time
    for name in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
        op = getattr(torch.ops.symm_mem, name)
        for _ in range(5):
            r = op(t[:266142], "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            r = op(t[:266142], "sum", dist.group.WORLD.group_name)
        b.record(); torch.cuda.synchronize()
        if rank == 0:
            print(name, "%.1f us per call (266k floats, host-launched back to back)" % (a.elapsed_time(b) / 50 * 1e3), flush=True)
        dist.barrier()
except Exception as ex:
    if rank == 0: print("ops failed:", repr(ex)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
