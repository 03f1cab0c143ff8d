"""NVLink primitive costs between two ranks: torchrun --nproc-per-node 2 profiles/nvl_probe.py
(build first: see the header of nvl_probe.cu)."""
import ctypes as C
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nvl_probe.so"))
flags = symm_mem.empty(64, dtype=torch.int32, device=dev); flags.zero_()
hf = symm_mem.rendezvous(flags, dist.group.WORLD)
N = 1 << 22
buf = symm_mem.empty(N, dtype=torch.float32, device=dev); buf.zero_()
hb = symm_mem.rendezvous(buf, dist.group.WORLD)
peer = 1 - rank
out = torch.zeros(2, dtype=torch.int64, device=dev)
sink = torch.zeros(2, dtype=torch.int32, device=dev)
scratch = torch.zeros(4096, device=dev)
dst = torch.zeros(N, device=dev)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda x: C.c_void_p(int(x))
torch.cuda.synchronize(); dist.barrier()
iters, base = 2000, 0
for mode, name in ((0, "relaxed.sys store + volatile poll"), (1, "st.release.sys + volatile poll"),
                   (2, "4 KB local writes + __threadfence_system + store")):
    L.nvl_pingpong(P(hf.buffer_ptrs[rank]), P(hf.buffer_ptrs[peer]), P(scratch.data_ptr()), rank, iters, mode, base,
                   P(out.data_ptr()), st())
    torch.cuda.synchronize(); dist.barrier()
    base += iters
    if rank == 0:
        print("ping-pong %-52s %.2f us round trip" % (name, out[0].item() / iters / 1e3), flush=True)
for tgt, name in ((peer, "peer"), (rank, "local")):
    L.nvl_chase(P(hb.buffer_ptrs[tgt]), 2000, P(out.data_ptr()), P(sink.data_ptr()), st())
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("dependent loads from %-6s memory: %.2f us per load" % (name, out[0].item() / 2000 / 1e3), flush=True)
for n in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
    for tgt, name in ((peer, "peer"), (rank, "local")):
        for blocks, threads in ((148, 512), (296, 512), (592, 256)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                L.nvl_stream(P(hb.buffer_ptrs[tgt]), P(dst.data_ptr()), C.c_int64(n), blocks, threads, st())
            a.record()
            for _ in range(20):
                L.nvl_stream(P(hb.buffer_ptrs[tgt]), P(dst.data_ptr()), C.c_int64(n), blocks, threads, st())
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1e3
            if rank == 0:
                print("stream read %8d floats from %-5s (%3d x %3d): %6.1f us  %7.1f GB/s" %
                      (n, name, blocks, threads, us, n * 4 / us / 1e3), flush=True)
            dist.barrier()
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
