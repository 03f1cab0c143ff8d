"""Run under torchrun with >= 2 GPUs: the fused peer-memory all-reduce + Adam-WN path must track the
NCCL all-reduce path (same data, same noise tapes) and keep all ranks' parameters bit-identical."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import util  # noqa: E402
from clvae_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = 24
    case = util.make_vrnn_case(5, B * world, 6, C=4, Z=2)
    sl = slice(rank * B, (rank + 1) * B)
    results = {}
    engines = []
    # nccl: step -> one all-reduce -> Adam-WN (three calls);  dp: clv_train_step_opt with the NCCL exchange
    # callback;  p2p: one-shot all-reduce kernels over peer memory (in-kernel flags) + the ordinary update;
    # p2p-fused: the exchange inside the Adam-WN kernels;  p2p-barrier: the round-1 form (host-side
    # symmetric-memory barriers around one kernel)
    # p2p-mc: the two-shot form through the NVSwitch multicast mapping (multimem.ld_reduce / multimem.st)
    modes = (("nccl", False), ("dp", False), ("dp", True), ("p2p", False), ("p2p", True), ("p2p-mc", False),
             ("p2p-mc", True), ("p2p-fused", False), ("p2p-fused", True), ("p2p-barrier", False))
    for mode, use_graph in modes:
        os.environ["CLV_P2P_MC"] = "1" if mode == "p2p-mc" else "0"
        e = Engine("vrnn", B, L=6, D=88, H=88, Z=2, n_classes=4, use_x_prev=True, world_size=world, rank=rank,
                   use_graph=use_graph,
                   p2p_allreduce=("fused" if mode == "p2p-fused" else mode.startswith("p2p")),
                   fused_optimizer=(mode in ("dp", "p2p", "p2p-fused")))
        engines.append(e)
        assert (e.symm is not None) == mode.startswith("p2p"), "symmetric memory set-up failed"
        if mode == "p2p-mc":
            assert e.p2p.mc_grads and e.p2p.mc_gsum, "no multicast mapping on this system"
        e.set_params({k: v.numpy() for k, v in case["p"].items()})
        e.stage_windows(torch.tensor(case["win"][sl]).cuda(), torch.tensor(case["labels"][sl]).cuda())
        e.eps_w.copy_(torch.tensor(case["eps_w"][sl], dtype=torch.float32).reshape(-1))
        e.eps_z.copy_(torch.tensor(case["eps_z"][sl], dtype=torch.float32).reshape(-1))
        losses = []
        for _ in range(4):
            e.run(train=True, gen_noise=False)
            losses.append(e.read_losses()["loss"])
        torch.cuda.synchronize()
        results[(mode, use_graph)] = (losses, e.params.clone())
        # replicas in lock-step: parameters bit-identical on every rank
        ref = e.params.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, e.params), "ranks diverged in mode %s" % mode
    base_l, base_p = results[("nccl", False)]
    for key in modes[1:]:
        l, p = results[key]
        assert np.allclose(l, base_l, rtol=2e-5), (key, l, base_l)
        err = float((p - base_p).abs().max() / base_p.abs().max())
        assert err < 2e-5, (key, err)
    # against the single-process oracle on the full batch (first step loss)
    if rank == 0:
        out, _ = util.oracle_vrnn(case)
        assert abs(base_l[0] - float(out["loss"])) < 1e-4 * float(out["loss"])
        print("P2P_CHECK_OK world=%d losses=%s" % (world, ["%.5f" % x for x in base_l]), flush=True)
    # CUDA graphs that captured NCCL kernels must be released before the communicator goes away
    torch.cuda.synchronize()
    dist.barrier()
    for e_ in engines:
        e_._graphs.clear()
    del engines, e
    torch.cuda.synchronize()
    import threading
    wd = threading.Timer(30.0, lambda: os._exit(0))      # fallback only: a stalled communicator teardown
    wd.daemon = True
    wd.start()
    dist.destroy_process_group()
    wd.cancel()


if __name__ == "__main__":
    main()
