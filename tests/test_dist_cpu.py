"""world_size-2 gloo test (CPU) of the data-parallel host logic: contiguous batch sharding, means over
the GLOBAL batch, sum-all-reduce of the flat [grads | losses] buffer.  The compute stand-in is the
oracle (the CUDA path needs a GPU); what is checked is that the sharded + all-reduced result equals
the single-process full-batch result, which is the invariant the NCCL path relies on."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from oracle import clv_oracle as O


def _worker(rank, world, port, case, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import clvae_b200  # noqa: F401
    from clvae_b200.parallel import shard_range, init_from_env, allreduce_sum_
    torch.set_num_threads(1)
    w, r, _ = init_from_env("gloo")
    lo, hi = shard_range(case["B"], w, r)
    sub = dict(case, B=hi - lo, win=case["win"][lo:hi], labels=case["labels"][lo:hi],
               eps_w=case["eps_w"][lo:hi], eps_z=case["eps_z"][lo:hi])
    out, g = util.oracle_vrnn(sub)
    # local means -> contributions to the global means (what clv_cfg.B_global does in the kernels)
    frac = (hi - lo) / case["B"]
    names = sorted(g)
    flat = torch.cat([g[k].reshape(-1) * frac for k in names] +
                     [torch.stack([out[k] * frac for k in ("vae", "w_kl", "w_rec", "z_kl", "acc")])])
    allreduce_sum_(flat)
    if r == 0:
        out_q.put(flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_equals_full_batch():
    case = util.make_vrnn_case(21, B=6, L=3, C=4, Z=2)
    out, g = util.oracle_vrnn(case)
    names = sorted(g)
    ref = torch.cat([g[k].reshape(-1) for k in names] +
                    [torch.stack([out[k] for k in ("vae", "w_kl", "w_rec", "z_kl", "acc")])]).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
