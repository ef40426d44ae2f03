"""World-size-2 `gloo` test of the sharded path's host logic (no GPU): shard the batch, per-shard statistics,
ONE all-reduce through the product's `allreduce_sums_`, closed forms on the global sums == full-batch value."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ecologysemanticsegmentation_b200 import distributed as D
        from ecologysemanticsegmentation_b200.synthetic import make_inputs
        from oracle import closed_form as cf, counts as oc
        z, g = make_inputs(6, 3, 16, 4242)
        p = torch.sigmoid(z)
        ps, gs = D.shard_batch(p, world, rank), D.shard_batch(g, world, rank)
        # per-channel leaves: sums are additive over shards
        sums = torch.tensor(np.stack([cf.leaf_sums(gs[:, c], ps[:, c]) for c in range(3)]))
        sums = D.allreduce_sums_(sums, D.WORLD)
        full = np.stack([cf.leaf_sums(g[:, c], p[:, c]) for c in range(3)])
        ok_sums = np.allclose(sums.numpy(), full, rtol=1e-12)
        losses = sum(cf.leaf_losses(sums[c].numpy(), 0.0, 2.0) for c in range(3))
        ref, _ = cf.plain_losses_and_grad(p, g)
        ok_loss = np.allclose(losses, ref, rtol=1e-12)
        # a mean of per-shard losses would be WRONG (Dice is non-linear in the sums)
        shard_losses = sum(cf.leaf_losses(cf.leaf_sums(gs[:, c], ps[:, c]), 0.0, 2.0) for c in range(3))
        t = torch.tensor(shard_losses)
        dist.all_reduce(t)
        differs = not np.allclose(t.numpy() / world, ref, rtol=1e-6)
        # scoring counts (int64) are additive too
        zs = D.shard_batch(z, world, rank)
        cnt = torch.tensor(oc.batch_counts(zs, gs, 0.8))
        cnt = D.allreduce_sums_(cnt, D.WORLD)
        ok_cnt = (cnt.numpy() == oc.batch_counts(z, g, 0.8)).all()
        # group=None never communicates
        same = D.allreduce_sums_(torch.ones(3), None)
        q.put((rank, bool(ok_sums), bool(ok_loss), bool(differs), bool(ok_cnt), bool((same == 1).all()), D.world_size(D.WORLD)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_sums_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1:] == (True, True, True, True, True, 2), r


def test_allreduce_requires_initialised_group():
    from ecologysemanticsegmentation_b200 import distributed as D
    if dist.is_initialized():
        pytest.skip("process group already initialised in this process")
    with pytest.raises(RuntimeError, match="not initialised"):
        D.allreduce_sums_(torch.ones(2), D.WORLD)
    assert D.world_size(None) == 1
