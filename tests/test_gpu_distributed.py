"""Multi-GPU sharded path (needs >= 2 CUDA devices; skipped otherwise): one process per GPU over NCCL, batch
sharded, only the statistics all-reduced.  Result must equal the single-device full-batch run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ecologysemanticsegmentation_b200 as eco
        from ecologysemanticsegmentation_b200 import distributed as D, fused, test_multiclass as tmc
        from ecologysemanticsegmentation_b200.synthetic import make_inputs
        up = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
        z, g = make_inputs(8 * world, 3, 64, 4242)
        zf, gf = z.cuda(), g.cuda()
        # single-device full batch
        zr = zf.clone().requires_grad_(True)
        np.random.seed(0)
        ref = eco.losses_fn(zr, gf, True, from_logits=True)
        sum(w * l for w, l in zip(up, ref) if w).backward()
        # sharded autograd path
        zs = D.shard_batch(zf, world, rank).clone().requires_grad_(True)
        gs = D.shard_batch(gf, world, rank)
        np.random.seed(0)
        ours = eco.losses_fn(zs, gs, True, from_logits=True, group=D.WORLD)
        sum(w * l for w, l in zip(up, ours) if w).backward()
        lo, hi = D.shard_bounds(zf.shape[0], world, rank)
        e_loss = max(abs(float(a) - float(b)) / max(abs(float(b)), 1e-30) for a, b in zip(ours[1:], ref[1:]))
        e_grad = float((zs.grad - zr.grad[lo:hi]).abs().max() / zr.grad.abs().max())
        # sharded step object (stats -> all-reduce -> closed forms -> grad)
        np.random.seed(0)
        step = fused.ShardedCompositeLossStep(up, group=D.WORLD)
        l2, dz = step(zs.detach(), gs)
        e_step = float((dz - zr.grad[lo:hi]).abs().max() / zr.grad.abs().max())
        # ONE-launch step with the all-reduce done in-kernel over NVLink peer memory (several calls: epochs, parity)
        np.random.seed(0)
        pstep = fused.PeerShardedCompositeLossStep(up, group=D.WORLD)
        e_peer = 0.0
        for _ in range(5):
            l3, dz3 = pstep(zs.detach(), gs)
            e_peer = max(e_peer, float((dz3 - zr.grad[lo:hi]).abs().max() / zr.grad.abs().max()),
                         max(abs(float(a) - float(b)) / max(abs(float(b)), 1e-30) for a, b in zip(l3[1:], ref[1:])))
        pstep.close()
        e_step = max(e_step, e_peer)
        # plain per-channel path and scoring
        zr2 = zf.clone().requires_grad_(True)
        r2 = eco.losses_fn(torch.sigmoid(zr2), gf)
        o2 = eco.losses_fn(torch.sigmoid(zs.detach()), gs, group=D.WORLD)
        e_plain = max(abs(float(a) - float(b)) / max(abs(float(b)), 1e-30) for a, b in zip(o2[1:], r2[1:]))
        d_full, c_full, _ = tmc.score_batch(zf, gf, 0.8, return_counts=True)
        d_sh, c_sh, _ = tmc.score_batch(zs.detach(), gs, 0.8, group=D.WORLD, return_counts=True)
        # frame stream (cfg5): sharded batches, ONE all-reduce of the whole counts buffer at the end
        sc, full = tmc.StreamScorer(3, 4, 0.8, group=D.WORLD), tmc.StreamScorer(3, 4, 0.8)
        for f in (1.0, 0.5, 2.0):
            sc.add(zs.detach() * f, gs)
            full.add(zf * f, gf)
        stream_eq = bool(torch.equal(sc.per_batch(), full.per_batch())) and bool(torch.equal(sc.result(), full.result()))
        q.put((rank, e_loss, e_grad, e_step, e_plain, bool(torch.equal(c_full, c_sh)), bool(torch.equal(d_full, d_sh)) and stream_eq))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_equals_single_device():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, e_loss, e_grad, e_step, e_plain, counts_eq, dice_eq in res:
        assert e_loss < 1e-6 and e_grad < 1e-6 and e_step < 1e-6 and e_plain < 1e-6, (rank, e_loss, e_grad, e_step, e_plain)
        assert counts_eq and dice_eq
