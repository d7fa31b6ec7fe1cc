"""CPU, world_size 2 over gloo: the slice-sharded driver (partition -> local compute -> one all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mridc_b200 import sharding

        g = torch.Generator().manual_seed(0)
        y = torch.randn(n_items, 3, 6, 5, 2, generator=g)
        S = torch.randn(n_items, 3, 6, 5, 2, generator=g)
        mask = torch.ones(1, 1, 1, 5, 1)
        calls = []

        def fake_recon(yl, Sl, ml):  # stands in for model.forward on the local block of slices
            calls.append(yl.shape[0])
            assert torch.equal(ml, mask)
            return torch.view_as_complex((yl * Sl).sum(1).contiguous())

        out = sharding.run_sharded(fake_recon, [y, S, mask], n_items=n_items)
        ref = torch.view_as_complex((y * S).sum(1).contiguous())
        a, b = sharding.partition(n_items, world)[rank]
        ok = out.shape == ref.shape and torch.equal(out, ref) and (calls[0] == max(b - a, 1))
        # gather to one rank only, asynchronously (what bench.py and the reference's test_epoch_end need)
        pend = sharding.run_sharded(fake_recon, [y, S, mask], n_items=n_items, dst=1, async_op=True)
        root = pend.wait()
        ok = ok and ((root is None) if rank != 1 else (root.shape == ref.shape and torch.equal(root, ref)))
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 4, 1])
def test_run_sharded_gloo_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == n_items for _, _, shape in res)


def test_single_process_passthrough():
    from mridc_b200 import sharding

    y = torch.randn(3, 2, 4, 4, 2)
    out = sharding.run_sharded(lambda t: t.sum(1), [y])
    assert torch.equal(out, y.sum(1))
