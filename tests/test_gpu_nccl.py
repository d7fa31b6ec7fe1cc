"""GPU, world_size 2 over NCCL (skipped on a 1-GPU box): the slice-sharded driver with the real CIRIM forward on each
rank -- the gathered reconstruction must equal the single-GPU reconstruction of the same slices BIT FOR BIT (slices are
independent units and every kernel's arithmetic is independent of the batch a slice is part of)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mridc_b200 as mb
        from mridc_b200 import sharding, synth

        cfg = synth.cirim_cfg("GRU", num_cascades=2)
        batch = synth.make_batch(n_items, 6, 96, 64)
        torch.manual_seed(5)
        model = mb.CIRIM(cfg).cuda().eval()
        y, S, m, tgt = (batch[k].cuda() for k in ("y", "sensitivity_maps", "mask", "target"))

        def recon(yl, Sl, ml, tl):
            return next(model(yl, Sl, ml, None, tl))[-1][-1]

        full = recon(y, S, m, tgt)  # every rank also reconstructs all slices alone: the single-GPU answer
        out_all = sharding.run_sharded(recon, [y, S, m, tgt], n_items=n_items)
        pend = sharding.run_sharded(recon, [y, S, m, tgt], n_items=n_items, dst=0, async_op=True)
        out_root = pend.wait()
        torch.cuda.synchronize()
        ok = out_all.shape == full.shape and torch.equal(out_all, full)
        ok = ok and ((out_root is None) if rank != 0 else torch.equal(out_root, full))
        q.put((rank, bool(ok), float((torch.view_as_real(out_all) - torch.view_as_real(full)).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 3])
def test_nccl_gather_equals_single_gpu(n_items):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
