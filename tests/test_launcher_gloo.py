"""CPU, world_size 2, gloo: the host-side multi-GPU logic (rank sharding, DDP wrapping with per-replica
BatchNorm, max-over-ranks timing, unused-parameter handling of SMOW_Net_LW).  The hot-path operators are
routed to the oracle restatement inside the worker processes (tests only) because the product has no CPU path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT
from smow_net_b200.runtime import synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_range_partitions_exactly():
    for total in (1, 7, 16, 128, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [synthetic.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_tile_crop_round_trip():
    t = torch.arange(2 * 3 * 512 * 768, dtype=torch.float32).view(2, 3, 512, 768)
    crops = synthetic.tiles_to_crops(t)
    assert crops.shape == (2 * 2 * 3, 3, 256, 256)
    assert torch.equal(crops[1], t[0, :, 0:256, 256:512]) and torch.equal(crops[3], t[0, :, 256:512, 0:256])
    back = synthetic.crops_to_tiles(crops[:, :1], 2, 512, 768)
    assert torch.equal(back, t[:, :1])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from oracle.cpu_model import reference_ops
    from smow_net_b200.runtime import launch, step as S
    r, lr, w, device = launch.init_distributed(backend="gloo")
    assert (r, w, device.type) == (rank, world, "cpu")
    res = {"max": launch.max_over_ranks(float(rank + 1), device, world)}
    with reference_ops():
        torch.manual_seed(0)                      # same init on both ranks, like DDP's broadcast would give
        model = launch.wrap_ddp(launch.build_model("lw", device), device, world).train()
        lo, hi = synthetic.shard_range(2, rank, world)
        a, b, y = synthetic.make_batch(2, seed=5)
        opt = S.make_optimizer(model)
        loss = S.train_step(model, opt, None, a[lo:hi], b[lo:hi], y[lo:hi])
    m = model.module
    res["loss"] = float(loss)
    res["grad"] = m.OFW.flow_make.weight.grad.clone()
    res["bn_mean"] = m.OFW.down[1].running_mean.clone()            # per-replica statistics: may differ
    res["frozen"] = all(p.grad is None for p in m.backbone.features[18].parameters())
    res["w"] = m.OFW.flow_make.weight.detach().clone()
    torch.save(res, os.path.join(out, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world))
    assert r0["max"] == r1["max"] == 2.0                             # MAX over ranks, not rank 0's own value
    assert r0["frozen"] and r1["frozen"]                             # unused features.18 excluded, DDP did not hang
    assert torch.equal(r0["grad"], r1["grad"])                       # gradients were all-reduced
    assert torch.equal(r0["w"], r1["w"])                             # replicas stay in lock-step after AdamW
    assert r0["loss"] != r1["loss"]                                  # each rank saw its own shard
    assert not torch.equal(r0["bn_mean"], r1["bn_mean"])             # BatchNorm statistics are per replica
