"""One-process-per-GPU launchers (torchrun): DDP training and collective-free sharded inference.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m smow_net_b200.runtime.launch train --model s --global-batch 128 --steps 50
    ... -m smow_net_b200.runtime.launch infer --model s --tiles 64 --tile-size 1024

The reference is single-GPU (train.py:2 pins CUDA_VISIBLE_DEVICES=0); this launcher is new functionality:
image pairs are sharded by rank, gradients are all-reduced by DDP over NCCL/NVLink (BatchNorm stays
per-replica, like the reference, so 8 x 16 pairs reproduces its batch-16 statistics), and inference
replicas never communicate.
"""
import argparse
import copy
import json
import os
import time

import torch
import torch.distributed as dist

from . import step as S
from . import synthetic


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend=None):
    rank, local_rank, world = dist_env()
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if use_cuda and backend in (None, "nccl"):
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    return rank, local_rank, world, device


def build_model(kind, device, pretrained=False):
    from ..models import SMOW_Net, SMOW_Net_LW
    if kind == "s":
        import torchvision
        weights = torchvision.models.ResNet18_Weights.DEFAULT if pretrained else None
        model = SMOW_Net(copy.deepcopy(torchvision.models.resnet18(weights=weights)))
    elif kind == "lw":
        model = S.freeze_unused(SMOW_Net_LW(pretrained=pretrained))
    else:
        raise ValueError("model must be 's' (SMOW_Net) or 'lw' (SMOW_Net_LW)")
    return model.to(device)


def align_pointwise_grad_strides(model):
    """1x1 (/1x1x1) convolution weights: both memory formats describe the same bytes, and cuDNN's NHWC weight-gradient
    kernels hand back the channels-last spelling (strides (Cin,1,Cin,Cin)) while the parameter carries the contiguous
    one (Cin,1,1,1).  DDP's reducer then warns "Grad strides do not match bucket view strides" and COPIES every such
    gradient into its bucket view.  A tensor hook re-spells the gradient's strides as the parameter's (a view of the same
    memory, no kernel), so gradient_as_bucket_view stays zero-copy."""
    for p in model.parameters():
        if p.requires_grad and p.dim() >= 4 and any(int(k) == 1 for k in p.shape):
            def fix(g, p=p):
                # strides may differ only where the size is 1 (depthwise (C,1,k,k,k) weights, 1x1 kernels): same bytes
                if g.stride() != p.stride() and g.shape == p.shape and all(
                        n == 1 or a == b for n, a, b in zip(g.shape, g.stride(), p.stride())):
                    return g.as_strided(g.shape, p.stride(), g.storage_offset())
                return g
            p.register_hook(fix)
    return model


def wrap_ddp(model, device, world):
    if world == 1:
        return model
    align_pointwise_grad_strides(model)
    from torch.nn.parallel import DistributedDataParallel as DDP
    ids = [device.index] if device.type == "cuda" else None
    # per-replica BatchNorm statistics (no SyncBN, no buffer broadcast) — the reference's regime
    # 5.5 M (LW) / 22 M (S-net) fp32 gradients: 8 MB buckets give 3 / 11 all-reduces that overlap the backward
    # (the default 25 MB would leave SMOW_Net_LW with a single, un-overlapped all-reduce at the very end)
    cap = float(os.environ.get("SMOW_DDP_BUCKET_MB", "8"))
    return DDP(model, device_ids=ids, broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=cap,
               static_graph=os.environ.get("SMOW_DDP_STATIC", "1") == "1")


def max_over_ranks(value, device, world):
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def train(args):
    rank, local_rank, world, device = init_distributed()
    synthetic.seed_everything(2022, rank)
    lo, hi = synthetic.shard_range(args.global_batch, rank, world)
    graphed = getattr(args, "graph", False) and device.type == "cuda"
    if graphed and world > 1:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):                      # DDP must be built off the default stream to be captured
            model = wrap_ddp(build_model(args.model, device), device, world).train()
        torch.cuda.current_stream().wait_stream(side)
    else:
        model = wrap_ddp(build_model(args.model, device), device, world).train()
    opt = S.make_optimizer(model, capturable=graphed)
    sched = S.make_scheduler(opt, args.steps + args.warmup)
    a, b, y = synthetic.make_batch(hi - lo, device=device, seed=2022 + rank)
    step = lambda: S.train_step(model, opt, sched, a, b, y)
    if graphed:       # forward + loss + backward (+ all-reduce) + clip + AdamW in ONE CUDA graph; the LR schedule steps outside
        from . import graph as G
        gs = G.GraphedStep(model, a, b, y, optimizer=opt, scheduler=sched, warmup=11 if world > 1 else 3)
        step = lambda: gs()
    for _ in range(args.warmup):
        step()
    if device.type == "cuda":
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    if device.type == "cuda":
        torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, device, world)
    if rank == 0:
        print(json.dumps({"mode": "train", "model": args.model, "n_gpus": world, "global_batch": args.global_batch,
                          "steps": args.steps, "ms_per_step": 1e3 * dt / max(1, args.steps),
                          "pairs_per_s": args.global_batch * args.steps / dt, "loss": float(loss.detach()),
                          "launch": "cuda-graph" if graphed else "eager"}), flush=True)
    if world > 1:
        if graphed:       # a captured graph keeps NCCL work alive; ProcessGroupNCCL's teardown would wait for ever
            dist.barrier()
            os._exit(0)
        dist.destroy_process_group()


@torch.no_grad()
def infer(args):
    """Config 4: `tiles` scene tiles of tile_size^2, contiguous shards per rank, 256^2 crops, no collectives
    in the data path (one MAX all-reduce of the elapsed time for reporting only)."""
    rank, local_rank, world, device = init_distributed()
    model = build_model(args.model, device).eval()
    lo, hi = synthetic.shard_range(args.tiles, rank, world)
    g = torch.Generator().manual_seed(2022 + rank)
    n_changed, t_total = 0, 0.0
    for i in range(args.warmup + 1):
        if device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_changed = 0
        for s in range(lo, hi, args.tiles_per_batch):
            n = min(args.tiles_per_batch, hi - s)
            ta = torch.randn(n, 3, args.tile_size, args.tile_size, generator=g).to(device, non_blocking=True)
            tb = torch.randn(n, 3, args.tile_size, args.tile_size, generator=g).to(device, non_blocking=True)
            prob = model(synthetic.tiles_to_crops(ta), synthetic.tiles_to_crops(tb))
            mask = synthetic.crops_to_tiles(prob > 0.5, n, args.tile_size, args.tile_size)   # test.py:134 threshold
            n_changed += int(mask.sum())
        if device.type == "cuda":
            torch.cuda.synchronize()
        t_total = time.perf_counter() - t0
    dt = max_over_ranks(t_total, device, world)
    crops = args.tiles * (args.tile_size // 256) ** 2
    if rank == 0:
        print(json.dumps({"mode": "infer", "model": args.model, "n_gpus": world, "tiles": args.tiles,
                          "tile_size": args.tile_size, "pairs_per_s": crops / dt, "s_total": dt,
                          "changed_px_rank0": n_changed}))
    if world > 1:
        dist.destroy_process_group()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="mode", required=True)
    t = sub.add_parser("train")
    t.add_argument("--model", default="s", choices=["s", "lw"])
    t.add_argument("--global-batch", type=int, default=128)
    t.add_argument("--steps", type=int, default=20)
    t.add_argument("--warmup", type=int, default=3)
    t.add_argument("--graph", action="store_true", help="replay the whole training step as one CUDA graph")
    t.set_defaults(fn=train)
    i = sub.add_parser("infer")
    i.add_argument("--model", default="s", choices=["s", "lw"])
    i.add_argument("--tiles", type=int, default=64)
    i.add_argument("--tile-size", type=int, default=1024)
    i.add_argument("--tiles-per-batch", type=int, default=2)
    i.add_argument("--warmup", type=int, default=1)
    i.set_defaults(fn=infer)
    args = ap.parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main()
