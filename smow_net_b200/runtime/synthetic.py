"""Synthetic LEVIR-CD-shaped batches: ImageNet-normalised-looking N(0,1) image pairs and
5 %-positive binary change masks; seeded per rank like the reference's seed_torch(2022) (train.py:42-48)."""
import torch


def seed_everything(seed=2022, rank=0):
    import random
    import numpy as np
    random.seed(seed + rank)
    np.random.seed(seed + rank)
    torch.manual_seed(seed + rank)
    torch.cuda.manual_seed_all(seed + rank)


def make_batch(batch, size=256, device="cpu", seed=2022, pin=False):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(batch, 3, size, size, generator=g)
    b = torch.randn(batch, 3, size, size, generator=g)
    y = (torch.rand(batch, size, size, generator=g) > 0.95).float()
    if pin:
        a, b, y = a.pin_memory(), b.pin_memory(), y.pin_memory()
    if str(device) != "cpu":
        a, b, y = a.to(device), b.to(device), y.to(device)
    return a, b, y


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` units for `rank`; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tiles_to_crops(tile, crop=256):
    """(B,3,H,W) scene tiles -> (B*n,3,crop,crop) non-overlapping crops, row-major (config 4: a 1024^2
    LEVIR-CD tile is 16 crops of 256^2, the only size the reference network accepts)."""
    b, c, h, w = tile.shape
    assert h % crop == 0 and w % crop == 0
    t = tile.view(b, c, h // crop, crop, w // crop, crop).permute(0, 2, 4, 1, 3, 5)
    return t.reshape(b * (h // crop) * (w // crop), c, crop, crop)


def crops_to_tiles(crops, tiles, h, w, crop=256):
    """Inverse of tiles_to_crops for (N,1,crop,crop) change maps -> (tiles,1,h,w)."""
    t = crops.view(tiles, h // crop, w // crop, crops.shape[1], crop, crop).permute(0, 3, 1, 4, 2, 5)
    return t.reshape(tiles, crops.shape[1], h, w)
