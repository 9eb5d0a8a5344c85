"""How the global batch is split over ranks: the index plan of the reference's
MyDistributedSampler (stage1/data/sampler.py:62-138) as one pure function plus a torch Sampler around it, so that a
data-parallel run of the native path sees exactly the per-rank image partition the reference's DDP run sees
(BatchNorm statistics are per rank, so the partition is part of the numerics)."""
from __future__ import annotations

import torch
from torch.utils.data import Sampler


def shard_indices(n: int, world: int, rank: int, epoch: int = 0, seed: int = 0, shuffle: bool = True, drop_last: bool = False,
                  padding: bool = True, pair: bool = False) -> list[int]:
    """Indices rank `rank` of `world` visits in `epoch` over a dataset of `n` items.
    shuffle: randperm(n) from a CPU generator seeded with seed + epoch (identical on every rank).  padding: the index list is
    brought to a multiple of world (2 * world with `pair`) -- extended by wrapping around (drop_last=False) or truncated
    (drop_last=True).  Rank r then takes every world-th entry starting at r (entries are index PAIRS with `pair`)."""
    if not 0 <= rank < world:
        raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {world - 1}]")
    group = world * 2 if pair else world
    total = n
    if padding:
        parts, rest = divmod(n, group)
        total = parts * group if drop_last else (parts + bool(rest)) * group
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g)
    else:
        idx = torch.arange(n)
    if drop_last:
        idx = idx[:total]
    elif padding:
        extra = total - n
        idx = torch.cat([idx, idx[:extra]]) if extra <= n else idx.repeat((total + n - 1) // n)[:total]
    if pair:
        idx = idx.view(-1, 2)
    return idx[rank:total:world].flatten().tolist()


class ShardedSampler(Sampler):
    """torch Sampler with MyDistributedSampler's constructor arguments and set_epoch()."""

    def __init__(self, dataset, num_replicas: int = 1, rank: int = 0, shuffle: bool = True, seed: int = 0, drop_last: bool = False,
                 padding: bool = True, pair: bool = False):
        self.n, self.world, self.rank = len(dataset), num_replicas, rank
        self.kw = dict(seed=seed, shuffle=shuffle, drop_last=drop_last, padding=padding, pair=pair)
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def __iter__(self):
        return iter(shard_indices(self.n, self.world, self.rank, self.epoch, **self.kw))

    def __len__(self):
        return len(shard_indices(self.n, self.world, self.rank, self.epoch, **self.kw))
