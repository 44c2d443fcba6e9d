"""Fold scheduler: runs many independent fold-trainings (the `--tables` sweeps of
mr_gan.py:244-341 / mr_nn.py:129-168) grouped per GPU and sharded over the GPUs of one box.

Each (modality, labeled %, fold) call of ``mr_gan()`` is independent (fresh models per call,
mr_gan.py:109-171), so the sweep shards by fold with NO collective: under ``torchrun`` rank r
takes the job groups r, r+W, r+2W, ... and rank 0 gathers one float per fold through
``torch.distributed`` (gloo object gather on the host) and prints in the reference's loop order."""
import os

import numpy as np


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def _ensure_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo")      # host-side exchange only (seed, one float per fold): no GPU collective on this path
    return dist


def shared_seed(seed=None):
    """The sweep's seed, identical on every rank.  It drives the synthetic data, the StratifiedKFold random_state and
    every job's streams, so ranks that drew their own would each train folds of a DIFFERENT partition.  ``seed=None``
    keeps the reference's 'Non Deterministic output' (mr_gan.py:74-75): rank 0 draws, everybody else receives."""
    rank, world, _ = dist_env()
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 31))
    if world > 1:
        box = [int(seed)]
        _ensure_group().broadcast_object_list(box, src=0)
        seed = box[0]
    return int(seed)


def make_groups(jobs, group_size, key=lambda j: 0):
    """Split jobs (kept in order) into groups of <= group_size whose members share key(job)."""
    groups, cur, cur_key = [], [], None
    for i, j in enumerate(jobs):
        k = key(j)
        if cur and (k != cur_key or len(cur) >= group_size):
            groups.append(cur)
            cur = []
        cur.append(i)
        cur_key = k
    if cur:
        groups.append(cur)
    return groups


def assign(groups, world, cost=None):
    """Static longest-first assignment of groups to ranks (cost ~ sum of D)."""
    order = sorted(range(len(groups)), key=(lambda g: -cost[g]) if cost else (lambda g: g))
    load = [0.0] * world
    owner = [0] * len(groups)
    for g in order:
        r = int(np.argmin(load))
        owner[g] = r
        load[r] += cost[g] if cost else 1.0
    return owner


def run_sharded(jobs, train_group, group_size=6, key=lambda j: 0, cost=lambda j: 1.0, init_dist=True):
    """Run train_group(list_of_jobs, device) -> list of results over all jobs; returns results in job order
    on every rank.  One process per GPU; no data-path collective."""
    rank, world, local = dist_env()
    groups = make_groups(jobs, group_size, key)
    owner = assign(groups, world, [sum(cost(jobs[i]) for i in g) for g in groups])
    mine = {}
    for g, idxs in enumerate(groups):
        if owner[g] != rank:
            continue
        res = train_group([jobs[i] for i in idxs], local)
        for i, r in zip(idxs, res):
            mine[i] = r
    if world > 1:
        import torch.distributed as dist
        if init_dist:
            _ensure_group()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        mine = {}
        for p in parts:
            mine.update(p)
    return [mine[i] for i in range(len(jobs))]
