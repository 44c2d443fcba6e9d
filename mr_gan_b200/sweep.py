"""Fold scheduler: runs many independent fold-trainings (the `--tables` sweeps of
mr_gan.py:244-341 / mr_nn.py:129-168) grouped per GPU and sharded over the GPUs of one box.

Each (modality, labeled %, fold) call of ``mr_gan()`` is independent (fresh models per call,
mr_gan.py:109-171), so the sweep shards by fold with NO collective: under ``torchrun`` the folds are dealt
longest-first to the least loaded rank (``plan``), every rank trains its own in groups, and rank 0 gathers one float per
fold through ``torch.distributed`` (gloo object gather on the host) and prints in the reference's loop order."""
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


def chain_count(nf):
    """Parallel fold chains an epoch of a group of nf folds is captured as (mrgan_api.cu: mrgan_create, MRGAN_CHAINS)."""
    env = os.environ.get("MRGAN_CHAINS")
    try:
        nch = int(env) if env else (4 if nf >= 32 else (2 if nf >= 8 else 1))
    except ValueError:
        nch = 1                                  # the library reads it with atoi(): garbage -> 0 -> one chain
    return max(1, min(nch, 16, nf))


def order_for_chains(idxs, c):
    """Order the folds of one group so that the library's chains -- contiguous fold ranges [nf ch / nch, nf (ch + 1) / nch)
    that run concurrently and join at the end of the epoch -- carry equal cost: longest-first into the least loaded chain
    that still has room.  Folds of equal cost keep their order."""
    nf = len(idxs)
    nch = chain_count(nf)
    if nch <= 1 or len({c[i] for i in idxs}) <= 1:
        return list(idxs)
    room = [nf * (ch + 1) // nch - nf * ch // nch for ch in range(nch)]
    load = [0.0] * nch
    bins = [[] for _ in range(nch)]
    for i in sorted(idxs, key=lambda i: (-c[i], i)):
        ch = min((ch for ch in range(nch) if len(bins[ch]) < room[ch]), key=lambda ch: (load[ch], ch))
        bins[ch].append(i)
        load[ch] += c[i]
    return [i for b in bins for i in b]


def plan(jobs, world, group_size, key=lambda j: 0, cost=lambda j: 1.0):
    """rank -> list of groups (lists of job indices); identical on every rank.

    One GPU: the jobs in their own order, cut into groups of <= group_size that share key(job) -- a table's folds of one
    modality train side by side.  Several GPUs: the JOBS (not the groups) are dealt longest-first to the least loaded rank,
    and each rank then groups what it got (same key, similar cost next to each other).  Dealing whole groups left table 1
    with 7 one-modality groups of very different width (D = 400 ... 3632) for 8 GPUs: one GPU idle and the sweep as slow as
    its widest modality.  Results do not depend on the grouping (every fold has its own streams), only the wall clock does."""
    n = len(jobs)
    c = [float(cost(j)) for j in jobs]
    if world <= 1:
        return {0: [order_for_chains(g, c) for g in make_groups(jobs, group_size, key)]}
    load = [0.0] * world
    mine = [[] for _ in range(world)]
    for i in sorted(range(n), key=lambda i: (-c[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        mine[r].append(i)
        load[r] += c[i]
    out = {}
    for r in range(world):
        keys = {}
        for i in mine[r]:
            keys.setdefault(key(jobs[i]), []).append(i)
        groups = []
        for k in sorted(keys, key=lambda k: min(keys[k])):          # key classes in the order they first appear
            idxs = sorted(keys[k], key=lambda i: (-c[i], i))
            groups += [order_for_chains(idxs[a:a + group_size], c) for a in range(0, len(idxs), group_size)]
        out[r] = groups
    return out


def run_sharded(jobs, train_group, group_size=6, key=lambda j: 0, cost=lambda j: 1.0, init_dist=True):
    """Run train_group(list_of_jobs, device) -> list of results over all jobs; returns results in job order
    on every rank.  One process per GPU; no data-path collective."""
    rank, world, local = dist_env()
    mine = {}
    for idxs in plan(jobs, world, group_size, key, cost).get(rank, []):
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
