"""One process per GPU: contiguous sharding of independent units (GPs, multistarts, replicas) and the single
collective of the path - the all-gather of per-unit (score, index) records (SURVEY.md par. 8e, K5).

torch.distributed is only the plumbing (rendezvous + NCCL communicator); on CPU the same code runs over gloo,
which is how the world_size-2 tests exercise it.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, **kw)
    return rank, local_rank, world


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_records(records, world=None, group=None):
    """records[count, 2] float64 (score, idx bit-cast to float64) of this rank -> [world*count, 2] on every rank.
    Equal counts per rank (pad with +inf scores otherwise)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    out = torch.empty((world * records.shape[0], 2), dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(out, records.contiguous(), group=group)
    return out


def unpack_records(records):
    """-> (score float64[count], idx int64[count])"""
    return records[:, 0].contiguous(), records[:, 1].contiguous().view(torch.int64)


def pack_records_cpu(score, idx):
    """Host-side twin of gprto_pack_best (used by the gloo tests)."""
    rec = torch.empty((score.numel(), 2), dtype=torch.float64)
    rec[:, 0] = score
    rec[:, 1] = idx.to(torch.int64).view(torch.float64)
    return rec


def global_argmin(records):
    """Arg-min over all gathered records with numpy.argmin semantics (NaN wins, first index on ties);
    computed redundantly on every rank - NCCL has no MINLOC.  Returns (score, global_idx, position)."""
    score, idx = unpack_records(records)
    nan = torch.isnan(score)
    if bool(nan.any()):
        pos = int(torch.nonzero(nan)[0])
    else:
        pos = int(torch.argmin(score))       # first minimum
    return float(score[pos]), int(idx[pos]), pos
