"""Host-side logic of the multi-GPU path on CPU: world_size 2 over gloo (SURVEY.md par. 8e).  The units (GPs) are
sharded in contiguous blocks, each rank computes its per-GP (score, index) records, one all-gather of 16-byte records
is the only collective, and the global arg-min is computed redundantly on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gp_rto_ei_b200 import parallel


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            edges = [parallel.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_records_roundtrip_and_argmin_semantics():
    score = torch.tensor([3.0, -1.0, -1.0, 2.0], dtype=torch.float64)
    idx = torch.tensor([10, 2 ** 40 + 3, 5, 7])
    rec = parallel.pack_records_cpu(score, idx)
    s2, i2 = parallel.unpack_records(rec)
    assert torch.equal(s2, score) and torch.equal(i2, idx)
    assert parallel.global_argmin(rec) == (-1.0, 2 ** 40 + 3, 1)          # first minimum wins
    rec[2, 0] = float('nan')
    v, i, pos = parallel.global_argmin(rec)
    assert np.isnan(v) and pos == 2                                       # numpy.argmin: NaN wins


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    parallel.init(backend='gloo')
    lo, hi = parallel.shard_range(total, rank, world)
    # per-GP best (score, candidate index) of this rank's shard: deterministic function of the GLOBAL GP id
    gid = torch.arange(lo, hi)
    score = torch.sin(gid.double() * 0.37) * 3.0
    cand = (gid * 7919) % 4096
    rec = parallel.pack_records_cpu(score, gid * 4096 + cand)            # global flat index = gp * m + candidate
    allrec = parallel.allgather_records(rec)
    torch.save((allrec, parallel.global_argmin(allrec)), os.path.join(out_dir, 'r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_allgather_and_global_argmin(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    total = 64                       # even split so that all_gather_into_tensor sees equal counts
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    r0, best0 = torch.load(os.path.join(tmp_path, 'r0.pt'))
    r1, best1 = torch.load(os.path.join(tmp_path, 'r1.pt'))
    assert torch.equal(r0, r1) and best0 == best1                        # every rank holds the full gathered set
    gid = torch.arange(total)
    score = torch.sin(gid.double() * 0.37) * 3.0
    s, i = parallel.unpack_records(r0)
    assert torch.equal(s, score) and torch.equal(i, gid * 4096 + (gid * 7919) % 4096)
    g = int(torch.argmin(score))
    assert best0[0] == float(score[g]) and best0[1] // 4096 == g
