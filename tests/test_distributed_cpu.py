"""CPU, world_size 2 over gloo: the N>1 path's host logic -- shard ranges and the
all-reduce of per-rank partial outputs reproduce the whole-tensor MTTKRP."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import factor_mats, random_coo, rel_fro


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import restate
        from splatt_b200 import parallel
        dims, inds, vals = random_coo((40, 30, 50), 5000, seed=2)
        R = 8
        mats = factor_mats(dims, R)
        ok = True
        for mode in range(3):
            # the stream of `mode` is sorted with the mode at the root (MINUSONE order)
            perms, _ = restate.csf_policy(dims, 2)
            perm = perms[mode]
            order = np.lexsort([inds[m] for m in reversed(perm)])
            first, count = parallel.shard_range(len(vals), rank, world)
            sel = order[first:first + count]
            part = restate.mttkrp_coo(dims, [i[sel] for i in inds], vals[sel], mats, mode)
            t = torch.from_numpy(part)
            parallel.all_reduce_output(t)                      # gloo all-reduce(sum)
            gold = restate.mttkrp_coo(dims, inds, vals, mats, mode)
            ok = ok and rel_fro(t.numpy(), gold) < 1e-12
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_sharded_partials_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.parametrize("nnz", [0, 1, 63, 64, 65, 1000, 12345, 10_000_000])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_ranges_tile_the_stream(nnz, world):
    from splatt_b200 import parallel
    pos = 0
    sizes = []
    for r in range(world):
        first, count = parallel.shard_range(nnz, r, world)
        assert first == pos
        assert first % 64 == 0                      # shares start on a descriptor chunk
        pos += count
        sizes.append(count)
    assert pos == nnz
    assert max(sizes) - min(sizes) <= 64            # nnz-balanced to within one chunk
