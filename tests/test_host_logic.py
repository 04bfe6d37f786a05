"""Host-side logic that needs no GPU: MC-sample sharding and the world_size-2 gather (gloo, CPU tensors)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_shard_samples_partitions_exactly():
    from mauv.inference.predictors import shard_samples
    for S in (1, 5, 30, 31, 100):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_samples(S, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == S
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == S
    assert [shard_samples(30, 8, r) for r in range(8)][:2] == [(0, 4), (4, 8)]
    assert shard_samples(30, 4, 3) == (23, 30)      # 8+8+7+7


def _worker(rank, world, port, S, B, C, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from mauv.inference.predictors import gather_sample_blocks, shard_samples
    full = torch.arange(S * B * C, dtype=torch.float32).view(S, B, C)
    lo, hi = shard_samples(S, world, rank)
    got = gather_sample_blocks(full[lo:hi].clone() if hi > lo else None, S, B, C, world, "cpu")
    q.put((rank, bool(torch.equal(got, full))))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("S", [5, 1])
def test_gather_sample_blocks_world2_gloo(S):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 3, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
