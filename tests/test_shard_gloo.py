"""Host-side multi-GPU logic on CPU: contiguous sharding of the problem list and the single
end-of-run gather, exercised with torch.distributed (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest

from nirrt_star_b200.shard import gather_lists, pack_rows, shard_bounds, unpack_rows


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 8, 511, 4096):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                b, e = shard_bounds(n, world, r)
                assert 0 <= b <= e <= n
                covered += list(range(b, e))
            assert covered == list(range(n))
            sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip_keeps_inf():
    lists = [[np.inf, np.inf, 3.5], [], [1.0], [np.inf] * 7]
    rows, lens = pack_rows(lists)
    assert rows.shape == (4, 7)
    assert unpack_rows(rows, lens) == lists


def _fake_result(i):
    rs = np.random.RandomState(i)
    n = int(rs.randint(0, 40))
    l = list(np.sort(rs.uniform(50, 90, n))[::-1])
    k = int(rs.randint(0, n + 1))
    return [np.inf] * k + l[k:]


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_bounds(n_total, world, rank)
    got = gather_lists([_fake_result(i) for i in range(b, e)], n_total)
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 9), (2, 8), (3, 10)])
def test_gather_lists_over_gloo(world, n_total):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [_fake_result(i) for i in range(n_total)]
    for rank, got in results:
        assert got == want, rank
