"""N>1 host logic on CPU with gloo, world size 2: the flat gradient bucket all-reduce
(DDP equivalence: mean of per-rank grads) and the no-communication pixel sharding of relighting."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rise_sdf_b200.relight import balanced_tile, my_pixels, my_tiles
    from rise_sdf_b200.train import FlatGradBucket
    torch.manual_seed(0)                                   # identical "weights" on every rank
    params = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.tensor(0.3))]
    bucket = FlatGradBucket(params)
    bucket.zero()
    x = torch.full((3,), float(rank + 1))                  # per-rank data
    loss = (params[0] @ x).sum() * (rank + 1) + (params[1] ** 2).sum() * (rank + 2) + params[2] * (rank + 5)
    loss.backward()
    assert params[0].grad.data_ptr() == bucket.flat.data_ptr()      # grads really live in the flat buffer
    local = bucket.flat.clone()
    bucket.all_reduce_mean()
    mine = torch.arange(640000)[my_pixels(640000, rank, world)]          # this rank's pixels of an 800x800 frame
    tiles = my_tiles(mine.numel(), balanced_tile(640000, world), 0, 1)
    # plain numpy payloads: torch tensors travel through a queue by file-descriptor passing, which needs the sender
    # alive when the receiver unpickles them
    q.put((rank, local.numpy().copy(), bucket.flat.numpy().copy(), (mine.numpy().copy(), tiles)))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_tile_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    import numpy as np
    mean = (res[0][1] + res[1][1]) / 2
    for r in range(world):
        assert np.allclose(res[r][2], mean, rtol=1e-6, atol=1e-7)             # == DDP's averaged gradient
    assert not np.allclose(res[0][1], res[1][1])
    pixels = np.concatenate([res[r][3][0] for r in range(world)])
    assert np.array_equal(np.sort(pixels), np.arange(640000))                 # disjoint, complete cover
    for r in range(world):
        mine, tiles = res[r][3]
        assert tiles[0][0] == 0 and tiles[-1][1] == mine.size
        assert all(a[1] == b[0] for a, b in zip(tiles[:-1], tiles[1:]))       # the shard is tiled without gaps
        rows = mine // 800
        assert rows.min() == 0 and rows.max() == 799                          # every rank sees the whole image
    assert len(res[0][3][1]) == len(res[1][3][1])                             # same tile count on every rank


def test_balanced_tile_counts():
    from rise_sdf_b200.relight import MAX_TILE, balanced_tile, my_pixels, my_tiles
    for world in (1, 2, 3, 4, 8):
        t = balanced_tile(640000, world)
        shard = [len(range(640000)[my_pixels(640000, r, world)]) for r in range(world)]
        counts = [len(my_tiles(n, t, 0, 1)) for n in shard]
        assert sum(shard) == 640000 and max(shard) - min(shard) <= 1
        assert t % 64 == 0 and t <= MAX_TILE and len(set(counts)) == 1 and counts[0] * t >= max(shard)
        assert (counts[0] - 1) * MAX_TILE < max(shard)                       # no more tiles than the cap requires
