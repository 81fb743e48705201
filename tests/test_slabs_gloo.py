"""CPU, world_size 2 (and 3) over gloo: the slab partition and the halo-exchange plumbing of the N>1 path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ballxworld_b200.slabs import agree_on_widths, exchange_halos, finish_exchange, rebalance_slabs, slab_partition, start_exchange


def test_slab_partition_covers_region():
    for total, world in [(128, 1), (128, 2), (512, 8), (10, 3), (7, 7)]:
        slabs = [slab_partition(total, world, r) for r in range(world)]
        assert slabs[0].x0 == 0 and sum(s.nx for s in slabs) == total
        for a, b in zip(slabs, slabs[1:]):
            assert a.x0 + a.nx == b.x0
        assert not slabs[0].has_lo and not slabs[-1].has_hi
    with pytest.raises(ValueError):
        slab_partition(3, 4, 0)


def _worker(rank, world, port, q, split=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slab = slab_partition(8 * world, world, rank)
    n = 64
    # a fake slab: blocks[x][k] with a shell column on each side; value encodes (global x, k)
    width = slab.nx * 32
    gx = np.arange(slab.x0 * 32 - 1, slab.x0 * 32 + width + 1)
    blocks = (gx[:, None] * 1000 + np.arange(n)[None, :]).astype(np.int64)
    blocks[0, :] = -1   # shells start unknown
    blocks[-1, :] = -1
    bufs = {k: torch.zeros(n, dtype=torch.int64) for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}

    def export_plane(face, t):
        t.copy_(torch.from_numpy(blocks[1 if face == 0 else width].copy()))

    def import_plane(face, t):
        blocks[0 if face == 0 else width + 1] = t.numpy()

    if split:  # the overlapped form: work between the two halves must not disturb the transfer
        works = start_exchange(slab, dist, bufs, export_plane)
        blocks[1:-1, :] += 0  # (the rest of the generation would run here)
        finish_exchange(slab, works, bufs, import_plane)
    else:
        exchange_halos(slab, dist, bufs, export_plane, import_plane)
    ok = True
    if slab.has_lo:
        ok &= bool((blocks[0] == (slab.x0 * 32 - 1) * 1000 + np.arange(n)).all())
    else:
        ok &= bool((blocks[0] == -1).all())
    if slab.has_hi:
        ok &= bool((blocks[-1] == (slab.x0 * 32 + width) * 1000 + np.arange(n)).all())
    else:
        ok &= bool((blocks[-1] == -1).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,split", [(2, False), (3, False), (2, True), (3, True)])
def test_halo_exchange_over_gloo(world, split):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, split)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


def test_rebalance_slabs_equalises_modelled_cost():
    # equal cost: nothing moves; the total never changes; a slab that costs 3x gives up columns
    assert rebalance_slabs([128] * 8, [1.0] * 8) == [128] * 8
    assert rebalance_slabs([128, 128], [1.0, 3.0]) == [171, 85]
    w = rebalance_slabs([128] * 4, [3.86, 3.59, 3.46, 4.13])
    assert sum(w) == 512 and w[3] < 128 < w[2]
    # under the model (cost spread evenly over a slab's columns) the new cut is balanced to within one column
    density = [3.86 / 128] * 128 + [3.59 / 128] * 128 + [3.46 / 128] * 128 + [4.13 / 128] * 128
    x, parts = 0, []
    for wk in w:
        parts.append(sum(density[x:x + wk]))
        x += wk
    assert max(parts) - min(parts) < 2 * max(density)
    # minimum width is respected and the order of the slabs is kept
    assert rebalance_slabs([16, 16], [1.0, 100.0]) == [24, 8]
    w = rebalance_slabs([64, 64, 64], [100.0, 1.0, 1.0], min_width=32)
    assert sum(w) == 192 and w[0] == 32 and min(w) >= 32
    for bad in (([128], [1.0, 2.0]), ([128, 0], [1.0, 1.0]), ([128, 128], [1.0, 0.0]), ([4, 4], [1.0, 1.0])):
        with pytest.raises(ValueError):
            rebalance_slabs(*bad)


def _rebalance_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    new_w, costs = agree_on_widths(dist, [128] * world, 1.0 + 2.0 * rank)
    q.put((rank, new_w, costs))
    dist.destroy_process_group()


def test_ranks_agree_on_widths_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] == [171, 85] and res[0][2] == res[1][2] == [1.0, 3.0]
