"""CPU, world_size 2 (and 3) over gloo: the slab partition and the halo-exchange plumbing of the N>1 path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ballxworld_b200.slabs import exchange_halos, finish_exchange, slab_partition, start_exchange


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
