"""Multi-GPU decomposition of a region: contiguous slabs of chunk columns along x, one process per GPU.

Generation is a pure function of (seed, position) and needs no communication; meshing needs the 1-voxel plane just
outside each slab boundary. exchange_halos() moves those planes between x-neighbour ranks with
torch.distributed point-to-point ops (NCCL over NVLink on GPUs; gloo in the CPU tests). There is no collective on
the data path.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    rank: int
    world: int
    x0: int  # first interior chunk column of this rank (region-relative)
    nx: int  # interior chunk columns of this rank

    @property
    def has_lo(self):
        return self.rank > 0

    @property
    def has_hi(self):
        return self.rank < self.world - 1


def slab_partition(total_nx, world, rank):
    """Split total_nx chunk columns into `world` contiguous slabs, remainders to the lowest ranks."""
    if not (0 <= rank < world) or total_nx < world:
        raise ValueError("bad slab partition")
    base, rem = divmod(total_nx, world)
    nx = base + (1 if rank < rem else 0)
    x0 = rank * base + min(rank, rem)
    return Slab(rank, world, x0, nx)


def start_exchange(slab, dist, bufs, export_plane):
    """First half of exchange_halos: pack this rank's boundary planes and start the point-to-point transfers. Returns the
    pending work handles; whatever the caller launches next (the rest of the generation) overlaps the transfer."""
    ops = []
    if slab.has_lo:
        export_plane(0, bufs["send_lo"])
        ops.append(dist.P2POp(dist.isend, bufs["send_lo"], slab.rank - 1))
        ops.append(dist.P2POp(dist.irecv, bufs["recv_lo"], slab.rank - 1))
    if slab.has_hi:
        export_plane(1, bufs["send_hi"])
        ops.append(dist.P2POp(dist.isend, bufs["send_hi"], slab.rank + 1))
        ops.append(dist.P2POp(dist.irecv, bufs["recv_hi"], slab.rank + 1))
    return dist.batch_isend_irecv(ops) if ops else []


def finish_exchange(slab, works, bufs, import_plane):
    """Second half: wait for the transfers (the current stream waits, not the host) and write the received planes into the
    shell."""
    for w in works:
        w.wait()
    if slab.has_lo:
        import_plane(0, bufs["recv_lo"])
    if slab.has_hi:
        import_plane(1, bufs["recv_hi"])


def exchange_halos(slab, dist, bufs, export_plane, import_plane):
    """One halo exchange, start to finish (nothing overlapped).

    bufs: dict with tensors send_lo, send_hi, recv_lo, recv_hi (same shape/dtype/device).
    export_plane(face, tensor): pack this rank's own boundary plane at face 0 (x-) / 1 (x+) into tensor.
    import_plane(face, tensor): write a received plane into this rank's shell at that face.
    The x- plane of rank r becomes the x+ shell of rank r-1 and vice versa.
    """
    finish_exchange(slab, start_exchange(slab, dist, bufs, export_plane), bufs, import_plane)


def rebalance_slabs(widths, costs, min_width=8):
    """Slab widths for the next pass from the measured cost of the last one.

    widths[r] chunk columns cost costs[r] (device time of rank r's own kernels for one step; the slabs hold different terrain,
    so equal widths are not equal work). Model: a slab's cost is spread evenly over its columns. The new boundaries cut the
    cumulative cost into equal parts; the total width is unchanged and no slab gets narrower than min_width. Every rank calls
    this with the same inputs and gets the same answer.
    """
    world, total = len(widths), sum(widths)
    if world != len(costs) or any(w <= 0 for w in widths) or any(c <= 0 for c in costs) or total < world * min_width:
        raise ValueError("bad rebalance input")
    per_column = []
    for w, c in zip(widths, costs):
        per_column += [c / w] * w
    target = sum(costs) / world
    bounds, acc, k = [0], 0.0, 1
    for x, c in enumerate(per_column):
        # boundary k goes to whichever side of column x is closer to k * target
        while k < world and acc + c >= k * target:
            bounds.append(x + 1 if (acc + c - k * target) <= (k * target - acc) else x)
            k += 1
        acc += c
    while len(bounds) < world:
        bounds.append(total)
    bounds.append(total)
    for k in range(1, world):  # minimum width, left to right, then right to left
        bounds[k] = max(bounds[k], bounds[k - 1] + min_width)
    for k in range(world - 1, 0, -1):
        bounds[k] = min(bounds[k], bounds[k + 1] - min_width)
    return [bounds[k + 1] - bounds[k] for k in range(world)]


def agree_on_widths(dist, widths, my_cost, device="cpu", min_width=8):
    """All ranks exchange their measured cost (one all_gather of a scalar, outside the data path) and compute the same new
    widths. Returns (new_widths, costs)."""
    import torch

    mine = torch.tensor([float(my_cost)], dtype=torch.float64, device=device)
    allc = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(len(widths))]
    dist.all_gather(allc, mine)
    costs = [float(c.item()) for c in allc]
    return rebalance_slabs(list(widths), costs, min_width), costs
