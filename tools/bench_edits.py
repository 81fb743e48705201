#!/usr/bin/env python
"""Streaming edit benchmark (BASELINE.json configs[4], SURVEY.md section 8d cfg 5): a load anchor sweeps +x one chunk per
step; every step applies N random block edits inside the anchor sphere (radius 10, src/config.rs:35-36) through
World::apply_voxel_changes semantics and re-meshes the dirty chunks. Reports dirty chunks re-meshed per second and the
edit->mesh latency (host wall clock around apply_changes + remesh_dirty, device synchronised).

Under torchrun (one rank per GPU) every rank owns an x-slab of --dims chunks with its own anchor; all ranks build the same
global change list (every rank's edits) and apply what falls inside their own storage box, shell included, so edits next to
a slab boundary dirty the neighbour's chunks without any exchange. The reported rate is the sum over ranks divided by the
slowest rank's time."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="96,16,32")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--edits", type=int, default=4096)
    ap.add_argument("--radius", type=int, default=10)
    a = ap.parse_args()
    nx, ny, nz = (int(x) for x in a.dims.split(","))
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bxw.Context(local)
    reg = bxw.standard_registry()
    region = bxw.Region((rank * nx, 0, 0), (nx, ny, nz), reg, ctx=ctx)
    region.generate(0)
    region.mesh()
    lat, dirty = [], 0
    for step in range(a.steps):
        parts = []
        for r in range(world):  # the same list on every rank
            rng = np.random.default_rng(1000 * step + r + 1)
            anchor = np.array([r * nx + a.radius + step, ny // 4, nz // 2])
            # random positions inside the anchor sphere, random standard blocks with random shapes/orientations
            off = rng.integers(-a.radius * 32, a.radius * 32, size=(a.edits * 2, 3))
            off = off[(off.astype(np.int64) ** 2).sum(axis=1) <= (a.radius * 32) ** 2][: a.edits]
            pos = anchor * 32 + off
            ch = np.zeros(len(pos), dtype=bxw.CHANGE_DTYPE)
            ch["x"], ch["y"], ch["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
            ids = rng.integers(1, 8, len(pos)).astype(np.uint32)
            meta = (rng.integers(0, 24, len(pos)) * 64 + rng.integers(0, 4, len(pos))).astype(np.uint32)
            ch["to"] = (ids << 16) | meta
            ch["from"] = 0  # place into air; edits whose target is not air are no-ops (compare-and-swap) but still dirty the chunk
            parts.append(ch)
        ch = np.concatenate(parts)
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        nd = ctx.region_apply_changes(ch)
        nr = ctx.region_remesh_dirty()
        ctx.synchronize()
        lat.append(time.perf_counter() - t0)
        dirty += nr
        assert nr == nd
    if dist is not None:
        import torch
        t = torch.tensor(lat, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lat = t.tolist()
        d = torch.tensor([dirty], dtype=torch.int64, device="cuda")
        dist.all_reduce(d)
        dirty = int(d.item())
    lat = np.array(lat[4:])
    if rank == 0:
        print(json.dumps({"n_gpus": world, "workload": f"{nx}x{nz}x{ny} chunks per GPU, {a.steps} steps x {a.edits} edits per GPU in a radius-{a.radius} sphere",
                          "dirty_chunks_remeshed": dirty, "dirty_chunks_per_s": dirty / sum(lat) * len(lat) / a.steps if len(lat) else None,
                          "edit_to_mesh_latency_ms_median": float(np.median(lat) * 1e3), "edit_to_mesh_latency_ms_p95": float(np.percentile(lat, 95) * 1e3)}))
    region.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
