"""Run under torchrun on >= 2 GPUs: x-slab decomposition + NCCL halo exchange give exactly the single-region meshes.

Each rank owns a slab of hashed random blocks (a pure function of position), deliberately ZEROES the shell chunks that
face its neighbours, exchanges the boundary voxel planes, meshes, and checks its slab-boundary chunks bit-exactly
against the oracle evaluated on the true (global) neighbourhood. Also checks edits crossing a slab boundary."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ballxworld_b200 as bxw  # noqa: E402
import oracle_lib as O  # noqa: E402
from ballxworld_b200.slabs import exchange_halos, slab_partition  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ny, nz, per = 2, 2, 3
    slab = slab_partition(per * world, world, rank)
    ctx = bxw.Context(local)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)  # NCCL orders its sends / receives against the current stream
    ctx.set_stream(stream.cuda_stream)
    reg = bxw.standard_registry()
    reg.bind(ctx)
    ctx.region_create(slab.x0, 0, 0, slab.nx, ny, nz)
    ctx.region_fill_synthetic(3, 150)
    zero = np.zeros(32768, np.uint32)
    for ly in range(-1, ny + 1):
        for lz in range(-1, nz + 1):
            if slab.has_lo:
                ctx.region_upload_chunk(-1, ly, lz, zero)
            if slab.has_hi:
                ctx.region_upload_chunk(slab.nx, ly, lz, zero)
    hb = ctx.region_halo_bytes()
    bufs = {k: torch.empty(hb // 4, dtype=torch.int32, device="cuda") for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}
    exchange_halos(slab, dist, bufs, lambda f, t: ctx.region_halo_export(f, t.data_ptr()),
                   lambda f, t: ctx.region_halo_import(f, t.data_ptr()))
    av, ai = ctx.region_mesh()
    spans = ctx.region_mesh_spans()
    V, I = ctx.region_mesh_fetch(av, ai)
    oreg = O.OracleRegistry()
    bad = 0
    for ix in sorted({0, slab.nx - 1}):
        for iy in range(ny):
            for iz in range(nz):
                d = [O.random_chunk(slab.x0 + ix + dx, iy + dy, iz + dz, seed=3, void_threshold=150)
                     for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                wv, wi = O.mesh_from_dense27(oreg, d)
                s = spans[ix + slab.nx * (iz + nz * iy)]
                gv = V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
                gi = I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])]
                if gv.tobytes() != wv.tobytes() or not np.array_equal(gi, wi):
                    bad += 1
                    print(f"rank {rank}: hashed-block chunk {(ix, iy, iz)} differs: {len(gv)} vs {len(wv)} vertices", flush=True)
    ctx.region_destroy()

    # Generated terrain with chunk summaries in play: deep uniform stone and sky are skipped without being read, so the
    # shell the neighbour rank fills in through the halo import must invalidate what the generator recorded there.
    ny2, oy = 4, -2
    ctx.set_gen_ids(*bxw.StdGenerator.resolve_ids(reg))
    ctx.region_create(9 + slab.x0, oy, 0, slab.nx, ny2, nz)
    ctx.region_generate(0, 64)
    for ly in range(-1, ny2 + 1):
        for lz in range(-1, nz + 1):
            if slab.has_lo:
                ctx.region_upload_chunk(-1, ly, lz, zero)
            if slab.has_hi:
                ctx.region_upload_chunk(slab.nx, ly, lz, zero)
    exchange_halos(slab, dist, bufs2 := {k: torch.empty(ctx.region_halo_bytes() // 4, dtype=torch.int32, device="cuda")
                                         for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")},
                   lambda f, t: ctx.region_halo_export(f, t.data_ptr()), lambda f, t: ctx.region_halo_import(f, t.data_ptr()))
    av, ai = ctx.region_mesh()
    spans = ctx.region_mesh_spans()
    V, I = ctx.region_mesh_fetch(av, ai)
    g = O.StdGen(0)
    for ix in sorted({0, slab.nx - 1}):
        for iy in range(ny2):
            for iz in range(nz):
                d = [g.generate_chunk(9 + slab.x0 + ix + dx, oy + iy + dy, iz + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                wv, wi = O.mesh_from_dense27(oreg, d)
                s = spans[ix + slab.nx * (iz + nz * iy)]
                gv = V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
                gi = I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])]
                if gv.tobytes() != wv.tobytes() or not np.array_equal(gi, wi):
                    bad += 1
                    print(f"rank {rank}: terrain chunk {(ix, iy, iz)} differs: {len(gv)} vs {len(wv)} vertices", flush=True)
    # Edits across the slab boundaries (SURVEY.md section 8d cfg 5): every rank receives the same global change list and applies
    # what falls inside its own storage box, shell included -- so a change in rank r's last voxel column also lands in rank
    # r+1's shell plane and dirties rank r+1's first chunk column, with no further exchange.
    world_chunks = {}

    def chunk_of(cx, cy, cz):
        if (cx, cy, cz) not in world_chunks:
            world_chunks[(cx, cy, cz)] = g.generate_chunk(cx, cy, cz).copy()
        return world_chunks[(cx, cy, cz)]

    def voxel(x, y, z):
        return chunk_of(x >> 5, y >> 5, z >> 5), (x & 31) + 32 * (z & 31) + 1024 * (y & 31)

    changes = []
    stone = O.STANDARD_GEN_IDS[3] << 16
    for r in range(world - 1):
        bx = (9 + slab_partition(per * world, world, r).x0 + slab_partition(per * world, world, r).nx) * 32  # first voxel column of rank r+1
        for x in (bx - 1, bx):
            for z in (5, 31, 32):
                h = g.params(x, z).height
                for (y, to) in ((h, 0), (h + 2, stone)):            # dig out the surface block, float a stone block above it
                    c, k = voxel(x, y, z)
                    changes.append((x, y, z, int(c[k]), to))
                    c[k] = to
    nd = ctx.region_apply_changes(np.array(changes, dtype=bxw.CHANGE_DTYPE))
    n_re = ctx.region_remesh_dirty()
    spans = ctx.region_mesh_spans()
    av = int((spans["v_off"] + spans["v_count"]).max())
    ai = int((spans["i_off"] + spans["i_count"]).max())
    V, I = ctx.region_mesh_fetch(av, ai)
    edited_seen = 0
    for ix in sorted({0, slab.nx - 1}):
        for iy in range(ny2):
            for iz in range(nz):
                d = [chunk_of(9 + slab.x0 + ix + dx, oy + iy + dy, iz + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                wv, wi = O.mesh_from_dense27(oreg, d)
                s = spans[ix + slab.nx * (iz + nz * iy)]
                gv = V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
                gi = I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])]
                if gv.tobytes() != wv.tobytes() or not np.array_equal(gi, wi):
                    bad += 1
                    print(f"rank {rank}: edited terrain chunk {(ix, iy, iz)} differs: {len(gv)} vs {len(wv)} vertices", flush=True)
    if world > 1 and (nd == 0 or n_re != nd):
        bad += 1
        print(f"rank {rank}: edits dirtied {nd} chunks, re-meshed {n_re}", flush=True)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MULTIGPU_CHECK", "OK" if int(t.item()) == 0 else f"FAILED ({int(t.item())} chunks differ)", f"world={world}")
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
