"""Batched, GPU-resident form of the chunk handlers.

The reference schedules one task per chunk through ChunkDataHandler::create_chunk_update_task
(bxw_world/src/worldmgr.rs:49-77; WorldBlocks generation.rs:96-148; MeshDataHandler voxrender.rs:224-418) and
applies edits with World::apply_voxel_changes (worldmgr.rs:312-357). A Region does the same work for a whole box of
chunks per call, keeping block storage and meshes on the device.
"""
import numpy as np

from .capi import CHANGE_DTYPE
from .world import ChunkBuffers, StdGenerator, default_context


class Region:
    def __init__(self, origin, dims, registry, ctx=None):
        """origin: world chunk coordinate of interior chunk (0,0,0); dims: (nx, ny, nz) interior chunks."""
        self.ctx = ctx or default_context()
        self.origin = tuple(origin)
        self.dims = tuple(dims)
        self.registry = registry
        registry.bind(self.ctx)
        self.ctx.region_create(*self.origin, *self.dims)
        self.arena = (0, 0)

    def generate(self, seed=0, precision=64):
        self.ctx.set_gen_ids(*StdGenerator.resolve_ids(self.registry))
        self.ctx.region_generate(seed, precision)

    def generate_part(self, part, seed=0, precision=64):
        """part 1: the chunk columns next to the slab's x faces (what the halo export reads); part 2: the rest. A multi-GPU
        step starts the halo transfer between the two so that it overlaps part 2 (slabs.start_exchange / finish_exchange)."""
        if part == 1:
            self.ctx.set_gen_ids(*StdGenerator.resolve_ids(self.registry))
        self.ctx.region_generate_part(seed, precision, part)

    def mesh(self):
        self.registry.bind(self.ctx)
        self.arena = self.ctx.region_mesh()
        return self.arena

    def mesh_sphere(self, anchor, radius):
        """Mesh the chunks within `radius` chunks of the load anchor (CLoadAnchor, worldmgr.rs:734-746), nearest first."""
        self.registry.bind(self.ctx)
        n, av, ai = self.ctx.region_mesh_sphere(anchor, radius)
        self.arena = (av, ai)
        return n

    def generate_and_mesh(self, seed=0, precision=64):
        self.registry.bind(self.ctx)
        self.ctx.set_gen_ids(*StdGenerator.resolve_ids(self.registry))
        self.arena = self.ctx.region_generate_and_mesh(seed, precision)
        return self.arena

    def spans(self):
        return self.ctx.region_mesh_spans()

    def fetch(self):
        return self.ctx.region_mesh_fetch(*self.arena)

    def chunk_buffers(self, spans, vertices, indices, ix, iy, iz):
        nx, ny, nz = self.dims
        s = spans[ix + nx * (iz + nz * iy)]
        return ChunkBuffers(vertices[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])],
                            indices[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])])

    def upload_chunk(self, lx, ly, lz, blocks):
        self.ctx.region_upload_chunk(lx, ly, lz, blocks)

    def download_chunk(self, lx, ly, lz):
        return self.ctx.region_download_chunk(lx, ly, lz)

    def get_data(self, lx, ly, lz):
        """WorldBlocks::get_data (generation.rs:80-84): the chunk as a VChunk, compressed on the device."""
        from .world import ChunkPosition, VChunk
        words, _ = self.ctx.region_export_rle([(lx, ly, lz)])
        o = self.ctx.region_origin
        return VChunk(ChunkPosition(o[0] + lx, o[1] + ly, o[2] + lz), words.copy())

    def serialize_data(self, lpos):
        """WorldBlocks::serialize_data (generation.rs:158-166) for a batch: one little-endian save blob per chunk."""
        words, offs = self.ctx.region_export_rle(lpos)
        return [words[int(offs[i]):int(offs[i + 1])].astype("<u4").tobytes() for i in range(len(offs) - 1)]

    def deserialize_data(self, lpos, blobs):
        """WorldBlocks::deserialize_data (generation.rs:168-194) for a batch of save blobs; decoded on the device."""
        for b in blobs:
            if len(b) % 4:
                raise ValueError("Invalid serialized data length")
        arrs = [np.frombuffer(b, dtype="<u4").astype(np.uint32) for b in blobs]
        offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([a.size for a in arrs])
        words = np.concatenate(arrs) if arrs else np.zeros(0, np.uint32)
        self.ctx.region_import_rle(lpos, words, offs)

    def apply_voxel_changes(self, changes):
        """changes: iterable of (x, y, z, from, to) global block positions, applied in order (worldmgr.rs:312-357)."""
        ch = np.array([tuple(c) for c in changes], dtype=CHANGE_DTYPE) if not isinstance(changes, np.ndarray) else changes
        return self.ctx.region_apply_changes(ch)

    def remesh_dirty(self):
        """Re-mesh the chunks the edits made dirty (worldmgr.rs:335-356). Meshes that no longer fit their old arena space move
        to the arena's end (or the arena is compacted), so the extent is refreshed: fetch() / chunk_buffers() after this call
        see the new meshes."""
        self.registry.bind(self.ctx)
        n = self.ctx.region_remesh_dirty()
        self.arena = self.ctx.region_arena_extent()
        return n

    def set_origin(self, origin):
        """Re-position the box (slab streaming / sliding window); the voxels are stale until the next generate."""
        self.ctx.region_set_origin(*origin)
        self.origin = tuple(origin)

    # ---- load anchors: World::check_load_deltas / recalculate_load_deltas / main_loop_tick (worldmgr.rs:407-732) ----
    def unload_all(self):
        self.ctx.region_unload_all()
        self.arena = (0, 0)

    def update_anchors(self, anchors):
        """anchors: (cx, cy, cz, radius, load_mesh) per CLoadAnchor entity -> (n_unload_deltas, n_load_deltas)."""
        return self.ctx.region_update_anchors(anchors)

    def tick(self, anchors, seed=0, precision=64, max_deltas=0):
        """One World::main_loop_tick for the box: rebuild the load deltas for the anchors and process them (nearest first):
        unloads, block loads (generator), mesh loads. Meshes of newly generated chunks follow on the next tick, as in the
        reference (their dependency is checked against the state before the tick)."""
        self.registry.bind(self.ctx)
        self.ctx.set_gen_ids(*StdGenerator.resolve_ids(self.registry))
        self.ctx.region_update_anchors(anchors)
        st = self.ctx.region_apply_deltas(seed, precision, max_deltas)
        self.arena = self.ctx.region_arena_extent()
        return st

    def close(self):
        self.ctx.region_destroy()
