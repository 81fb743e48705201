"""Host-side mirror of the reference types and functions on the chunk path, over the CUDA C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to the reference tree):
  VoxelDatum / UncompressedChunk / VChunk          bxw_world/src/lib.rs:173-261, 549-582
  StdMeta                                          bxw_world/src/blocks/stdshapes.rs:8-44
  VoxelRegistry / register_standard_blocks         bxw_world/src/voxregistry.rs:94-150, bxw_world/src/blocks/mod.rs:6-65
  StdGenerator::{new, generate_chunk}              bxw_world/src/stdgen.rs:297-374
  mesh_from_chunk / is_chunk_trivial / ChunkBuffers src/client/render/voxmesh.rs:16-190, voxrender.rs:34-37
  iter_neighbors                                   bxw_world/src/worldmgr.rs:748-759
All compute happens in libbxw_cuda.so; nothing here evaluates noise, RLE or geometry on the CPU.
"""
import os
from dataclasses import dataclass, field

import itertools

import numpy as np

from .capi import Context

CHUNK_DIM = 32
CHUNK_DIM2 = CHUNK_DIM * CHUNK_DIM
CHUNK_DIM3 = CHUNK_DIM * CHUNK_DIM * CHUNK_DIM

_default_ctx = None


def default_context():
    """Process-wide context on cuda:LOCAL_RANK (one process per GPU)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


class VoxelDatum:
    """u32 = id << 16 | meta (lib.rs:173-206)."""

    @staticmethod
    def new(id, meta=0):
        return ((int(id) & 0xFFFF) << 16) | (int(meta) & 0xFFFF)

    @staticmethod
    def id(datum):
        return (int(datum) >> 16) & 0xFFFF

    @staticmethod
    def meta(datum):
        return int(datum) & 0xFFFF


class StdMeta:
    """meta = orientation * 64 + shape (stdshapes.rs:8-44)."""

    SHAPE_CUBE, SHAPE_SLOPE, SHAPE_CORNER, SHAPE_INNER_CORNER = 0, 1, 2, 3

    @staticmethod
    def from_parts(shape, orientation):
        if shape >= 64 or orientation >= 24:
            return None
        return orientation * 64 + shape

    @staticmethod
    def from_meta(meta):
        return meta % 64, meta // 64  # (shape, orientation)


@dataclass(frozen=True)
class ChunkPosition:
    x: int
    y: int
    z: int

    def __add__(self, o):
        return ChunkPosition(self.x + o.x, self.y + o.y, self.z + o.z)


def iter_neighbors(cpos, include_self):
    """worldmgr.rs:748-759: y outer, z, x inner."""
    for y in range(cpos.y - 1, cpos.y + 2):
        for z in range(cpos.z - 1, cpos.z + 2):
            for x in range(cpos.x - 1, cpos.x + 2):
                if include_self or (x, y, z) != (cpos.x, cpos.y, cpos.z):
                    yield ChunkPosition(x, y, z)


@dataclass
class UncompressedChunk:
    """lib.rs:222-235: blocks_yzx[x + 32 z + 1024 y], zero-initialised, dirty = 1."""

    position: ChunkPosition = ChunkPosition(0, 0, 0)
    blocks_yzx: np.ndarray = field(default_factory=lambda: np.zeros(CHUNK_DIM3, dtype=np.uint32))
    dirty: int = 1


class VChunk:
    """lib.rs:237-261, 549-582: RLE-compressed chunk. Default data is the placeholder [0, 0, 32768]."""

    def __init__(self, position=ChunkPosition(0, 0, 0), vox=None):
        self.position = position
        self.vox = np.array([0, 0, CHUNK_DIM3], dtype=np.uint32) if vox is None else np.asarray(vox, dtype=np.uint32)

    def compress(self, from_chunk, ctx=None):
        """VChunk::compress (lib.rs:555-559), compress_rle on the device."""
        assert self.position == from_chunk.position
        ctx = ctx or default_context()
        words, _ = ctx.rle_compress(from_chunk.blocks_yzx.reshape(1, -1))
        self.vox = words

    def decompress(self, ctx=None):
        """VChunk::decompress (lib.rs:562-567); raises on an invalid stream like the reference's expect()."""
        ctx = ctx or default_context()
        dense = ctx.rle_decompress(self.vox, np.array([0, self.vox.size], dtype=np.uint64))
        return UncompressedChunk(self.position, dense[0].copy())

    def serialize(self):
        """WorldBlocks::serialize_data (generation.rs:158-166): the words as little-endian bytes."""
        return self.vox.astype("<u4").tobytes()

    @staticmethod
    def deserialize(position, data):
        """WorldBlocks::deserialize_data (generation.rs:168-194)."""
        if len(data) % 4 != 0:
            raise ValueError("Invalid serialized data length")
        return VChunk(position, np.frombuffer(data, dtype="<u4").copy())


@dataclass
class VoxelDefinition:
    id: int
    name: str
    mesh: str  # "None" | "CubeAndSlopes" (lib.rs:651-655)
    debug_color: tuple
    texture_mapping: tuple  # 6 ids by signed axis index (lib.rs:624-626)


class VoxelDefinitionBuilder:
    """voxregistry.rs:16-92. The id is taken when the builder is created (build_definition, :118-133), so an abandoned
    builder consumes an id exactly as in the reference."""

    def __init__(self, registry, id):
        self.registry = registry
        self.id = id
        self._name = ""
        self._mesh = "CubeAndSlopes"
        self._debug_color = (1.0, 1.0, 1.0)
        self._texture_mapping = (0,) * 6

    def name(self, v):
        self._name = str(v)
        return self

    def debug_color(self, r, g, b):
        self._debug_color = (float(r), float(g), float(b))
        return self

    def set_mesh(self, mesh):
        self._mesh = mesh
        return self

    def texture_mapping(self, t):
        self._texture_mapping = tuple(t)
        return self

    def finish(self):
        """voxregistry.rs:63-82: AlreadyExists only when the id slot is occupied; a duplicate NAME simply re-points the
        name lookup at the new definition."""
        reg = self.registry
        d = VoxelDefinition(self.id, self._name, self._mesh, self._debug_color, self._texture_mapping)
        if len(reg.definitions) <= d.id:
            reg.definitions.extend([None] * (d.id * 2 + 1 - len(reg.definitions)))
        elif reg.definitions[d.id] is not None:
            raise ValueError("AlreadyExists")
        reg.name_lut[d.name] = d.id
        reg.definitions[d.id] = d
        reg.version += 1
        return d


_registry_tokens = itertools.count(1)


class VoxelRegistry:
    """voxregistry.rs:94-150. Default registry holds core:void (id 0, mesh None)."""

    def __init__(self):
        self.definitions = []  # Vec<Option<VoxelDefinition>>: None = free slot
        self.name_lut = {}
        self.last_free_id = 0
        self.version = 0
        self.token = next(_registry_tokens)  # process-wide unique: id() values are recycled after garbage collection
        self.build_definition().name("core:void").debug_color(0.0, 0.0, 0.0).set_mesh("None").finish()

    def build_definition(self):
        b = VoxelDefinitionBuilder(self, self.last_free_id)
        self.last_free_id += 1
        return b

    def register(self, name, mesh="CubeAndSlopes", debug_color=(1.0, 1.0, 1.0), texture_mapping=(0,) * 6):
        """build_definition().name(..)...finish() in one call."""
        return self.build_definition().name(name).set_mesh(mesh).debug_color(*debug_color).texture_mapping(texture_mapping).finish()

    def get_definition_from_id(self, id):
        d = self.definitions[id] if 0 <= id < len(self.definitions) else None
        if d is None:
            raise IndexError(f"no definition for block id {id}")  # the reference's unwrap() panic (voxregistry.rs:141-143)
        return d

    def get_definition_from_datum(self, datum):
        return self.get_definition_from_id(VoxelDatum.id(datum))

    def get_definition_from_name(self, name):
        i = self.name_lut.get(name)
        return None if i is None else self.definitions[i]

    def as_defs(self):
        n = max((d.id for d in self.definitions if d is not None), default=-1) + 1
        out = []
        for d in self.definitions[:n]:
            if d is None:
                out.append(dict(present=0, mesh_kind=0, debug_color=(0.0, 0.0, 0.0), tex=(0,) * 6))
            else:
                out.append(dict(present=1, mesh_kind=0 if d.mesh == "None" else 1, debug_color=d.debug_color, tex=d.texture_mapping))
        return out

    def bind(self, ctx):
        key = (self.token, self.version)
        if getattr(ctx, "_bound_registry", None) != key:
            ctx.set_registry(self.as_defs())
            ctx._bound_registry = key


def _tsb(top, side, bottom):
    return (side, side, bottom, top, side, side)  # TextureMapping::new_tsb (lib.rs:602-606)


def register_standard_blocks(vxreg, texmapper):
    """blocks/mod.rs:6-65."""
    t = texmapper
    vxreg.register("core:grass", texture_mapping=_tsb(t("grass_top"), t("dirt_grass"), t("dirt")))
    vxreg.register("core:snow_grass", texture_mapping=_tsb(t("snow"), t("dirt_snow"), t("dirt")))
    vxreg.register("core:dirt", texture_mapping=(t("dirt"),) * 6)
    vxreg.register("core:stone", texture_mapping=(t("stone"),) * 6)
    vxreg.register("core:diamond_ore", texture_mapping=(t("stone_diamond"),) * 6)
    vxreg.register("core:debug", texture_mapping=tuple(t(n) for n in ("dbg_left", "dbg_right", "dbg_down", "dbg_up", "dbg_front", "dbg_back")))
    vxreg.register("core:table", texture_mapping=_tsb(t("table"), t("wood"), t("table")))


# Texture ids of the client path: rank of the name among the byte-wise sorted keys of res/manifest.toml
# (src/client/render/resources.rs:149-182 iterates a toml 0.5 BTreeMap). Only the names the standard blocks use.
CLIENT_TEXTURE_IDS = {
    "dbg_back": 9, "dbg_down": 10, "dbg_front": 11, "dbg_left": 12, "dbg_right": 13, "dbg_up": 14, "dirt": 15,
    "dirt_grass": 16, "dirt_snow": 18, "grass_top": 29, "snow": 55, "stone": 56, "stone_diamond": 61, "table": 73,
    "wood": 90,
}


def standard_registry(texmapper=None):
    reg = VoxelRegistry()
    register_standard_blocks(reg, texmapper or (lambda n: CLIENT_TEXTURE_IDS[n]))
    return reg


class StdGenerator:
    """stdgen.rs:285-375."""

    def __init__(self, seed=0, ctx=None):
        self.seed = seed
        self.ctx = ctx

    @staticmethod
    def resolve_ids(registry):
        ids = []
        for name, what in (("core:void", "air"), ("core:grass", "grass"), ("core:dirt", "dirt"),
                           ("core:stone", "stone"), ("core:snow_grass", "snow grass")):
            d = registry.get_definition_from_name(name)
            if d is None:  # stdgen.rs:307-326 .expect(...)
                raise LookupError(f"No standard {what} block definition found")
            ids.append(d.id)
        return ids

    def generate_chunk(self, chunk, registry):
        ctx = self.ctx or default_context()
        ctx.set_gen_ids(*self.resolve_ids(registry))
        p = chunk.position
        ctx.gen_chunk(self.seed, p.x, p.y, p.z, out=chunk.blocks_yzx)


@dataclass
class ChunkBuffers:
    """voxrender.rs:34-37."""

    vertices: np.ndarray
    indices: np.ndarray


def is_chunk_trivial(chunk, registry, ctx=None):
    """voxmesh.rs:16-25."""
    ctx = ctx or default_context()
    registry.bind(ctx)
    return ctx.is_chunk_trivial(chunk.vox)


def mesh_from_chunk(registry, chunks, _texture_dim=None, ctx=None):
    """voxmesh.rs:45-190. chunks: 27 VChunk in iter_neighbors(cpos, True) order. Always returns ChunkBuffers."""
    assert len(chunks) == 27
    ctx = ctx or default_context()
    registry.bind(ctx)
    v, i = ctx.mesh_chunk27_rle([c.vox for c in chunks])
    return ChunkBuffers(v, i)
