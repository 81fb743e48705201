"""ballxworld_b200 — B200-native chunk pipeline (generate -> GPU-resident store -> mesh) of BallX World.

Host-side mirror of the reference's generator / mesher / chunk-handler interfaces over the C ABI of
csrc/libbxw_cuda.so (include/bxw_cuda.h). There is no CPU fallback: importing the bindings without the built
CUDA library raises, and every compute call runs sm_100a kernels.
"""
from .capi import ANCHOR_DTYPE, DELTA_DTYPE, BxwError, Context, VERTEX_DTYPE, SPAN_DTYPE, CHANGE_DTYPE, library_path  # noqa: F401
from .world import (  # noqa: F401
    CHUNK_DIM,
    CHUNK_DIM3,
    ChunkBuffers,
    ChunkPosition,
    StdGenerator,
    StdMeta,
    UncompressedChunk,
    VChunk,
    VoxelDatum,
    VoxelRegistry,
    is_chunk_trivial,
    iter_neighbors,
    mesh_from_chunk,
    register_standard_blocks,
    standard_registry,
)
from .region import Region  # noqa: F401
