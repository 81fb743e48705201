"""ctypes binding of oracle/libbxw_oracle.so — the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under ballxworld_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "libbxw_oracle.so")

CHUNK_DIM3 = 32 * 32 * 32

VERTEX_DTYPE = np.dtype(
    [
        ("position", "<u4"),
        ("color", "<u4"),
        ("texcoord", "<f4", (3,)),
        ("index", "<i4"),
        ("barycentric_color_offset", "<f4", (4,)),
    ]
)
assert VERTEX_DTYPE.itemsize == 40


class BlockDef(C.Structure):
    _fields_ = [
        ("present", C.c_uint8),
        ("mesh_kind", C.c_uint8),
        ("debug_color", C.c_float * 3),
        ("tex", C.c_uint32 * 6),
    ]


class Registry(C.Structure):
    _fields_ = [("n", C.c_uint32), ("defs", C.POINTER(BlockDef))]


class Mesh(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("n_vertices", C.c_size_t),
        ("cap_vertices", C.c_size_t),
        ("indices", C.c_void_p),
        ("n_indices", C.c_size_t),
        ("cap_indices", C.c_size_t),
    ]


class VoxelParams(C.Structure):
    _fields_ = [
        ("height", C.c_int32),
        ("elevation", C.c_int32),
        ("height_f64", C.c_double),
        ("ec", C.c_double),
        ("cec", C.c_double),
    ]


class BaselineResult(C.Structure):
    _fields_ = [
        ("gen_seconds", C.c_double),
        ("mesh_seconds", C.c_double),
        ("chunks", C.c_uint64),
        ("gen_chunks", C.c_uint64),
        ("total_vertices", C.c_uint64),
        ("total_indices", C.c_uint64),
        ("nontrivial", C.c_uint64),
        ("threads", C.c_int),
    ]


class RleIter(C.Structure):
    _fields_ = [
        ("rem", C.c_void_p),
        ("rem_len", C.c_size_t),
        ("state", C.c_int),
        ("prev", C.c_uint32),
        ("what", C.c_uint32),
        ("count", C.c_uint32),
        ("pos", C.c_uint32),
    ]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "libbxw_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u32p = C.POINTER(C.c_uint32)
    L.bxo_lh_cross.restype = C.c_int
    L.bxo_orient_from_index.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.bxo_orient_matrix.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.bxo_standard_registry.argtypes = [C.POINTER(BlockDef)]
    L.bxo_compress_rle.argtypes = [u32p, C.c_size_t, u32p]
    L.bxo_compress_rle.restype = C.c_size_t
    L.bxo_decompress_rle.argtypes = [u32p, C.c_size_t, u32p]
    L.bxo_rle_iter_init.argtypes = [C.POINTER(RleIter), u32p, C.c_size_t]
    L.bxo_rle_iter_next.argtypes = [C.POINTER(RleIter), u32p, u32p]
    L.bxo_rle_iter_skip_until.argtypes = [C.POINTER(RleIter), C.c_uint32, u32p, u32p]
    L.bxo_mesh_free.argtypes = [C.POINTER(Mesh)]
    L.bxo_is_chunk_trivial.argtypes = [C.POINTER(Registry), u32p, C.c_size_t]
    L.bxo_mesh_from_chunk.argtypes = [C.POINTER(Registry), C.POINTER(u32p), C.POINTER(C.c_size_t), C.POINTER(Mesh)]
    L.bxo_mesh_from_dense27.argtypes = [C.POINTER(Registry), C.POINTER(u32p), C.POINTER(Mesh)]
    L.bxo_permutation_table.argtypes = [C.c_uint32, C.POINTER(C.c_uint8)]
    L.bxo_super_simplex_2d.argtypes = [C.POINTER(C.c_uint8), C.c_double, C.c_double]
    L.bxo_super_simplex_2d.restype = C.c_double
    L.bxo_stdgen_new.argtypes = [C.c_uint64]
    L.bxo_stdgen_new.restype = C.c_void_p
    L.bxo_stdgen_free.argtypes = [C.c_void_p]
    L.bxo_stdgen_perm.argtypes = [C.c_void_p, C.c_int]
    L.bxo_stdgen_perm.restype = C.POINTER(C.c_uint8)
    L.bxo_calc_voxel_params.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(VoxelParams)]
    L.bxo_generate_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint16), u32p]
    L.bxo_tg_simplex_new.argtypes = [C.c_uint64]
    L.bxo_tg_simplex_new.restype = C.c_void_p
    L.bxo_tg_simplex_free.argtypes = [C.c_void_p]
    L.bxo_tg_simplex_noise2.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.bxo_tg_simplex_noise2.restype = C.c_double
    L.bxo_tg_simplex_perm.argtypes = [C.c_void_p]
    L.bxo_tg_simplex_perm.restype = C.POINTER(C.c_uint16)
    L.bxo_tg_splitmix64.argtypes = [C.c_uint64, C.POINTER(C.c_uint64)]
    L.bxo_tg_splitmix64.restype = C.c_uint64
    L.bxo_splitmix64_next_u32.argtypes = [C.POINTER(C.c_uint64)]
    L.bxo_splitmix64_next_u32.restype = C.c_uint32
    L.bxo_splitmix64_next_u64.argtypes = [C.POINTER(C.c_uint64)]
    L.bxo_splitmix64_next_u64.restype = C.c_uint64
    L.bxo_xoshiro256ss_seed_from_u64.argtypes = [C.POINTER(C.c_uint64 * 4), C.c_uint64]
    L.bxo_xoshiro256ss_next_u32.argtypes = [C.POINTER(C.c_uint64 * 4)]
    L.bxo_xoshiro256ss_next_u32.restype = C.c_uint32
    L.bxo_baseline_region.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(BaselineResult)]
    _lib = L
    return L


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


# GEN_IDS order of the C ABI / oracle: void, grass, dirt, stone, snow_grass (stdgen.rs:307-326)
STANDARD_GEN_IDS = (0, 1, 3, 4, 2)


def standard_registry_defs():
    defs = (BlockDef * 8)()
    lib().bxo_standard_registry(defs)
    return defs


def defs_to_numpy(defs):
    """-> list of dicts usable by the product API's set_registry."""
    out = []
    for d in defs:
        out.append(dict(present=int(d.present), mesh_kind=int(d.mesh_kind),
                        debug_color=[float(x) for x in d.debug_color], tex=[int(x) for x in d.tex]))
    return out


class OracleRegistry:
    def __init__(self, defs=None):
        self.defs = defs if defs is not None else standard_registry_defs()
        self.reg = Registry(len(self.defs), C.cast(self.defs, C.POINTER(BlockDef)))


def compress_rle(dense):
    dense = np.ascontiguousarray(dense, dtype=np.uint32)
    out = np.empty(dense.size * 2 + 4, dtype=np.uint32)
    n = lib().bxo_compress_rle(_u32p(dense), dense.size, _u32p(out))
    return out[:n].copy()


def decompress_rle(words):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.empty(CHUNK_DIM3, dtype=np.uint32)
    rc = lib().bxo_decompress_rle(_u32p(words), words.size, _u32p(out))
    return rc, out


def rle_iterate(words):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    it = RleIter()
    lib().bxo_rle_iter_init(C.byref(it), _u32p(words), words.size)
    out = []
    d = C.c_uint32()
    i = C.c_uint32()
    while lib().bxo_rle_iter_next(C.byref(it), C.byref(d), C.byref(i)):
        out.append(d.value)
    return np.array(out, dtype=np.uint32)


def _mesh_out(m):
    nv, ni = m.n_vertices, m.n_indices
    v = np.empty(nv, dtype=VERTEX_DTYPE)
    ix = np.empty(ni, dtype=np.uint32)
    if nv:
        C.memmove(v.ctypes.data, m.vertices, nv * 40)
    if ni:
        C.memmove(ix.ctypes.data, m.indices, ni * 4)
    lib().bxo_mesh_free(C.byref(m))
    return v, ix


def mesh_from_dense27(reg, dense27):
    """dense27: sequence of 27 uint32[32768] arrays in iter_neighbors order (slot = rx + 3*rz + 9*ry)."""
    arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in dense27]
    ptrs = (C.POINTER(C.c_uint32) * 27)(*[_u32p(a) for a in arrs])
    m = Mesh()
    rc = lib().bxo_mesh_from_dense27(C.byref(reg.reg), ptrs, C.byref(m))
    if rc:
        raise RuntimeError(f"oracle mesh_from_dense27 failed rc={rc}")
    return _mesh_out(m)


def mesh_from_chunk(reg, rle27):
    arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in rle27]
    ptrs = (C.POINTER(C.c_uint32) * 27)(*[_u32p(a) for a in arrs])
    lens = (C.c_size_t * 27)(*[a.size for a in arrs])
    m = Mesh()
    rc = lib().bxo_mesh_from_chunk(C.byref(reg.reg), ptrs, lens, C.byref(m))
    if rc:
        raise RuntimeError(f"oracle mesh_from_chunk failed rc={rc}")
    return _mesh_out(m)


def is_chunk_trivial(reg, words):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    return bool(lib().bxo_is_chunk_trivial(C.byref(reg.reg), _u32p(words), words.size))


class StdGen:
    def __init__(self, seed):
        self.h = lib().bxo_stdgen_new(seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().bxo_stdgen_free(self.h)
            self.h = None

    def perm(self, which):
        p = lib().bxo_stdgen_perm(self.h, which)
        return np.ctypeslib.as_array(p, shape=(256,)).copy()

    def params(self, x, z):
        vp = VoxelParams()
        lib().bxo_calc_voxel_params(self.h, x, z, C.byref(vp))
        return vp

    def generate_chunk(self, cx, cy, cz, ids=STANDARD_GEN_IDS):
        out = np.empty(CHUNK_DIM3, dtype=np.uint32)
        idarr = (C.c_uint16 * 5)(*ids)
        lib().bxo_generate_chunk(self.h, cx, cy, cz, idarr, _u32p(out))
        return out


def super_simplex(perm, x, y):
    perm = np.ascontiguousarray(perm, dtype=np.uint8)
    return lib().bxo_super_simplex_2d(perm.ctypes.data_as(C.POINTER(C.c_uint8)), x, y)


def permutation_table(seed):
    out = np.empty(256, dtype=np.uint8)
    lib().bxo_permutation_table(seed, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def baseline_region(seed, origin, dims, threads):
    r = BaselineResult()
    rc = lib().bxo_baseline_region(seed, origin[0], origin[1], origin[2], dims[0], dims[1], dims[2], threads, C.byref(r))
    if rc:
        raise RuntimeError(f"baseline rc={rc}")
    return r


# ---------------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench: hashed random blocks incl. shapes.
# ---------------------------------------------------------------------------------------------------

def splitmix64_np(x):
    """bxw_terragen/src/math.rs:24-30 (.0 of the pair), vectorised over uint64 arrays."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def random_chunk(cx, cy, cz, seed=1, void_threshold=128):
    """SURVEY.md §8d cfg 2: datum for every voxel of chunk (cx,cy,cz) as a pure function of global position."""
    with np.errstate(over="ignore"):
        x = (np.arange(32, dtype=np.int64) + cx * 32)
        y = (np.arange(32, dtype=np.int64) + cy * 32)
        z = (np.arange(32, dtype=np.int64) + cz * 32)
        gx = (x & 0x1FFFFF).astype(np.uint64)[None, None, :]
        gy = (y & 0x1FFFFF).astype(np.uint64)[:, None, None]
        gz = (z & 0x1FFFFF).astype(np.uint64)[None, :, None]
        k = gx | (gy << np.uint64(21)) | (gz << np.uint64(42))
        r = splitmix64_np(np.uint64(seed) ^ k)
        void = (r & np.uint64(0xFF)) < np.uint64(void_threshold)
        bid = np.uint64(1) + (r >> np.uint64(8)) % np.uint64(7)
        s = (r >> np.uint64(16)) & np.uint64(15)
        shape = np.where(s < 10, 0, np.where(s < 12, 1, np.where(s < 14, 2, 3))).astype(np.uint64)
        orient = (r >> np.uint64(20)) % np.uint64(24)
        datum = (bid << np.uint64(16)) | (orient * np.uint64(64) + shape)
        datum = np.where(void, np.uint64(0), datum)
    return datum.astype(np.uint32).reshape(-1)  # [y][z][x] order = x + 32 z + 1024 y
